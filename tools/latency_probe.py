"""Single-frame latency of the blocking host call (what ORBextractor::operator() costs per image) and small batches.
usage: python tools/latency_probe.py   (GPU)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch

W, H = 640, 480
for B in (1, 2, 4, 8, 16, 32):
    ex = _lib.Extractor(1000, 1.2, 8, 20, 7, W, H, B, 0)
    cap = ex.capacity
    hf = torch.from_numpy(synth_batch(range(B), W, H)).pin_memory()
    hk = torch.empty((B, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.empty(B, dtype=torch.int32).pin_memory()
    for _ in range(20):
        ex.extract_host_ptr(hf.data_ptr(), W, W * H, W, H, B, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        ex.extract_host_ptr(hf.data_ptr(), W, W * H, W, H, B, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
    dt = (time.perf_counter() - t0) / n
    print("batch %2d: %.1f us per call, %.1f us per frame, %.0f frames/s" % (B, dt * 1e6, dt * 1e6 / B, B / dt), flush=True)
    ex.close()

# stage breakdown of one single-frame call (library stage events) and the pure device time
ex = _lib.Extractor(1000, 1.2, 8, 20, 7, W, H, 1, 0)
cap = ex.capacity
dev = torch.device("cuda:0")
d_f = torch.from_numpy(synth_batch(range(1), W, H)).to(dev)
d_k = torch.empty((1, cap, 7), dtype=torch.float32, device=dev); d_d = torch.empty((1, cap, 32), dtype=torch.uint8, device=dev)
d_c = torch.empty(1, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def dstep():
    ex.extract_device(d_f.data_ptr(), W, W * H, W, H, 1, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), st.cuda_stream)
for _ in range(20):
    dstep()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    dstep()
e1.record(); torch.cuda.synchronize()
print("device path, 1 frame, back to back: %.1f us per call" % (e0.elapsed_time(e1) * 10))
t0 = time.perf_counter()
for _ in range(100):
    dstep(); st.synchronize()
print("device path, 1 frame, sync each call: %.1f us per call" % ((time.perf_counter() - t0) * 1e4))
ex.set_profiling(True)
for _ in range(50):
    dstep()
torch.cuda.synchronize()
print("stage us (level0, resize, fast, octree, blur, describe):", np.round(ex.stage_times() * 1e3, 1))
