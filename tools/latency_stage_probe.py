"""Device time and stage times of small batches (low-latency path).  usage: python tools/latency_stage_probe.py [B ...]   (GPU)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch

W, H = 640, 480
dev = torch.device("cuda:0")
for B in [int(a) for a in sys.argv[1:]] or [1, 4, 8]:
    ex = _lib.Extractor(1000, 1.2, 8, 20, 7, W, H, B, 0)
    cap = ex.capacity
    d_f = torch.from_numpy(synth_batch(range(B), W, H)).to(dev)
    d_k = torch.empty((B, cap, 7), dtype=torch.float32, device=dev); d_d = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_c = torch.empty(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    def dstep():
        ex.extract_device(d_f.data_ptr(), W, W * H, W, H, B, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), st.cuda_stream)
    for _ in range(20):
        dstep()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        dstep(); st.synchronize()
    dt = (time.perf_counter() - t0) / 200
    ex.set_profiling(True)
    for _ in range(50):
        dstep()
    torch.cuda.synchronize()
    print("batch %2d: device path, sync each call %.1f us; stage us (level0, resize, fast, octree, blur, describe): %s"
          % (B, dt * 1e6, np.round(ex.stage_times() * 1e3, 1)), flush=True)
    ex.close()
