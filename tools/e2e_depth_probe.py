"""End-to-end throughput of orbx_extract_host_begin/_end with 2, 3 or 4 batches in flight (one handle each).
usage: python tools/e2e_depth_probe.py   (GPU)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch

W, H, B, K = 640, 480, 256, 24
frames = np.ascontiguousarray(np.concatenate([synth_batch(range(32), W, H)] * (B // 32)))
slots = []
for i in range(4):
    ex = _lib.Extractor(1000, 1.2, 8, 20, 7, W, H, B, 0)
    cap = ex.capacity
    slots.append((ex, torch.from_numpy(frames).pin_memory(), torch.empty((B, cap, 7), dtype=torch.float32).pin_memory(),
                  torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()))


def run(depth, n):
    def begin(i):
        e, hf, hk, hd, hc = slots[i % depth]
        e.extract_host_begin(hf.data_ptr(), W, W * H, W, H, B, hk.data_ptr(), hd.data_ptr(), hc.data_ptr())
    for i in range(min(depth - 1, n)):
        begin(i)
    for i in range(n):
        if i + depth - 1 < n:
            begin(i + depth - 1)
        slots[i % depth][0].extract_host_end()


for depth in (2, 3, 4, 2, 3, 4):
    run(depth, 6)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(depth, K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("depth %d: %.3f ms/step  %.0f frames/s  (H2D %.1f GB/s)" % (depth, dt / K * 1e3, B * K / dt, frames.nbytes * K / dt / 1e9), flush=True)
