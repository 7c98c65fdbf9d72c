"""Does overlapping consecutive device-resident steps (two handles, two streams) beat back-to-back steps on one stream?
usage: python tools/device_overlap_probe.py  (GPU; prints frames/s for both arrangements)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch

W, H, B, K = 640, 480, 256, 20
dev = torch.device("cuda", 0)
frames = np.ascontiguousarray(np.concatenate([synth_batch(range(32), W, H)] * (B // 32)))
d_frames = torch.from_numpy(frames).to(dev)
hs = []
for i in range(4):
    ex = _lib.Extractor(1000, 1.2, 8, 20, 7, W, H, B, 0)
    cap = ex.capacity
    hs.append((ex, torch.empty((B, cap, 7), dtype=torch.float32, device=dev), torch.empty((B, cap, 32), dtype=torch.uint8, device=dev),
               torch.empty(B, dtype=torch.int32, device=dev), torch.cuda.Stream(device=dev)))


def step(i, two):
    ex, k, d, c, s = hs[i % two] if two else hs[0]
    ex.extract_device(d_frames.data_ptr(), W, W * H, W, H, B, k.data_ptr(), d.data_ptr(), c.data_ptr(), s.cuda_stream)


for two in (0, 2, 3, 4, 0, 2, 3, 4):
    for i in range(4):
        step(i, two)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream(dev))
    for _, _, _, _, s in hs:
        s.wait_event(e0)
    for i in range(K):
        step(i, two)
    for _, _, _, _, s in hs:
        torch.cuda.default_stream(dev).wait_stream(s)
    e1.record(torch.cuda.default_stream(dev))
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%d handles/streams" % (two or 1), "%.3f ms/step  %.0f frames/s" % (ms / K, B * K / ms * 1e3), flush=True)
