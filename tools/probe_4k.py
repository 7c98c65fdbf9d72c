"""Timing probe for the 3840x2160 / nfeatures=8000 configuration (device-resident batch)."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch
W, H, B = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 16
base = synth_batch(range(4), W, H); fr = np.ascontiguousarray(np.concatenate([base] * (B // 4)))
ex = _lib.Extractor(8000, 1.2, 8, 20, 7, W, H, B, 0); cap = ex.capacity
d = torch.from_numpy(fr).cuda(); dk = torch.empty((B, cap, 7), device='cuda'); dd = torch.empty((B, cap, 32), dtype=torch.uint8, device='cuda'); dc = torch.empty(B, dtype=torch.int32, device='cuda')
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
def run(): ex.extract_device(d.data_ptr(), W, W * H, W, H, B, dk.data_ptr(), dd.data_ptr(), dc.data_ptr(), s.cuda_stream)
for _ in range(2): run()
torch.cuda.synchronize(); ex.set_profiling(True)
t = time.perf_counter()
for _ in range(5): run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print('4K batch', B, 'ms', round(dt * 1e3, 2), 'fps', round(B / dt, 1), 'kp/frame', int(dc.sum()) / B, 'stages', dict(zip(['level0', 'resize', 'fast', 'octree', 'blur', 'describe'], ex.stage_times().round(3))))
