"""One extraction step for profilers: WARM un-profiled steps, then ONE step of orbx_extract_device on a resident batch
(device split 1, so every kernel covers the whole batch: 12 launches per step = level0, 7 x resize, FAST, octree, blur, describe).
usage: python tools/prof_step.py [--workload vga|kitti|4k] [--frames N] [--warm W] [--steps K]
  ncu ... -s $((12*W)) -c 12 python tools/prof_step.py --warm W
Prints per-stage CUDA-event times of K further steps (profiling off path: serial stages) for reference."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch

WL = {"vga": (640, 480, 1000, 256), "kitti": (1241, 376, 2000, 128), "4k": (3840, 2160, 8000, 32)}
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="vga"); ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--warm", type=int, default=2); ap.add_argument("--steps", type=int, default=0)
ap.add_argument("--split", type=int, default=1)
a = ap.parse_args()
W, H, NF, B = WL[a.workload]
B = a.frames or B
nu = min(B, 32 if W * H <= 1 << 20 else 4)
base = synth_batch(range(nu), W, H)
frames = np.ascontiguousarray(np.concatenate([base] * ((B + nu - 1) // nu))[:B])
dev = torch.device("cuda:0")
ex = _lib.Extractor(NF, 1.2, 8, 20, 7, W, H, B, 0)
ex.set_device_split(a.split)
cap = ex.capacity
d_f = torch.from_numpy(frames).to(dev)
d_k = torch.empty((B, cap, 7), dtype=torch.float32, device=dev); d_d = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
d_c = torch.empty(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def step():
    ex.extract_device(d_f.data_ptr(), W, W * H, W, H, B, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), st.cuda_stream)
for _ in range(a.warm + 1):
    step()
torch.cuda.synchronize()
print("keypoints:", int(d_c.sum().item()), "launches:", ex.launches, "build_id:", _lib.build_id())
if a.steps:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    print("ms per step (unprofiled, split %d): %.4f" % (a.split, e0.elapsed_time(e1) / a.steps))
    ex.set_profiling(True)
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    print("stage ms (level0, resize, fast, octree, blur, describe):", np.round(ex.stage_times(), 4))
