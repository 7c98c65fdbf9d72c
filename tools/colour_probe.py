import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch
W,H,B=640,480,256
base=synth_batch(range(32),W,H)
frames=np.ascontiguousarray(np.concatenate([base]*8))
for ch in (1,3,4):
    ex=_lib.Extractor(1000,1.2,8,20,7,W,H,B,0); ex.set_device_split(1)
    if ch>1: ex.set_input_format(ch, 1)
    f=np.repeat(frames[...,None],ch,axis=-1) if ch>1 else frames
    d_f=torch.from_numpy(np.ascontiguousarray(f)).cuda(); cap=ex.capacity
    d_k=torch.empty((B,cap,7),dtype=torch.float32,device='cuda'); d_d=torch.empty((B,cap,32),dtype=torch.uint8,device='cuda'); d_c=torch.empty(B,dtype=torch.int32,device='cuda')
    st=torch.cuda.Stream(); torch.cuda.set_stream(st)
    def step(): ex.extract_device(d_f.data_ptr(), W*ch, W*H*ch, W, H, B, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), st.cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize(); ex.set_profiling(True)
    for _ in range(10): step()
    torch.cuda.synchronize()
    print('channels',ch,'keypoints',int(d_c.sum()),'stage ms',np.round(ex.stage_times(),4))
