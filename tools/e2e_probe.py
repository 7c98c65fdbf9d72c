import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch
from orbslam_in_practice_b200 import _lib
from orbslam_in_practice_b200.synth import synth_batch
W,H,B=640,480,256
base=synth_batch(range(32)); fr=np.ascontiguousarray(np.concatenate([base]*8))
ex=_lib.Extractor(1000,1.2,8,20,7,W,H,B,0); cap=ex.capacity
hf=torch.from_numpy(fr).pin_memory(); hk=torch.empty((B,cap,7)).pin_memory(); hd=torch.empty((B,cap,32),dtype=torch.uint8).pin_memory(); hc=torch.empty(B,dtype=torch.int32).pin_memory()
def run(): ex.extract_host_ptr(hf.data_ptr(),W,W*H,W,H,B,hk.data_ptr(),hd.data_ptr(),hc.data_ptr())
for ch in (3,4,5,6,8):
    os.environ['ORBX_HOST_CHUNKS']=str(ch)
    for _ in range(3): run()
    t=time.perf_counter()
    for _ in range(10): run()
    dt=(time.perf_counter()-t)/10
    print('chunks',ch,'ms',round(dt*1e3,3),'fps',round(B/dt))
# raw copies
d=torch.empty_like(hf,device='cuda'); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): d.copy_(hf,non_blocking=True)
torch.cuda.synchronize(); print('H2D ms',(time.perf_counter()-t)/10*1e3)
dk=torch.empty((B,cap,15),dtype=torch.float32,device='cuda'); hk2=torch.empty((B,cap,15)).pin_memory(); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): hk2.copy_(dk,non_blocking=True)
torch.cuda.synchronize(); print('D2H ms',(time.perf_counter()-t)/10*1e3)
