import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from cubiquity_b200 import api
import bench
W,H=1920,1080
sc = api.Scene("terrain", 12, 1)
ctx = api.Context(0); ctx.upload(sc.nodes, sc.root, sc.colours)
class B: pass
b=B(); b.lower,b.upper=sc.lower,sc.upper
cam,pos,yaw = bench.orbit_camera(api,b,0)
dev=torch.device("cuda",0); stream=torch.cuda.current_stream().cuda_stream
hits=torch.zeros(W*H*10,dtype=torch.int32,device=dev)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
def run(mf, reps=12):
    for _ in range(5): ctx.raycast_frame_device(cam,W,H,hits.data_ptr(),True,mf,stream)
    torch.cuda.synchronize(); t=[]
    for _ in range(reps):
        flush.zero_()
        a,c=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); ctx.raycast_frame_device(cam,W,H,hits.data_ptr(),True,mf,stream); c.record(); torch.cuda.synchronize(); t.append(a.elapsed_time(c))
    return float(np.mean(t))
r={}
for ao in (1,0):
    ctx.set_option("adaptive_order", ao)
    for mf in (-1.0, 0.0035):
        ms=run(mf); r["adaptive_order=%d mf=%g"%(ao,mf)]={"ms":round(ms,4),"grays":round(W*H/ms/1e6,3)}
print(json.dumps(r))
