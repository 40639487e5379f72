import sys, time, numpy as np, torch, ctypes
sys.path.insert(0, '/root/repo')
from liodom_b200 import api, synth
B=int(sys.argv[1]) if len(sys.argv)>1 else 128
K=16
scans,_=synth.sequence("hdl64",1000,K+3)
dev=torch.device("cuda",0)
host_steps=[];host_ptrs=[];cnts=[]
for f in range(K+3):
    n=len(scans[f]); buf=torch.empty((n*B,4),dtype=torch.float32).pin_memory()
    for l in range(B): buf[l*n:(l+1)*n]=torch.from_numpy(scans[f])
    host_steps.append(buf); host_ptrs.append([buf.data_ptr()+l*n*16 for l in range(B)]); cnts.append([n]*B)
torch.cuda.synchronize()
# raw link
probe=torch.empty((max(len(h) for h in host_steps),4),dtype=torch.float32,device=dev)
for rep in range(2):
    t=time.perf_counter()
    for f in range(3,3+8): probe[:len(host_steps[f])].copy_(host_steps[f],non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
print("link alone GB/s", sum(host_steps[f].numel()*4 for f in range(3,11))/dt/1e9)
for mode in ("pipelined","sync"):
    ctx=api.Context(batch=B,prev_frames=15,max_points=131072)
    for f in range(3):
        ctx.scan_batch_ptrs(host_ptrs[f],cnts[f],16,on_device=False); ctx.results()
    t=time.perf_counter()
    for f in range(3,3+K):
        ctx.scan_batch_ptrs(host_ptrs[f],cnts[f],16,on_device=False)
        if mode=="sync": ctx.results()
        elif f>3: ctx.results(age=1)
    ctx.results(); dt=time.perf_counter()-t
    print(mode,"ms/step",dt/K*1e3,"GB/s",sum(host_steps[f].numel()*4 for f in range(3,3+K))/dt/1e9)
    ctx.close()
