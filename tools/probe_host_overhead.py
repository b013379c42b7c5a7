import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
dev=torch.device('cuda:0')
scs=[synthetic(100,100,seed=1000+k) for k in range(64)]
B=4096
env=BatchedWRSN(scs,num_agent=3,num_envs=B,device=dev)
obs=torch.zeros((B,4,100,100),dtype=torch.float32,device=dev)
env.reset(); env.get_state(out=obs)
act=torch.rand((B,3),dtype=torch.float64,device=dev)*torch.tensor([1,1,0.05],dtype=torch.float64,device=dev)
def step():
    req=env.req; aid=req.agent_id; m=aid>=0
    env.step(aid,act,mask=m); 
def sync(): torch.cuda.synchronize()
for it in range(3):
    sync(); t0=time.perf_counter(); step(); t1=time.perf_counter(); sync(); t2=time.perf_counter()
    env.get_state(out=obs); t3=time.perf_counter(); sync(); t4=time.perf_counter()
    done=env.req.agent_id<0; env.reset(mask=done); t5=time.perf_counter(); sync(); t6=time.perf_counter()
    print("step launch %.3f ms, gpu %.3f | obs launch %.3f gpu %.3f | reset launch %.3f gpu %.3f"%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3,(t6-t5)*1e3))
# full loop with one sync per step
host=torch.zeros(B,dtype=torch.int32).pin_memory()
for rep in range(2):
    sync(); t0=time.perf_counter()
    for k in range(20):
        req=env.req; aid=req.agent_id; m=aid>=0
        env.step(aid,act,mask=m); env.get_state(out=obs); done=req.agent_id<0; env.reset(mask=done)
        env.get_state(out=obs, agent_id=torch.where(done, req.agent_id, torch.full_like(aid,-1)))
        host.copy_(req.agent_id, non_blocking=True); torch.cuda.current_stream().synchronize()
    t1=time.perf_counter(); print("sync-per-step loop: %.3f ms/step"%((t1-t0)/20*1e3))
    sync(); t0=time.perf_counter()
    for k in range(20):
        req=env.req; aid=req.agent_id; m=aid>=0
        env.step(aid,act,mask=m); env.get_state(out=obs); done=req.agent_id<0; env.reset(mask=done)
        env.get_state(out=obs, agent_id=torch.where(done, req.agent_id, torch.full_like(aid,-1)))
    sync(); t1=time.perf_counter(); print("async loop: %.3f ms/step"%((t1-t0)/20*1e3))
print(env.counters())
