"""Per-step kernel times over a long rollout (diagnostic; run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
dev = torch.device('cuda:0')
scs = [synthetic(100, 100, seed=1000 + k) for k in range(64)]
B = 4096
env = BatchedWRSN(scs, num_agent=3, num_envs=B, device=dev)
obs = torch.zeros((B, 4, 100, 100), dtype=torch.float32, device=dev)
env.reset(); env.get_state(out=obs)
gen = torch.Generator(device=dev); gen.manual_seed(1)
scale = torch.tensor([1, 1, 0.05], dtype=torch.float64, device=dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(S)]
stats = []
for k in range(S):
    act = torch.rand((B, 3), generator=gen, dtype=torch.float64, device=dev) * scale
    req = env.req; aid = req.agent_id; m = aid >= 0
    now0 = req.now.clone()
    slow0 = env.hdr("NSLOW").sum()
    evs[k][0].record(); env.step(aid, act, mask=m); evs[k][1].record()
    env.get_state(out=obs); evs[k][2].record()
    dt = (req.now - now0)
    done = req.agent_id < 0
    stats.append((dt.mean(), dt.max(), done.sum(), env.hdr("NSLOW").sum() - slow0))
    env.reset(mask=done)
    env.get_state(out=obs, agent_id=torch.where(done, req.agent_id, torch.full_like(aid, -1))); evs[k][3].record()
torch.cuda.synchronize()
for k in range(S):
    s = stats[k]
    print("step %3d: step %.3f ms obs %.3f ms rest %.3f ms | mean dt %.1f max dt %.1f done %d serial ticks %d" % (
        k, evs[k][0].elapsed_time(evs[k][1]), evs[k][1].elapsed_time(evs[k][2]), evs[k][2].elapsed_time(evs[k][3]),
        float(s[0]), float(s[1]), int(s[2]), int(s[3])))
