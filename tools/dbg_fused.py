import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
DEV="cuda:0"
scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + k) for k in range(4)]
B=512
env = BatchedWRSN(scs, num_agent=3, num_envs=B, device=DEV)
env.reset()
g = torch.Generator(device=DEV); g.manual_seed(7)
for _ in range(80):
    a = torch.rand((B, 3), dtype=torch.float64, device=DEV, generator=g); a[:, 2] *= 0.1
    env.rollout_step(a)
obs = env.get_state(dtype=torch.float32)
agents = torch.zeros(B, dtype=torch.int32, device=DEV)
dm = (obs[:, 0] + obs[:, 1] - 10.0 * obs[:, 2] + obs[:, 3]).contiguous()
o = obs.cpu().numpy()
emu = ((o[:,0] + o[:,1]) - np.float32(10.0) * o[:,2]) + o[:,3]
print("map torch-gpu vs numpy emulation: mismatching cells", int((dm.cpu().numpy() != emu).sum()), "of", emu.size)
want = env.density_map_to_action(dm, agent_id=agents)
got = env.linear_controller_action(obs, (1.0, 1.0, -10.0, 1.0), agent_id=agents)
d = (want - got).abs()
print("rows differing per component:", (d > 0).sum(0).tolist(), "max abs diff:", d.max(0).values.tolist())
bad = torch.nonzero((d > 0).any(1)).flatten()[:5].tolist()
for b in bad:
    print(b, want[b].tolist(), got[b].tolist(), "argmax map", int(dm[b].argmax()), "nan in map", bool(torch.isnan(dm[b]).any()), "inf", bool(torch.isinf(dm[b]).any()))
