"""Three launches of the density-map decoder on 512 environments (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + k) for k in range(8)]
env = BatchedWRSN(scs, num_agent=3, num_envs=B, device=dev)
env.reset()
maps = [torch.rand((B, 100, 100), dtype=torch.float32, device=dev) for _ in range(3)]
aid = torch.zeros(B, dtype=torch.int32, device=dev)
for m in maps:
    out = env.density_map_to_action(m, agent_id=aid)
torch.cuda.synchronize()
print("ok", out[0].tolist())
