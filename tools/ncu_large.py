"""One steady-state step launch of BASELINE.json configs[4] (1000 nodes / 10 chargers, the multi-warp build gany::k_env<STEP>) for
`ncu --set full --profile-from-start off -k regex:^k_`: python tools/ncu_large.py [envs]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
big = [synthetic(num_nodes=1000, num_targets=1000, seed=1000 + k, num_gateways=25) for k in range(4)]
env = BatchedWRSN(big, num_agent=10, num_envs=B, device=dev, step_budget=100)
obs = torch.zeros((B, 4, 100, 100), dtype=torch.float32, device=dev)
env.reset(); env.get_state(out=obs)
for k in range(60):
    env.rollout_step(env.linear_controller_action(obs, (1.0, 1.0, -10.0, 1.0)), obs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
env.rollout_step(env.linear_controller_action(obs, (1.0, 1.0, -10.0, 1.0)), obs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok threads", env.dims.threads, env.counters())
