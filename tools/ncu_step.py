"""Short steady-state rollout for ncu: B environments, `pre` untimed rollout steps, then `n` more (profile the last ones with -s/-c)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + k) for k in range(64)]
env = BatchedWRSN(scs, num_agent=3, num_envs=B, device=dev, threads=int(os.environ.get("WRSN_THREADS", "0")))
env.dims.step_budget = int(os.environ.get("WRSN_BUDGET", "0"))
env.dims.step_rounds = int(os.environ.get("WRSN_ROUNDS", "0"))
FUSED = os.environ.get("WRSN_FUSED", "1") == "1"
obs = torch.zeros((B, 4, 100, 100), dtype=torch.float32, device=dev)
env.reset()
g = torch.Generator(device=dev); g.manual_seed(0)
RC = os.environ.get("WRSN_ACTIONS", "uniform") == "rc"      # the reference's RandomController map, decoded on the device
env.get_state(out=obs)
for k in range(pre + n):
    if RC:
        a = env.linear_controller_action(obs, (1.0, 1.0, -10.0, 1.0)) if FUSED else env.density_map_to_action(obs[:, 0] + obs[:, 1] - 10.0 * obs[:, 2] + obs[:, 3])
    else:
        a = torch.rand((B, 3), dtype=torch.float64, device=dev, generator=g); a[:, 2] *= 0.05
    env.rollout_step(a, obs)
torch.cuda.synchronize()
print("ok", env.counters())
