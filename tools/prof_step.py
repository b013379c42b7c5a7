"""Where do the SM cycles of a step launch go?  Needs the profiling build:
    nvcc ... -DWRSN_PROF -o multi_agent_rl_wrsn_b200/csrc/libwrsn_b200_prof.so   (tools/build_prof.sh)
Prints, per rollout step of B environments, the mean and the maximum over environments of the cycles spent in the
whole step kernel, the serial (possible death) ticks, the whole-cycle batches, BFS + routing tree, and fitness."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from multi_agent_rl_wrsn_b200 import _lib, BatchedWRSN, synthetic
VARIANT = int(sys.argv[3]) if len(sys.argv) > 3 else 1
import ctypes
_lib._lib = _lib._bind(ctypes.CDLL(os.path.join(REPO, "multi_agent_rl_wrsn_b200", "csrc", "libwrsn_b200_prof%d.so" % VARIANT)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
NODES = int(sys.argv[4]) if len(sys.argv) > 4 else 100
CHARGERS = int(sys.argv[5]) if len(sys.argv) > 5 else 3
scs = [synthetic(num_nodes=NODES, num_targets=NODES, seed=1000 + k, num_gateways=max(3, NODES // 40)) if NODES > 100 else
       synthetic(num_nodes=NODES, num_targets=NODES, seed=1000 + k) for k in range(64 if NODES <= 100 else 8)]
env = BatchedWRSN(scs, num_agent=CHARGERS, num_envs=B, device="cuda:0", step_budget=int(os.environ.get("WRSN_BUDGET", "0")))
env.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
names = ["total", "serial", "batch", "bfs", "fitness"] if VARIANT == 1 else (
    ["total", "charger_events", "lazy_replay", "grid_events", "slot_scan"] if VARIANT == 2 else
    ["total", "batch_pass1", "batch_all_at_once", "batch_cycle_loop", "update_reward_in_loop"])
idx = [env.E["WRSN_H_PROF%d" % k] for k in range(5)]
acc = []
RC = os.environ.get("WRSN_ACTIONS", "uniform") == "rc"      # the reference's RandomController map, decoded on the device
obs = torch.zeros((B, 4, 100, 100), dtype=torch.float32, device="cuda:0")
env.get_state(out=obs)
for k in range(steps):
    if RC:
        a = env.density_map_to_action(obs[:, 0] + obs[:, 1] - 10.0 * obs[:, 2] + obs[:, 3])
    else:
        a = torch.rand((B, 3), dtype=torch.float64, device="cuda:0", generator=g); a[:, 2] *= 0.05
    env.rollout_step(a, obs if RC else None)
    if k >= steps - 50:
        h = env.view("hdr")[:, idx].cpu().numpy()
        acc.append(h)
acc = np.stack(acc)            # [steps, B, 5]
print("per-step mean over envs (kcycles):", dict(zip(names, np.round(acc.mean((0, 1)) / 1e3, 1))))
print("per-step MAX over envs, averaged over steps (kcycles):", dict(zip(names, np.round(acc.max(1).mean(0) / 1e3, 1))))
tot = acc[..., 0]
print("total kcycles percentiles 50/90/99/99.9/max:", np.round(np.percentile(tot, [50, 90, 99, 99.9, 100]) / 1e3, 1))
srt = np.argsort(-tot.reshape(-1))[:10]
print("slowest rows (kcycles, same columns):")
for r in srt:
    print("   ", np.round(acc.reshape(-1, 5)[r] / 1e3, 1))
print(env.counters())
