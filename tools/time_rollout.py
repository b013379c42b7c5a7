"""Quick device-timed rollout throughput (decisions/s) for tuning: python tools/time_rollout.py [--threads T] [--budget W] [--groups G]
[--actions rc|uniform] [--envs B] [--steps K] [--nodes N] [--chargers M].  Not the bench (no e2e, no clocks): bench.py is."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic, _lib
if os.environ.get("WRSN_LIB"):                      # a tuning build of the library
    import ctypes
    _lib._lib = _lib._bind(ctypes.CDLL(os.path.abspath(os.environ["WRSN_LIB"])))
p = argparse.ArgumentParser()
p.add_argument("--threads", type=int, default=0); p.add_argument("--budget", type=int, default=0)
p.add_argument("--groups", type=int, default=1); p.add_argument("--actions", default="rc")
p.add_argument("--envs", type=int, default=4096); p.add_argument("--steps", type=int, default=60)
p.add_argument("--pre", type=int, default=160); p.add_argument("--nodes", type=int, default=100)
p.add_argument("--chargers", type=int, default=3); p.add_argument("--topologies", type=int, default=64)
p.add_argument("--rounds", type=int, default=0)
a = p.parse_args()
dev = torch.device("cuda:0")
kw = dict(num_gateways=max(3, a.nodes // 40)) if a.nodes > 100 else {}
scs = [synthetic(num_nodes=a.nodes, num_targets=a.nodes, seed=1000 + k, **kw) for k in range(a.topologies)]
G, Bg = a.groups, a.envs // a.groups
envs = [BatchedWRSN(scs, num_agent=a.chargers, num_envs=Bg, device=dev, threads=a.threads,
                    scenario_index=(np.arange(Bg) + g * Bg) % len(scs)) for g in range(G)]
for e in envs:
    e.dims.step_budget = a.budget; e.dims.step_rounds = a.rounds
streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
obs = [torch.zeros((Bg, 4, 100, 100), dtype=torch.float32, device=dev) for _ in range(G)]
act = [torch.zeros((Bg, 3), dtype=torch.float64, device=dev) for _ in range(G)]
gen = torch.Generator(device=dev); gen.manual_seed(0)
for g in range(G):
    envs[g].reset(); envs[g].get_state(out=obs[g])
torch.cuda.synchronize()
def step(g):
    with torch.cuda.stream(streams[g]):
        if a.actions == "rc":
            o = obs[g]
            envs[g].density_map_to_action(o[:, 0] + o[:, 1] - 10.0 * o[:, 2] + o[:, 3], out=act[g])
        else:
            act[g].copy_(torch.rand((Bg, 3), dtype=torch.float64, device=dev, generator=gen)); act[g][:, 2] *= 0.05
        envs[g].rollout_step(act[g], obs[g])
def totals():
    st = torch.stack([e.req.stats.sum(0) for e in envs]).sum(0)
    return float(st[0]), float(st[1])
for k in range(a.pre):
    for g in range(G): step(g)
torch.cuda.synchronize()
d0, s0 = totals()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for st in streams: st.wait_stream(torch.cuda.current_stream())
for k in range(a.steps):
    for g in range(G): step(g)
for st in streams: torch.cuda.current_stream().wait_stream(st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); d1, s1 = totals()
print("threads=%d budget=%d rounds=%d groups=%d actions=%s nodes=%d: %.3f M decisions/s, %.3f ms/step, %.1f sim-s/decision, %.2f decisions/env-step"
      % (envs[0].dims.threads, a.budget, a.rounds, G, a.actions, a.nodes, (d1 - d0) / ms / 1e3, ms / a.steps, (s1 - s0) / max(d1 - d0, 1), (d1 - d0) / (a.steps * a.envs)), flush=True)
