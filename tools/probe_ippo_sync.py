"""Diagnostic: run IPPORollout.collect under torch's sync debug mode and report host synchronisations."""
import math
import sys
import warnings

import torch

sys.path.insert(0, ".")
from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
from multi_agent_rl_wrsn_b200.controllers import IPPORollout

scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + s) for s in range(4)]
env = BatchedWRSN(scs, num_agent=3, num_envs=256, device="cuda")
env.reset()
ro = IPPORollout(env, 4)
sigma = 1e-3


def pol(agent_id, o):
    mean = o[:, 0] + o[:, 1] - 10.0 * o[:, 2] + o[:, 3]
    x = torch.randn_like(mean).mul_(sigma).add_(mean)
    lp = (-0.5 * ((x - mean) / sigma) ** 2).sum((1, 2)) - 1e4 * (math.log(sigma) + 0.5 * math.log(2.0 * math.pi))
    return x, lp


ro.collect(pol)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    ro.carry_over()
    ro.collect(pol)
torch.cuda.set_sync_debug_mode("default")
print("sync warnings:", len(w))
for x in w[:10]:
    print(x.filename, x.lineno, str(x.message)[:100])
