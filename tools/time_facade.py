"""BASELINE.json configs[0] through the drop-in facade: ONE environment, the reference's own loop (runner/checkRL.py:33-36:
RandomController.make_action -> WRSN.step, density_map=True) on the reference's hanoi1000n50 (arrays from the committed
fixture), request dicts with host arrays.  Prints decisions/s — launch- and transfer-latency bound by construction: one
environment keeps one SM busy; the batched API is the throughput path."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_rl_wrsn_b200.wrsn import WRSN
from tests import parity_cases as pc
from tests.helpers import golden
name = sys.argv[1] if len(sys.argv) > 1 else "net_hanoi1000n50"
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
env = WRSN(pc.sc_from_golden(golden(name)), None, 3, density_map=True, device="cuda:0")
req = env.reset()
n, resets, t0 = 0, 0, time.perf_counter()
while time.perf_counter() - t0 < seconds:
    if req is None or req["terminal"]:
        req = env.reset(); resets += 1
        continue
    s = req["state"]
    req = env.step(req["agent_id"], np.copy(s[0] + s[1] - 10 * s[2] + s[3]))
    if req is not None and req["agent_id"] is not None:
        n += 1
dt = time.perf_counter() - t0
print(json.dumps(dict(metric="agent-decisions/sec", value=n / dt, unit="decisions/s", config="configs[0]: %s, 3 chargers, single environment "
                      "through the WRSN facade (request dicts with host float64 arrays, density maps decoded on the device)" % name.replace("net_", ""),
                      decisions=n, episodes=resets, seconds=dt, ms_per_decision=1e3 * dt / max(n, 1))))
