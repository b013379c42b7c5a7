"""TEST INFRASTRUCTURE (not collected by pytest; run by hand): random scenarios, charger counts, action scales and episode
lengths — the engine (host emulation of the kernel source, or the CUDA library with --device cuda) against the C restatement
of the reference, decision by decision, every fourth configuration with observations.

    python tests/fuzz_engine_vs_oracle.py [--device cpu|cuda] [--first-seed 1000] [--count 400] [--seconds 400]

End of round 1: seeds 1000-1399 on the emulation, 45 323 decisions compared, no difference."""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from multi_agent_rl_wrsn_b200 import _lib, synthetic  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--device", default="cpu")
    p.add_argument("--first-seed", type=int, default=1000)
    p.add_argument("--count", type=int, default=400)
    p.add_argument("--seconds", type=float, default=400.0)
    a = p.parse_args()
    if a.device == "cpu":
        from tests import helpers
        helpers.use_host_build()
    from tests import parity_cases as pc
    t0, fails, n = time.time(), 0, 0
    for seed in range(a.first_seed, a.first_seed + a.count):
        rng = np.random.default_rng(seed)
        N, T, M = int(rng.integers(20, 90)), int(rng.integers(10, 150)), int(rng.integers(1, 5))
        gw, scale2 = int(rng.integers(2, 5)), float(rng.choice([0.02, 0.1, 0.5, 1.0]))
        steps = int(rng.integers(30, 160))
        try:
            sc = synthetic(num_nodes=N, num_targets=T, seed=seed, num_gateways=gw)
            n_dec, _ = pc.check_vs_oracle(sc, a.device, num_envs=3, steps=steps, seed=seed, num_agent=M, scale2=scale2,
                                          check_obs=seed % 4 == 0)
            n += n_dec
        except AssertionError as e:
            fails += 1
            print("FAIL seed", seed, (N, T, M, gw, scale2, steps), str(e)[:300], flush=True)
        if time.time() - t0 > a.seconds:
            print("time up at seed", seed)
            break
    print("%d decisions compared, %d failing configurations, %.0f s" % (n, fails, time.time() - t0))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
