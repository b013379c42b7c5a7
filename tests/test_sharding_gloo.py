"""N > 1 path on CPU: two processes (gloo, 127.0.0.1), each advancing its shard of the environments through the host
emulation of the kernel source; the concatenation of the shards must equal the unsharded job bit for bit, and the
all-reduced rollout statistics must equal the unsharded totals."""
import os
import socket
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_rl_wrsn_b200 import _lib, synthetic
from multi_agent_rl_wrsn_b200.sharding import reduce_stats, shard_range, shard_scenario_index
from tests.helpers import REPO

EMU = os.path.join(REPO, "tests", "emu", "libwrsn_emu.so")
NUM_ENVS, STEPS, SEEDS = 7, 25, (31, 32, 33)


def _run(rank, world, lo_hi=None):
    from multi_agent_rl_wrsn_b200 import BatchedWRSN
    from tests import helpers
    helpers.use_host_build(EMU)
    scs = [synthetic(num_nodes=40, num_targets=60, seed=s) for s in SEEDS]
    lo, hi = shard_range(NUM_ENVS, rank, world)
    env = BatchedWRSN(scs, num_agent=2, num_envs=hi - lo, device="cpu",
                      scenario_index=shard_scenario_index(NUM_ENVS, len(scs), rank, world))
    rng = np.random.default_rng(0)
    acts = rng.uniform(0, 1, size=(STEPS, NUM_ENVS, 3))
    acts[..., 2] *= 0.2
    env.reset()
    trace = []
    for k in range(STEPS):
        env.rollout_step(torch.as_tensor(acts[k, lo:hi].copy()))
        trace.append((env.req.agent_id.clone(), env.req.now.clone(), env.view("energy").clone()))
    return env, trace


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    env, trace = _run(rank, world)
    dec, sim = reduce_stats(env)
    torch.save(dict(trace=trace, dec=dec, sim=sim), os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    for n in (1, 7, 4096, 16385):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_job_equals_single_process(tmp_path):
    subprocess.check_call(["make", "-C", os.path.dirname(EMU), "libwrsn_emu.so"], stdout=subprocess.DEVNULL)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    prev = _lib._lib
    try:
        env, full = _run(0, 1)
    finally:
        _lib._lib = prev
    parts = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(2)]
    for k in range(STEPS):
        for f in range(3):
            cat = torch.cat([parts[0]["trace"][k][f], parts[1]["trace"][k][f]])
            assert torch.equal(cat, full[k][f]), (k, f)
    tot = env.req.stats.sum(0)
    assert parts[0]["dec"] == parts[1]["dec"] == float(tot[0]) and parts[0]["dec"] >= NUM_ENVS * STEPS * 0.9
    assert abs(parts[0]["sim"] - float(tot[1])) < 1e-6 * float(tot[1])
