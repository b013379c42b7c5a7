"""CPU-side tests of the engine logic: the kernel source compiled single-lane for the host (tests/emu) is driven
through the same C ABI and host layer as the CUDA library and checked against the reference's golden fixtures and
the C oracle.  The parity tests proper (CUDA path) are tests/test_gpu_parity.py (-m gpu)."""
import os
import subprocess

import numpy as np
import pytest

from tests import helpers

from multi_agent_rl_wrsn_b200 import _lib, synthetic
from tests import parity_cases as pc
from tests.helpers import REPO, golden_names

EMU_DIR = os.path.join(REPO, "tests", "emu")


@pytest.fixture(scope="module", autouse=True)
def emu_library():
    subprocess.check_call(["make", "-C", EMU_DIR, "libwrsn_emu.so"], stdout=subprocess.DEVNULL)
    helpers.use_host_build()
    yield
    helpers.use_cuda_build_lazy()


@pytest.mark.parametrize("name", golden_names("net_"))
def test_pure_network_golden(name):
    exact, total = pc.check_pure_network(name, "cpu")
    assert exact == total          # energies bit-identical to the reference at every snapshot


@pytest.mark.parametrize("name", golden_names("ep_"))
def test_episode_golden(name):
    pc.check_episode(name, "cpu")


def test_batched_replicas_identical():
    pc.check_episode("ep_edge_n50", "cpu", replicas=3)


@pytest.mark.parametrize("seed", [0, 1])
def test_random_episodes_vs_oracle(seed):
    sc = synthetic(num_nodes=60, num_targets=70, seed=10 + seed)
    n_dec, cnt = pc.check_vs_oracle(sc, "cpu", num_envs=6, steps=40, seed=seed)
    assert n_dec > 100


def test_heterogeneous_scenarios_vs_oracle():
    scs = [synthetic(num_nodes=48, num_targets=48, seed=s) for s in (21, 22, 23)]
    pc.check_vs_oracle(scs, "cpu", num_envs=6, steps=25, seed=5)


def test_long_charging_and_deaths_vs_oracle():
    # long charge phases (capacity clamp, Q19) and episodes that run into node deaths (death ticks in pieces vs the oracle)
    sc = synthetic(num_nodes=40, num_targets=120, seed=3, num_gateways=2)
    n_dec, cnt = pc.check_vs_oracle(sc, "cpu", num_envs=4, steps=60, seed=9, scale2=0.5)
    assert cnt["serial_ticks"] >= 1 and cnt["split_death_ticks"] >= 1


def test_small_charger_exhaustion_vs_oracle():
    mc = dict(capacity=2500, threshold=0, velocity=5, pm=1, charging_range=27, alpha=4500, beta=30, epsilon=1e-10)
    sc = synthetic(num_nodes=50, num_targets=50, seed=4)
    pc.check_vs_oracle(sc, "cpu", num_envs=4, steps=50, seed=2, mc=mc, scale2=0.01)


def test_deaths_after_network_operate_stopped():
    """hanoi1000n50 far past the end of Network.operate: four more nodes die while the levels are stale."""
    from tests.helpers import golden
    sc = pc.sc_from_golden(golden("net_hanoi1000n50"))
    cnt, dead = pc.check_network_after_operate_stopped(sc, "cpu", horizon=12000.0, every=100.0)
    assert dead >= 5 and cnt["stale_rebuilds"] >= 4

def test_rollout_step_equals_manual_loop():
    scs = [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]
    resets = pc.check_rollout_step(scs, "cpu", num_envs=6, steps=120, seed=1, with_obs=False)
    assert resets >= 3


def test_batches_equal_event_path():
    scs = [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]
    cnt = pc.check_batches_equal_event_path(scs, "cpu", num_envs=6, steps=150, seed=1)
    assert cnt["batched_ticks"] > 0.1 * cnt["ticks"] and cnt["episode_ends"] >= 3


def test_batches_equal_event_path_short_charges():
    sc = synthetic(num_nodes=60, num_targets=70, seed=11)
    cnt = pc.check_batches_equal_event_path(sc, "cpu", num_envs=4, steps=120, seed=2, scale2=0.05)
    assert cnt["batched_ticks"] > 0.25 * cnt["ticks"]


def test_pure_network_batches():
    from tests.helpers import golden
    sc = pc.sc_from_golden(golden("net_hanoi1000n50"))
    cnt = pc.check_pure_network_batches(sc, "cpu", horizon=12500.0, every=37.0)
    assert cnt["batched_ticks"] > 0.6 * cnt["ticks"]
    # Network.operate ends at t = 4048; later deaths re-route on stale levels: the piecewise death tick against the serial one
    assert cnt["split_death_ticks"] >= 5 and cnt["stale_rebuilds"] >= 4


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_fragile_networks_die_out_identically(seed):
    """Two gateways for 120 targets: node after node dies (also in the same second, also after Network.operate has ended);
    the piecewise death tick and the batches against the event-by-event path with the plain serial tick, all the way."""
    sc = synthetic(num_nodes=40, num_targets=120, seed=seed, num_gateways=2)
    cnt = pc.check_pure_network_batches(sc, "cpu", horizon=20000.0, every=211.0)
    assert cnt["serial_ticks"] >= 2 and cnt["split_death_ticks"] >= 1


@pytest.mark.parametrize("nodes,targets,chargers,envs,steps", [(500, 500, 5, 2, 30), (1000, 1000, 10, 1, 45)])
def test_large_configs_vs_oracle(nodes, targets, chargers, envs, steps):
    sc = synthetic(num_nodes=nodes, num_targets=targets, seed=nodes, num_gateways=max(3, nodes // 40))
    n_dec, cnt = pc.check_vs_oracle(sc, "cpu", num_envs=envs, steps=steps, seed=4, num_agent=chargers)
    assert n_dec >= envs * 15


def test_charge_kernel_restatement_vs_reference_statements():
    """The emulation's wrsn_k_charge (host restatement) and the harness of the GPU test against the reference's statements."""
    scs = [synthetic(num_nodes=100, num_targets=100, seed=1000 + s) for s in range(2)]
    assert pc.check_charge_kernel(scs, "cpu", num_envs=4, steps=30, seed=2) > 20


@pytest.mark.parametrize("name", ["ep_basic_n50", "ep_map32_n50", "ep_two_mc_sonla"])
def test_emulated_raster_vs_reference_observations(name):
    """The emulation's host raster (what lets the CPU tests of the trainers' loop see observations) against the reference's
    golden get_state maps; the CUDA rasters are checked against the same fixtures in tests/test_gpu_parity.py."""
    pc.check_episode(name, "cpu", check_obs=True)


@pytest.mark.parametrize("budget,scale2,rounds", [(12, 0.3, 0), (40, 0.05, 0), (3, 0.3, 0), (12, 0.3, 1), (30, 0.1, 2), (5, 0.3, 3)])
def test_step_budget_only_cuts_steps_into_launches(budget, scale2, rounds):
    """wrsn_dims.step_budget / step_rounds: interrupted steps continue where they stopped — in the same kernel or, split by
    kind of work, alternating between the events kernel and the batch kernel; requests and records are those of the
    unbudgeted run (episodes end and restart on the way)."""
    scs = [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]
    n = pc.check_budget_equals_unbudgeted(scs, "cpu", num_envs=4, calls=60, seed=5, budget=budget, scale2=scale2, rounds=rounds)
    assert n > 20


def test_tick_kernels_across_a_death_tick():
    """Ladder L1 / L2 on the host build: the per-tick entry points, tick by tick, across hanoi1000n200's death tick (518.5)."""
    assert pc.check_tick_kernels("net_hanoi1000n200", "cpu", 512, 524) >= 1


def test_reward_tick_vs_reference_statements():
    scs = [synthetic(num_nodes=60, num_targets=70, seed=s) for s in (11, 12)]
    assert pc.check_reward_kernel(scs, "cpu", num_envs=4, steps=25, seed=3) > 10


def test_sharded_equals_unsharded_records():
    scs = [synthetic(num_nodes=40, num_targets=60, seed=s) for s in (31, 32, 33)]
    pc.check_sharded_equals_unsharded(scs, "cpu", num_envs=7, steps=40, seed=2, world=2, num_agent=2)
    pc.check_sharded_equals_unsharded(scs, "cpu", num_envs=7, steps=40, seed=2, world=3, num_agent=2, budget=10)
    pc.check_sharded_equals_unsharded(scs, "cpu", num_envs=7, steps=40, seed=2, world=2, num_agent=2, budget=20, rounds=2)


def test_sticky_flags_survive_a_reset():
    """wrsn_request.sticky: the fixture whose 700 J chargers all die ends on an implicit None; one more `step` starts with no alive
    charger (the reference would never return, Q1): `flags` bit 0.  A reset overwrites the row's `flags` — `sticky` keeps the bit,
    `BatchedWRSN.raise_on_error()` counts those environments and raises on an engine error (bit 1), also one that was reset since."""
    import torch
    from multi_agent_rl_wrsn_b200 import BatchedWRSN
    from tests.helpers import golden, mc_dict_of
    g = golden("ep_deadmc_n50")
    R = 2
    env = BatchedWRSN(pc.sc_from_golden(g), num_agent=int(g["num_agent"]), mc_type=mc_dict_of(g), num_envs=R, device="cpu")
    env.reset()
    for i in range(1, int(g["n"])):
        env.step(np.full(R, int(g["fed_agent"][i]), np.int32), np.tile(g["fed_action"][i], (R, 1)))
    assert int(env.req.agent_id[0]) == -2 and env.raise_on_error() == 0
    env.step(np.full(R, -1, np.int32), np.zeros((R, 3)))
    assert int((env.req.flags & 1).sum()) == R
    env.reset()
    assert int((env.req.flags & 1).sum()) == 0 and env.raise_on_error() == R
    env.req.sticky[1] |= 2
    with pytest.raises(RuntimeError, match="engine error"):
        env.raise_on_error()
