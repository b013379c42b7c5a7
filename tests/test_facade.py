"""The single-environment façade (multi_agent_rl_wrsn_b200.wrsn.WRSN) against the reference's request dicts."""
import numpy as np
import pytest
import torch

from multi_agent_rl_wrsn_b200 import _lib
from tests import parity_cases as pc
from tests.helpers import golden, mc_dict_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def cuda_library():
    from tests import helpers
    helpers.use_cuda_build()
    yield


def test_request_dict_matches_reference_episode():
    from multi_agent_rl_wrsn_b200.wrsn import WRSN
    g = golden("ep_edge_n50")
    env = WRSN(pc.sc_from_golden(g), mc_dict_of(g), int(g["num_agent"]), map_size=100, device="cuda:0")
    assert env.num_agent == 3 and env.observation_space.shape == (4, 100, 100) and env.action_space.shape == (3,)
    req = env.reset()
    full = {int(i): k for k, i in enumerate(g["full_state_idx"])}
    for i in range(int(g["n"])):
        if i > 0:
            req = env.step(int(g["fed_agent"][i]), list(g["fed_action"][i]))
        assert set(req) >= {"agent_id", "prev_state", "input_action", "action", "reward", "state", "terminal", "info", "detailed_rewards"}
        assert req["agent_id"] == int(g["agent_id"][i]) and req["terminal"] == bool(g["terminal"][i])
        assert env.env.now == float(g["now"][i])
        assert env.net.targets_active == [int(v) for v in g["targets_active"][i]]
        np.testing.assert_array_equal(req["action"], g["action"][i])
        if i > 0:
            np.testing.assert_allclose(req["reward"], float(g["reward"][i]), rtol=1e-9, atol=1e-18)
            np.testing.assert_allclose(req["detailed_rewards"][2], req["reward"])
        assert req["state"].shape == (4, 100, 100) and req["state"].dtype == np.float64
        if i in full:
            np.testing.assert_allclose(req["state"], g["full_state"][full[i]], rtol=1e-9, atol=1e-12)


def test_prev_state_is_last_state_of_that_agent():
    from multi_agent_rl_wrsn_b200.wrsn import WRSN
    g = golden("ep_basic_n50")
    env = WRSN(pc.sc_from_golden(g), mc_dict_of(g), 3, device="cuda:0")
    req = env.reset()
    want = {int(i): k for k, i in enumerate(g["full_prev_state_idx"])}
    for i in range(1, max(want) + 1):
        req = env.step(int(g["fed_agent"][i]), list(g["fed_action"][i]))
        if i in want:
            np.testing.assert_allclose(req["prev_state"], g["full_prev_state"][want[i]], rtol=1e-9, atol=1e-12)


def test_density_map_device_decode_follows_the_reference():
    """The same golden density-map episode with the map decoded on the device (wrsn_decode_density_map): the reference's
    decoded actions at 1e-6 wherever scipy's L-BFGS-B stops at (or one gradient step from) the box centre — every decision
    of this fixture — and the episode stays on the reference's trajectory."""
    from multi_agent_rl_wrsn_b200.wrsn import WRSN
    g = golden("dmap_random_n50")
    env = WRSN(pc.sc_from_golden(g), None, int(g["num_agent"]), density_map=True, device="cuda:0")
    req = env.reset()
    for i in range(int(g["n"])):
        st = req["state"]
        aid = req["agent_id"]
        assert aid == int(g["fed_agent"][i])
        # the views of `info` read one host copy of the record per request
        net, agents = req["info"]
        assert net.listNodes[3].energy == float(env._b.view("energy")[0, 3].item()) and agents[aid].status in (0, 1)
        assert len(net.targets_active) == env._b.T and net.alive in (0, 1)
        req = env.step(aid, np.copy(st[0] + st[1] - 10 * st[2] + st[3]))
        for k, name in enumerate(("ACT0", "ACT1", "ACT2")):
            np.testing.assert_allclose(env._b.mc(name)[0, aid].item(), g["action"][i][k], rtol=1e-6, atol=1e-9)
        assert req["agent_id"] == int(g["agent_id"][i])
        np.testing.assert_allclose(env.env.now, float(g["now"][i]), rtol=1e-7)
        np.testing.assert_allclose(env._b.view("energy")[0].cpu().numpy(), g["energy"][i], rtol=1e-5)


def test_batched_random_controller_rollout_equals_the_single_environment_loop():
    """runner/checkRL.py's loop (RandomController map -> step) through the batched caller-side helpers: every
    environment of a replicated batch walks exactly the trajectory of the single-environment façade with the device
    decoder (same kernels, deterministic), and the returned statistics count its decisions."""
    import torch
    from multi_agent_rl_wrsn_b200 import BatchedWRSN, BatchedRandomController, rollout
    from multi_agent_rl_wrsn_b200.wrsn import WRSN
    g = golden("dmap_random_n50")
    sc = pc.sc_from_golden(g)
    steps = 6
    single = WRSN(sc, None, 3, density_map=True, device="cuda:0")
    req = single.reset()
    nows, agents = [], []
    for _ in range(steps):
        st = req["state"]
        req = single.step(req["agent_id"], np.copy(st[0] + st[1] - 10 * st[2] + st[3]))
        nows.append(single.env.now); agents.append(req["agent_id"])
    env = BatchedWRSN(sc, num_agent=3, num_envs=5, device="cuda:0")
    env.reset()
    obs = torch.zeros((5, 4, env.S, env.S), dtype=torch.float64, device="cuda:0")     # float64 maps, as the façade feeds them
    env.get_state(out=obs)
    ctl = BatchedRandomController()
    for k in range(steps):
        obs, stats = rollout(env, ctl, 1, obs=obs)
        assert stats["decisions"] == 5.0
        assert bool((env.req.now == nows[k]).all()) and bool((env.req.agent_id == agents[k]).all())
