"""Batched IPPO rollout record (controllers.IPPORollout) against the reference's roll_out loop
(controller/ippo/IPPO.py:128-210) written on the single-environment façade.  CPU: through the host emulation of the
engine (tests/emu, test infrastructure); -m gpu: through the CUDA library."""
import os
import subprocess

import numpy as np
import pytest
import torch

from multi_agent_rl_wrsn_b200 import _lib, synthetic
from multi_agent_rl_wrsn_b200.controllers import select_batch
from tests import parity_cases as pc
from tests.helpers import REPO

EMU_DIR = os.path.join(REPO, "tests", "emu")


@pytest.fixture()
def emu_library():
    subprocess.check_call(["make", "-C", EMU_DIR, "libwrsn_emu.so"], stdout=subprocess.DEVNULL)
    prev = _lib._lib
    _lib.use_library(os.path.join(EMU_DIR, "libwrsn_emu.so"))
    yield
    _lib._lib = prev


def _scenarios():
    return [synthetic(num_nodes=40, num_targets=120, seed=s, num_gateways=2) for s in (3, 4)]


def test_ippo_rollout_equals_reference_loop_emu(emu_library):
    n_tr, episodes = pc.check_ippo_rollout(_scenarios(), "cpu", num_envs=4, steps=70, with_obs=False)
    assert n_tr > 200 and episodes >= 2


def test_windows_carry_over(emu_library):
    """A record is reused window after window: carry_over on a fresh record changes nothing; after a window it starts from
    the open requests and drops the links into the finished window."""
    from multi_agent_rl_wrsn_b200 import BatchedWRSN
    from multi_agent_rl_wrsn_b200.controllers import IPPORollout
    env = BatchedWRSN(_scenarios(), num_agent=3, num_envs=4, device="cpu")
    ro = IPPORollout(env, 12, action_shape=(3,), with_obs=False)            # resets the never-reset environments itself
    first = ro.agent[0].clone()
    assert bool((first >= 0).all())
    ro.carry_over()
    assert torch.equal(ro.agent[0], first)
    pol = pc._toy_policy(np.arange(4), lambda: env.req.now, False)
    ro.collect(pol)
    n1 = sum(int(ro.transitions(i)[0].numel()) for i in range(3))
    last_agent, last_now = ro.agent[12].clone(), ro.now[12].clone()
    ro.carry_over()
    assert torch.equal(ro.agent[0], last_agent) and torch.equal(ro.now[0], last_now)
    assert int((ro.link >= 0).sum()) == 0 and int((ro.last >= 0).sum()) == 0
    ro.collect(pol)
    n2 = sum(int(ro.transitions(i)[0].numel()) for i in range(3))
    assert n1 > 0 and n2 > 0
    for i in range(3):
        t, b, tp = ro.transitions(i)
        assert bool((tp < t).all()) and bool((ro.agent[tp, b] == i).all()) and bool((ro.agent[t, b] == i).all())


@pytest.mark.gpu
def test_ippo_rollout_equals_reference_loop_gpu():
    n_tr, episodes = pc.check_ippo_rollout(_scenarios(), "cuda", num_envs=6, steps=90)
    assert n_tr > 400 and episodes >= 2


def test_select_batch():
    """IPPO.py:193-200: the half with the most unusual rewards, the rest drawn from the first len - half indices."""
    g = torch.Generator().manual_seed(0)
    r = torch.randn(300, generator=g)
    idx = select_batch(r, 64, generator=g)
    assert idx.shape == (64,)
    top = np.argsort(np.abs(r.numpy() - r.numpy().mean()), kind="stable")[-32:]
    assert np.array_equal(idx[:32].numpy(), top)
    rest = idx[32:].numpy()
    assert len(set(rest.tolist())) == 32 and rest.max() < 300 - 32 and rest.min() >= 0
    with pytest.raises(ValueError):
        select_batch(torch.zeros(40), 64)
