"""The C-ABI library loads on a box without a GPU and exports every symbol include/wrsn_b200.h declares
(no compute calls here); layouts are consistent; the host layer refuses to run without a device."""
import ctypes
import os
import re

import pytest

from multi_agent_rl_wrsn_b200 import _lib
from tests.helpers import REPO


def _declared_functions():
    with open(os.path.join(REPO, "include", "wrsn_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(wrsn_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def cuda_so():
    import __graft_entry__ as g
    g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_exports_every_declared_symbol(cuda_so):
    names = _declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(cuda_so, n), n


def test_emulation_exports_the_same_abi():
    emu = ctypes.CDLL(os.path.join(REPO, "tests", "emu", "libwrsn_emu.so"))
    for n in _declared_functions():
        assert hasattr(emu, n), n


def test_layout_and_enums(cuda_so):
    L = _lib._bind(cuda_so)
    e = _lib.enums()
    d = _lib.Dims()
    d.B, d.N, d.T, d.M, d.S, d.Emax, d.TEmax, d.n_scen = 4096, 100, 100, 3, 100, 240, 130, 4
    assert L.wrsn_dims_finalize(ctypes.byref(d)) == 0
    assert d.Npad == 112 and d.W == 4 and d.Tw == 4 and d.threads == 32
    off = (ctypes.c_int64 * e["WRSN_F_COUNT"])()
    assert L.wrsn_state_layout(ctypes.byref(d), off) == 0
    offs = list(off)
    assert offs == sorted(offs) and all(o % 16 == 0 for o in offs)
    assert offs[e["WRSN_F_LOGTICK"]] == d.state_resident_bytes and d.state_bytes > d.state_resident_bytes
    assert d.smem_bytes <= 227 * 1024
    soff = (ctypes.c_int64 * e["WRSN_S_COUNT"])()
    assert L.wrsn_scen_layout(ctypes.byref(d), soff) == 0
    assert list(soff) == sorted(soff)
    bad = _lib.Dims()
    bad.B, bad.N, bad.T, bad.M, bad.S = 1, 100, 100, 99, 100
    assert L.wrsn_dims_finalize(ctypes.byref(bad)) == 0
    assert L.wrsn_state_layout(ctypes.byref(bad), off) != 0            # M > WRSN_MAX_MC is rejected
    assert b"bad dims" in L.wrsn_last_error()


def test_no_cpu_path():
    """Without a GPU the product host layer must refuse loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multi_agent_rl_wrsn_b200 import BatchedWRSN, synthetic
    from tests import helpers
    helpers.use_cuda_build_lazy()
    with pytest.raises(RuntimeError):
        BatchedWRSN(synthetic(num_nodes=20, num_targets=20, seed=0), num_agent=1, num_envs=1)
    with pytest.raises(RuntimeError):
        BatchedWRSN(synthetic(num_nodes=20, num_targets=20, seed=0), num_agent=1, num_envs=1, device="cpu")
