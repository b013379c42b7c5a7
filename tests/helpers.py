"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import glob
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def scenario_of(g):
    """Scenario dict (oracle/loader schema) from the arrays stored inside a fixture."""
    return dict(nodes=np.array(g["sc_nodes"], np.float64).reshape(-1, 2),
                targets=np.array(g["sc_targets"], np.float64).reshape(-1, 2),
                bs=np.array(g["sc_bs"], np.float64), par=np.array(g["sc_par"], np.float64))


def mc_dict_of(g):
    p = g["mc_par"]
    return dict(capacity=p[0], threshold=p[1], velocity=p[2], pm=p[3], charging_range=p[4], alpha=p[5], beta=p[6],
                epsilon=p[7])


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


# ------------------------------------------------------------------ the tests' host build of the kernel source
EMU_SO = os.path.join(REPO, "tests", "emu", "libwrsn_emu.so")
_saved = {}


def use_host_build(path=EMU_SO):
    """TEST INFRASTRUCTURE.  Point the product's host layer (ctypes binding + BatchedWRSN) at tests/emu/libwrsn_emu.so — the
    kernel source compiled single-lane for the host — so that the `-m "not gpu"` tests can drive the engine's event logic
    through the same C ABI and the same Python code on a box without a GPU.  The product package has no such switch: this
    function patches it from the outside (the bound library object, and the three methods of BatchedWRSN that touch CUDA)."""
    import ctypes as C
    import torch
    from multi_agent_rl_wrsn_b200 import _lib, batched
    if not _saved:
        for k in ("_require_device", "_stream", "_sync", "_call"):
            _saved[k] = getattr(batched.BatchedWRSN, k)
    _lib._lib = _lib._bind(C.CDLL(path))
    B = batched.BatchedWRSN
    B._require_device = lambda self, device: torch.device("cpu")
    B._stream = lambda self: None
    B._sync = lambda self: None
    B._call = lambda self, fn, *args: _lib.check(fn(*args), self.L)
    return _lib._lib


def use_cuda_build():
    """Undo use_host_build(): the product as shipped (libwrsn_b200.so, CUDA device required)."""
    from multi_agent_rl_wrsn_b200 import _lib, batched
    for k, v in _saved.items():
        setattr(batched.BatchedWRSN, k, v)
    _lib._lib = None
    return _lib.lib()


def use_cuda_build_lazy():
    """Undo use_host_build() without loading the CUDA library (CPU test runs never need it)."""
    from multi_agent_rl_wrsn_b200 import _lib, batched
    for k, v in _saved.items():
        setattr(batched.BatchedWRSN, k, v)
    _lib._lib = None


def is_host_build(L):
    return hasattr(L, "wrsn_is_emulation")
