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
