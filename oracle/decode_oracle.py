"""TEST INFRASTRUCTURE — CPU restatement of the reference's density-map action decode.

Follows ``rl_env/WRSN.py``: the map normalisation of ``WRSN.step`` (:293-296) and ``WRSN.density_map_to_action``
(:229-287), statement by statement, with the reference's own dependencies (numpy's ``argmax`` / ``percentile``, scipy's
``minimize(method='L-BFGS-B')`` and ``scipy.spatial.distance.euclidean``).  Only ``tests/`` may import this module: it is
the checker of ``wrsn_decode_density_map`` (CUDA), never part of the product path.

Pinned by ``tests/test_decode.py::test_decode_oracle_matches_reference_golden`` against the decoded actions the
unmodified reference produced here (``tests/golden/dmap_random_n50.npz``, ``oracle/gen_golden.py``).  scipy's L-BFGS-B
has changed since the reference's pinned 1.10.1: locations are compared at 1e-6, as in that fixture's
own façade test.
"""
import numpy as np
from scipy.optimize import minimize
from scipy.spatial.distance import euclidean


def normalise_map(action, epsilon=1e-9):
    """WRSN.step :293-296."""
    action = np.array(action)
    if not (np.all((action >= 0) & (action <= 1)) and np.isclose(np.sum(action), 1)):
        action = np.exp(action)
        action = action / (np.sum(action) + epsilon)
    return action


def density_map_to_action(dmap, frame, xy, status, energy, cs, threshold, charging_range, alpha, beta, map_size,
                          return_result=False):
    """WRSN.density_map_to_action :229-287 (``dmap`` already normalised by ``normalise_map``)."""
    unit = 1.0 / map_size
    f = frame

    def up_mapping(dm):                                  # :91-93
        return np.array([dm[0] * (f[1] - f[0]) + f[0], dm[1] * (f[3] - f[2]) + f[2]])

    def down_mapping(loc):                               # :86-88
        return np.array([(loc[0] - f[0]) / (f[1] - f[0]), (loc[1] - f[2]) / (f[3] - f[2])])

    max_index = np.unravel_index(np.argmax(dmap), dmap.shape)
    lower = up_mapping([(max_index[0] + 0.5) * unit - charging_range / (f[1] - f[0]),
                        (max_index[1] + 0.5) * unit - charging_range / (f[3] - f[2])])
    upper = up_mapping([(max_index[0] + 0.5) * unit + charging_range / (f[1] - f[0]),
                        (max_index[1] + 0.5) * unit + charging_range / (f[3] - f[2])])
    bounds = [(lower[0], upper[0]), (lower[1], upper[1])]
    alive = [n for n in range(len(status)) if status[n] != 0]

    def objective(loc):
        loc = np.array(loc)
        res = 0
        for n in alive:
            d = euclidean(loc, xy[n])
            res += int(d <= charging_range) * (cs[n] / (energy[n] - threshold)) * alpha / ((d + beta) ** 2)
        return -res

    x0 = [(lower[0] + upper[0]) / 2, (lower[1] + upper[1]) / 2]
    result = minimize(objective, x0, bounds=bounds, method="L-BFGS-B")
    prob = np.copy(dmap)
    flat = prob.flatten()
    thr = np.percentile(flat, 99.9)
    flat[flat < thr] = 0
    prob = flat.reshape(prob.shape)
    prob = prob / np.sum(prob)
    loc = down_mapping(np.array(result.x))
    out = np.array([loc[0], loc[1], prob[max_index[0]][max_index[1]]])
    if return_result:
        return out, result, np.array(x0), objective
    return out
