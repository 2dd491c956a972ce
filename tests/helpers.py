import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def hpm_from_golden(g):
    return {k[4:]: int(v) for k, v in g.items() if k.startswith("hpm_")}


def random_labels(rng, n, p_switch=0.08):
    """Binary label sequence with geometric run lengths (mean 1 / p_switch)."""
    out = np.empty(n, np.int64)
    i, cur = 0, int(rng.integers(0, 2))
    while i < n:
        run = int(rng.geometric(p_switch))
        out[i:i + run] = cur
        i += run
        cur ^= 1
    return out


def allowed_label_flips(p_ref, threshold=0.5, band=1e-3):
    """Positions where the contract allows a different label: reference p within `band` of the threshold."""
    return np.abs(np.asarray(p_ref, np.float64) - threshold) <= band
