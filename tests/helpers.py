import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def hpm_from_golden(g):
    return {k[4:]: int(v) for k, v in g.items() if k.startswith("hpm_")}


def random_bn_weights(seed):
    """Seeded ResNetRNN weights (shipped hyper-parameters) with non-trivial BN gamma/beta/moving stats."""
    from catfish_b200 import weights
    w = weights.random_init("ResNetRNN", seed=seed, layer_size=64, n_layers=3, layer_size_res=32, n_layers_res=2)
    rng = np.random.default_rng(seed + 1000)
    for k in sorted(w):
        if k.endswith("moving_mean"):
            w[k] = rng.normal(0, 0.3, size=w[k].shape).astype(np.float32)
        elif k.endswith("moving_variance"):
            w[k] = rng.uniform(0.3, 2.0, size=w[k].shape).astype(np.float32)
        elif k.endswith("beta"):
            w[k] = rng.normal(0, 0.2, size=w[k].shape).astype(np.float32)
        elif k.endswith("gamma"):
            w[k] = rng.uniform(0.5, 1.5, size=w[k].shape).astype(np.float32)
    return w


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
