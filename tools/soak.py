#!/usr/bin/env python3
"""Soak test: many batch shapes through the tcgen05 engine, each checked against the fp32 SIMT engine.

Targets the places where barrier bookkeeping could go wrong: tile counts around the grid size
(1, 2, 3, 73..75, 147..149, 295..297), around the internal pass size (1183..1186, 2368..2370), odd
numbers of tiles per chain, and random ragged mixes.  Exits non-zero on the first mismatch.

    python tools/soak.py [n_random]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import infer, neural_network, synth  # noqa: E402


def reads_for_tiles(tiles, rng):
    """A ragged batch whose window count lands in the last few windows of `tiles` tiles."""
    windows = tiles * 128 - int(rng.integers(0, 5))
    lengths = []
    while windows > 0:
        w = int(min(windows, rng.integers(1, 4000)))
        lengths.append((w - 1) * 35 + int(rng.integers(0, 35)))          # L // 35 + 1 == w
        windows -= w
    return [max(1, n) for n in lengths]


if __name__ == "__main__":
    n_random = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    fast = neural_network.load_network("ResNetRNN", None, 30000, engine="auto")
    slow = neural_network.load_network("ResNetRNN", None, 30000, engine="simt")
    assert fast.resolved_engine == "tcgen05"
    rng = np.random.default_rng(77)
    cases = [reads_for_tiles(t, rng) for t in (1, 2, 3, 4, 73, 74, 75, 147, 148, 149, 150, 295, 296, 297, 591, 592, 593,
                                               1183, 1184, 1185, 1186, 2367, 2368, 2369, 2370)]
    for _ in range(n_random):
        n = int(rng.integers(1, 200))
        kind = rng.integers(0, 3)
        hi = [300, 20000, 120000][kind]
        cases.append([int(x) for x in rng.integers(1, hi, size=n)])
    t0 = time.time()
    worst = 0.0
    for ci, lengths in enumerate(cases):
        reads = synth.synth_reads(lengths, base_seed=10_000 * ci)
        h1, l1, s1 = infer.infer_reads(reads, fast, return_scores=True)
        h2, l2, s2 = infer.infer_reads(reads, slow, return_scores=True)
        assert l1 == l2
        for r, (a, b, sa, sb) in enumerate(zip(h1, h2, s1, s2)):
            if len(sa) == 0:
                continue
            fin = np.isfinite(sb)
            assert np.array_equal(np.isfinite(sa), fin), (ci, r)
            d = float(np.abs(sa[fin] - sb[fin]).max()) if fin.any() else 0.0
            worst = max(worst, d)
            assert d < 1e-3, (ci, r, d)
            if a != b:
                flips = (sa >= 0.5) != (sb >= 0.5)
                assert np.all(np.abs(sb[flips].astype(np.float64) - 0.5) <= 1e-3), (ci, r)
        windows = sum(n // 35 + 1 for n in lengths)
        print("case %3d: %4d reads, %8d samples, %5d tiles ok (worst |dp| so far %.2e, %.0f s)"
              % (ci, len(lengths), sum(lengths), -(-windows // 128), worst, time.time() - t0), flush=True)
    print("soak ok: %d cases, worst |dp| %.2e" % (len(cases), worst))
