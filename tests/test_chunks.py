"""Chunk merging (next row N1): oracle pinned by the reference's own source lines; GPU kernel vs both."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import chunks as ochunks


def _cases():
    with open(os.path.join(GOLDEN, "chunks.json")) as f:
        return json.load(f)


def _norm(x):
    return json.loads(json.dumps(x))


def _random_intervals(rng, len_read):
    out, pos = [], int(rng.integers(-11, 200))
    while pos < len_read - 30 and len(out) < 400:
        n = int(rng.integers(15, 400))
        out.append([pos - 11, pos + n + 16])
        pos += n + int(rng.integers(1, 1500))
    return out


def test_oracle_matches_reference_lines():
    for c in _cases():
        merged, nonhp = ochunks.merge_read(c["hp_positions"], c["len_read"], c["chunk_size"])
        assert merged == c["merged"] and _norm(nonhp) == c["nonhp"]


@pytest.mark.gpu
def test_gpu_merge_matches_golden_and_oracle():
    from catfish_b200 import chunks
    cases = _cases()
    for chunk in (1000, 300):
        sel = [c for c in cases if c["chunk_size"] == chunk]
        hp, non = chunks.merge_reads([c["hp_positions"] for c in sel], [c["len_read"] for c in sel], chunk)
        for c, h, n in zip(sel, hp, non):
            assert h == c["merged"] and _norm(n) == c["nonhp"]
    rng = np.random.default_rng(8)
    lens = [int(rng.integers(200, 60000)) for _ in range(300)]
    ivs = [_random_intervals(rng, n) if i % 11 else [] for i, n in enumerate(lens)]
    for chunk in (1000, 100, 5000):
        hp, non = chunks.merge_reads(ivs, lens, chunk)
        for iv, n, h, x in zip(ivs, lens, hp, non):
            want_h, want_n = ochunks.merge_read(iv, n, chunk)
            assert h == want_h and _norm(x) == _norm(want_n)
    h1, n1 = chunks.merge_read([[-11, 46]], 500)
    assert (h1, n1) == ochunks.merge_read([[-11, 46]], 500)
    m = [[10, 50]]
    assert chunks.center_hp(m, 2000, 1000) == ochunks.center_hp([[10, 50]], 2000, 1000)


@pytest.mark.gpu
def test_gpu_split_raw_matches_numpy_slicing():
    """Next row N2 (split_f5.py:36,64): pieces = signal[s:e] with numpy slice semantics."""
    from catfish_b200 import split_f5, synth
    rng = np.random.default_rng(4)
    raws = synth.synth_reads([1, 7, 100, 4097, 30000], base_seed=9)
    ranges = []
    for r in raws:
        n = len(r)
        rr = [(0, n), (0, 0), (n, n), (-5, n), (3, -2), (n + 10, n + 20), (5, 2), (-n - 3, 4), (1, 2)]
        rr += [tuple(sorted(rng.integers(-10, n + 10, size=2).tolist())) for _ in range(20)]
        ranges.append(rr)
    got = split_f5.split_raw(raws, ranges)
    for r, rr, pieces in zip(raws, ranges, got):
        assert len(pieces) == len(rr)
        for (s, e), p in zip(rr, pieces):
            np.testing.assert_array_equal(p, r[s:e])
            assert p.dtype == np.int16
    assert split_f5.split_raw(raws, [[] for _ in raws]) == [[] for _ in raws]
    # end to end with the chunk merging: the pieces of a read tile it (HP chunks + complement)
    from catfish_b200 import chunks
    hp = [[-11, 46], [300, 420], [2500, 2600], [9000, 9100]]
    merged, nonhp = chunks.merge_read(hp, 30000, 1000)
    pieces = split_f5.split_raw([raws[4]], [[tuple(m) for m in merged] + [tuple(x) for x in nonhp]])[0]
    assert sum(len(p) for p in pieces[:len(merged)]) == sum(max(0, min(e, 30000) - max(s, 0)) for s, e in merged)
