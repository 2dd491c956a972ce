"""Synthetic nanopore raw-signal reads (no real FAST5 data is available offline).

The generator is the one fixed in SURVEY.md section 8(d): per read ``i`` a
``numpy.random.default_rng(base_seed + i)`` stream draws event lengths
``U{3..19}``, event levels ``N(500, 80)`` and per-sample noise ``N(0, 8)``;
the signal is rounded to int16 DAC values and truncated to the target length.
With the shipped weights this yields ~13 % positive positions and ~46 intervals
per 10 k samples, i.e. a non-trivial interval output.
"""

import numpy as np


def synth_read(length, seed):
    rng = np.random.default_rng(seed)
    n_events = length // 3 + 2
    ev_len = rng.integers(3, 20, size=n_events)
    ev_level = rng.normal(500.0, 80.0, size=n_events)
    signal = np.repeat(ev_level, ev_len)[:length]
    signal = signal + rng.normal(0.0, 8.0, size=length)
    return np.clip(np.rint(signal), -32768, 32767).astype(np.int16)


def synth_reads(lengths, base_seed=0):
    return [synth_read(int(n), base_seed + i) for i, n in enumerate(lengths)]


def ragged_lengths(n_reads, lo, hi, seed=0):
    """Seeded read lengths ``U{lo..hi}`` (config 4: 50 000..200 000)."""
    return np.random.default_rng(seed).integers(lo, hi + 1, size=n_reads).astype(np.int64)


def concat_reads(reads):
    """Concatenate reads into (int16 array, int64 offsets[R+1])."""
    offsets = np.zeros(len(reads) + 1, np.int64)
    if reads:
        offsets[1:] = np.cumsum([len(r) for r in reads])
        raw = np.concatenate(reads).astype(np.int16, copy=False)
    else:
        raw = np.zeros(0, np.int16)
    return raw, offsets
