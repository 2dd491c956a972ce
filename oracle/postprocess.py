"""ORACLE (test infrastructure, not product code): pre/post-processing of catfish/infer.py.

Restates, in plain Python/numpy, the functions of
/root/reference/catfish/infer.py that surround the network call:

* normalize_raw_signal  infer.py:96-105
* pad + reshape_input   infer.py:32-43, 108-124
* class_from_threshold  infer.py:128-138
* correct_short         infer.py:174-198
* hp_in_pred            infer.py:141-162
* infer_class_from_signal (array-level, without the HDF5 read) infer.py:12-51

Pinned: tests/test_oracle_vs_reference.py runs these against the reference's own
functions imported unmodified (oracle/ref_infer.py) on seeded inputs, and
tests/golden/ holds vectors generated from the reference functions by
tests/tools/make_golden.py.

The ``*_loops`` variants follow the reference statement by statement (small
cases); the vectorised variants return identical results and are used for the
larger parity sizes and the timed CPU baseline.
"""

import numpy as np

WINDOW = 35


def normalize_raw_signal(raw, norm_method="median"):
    """infer.py:96-105: shift = median, scale = median absolute deviation, float64."""
    if norm_method == "median":
        shift = np.median(raw)
        scale = np.median(np.abs(raw - shift))
    else:
        raise ValueError("norm_method not recognized")
    return (raw - shift) / scale


def padding_size(length, window_size=WINDOW):
    """infer.py:32-36: a read whose length divides evenly still gets a full extra window."""
    if not (length / window_size).is_integer():
        return window_size - (length - (length // window_size * window_size))
    return 35


def pad_and_window(norm, window_size=WINDOW):
    """infer.py:32-43: zero padding after normalisation, reshape (-1, window, 1)."""
    pad = padding_size(len(norm), window_size)
    raw = np.hstack((norm, np.array(pad * [0])))
    return np.reshape(raw, (-1, window_size, 1)), pad


def class_from_threshold_loops(predicted_scores, threshold=0.5):
    """infer.py:128-138."""
    return [1 if y >= threshold else 0 for y in predicted_scores]


def class_from_threshold(predicted_scores, threshold=0.5):
    return (np.asarray(predicted_scores, dtype=float) >= threshold).astype(np.int64)


def _runs(values):
    """Run-length encode a 1-D array: (run_values, run_starts, run_lengths)."""
    v = np.asarray(values)
    n = len(v)
    if n == 0:
        raise IndexError("list index out of range")            # infer.py:151,184 index [0]
    change = np.flatnonzero(v[1:] != v[:-1]) + 1
    starts = np.concatenate(([0], change))
    lengths = np.diff(np.concatenate((starts, [n])))
    return v[starts], starts, lengths


def correct_short_loops(predictions, threshold=15):
    """infer.py:174-198, element by element: run-length encode, zero the short non-zero runs, expand."""
    runs = []                                   # [label, length]
    first = predictions[0]                      # IndexError on empty input, like the reference
    del first
    for label in predictions:
        if runs and runs[-1][0] == label:
            runs[-1][1] += 1
        else:
            runs.append([label, 1])
    pieces = [np.repeat(0 if (label != 0 and length < threshold) else label, length) for label, length in runs]
    return np.concatenate(pieces)


def correct_short(predictions, threshold=15):
    vals, _, lengths = _runs(predictions)
    vals = np.where((vals != 0) & (lengths < threshold), 0, vals)
    return np.repeat(vals, lengths)


def hp_in_pred_loops(predictions, extension_left=11, extension_right=16, label=1):
    """infer.py:141-162, element by element: (label, start, length) of every run, keep those of ``label``."""
    predictions[0]                              # IndexError on empty input, like the reference
    out = []
    start = 0
    for pos in range(1, len(predictions) + 1):
        if pos == len(predictions) or predictions[pos] != predictions[start]:
            if predictions[start] == label:
                out.append([start - extension_left, pos + extension_right])      # start + length = pos
            start = pos
    return out


def hp_in_pred(predictions, extension_left=11, extension_right=16, label=1):
    vals, starts, lengths = _runs(predictions)
    keep = vals == label
    return [[int(s) - extension_left, int(s) + int(n) + extension_right]
            for s, n in zip(starts[keep], lengths[keep])]


def infer_read(raw, model_infer, threshold=0.5, window_size=WINDOW, loops=False):
    """Array-level infer_class_from_signal (infer.py:30-51), first_sample already trimmed.

    raw: int16 samples of one read; model_infer: callable [B,35,1] -> float64 [B*35].
    Returns (predicted_hps, len_read, scores)."""
    norm = normalize_raw_signal(np.asarray(raw), "median")
    raw_in, pad = pad_and_window(norm, window_size)
    scores = model_infer(raw_in)
    scores = scores[:-pad]
    if loops:
        labels = correct_short_loops(class_from_threshold_loops(scores, threshold))
        hps = hp_in_pred_loops(labels)
    else:
        labels = correct_short(class_from_threshold(scores, threshold))
        hps = hp_in_pred(labels)
    return hps, len(labels), scores
