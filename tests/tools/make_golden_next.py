#!/usr/bin/env python3
"""Generate tests/golden/events.json and confusion.json ("next" rows N3 / N4; build container only).

    python tests/tools/make_golden_next.py

Source of truth: the REFERENCE's own functions executed from /root/reference through
oracle/ref_infer.py - ``correct_events`` (networks/correct_output.py:14-76, results recovered from
what it prints) and ``confusion_matrix`` (networks/trainingDB/metrics.py:10-37).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from oracle import ref_infer  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def event_cases():
    rng = np.random.default_rng(20261018)
    cases = []
    for trial in range(120):
        n_ev = int(rng.integers(1, 80))
        ev = rng.integers(0 if trial % 9 == 0 else 1, 14, n_ev).tolist()
        total = sum(ev)
        n_scores = int(rng.integers(max(1, total // 2), total + 20))
        if trial % 5 == 0:
            scores = [0.5] * n_scores                       # every mean is an exact tie
        elif trial % 5 == 1:
            scores = rng.choice([0.25, 0.75, 0.5], n_scores).tolist()   # frequent ties
        else:
            scores = rng.random(n_scores).astype(np.float32).astype(float).tolist()
        start = int(rng.integers(0, max(1, total // 2)))
        length = int(rng.integers(0, total + 10))
        if trial % 4 == 0:                                   # make the `start + length` exit reachable
            cum = np.cumsum(ev)
            length = int(cum[int(rng.integers(0, n_ev))])
            start = 0
        case = {"scores": scores, "events": ev, "start": start, "length": length}
        try:
            classes, voted, first, last = ref_infer.run_correct_events(scores, ev, start, length)
            case.update(classes=classes, voted=voted, start_event=first, final_event=last, raises=None)
        except (ZeroDivisionError, UnboundLocalError, NameError) as e:
            case["raises"] = "ZeroDivisionError" if isinstance(e, ZeroDivisionError) else "UnboundLocalError"
        cases.append(case)
    return cases


def confusion_cases():
    rng = np.random.default_rng(7)
    metrics = ref_infer.load_metrics()
    cases = []
    for n in (0, 1, 35, 1000, 4099):
        truth = rng.integers(0, 2, n).tolist()
        call = rng.integers(0, 2, n).tolist()
        if n == 1000:
            truth[5] = 2                                     # odd label values land in fp / fn
            truth[6] = 2
            call[5], call[6] = 1, 0
        cases.append({"true": truth, "pred": call, "counts": list(metrics.confusion_matrix(truth, call))})
    return cases


def main():
    with open(os.path.join(OUT, "events.json"), "w") as f:
        json.dump(event_cases(), f)
    with open(os.path.join(OUT, "confusion.json"), "w") as f:
        json.dump(confusion_cases(), f)
    print("wrote events.json, confusion.json")


if __name__ == "__main__":
    main()
