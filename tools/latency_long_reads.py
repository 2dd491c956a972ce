#!/usr/bin/env python3
"""BASELINE configs[4]: end-to-end latency / throughput on ultra-long (1M-sample) reads.

    python tools/latency_long_reads.py [n_reads ...]

Prints, per batch size, the end-to-end time of one cf_infer_reads_host call (median of 5 after a
warm-up) and the per-kernel-class device times.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import _cabi, infer, neural_network, synth  # noqa: E402

if __name__ == "__main__":
    import torch
    sizes = [int(a) for a in sys.argv[1:]] or [1, 8, 64]
    model = neural_network.load_network("ResNetRNN", None, 30000)
    for n in sizes:
        raw, off = synth.concat_reads(synth.synth_reads([1_000_000] * n, base_seed=5000))
        infer.infer_concatenated(raw, off, model)
        times = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            iv, ioff = infer.infer_concatenated(raw, off, model)
            times.append(time.perf_counter() - t0)
        _cabi.profile_enable(model.handle, True)
        infer.infer_concatenated(raw, off, model)
        prof = _cabi.profile_read(model.handle)
        _cabi.profile_enable(model.handle, False)
        t = float(np.median(times))
        print("reads %3d x 1M: e2e %.2f ms  (%.3g samples/s, %d intervals)  kernels: %s"
              % (n, 1e3 * t, n * 1e6 / t, len(iv), ", ".join("%s %.3f" % (k, v[0]) for k, v in prof.items() if v[0] > 0)))
