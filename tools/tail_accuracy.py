#!/usr/bin/env python3
"""Tail of the probability error at scale: tcgen05 engine vs the fp32 CUDA-core engine on the same reads.

    python tools/tail_accuracy.py [n_batches] [reads_per_batch]

The CPU oracle is too slow for 1e9 positions; the fp32 SIMT engine (itself within 1e-6 of the oracle,
tests/test_gpu_forward.py) stands in.  Prints the running maximum and how many positions exceed 2.5e-4 / 5e-4.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import infer, neural_network, synth  # noqa: E402

if __name__ == "__main__":
    n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    fast = neural_network.load_network("ResNetRNN", None, 30000, engine="auto")
    slow = neural_network.load_network("ResNetRNN", None, 30000, engine="simt")
    print("operand format", fast.operand_format)
    worst, total, over25, over50 = 0.0, 0, 0, 0
    t0 = time.time()
    for b in range(n_batches):
        lengths = synth.ragged_lengths(n_reads, 50_000, 200_000, seed=500 + b)
        raw, off = synth.concat_reads(synth.synth_reads(lengths, base_seed=700_000 + 1000 * b))
        _, _, s1 = infer.infer_concatenated(raw, off, fast, return_scores=True)
        _, _, s2 = infer.infer_concatenated(raw, off, slow, return_scores=True)
        d = np.abs(s1 - s2)
        worst = max(worst, float(d.max()))
        total += d.size
        over25 += int(np.count_nonzero(d > 2.5e-4))
        over50 += int(np.count_nonzero(d > 5e-4))
        print("batch %d: positions so far %.3e  max |dp| %.3e  >2.5e-4: %d  >5e-4: %d  (%.0f s)"
              % (b, total, worst, over25, over50, time.time() - t0), flush=True)
