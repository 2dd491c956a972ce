#!/usr/bin/env python3
"""Small fixed workload for ncu / compute-sanitizer runs: one pass of the hot path.

    python tools/profile_workload.py [n_reads] [repeats] [engine] [ResNetRNN|RNN|ResNet]

ResNetRNN runs the shipped checkpoint; RNN / ResNet (the reference's model variants, neural_network.py:17-18 and
resnet_class.py:23) run seeded random-init weights of the shipped sizes.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import infer, neural_network, synth, weights  # noqa: E402

if __name__ == "__main__":
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    engine = sys.argv[3] if len(sys.argv) > 3 else "auto"
    kind = sys.argv[4] if len(sys.argv) > 4 else "ResNetRNN"
    if kind == "ResNetRNN":
        model = neural_network.load_network("ResNetRNN", None, 30000, engine=engine)
    else:
        model = neural_network.build_model(kind, engine=engine, **weights.SHIPPED_HPARAMS)
        model.set_weights(weights.random_init(kind, seed=11, layer_size=64, n_layers=3) if kind == "RNN"
                          else weights.random_init(kind, seed=12, layer_size_res=32, n_layers_res=2))
    lengths = synth.ragged_lengths(n_reads, 50_000, 200_000, seed=1)
    raw, off = synth.concat_reads(synth.synth_reads(lengths, base_seed=17))
    for _ in range(repeats):
        iv, ioff = infer.infer_concatenated(raw, off, model)
    print("engine", model.resolved_engine, "reads", n_reads, "samples", int(off[-1]), "intervals", len(iv))
