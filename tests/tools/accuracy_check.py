#!/usr/bin/env python3
"""max / percentile |dp| of the GPU path against the fp32 CPU oracle on synthetic 10k-sample reads.

    python tests/tools/accuracy_check.py [n_reads]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from catfish_b200 import infer, neural_network, synth, weights
from oracle import postprocess, tf_graph
m = neural_network.load_network("ResNetRNN", None, 30000)
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reads = synth.synth_reads([10000]*n_reads, base_seed=300)
hps, lens, scores = infer.infer_reads(reads, m, return_scores=True)
g = tf_graph.TorchGraph(weights.load_shipped())
errs=[]; flips=0; out=0
for r, s in zip(reads, scores):
    _, _, want = postprocess.infer_read(r, g.infer)
    d = np.abs(s - want); errs.append(d)
    f = ((s>=0.5) != (want>=0.5)); flips += f.sum(); out += (f & (np.abs(want-0.5)>1e-3)).sum()
e = np.concatenate(errs)
print("operand format", m.operand_format, "positions", e.size, "max %.3e p99.9 %.3e p99 %.3e mean %.3e flips %d outside-band %d" % (e.max(), np.quantile(e,0.999), np.quantile(e,0.99), e.mean(), flips, out))
