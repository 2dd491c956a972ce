#!/usr/bin/env python3
"""Export the inference tensors of the reference's shipped checkpoint to an .npz.

Run in the build container, where /root/reference is mounted:

    python tools/export_checkpoint.py

Reads  /root/reference/catfish/ResNetRNN/checkpoints/ckpnt-30000.{index,data-*}
(CRC32C-verified, optimizer slots dropped) and writes
catfish_b200/data/ResNetRNN_ckpnt-30000.npz so that tests, smoke() and bench.py
can use the shipped weights on machines where the reference tree is absent.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import weights  # noqa: E402

PREFIX = "/root/reference/catfish/ResNetRNN/checkpoints/ckpnt-30000"

if __name__ == "__main__":
    w = weights.load_tf_checkpoint(PREFIX)
    w = weights.check_weights(w, "ResNetRNN", layer_size=64, n_layers=3, layer_size_res=32, n_layers_res=2)
    weights.save_npz(weights.SHIPPED_NPZ, w)
    print("wrote %s: %d tensors, %d floats" % (weights.SHIPPED_NPZ, len(w), sum(v.size for v in w.values())))
