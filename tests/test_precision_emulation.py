"""CPU: the operand-format emulation behind DESIGN.md section 4 keeps its ordering on a small sample
(single 16-bit passes miss the contract's margin, the split formats do not)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))


def test_format_ordering(shipped_weights):
    import precision_emulation as pe
    from catfish_b200 import synth
    from oracle import postprocess
    raw = synth.synth_reads([2100], base_seed=300)[0]
    x = postprocess.pad_and_window(postprocess.normalize_raw_signal(raw))[0]
    ref = pe.forward(shipped_weights, x, pe.make_mm("exact"))
    err = {s: float(np.abs(pe.forward(shipped_weights, x, pe.make_mm(s)) - ref).max())
           for s in ("bf16x3", "fp16+e5m2", "fp16x1", "bf16x1")}
    assert err["bf16x3"] < 1e-4 and err["fp16+e5m2"] < 2e-4
    assert err["fp16+e5m2"] < err["fp16x1"] < err["bf16x1"]
    assert err["bf16x1"] > 1e-3                      # a single bf16 pass breaks the 1e-3 contract
