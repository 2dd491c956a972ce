#!/usr/bin/env python3
"""Diagnose which partial products the f16e5 self-test kernel produces (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from catfish_b200 import _cabi

def rnd(x, dt):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dt).to(torch.float32).numpy().astype(np.float64)

lib = _cabi.load_library()
for mode in (0, 1):
    for k, n in ((16, 32), (32, 64), (64, 128), (16, 128), (64, 32)):
        rng = np.random.default_rng(k + n)
        a = rng.normal(0, 1, size=(128, k)).astype(np.float32)
        w = rng.normal(0, 0.3, size=(k, n)).astype(np.float32)
        ad = torch.from_numpy(a).cuda()
        out = torch.full((128, n), float("nan"), dtype=torch.float32, device="cuda")
        _cabi.check(lib.cf_selftest_f16e5(0, ad.data_ptr(), k, n, w.ctypes.data, mode, out.data_ptr(), 0))
        got = out.cpu().numpy().astype(np.float64)
        s = 64.0
        ah, wh = rnd(a, torch.float16), rnd(w, torch.float16)
        al, wl = a - ah, w - wh
        e5 = torch.float8_e5m2
        A1, B1 = rnd(al * s, e5), rnd(wh / s, e5)
        A2, B2 = rnd(ah / s, e5), rnd(wl * s, e5)
        main = ah @ wh
        hyp = {"main": main, "main+c1": main + A1 @ B1, "main+c2": main + A2 @ B2, "full": main + A1 @ B1 + A2 @ B2,
               "swapB": main + A1 @ B2 + A2 @ B1, "exact": a.astype(np.float64) @ w.astype(np.float64)}
        scale = np.abs(a).astype(np.float64) @ np.abs(w).astype(np.float64)
        print("mode", mode, "k", k, "n", n, {h: "%.1e" % (np.abs(got - v) / scale).max() for h, v in hyp.items()})
