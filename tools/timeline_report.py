#!/usr/bin/env python3
"""Per-phase cycle report from a CF_TC_TRACE timeline of the fused GRU-layer kernel (block 0).

    CF_TC_TRACE=gpurun_out/trace.txt python tools/profile_workload.py 96 1
    python tools/timeline_report.py gpurun_out/trace.txt

Regions 0/1 = MMA issuer of chain 0/1, 2/3 = epilogue of chain 0/1.  Tags (tc_engine.cu, CF_TR):
issuer 10 x-part start, 60+kk chunk kk's data there, 20+kk chunk kk issued, 30 h ready, 31 gate MMAs issued, 40 r*h ready, 41 candidate
MMAs issued; epilogue (warp 0, which serves both chains; regions 2/3 = its work for chain 0/1) 49 phase R entered,
50 gates landed, 51 r*h handed over | 52 phase C: update gate done, 53 candidate landed, 55 new state handed over,
56 phase C left (layer output stored).
"""
import sys
from collections import defaultdict

rows = [tuple(int(x) for x in l.split()) for l in open(sys.argv[1]) if l.strip()]
by_region = defaultdict(list)
for rg, tagt, clk in rows:
    by_region[rg].append((tagt % 1000, clk))
for rg in sorted(by_region):
    ev = by_region[rg]
    print("region", rg, "events", len(ev))
    # deltas between consecutive events, averaged per (tag_from -> tag_to)
    acc = defaultdict(list)
    for (t0, c0), (t1, c1) in zip(ev, ev[1:]):
        acc[(t0, t1)].append(c1 - c0)
    for (t0, t1), v in acc.items():
        print("   %3d -> %3d : n=%2d mean %7.0f cycles (min %d max %d)" % (t0, t1, len(v), sum(v) / len(v), min(v), max(v)))
    first = [c for t, c in ev if t == ev[0][0]]
    if len(first) > 1:
        print("   step period: %.0f cycles" % ((first[-1] - first[0]) / (len(first) - 1)))
