#!/usr/bin/env python3
"""Per-kernel totals and shares from an ncu launch list (--metrics gpu__time_duration.sum --csv).

    python tools/launch_summary.py profiles/r1_ncu_launches.csv > profiles/r1_ncu_launch_summary.txt
"""
import csv
import re
import sys
from collections import OrderedDict

if __name__ == "__main__":
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = OrderedDict()
    for r in rd:
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("cf::", "").strip()
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1.0)
        acc.setdefault(name, []).append(v)
    total = sum(sum(v) for v in acc.values())
    n = sum(len(v) for v in acc.values())
    print("%d launches captured, %.3f ms total" % (n, total / 1e3))
    for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        print("%-42s n=%4d total %9.3f ms  avg %8.1f us  share %5.1f%%" % (name, len(v), sum(v) / 1e3, sum(v) / len(v), 100 * sum(v) / total))
