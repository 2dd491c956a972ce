#!/usr/bin/env python3
"""Per-launch summary of an `ncu --metrics ... --csv` log (one block per launch, one line per metric).

    python tools/ncu_csv_summary.py gpurun_out/r2_ncu_rnn_only.csv > profiles/r2_ncu_rnn_only_summary.txt
"""
import csv
import sys
from collections import OrderedDict

launches = OrderedDict()
for row in csv.reader(open(sys.argv[1])):
    if len(row) < 15 or not row[0].isdigit():
        continue
    launches.setdefault(row[0], (row[4], []))[1].append((row[12], row[13], row[14]))
for _, (name, metrics) in launches.items():
    print("== kernel: %s" % name[:100])
    for m, unit, val in sorted(metrics):
        print("  %-70s %-10s %s" % (m, unit, val))
