#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread ",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    it = hdr.index("gpu__time_duration.sum")
    tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    for r in rows[2:]:
        if float(r[it]) * tu.get(units[it], 1.0) < 20.0:
            print("== kernel (early exit, %s %s):" % (r[it], units[it]), r[hdr.index("Kernel Name")][:70])
            continue
        print("== kernel:", r[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, units, r):
            if any(k.strip() in h for k in KEYS if not k.endswith(" ")) or any(h == k.strip() for k in KEYS if k.endswith(" ")):
                if "stalled" in h and "--stalls" not in sys.argv:
                    continue
                if "stalled" in h and float(v or 0) < 0.2:
                    continue
                print("  %-85s %-12s %s" % (h, u, v))


if __name__ == "__main__":
    main()
