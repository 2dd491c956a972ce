#!/usr/bin/env python3
"""Per-launch DRAM traffic of the main kernels from an `ncu --set full` report -> profiles/r2_traffic.json.

    python tools/extract_traffic.py gpurun_out/r2_prof_main.ncu-rep

bench.py reads the JSON to fill roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch,
averaged over the captured launches of a kernel class).  The file is stamped with the git commit and with a
fingerprint of catfish_b200/csrc/ (bench.source_fingerprint): bench.py refuses the figure when the kernels
it times are not the ones the capture was taken from.
"""
import csv
import json
import os
import subprocess
import sys

CLASSES = {"tc_gru_fused2_kernel": "k4_gru_recurrence", "tc_conv4_kernel": "k2_conv_stack", "tc_conv2_kernel": "k2_conv_stack"}
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

if __name__ == "__main__":
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    it = hdr.index("gpu__time_duration.sum")
    TU = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
    acc = {}
    for r in rows[2:]:
        name = r[ik]
        cls = next((v for k, v in CLASSES.items() if k in name), None)
        if cls is None:
            continue
        if float(r[it]) * TU.get(units[it], 1.0) < 20.0:      # early-exit twin of the other operand format
            continue
        b = float(r[ir]) * UNITS[units[ir]] + float(r[iw]) * UNITS[units[iw]]
        acc.setdefault(cls, []).append(b)
    res = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v)} for k, v in acc.items()}
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    sys.path.insert(0, root)
    import bench
    res["source"] = os.path.basename(rep)
    stamp = os.path.join(os.path.dirname(os.path.abspath(rep)), os.path.basename(rep).split("_")[0] + "_source_sha1.txt")
    res["source_sha1"] = open(stamp).read().strip() if os.path.exists(stamp) else bench.source_fingerprint()    # recorded at capture time
    res["commit"] = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() \
        + ("+dirty" if subprocess.run(["git", "-C", root, "status", "--porcelain", "catfish_b200/csrc"], capture_output=True,
                                       text=True).stdout.strip() else "")
    res["samples_per_launch_note"] = "tools/profile_workload.py 96 reads: 12 182 812 samples, first engine pass = 2368 tiles"
    path = os.path.join(root, "profiles", "r2_traffic.json")
    with open(path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
