#!/usr/bin/env python3
"""Throughput benchmark of the catfish inference hot path (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[3], the configuration the headline metric is quoted on):
ResNetRNN with the shipped checkpoint over synthetic raw-signal reads of ragged length
U{50 000 .. 200 000} samples, sharded by read.  One "step" is one pass of the whole hot path
(median/MAD normalisation, windowing, conv stack, GRU stack, dense+sigmoid, threshold /
short-run removal / interval emission) over one batch of --reads-per-step reads per GPU; the
100 000-read job of configs[3] is a sequence of such batches.  Every rank works on its own
batch (weak scaling), there is no collective on the data path.

* value : samples/s with the int16 signal already resident in HBM (cf_infer_reads)
* e2e   : samples/s through the host-buffer C-ABI call (cf_infer_reads_host): pinned host
          signal -> device, compute, intervals -> host, inside the timed region
* roofline : the dominant kernel class, timed with CUDA events on the launching stream
          during the timed steps
* cpu_baseline : the oracle's torch-fp32 restatement of the TF graph + the reference's
          post-processing, on all host cores, on a bounded sample (rank 0, N=1)

--impl reference times that same CPU restatement (TensorFlow itself is not installable
offline) on bounded samples of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "raw_signal_samples_per_sec_resnetrnn_infer"
UNIT = "samples/s"
LEN_LO, LEN_HI = 50_000, 200_000

# algorithmic work per real sample (SURVEY.md section 8d / BASELINE.md section 4)
FLOPS_PER_SAMPLE = {"k2_conv_stack": 20608.0, "k3_gru_input_proj": 221184.0, "k4_gru_recurrence": 147456.0,
                    "k5_head": 256.0}
BYTES_PER_SAMPLE = {"k1_stats": 2.0, "k6_intervals": 4.0, "k1_window_table": 0.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=512)
    ap.add_argument("--engine", default="auto", choices=["auto", "tcgen05", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reads", type=int, default=100, help="10k-sample reads of the cpu_baseline sample")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler(object):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for ts, r in self.rows if t0 <= ts <= t1] or [r for _, r in self.rows]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples taken while a kernel was running (power above half of the maximum seen);
        # samples that fall into the host-side gap between steps show the idle clock
        if power:
            thr = 0.5 * max(power)
            loaded = [c for c, p in zip(sm, power) if p >= thr] or sm
        else:
            loaded = sm
        return {"sm_mhz": float(np.median(loaded)) if loaded else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(loaded),
                "power_w_max": max(power) if power else None}


def make_batch(n_reads, seed):
    from catfish_b200 import synth
    lengths = synth.ragged_lengths(n_reads, LEN_LO, LEN_HI, seed=seed)
    reads = synth.synth_reads(lengths, base_seed=seed * 1_000_003)
    return synth.concat_reads(reads)


# ------------------------------------------------------------------------------------ CPU arm
def cpu_run(weights, raw, offsets):
    """One pass of the reference path restated on the CPU over the given reads; returns seconds."""
    import torch
    from oracle import postprocess, tf_graph
    if torch.get_num_threads() < (os.cpu_count() or 1):
        torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use all host cores
    graph = tf_graph.TorchGraph(weights)
    t0 = time.perf_counter()
    n_int = 0
    for r in range(len(offsets) - 1):
        hps, _, _ = postprocess.infer_read(raw[offsets[r]:offsets[r + 1]], graph.infer)
        n_int += len(hps)
    return time.perf_counter() - t0, n_int, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from catfish_b200 import weights as W
    w = W.load_shipped()
    n_reads = 4                                   # bounded sample of the configs[3] length distribution
    batches = [make_batch(n_reads, 9000 + i) for i in range(2)]
    cpu_run(w, batches[0][0][:35 * 200], np.array([0, 35 * 200]))         # page in torch
    for i in range(args.warmup):
        cpu_run(w, *batches[i % 2])
    total_s, total_samples, threads = 0.0, 0, 1
    for i in range(args.steps):
        raw, off = batches[i % 2]
        s, _, threads = cpu_run(w, raw, off)
        total_s += s
        total_samples += int(off[-1])
    value = total_samples / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n_reads, "cpu"),
        "reads_per_sec": n_reads * args.steps / total_s,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d reads U{50k..200k} samples per step (TF-graph CPU restatement, torch fp32; "
                                   "TensorFlow is not installable offline)" % n_reads},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def workload_config(reads_per_step, engine):
    return {"workload": "ResNetRNN infer, shipped checkpoint ckpnt-30000, synthetic raw-signal reads of ragged length "
                        "U{50k..200k} samples (BASELINE configs[3]), sharded by read",
            "reads_per_step_per_gpu": reads_per_step, "engine": engine, "window": 35,
            "l2": "inputs larger than L2 (signal + intermediates of a step exceed 126 MB); two alternating batches"}


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from catfish_b200 import _cabi, neural_network

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: relaunch ourselves one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _cabi.load_library()
    model = neural_network.load_network("ResNetRNN", None, 30000, device=local_rank, engine=args.engine)
    engine = model.resolved_engine
    handle = model.handle
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # two alternating batches per rank, staged in pinned host memory and resident on the device
    batches = []
    for b in range(2):
        raw, off = make_batch(args.reads_per_step, 1000 * (rank + 1) + b)
        raw_pin = torch.from_numpy(raw).pin_memory()
        batches.append(dict(raw_pin=raw_pin, raw_dev=raw_pin.to("cuda", non_blocking=True), off=off,
                            n=int(off[-1])))
    n_reads = args.reads_per_step
    cap = max(int(lib.cf_max_intervals(b["n"], n_reads, 15)) for b in batches)
    iv_dev = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
    ioff_dev = torch.empty(n_reads + 1, dtype=torch.int64, device="cuda")
    iv_host = torch.empty((cap, 2), dtype=torch.int64).pin_memory()
    ioff_host = torch.empty(n_reads + 1, dtype=torch.int64).pin_memory()
    found = ctypes.c_int64(0)
    _cabi.check(lib.cf_model_reserve(handle, max(b["n"] for b in batches), n_reads))

    def step_device(i):
        b = batches[i % 2]
        _cabi.check(lib.cf_infer_reads(handle, b["raw_dev"].data_ptr(), b["off"].ctypes.data_as(_cabi.c_i64_p), n_reads,
                                       None, iv_dev.data_ptr(), ioff_dev.data_ptr(), cap, 0.5, 15, 11, 16, sp))

    def step_host(i):
        b = batches[i % 2]
        _cabi.check(lib.cf_infer_reads_host(handle, b["raw_pin"].data_ptr(), b["off"].ctypes.data_as(_cabi.c_i64_p),
                                            n_reads, None, iv_host.data_ptr(), ioff_host.data_ptr(), cap, 0.5, 15, 11,
                                            16, ctypes.byref(found), sp))
        return int(found.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for i in range(steps):
            step_fn(i)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(3, args.warmup)):
        step_device(i)
    torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs, per-kernel-class events on
    _cabi.profile_enable(handle, True)
    launches0 = _cabi.launch_count()
    # address the GPU by UUID: nvidia-smi's index order need not be CUDA's (visible-device remapping)
    try:
        gpu_id = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_id = local_rank
    sampler = ClockSampler(gpu_id)
    t0 = time.time()
    ms_dev = timed(step_device, args.steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    launches = _cabi.launch_count() - launches0
    prof = _cabi.profile_read(handle)
    _cabi.profile_enable(handle, False)

    # ---- timed region 2: host buffers through the public C-ABI call
    for i in range(2):
        step_host(i)
    n_intervals = step_host(0)
    ms_e2e = timed(step_host, args.steps)

    samples_rank = sum(batches[i % 2]["n"] for i in range(args.steps))
    h2d = int(np.mean([2 * batches[i % 2]["n"] + 16 * (n_reads + 1) for i in range(args.steps)]))
    d2h = int(16 * n_intervals + 8 * (n_reads + 1))
    if world > 1:
        t = torch.tensor([samples_rank, launches], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        samples_all, launches_all = float(t[0].item()), int(t[1].item())
    else:
        samples_all, launches_all = float(samples_rank), int(launches)
    value = samples_all / (ms_dev * 1e-3)
    e2e = samples_all / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        # dominant kernel class of this rank (time inside the timed region, CUDA events on the stream)
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        name, (dom_ms, dom_launches) = dom
        per_launch_s = dom_ms * 1e-3 / max(1, dom_launches)
        units_per_launch = samples_rank / max(1, dom_launches)
        if name in FLOPS_PER_SAMPLE:
            achieved = FLOPS_PER_SAMPLE[name] * units_per_launch / per_launch_s / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tensor_sustained"], "traffic": None}
        else:
            achieved = BYTES_PER_SAMPLE.get(name, 0.0) * units_per_launch / per_launch_s / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm"], "traffic": None}
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if name in tj:      # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture
                roof["traffic"] = tj[name]["dram_bytes_per_launch"]
                roof["traffic_source"] = "profiles/r1_traffic.json (%s)" % tj.get("source", "ncu")
        roof.update({"kernel": name, "launches": dom_launches, "avg_launch_ms": dom_ms / max(1, dom_launches),
                     "share_of_step": dom_ms / ms_dev, "peak_source": peaks["source"] + " (sustained)",
                     "algorithmic_per_sample": FLOPS_PER_SAMPLE.get(name, BYTES_PER_SAMPLE.get(name))})
        if prof.get("k3_gru_input_proj", (0.0, 0))[0] == 0.0:
            # fused GRU layer kernel: projection + recurrence FLOPs both run inside k4
            FLOPS_PER_SAMPLE["k4_gru_recurrence"] = 221184.0 + 147456.0
            if name == "k4_gru_recurrence":
                achieved = FLOPS_PER_SAMPLE[name] * units_per_launch / per_launch_s / 1e12
                roof.update({"achieved": achieved, "frac": achieved / peaks["tensor_sustained"],
                             "algorithmic_per_sample": FLOPS_PER_SAMPLE[name],
                             "note": "fused GRU layer: input projection + recurrence"})
        kernels = {}
        for k, (ms, cnt) in prof.items():
            ent = {"ms": ms, "launches": cnt, "share": ms / ms_dev}
            if ms > 0 and k in FLOPS_PER_SAMPLE:
                ent["tflops"] = FLOPS_PER_SAMPLE[k] * samples_rank / (ms * 1e-3) / 1e12
                ent["frac_of_tensor_peak"] = ent["tflops"] / peaks["tensor_sustained"]
            elif ms > 0 and BYTES_PER_SAMPLE.get(k):
                ent["gbs"] = BYTES_PER_SAMPLE[k] * samples_rank / (ms * 1e-3) / 1e9
                ent["frac_of_hbm_peak"] = ent["gbs"] / peaks["hbm"]
            kernels[k] = ent
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"f16e5": "f16+e5m2 (fp16 product + e5m2 correction product, fp32 accumulate)",
                      "bf16x3": "bf16x3 (split bf16 operands, fp32 accumulate)"}.get(model.operand_format, "f32"),
            "data": "synthetic", "config": workload_config(n_reads, engine),
            "reads_per_sec": world * n_reads * args.steps / (ms_dev * 1e-3),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_all, "clocks": clocks, "roofline": roof, "kernels": kernels,
            "intervals_per_step": n_intervals,
        }
        if world == 1 and not args.no_cpu_baseline:
            from catfish_b200 import synth, weights as W
            reads = synth.synth_reads([10000] * args.cpu_reads, base_seed=77)
            raw, off = synth.concat_reads(reads)
            w = W.load_shipped()
            cpu_run(w, raw[:7000], np.array([0, 7000]))
            secs, _, threads = cpu_run(w, raw, off)
            line["cpu_baseline"] = {"value": len(raw) / secs, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d synthetic reads x 10 000 samples (BASELINE configs[0] style), "
                                              "TF-graph CPU restatement (torch fp32) + reference post-processing, "
                                              "%.1f s" % (args.cpu_reads, secs),
                                    "host_cores": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
