#!/bin/bash
# Round-end measurement set (run under gpurun on one B200): GPU tests, the bench line, the ncu launch
# list of a shortened bench, and one --set full capture of the conv stack + the three GRU-layer launches.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-reads 400 > gpurun_out/bench_full.txt 2> gpurun_out/bench_full.err
tail -c 600 gpurun_out/bench_full.txt
timeout 600 python bench.py --steps 2 --warmup 3 --reads-per-step 128 --no-cpu-baseline > gpurun_out/bench_short.txt 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --reads-per-step 128 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/profile_workload.py 96 2 > gpurun_out/prof_workload.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gru_fused2_kernel|tc_conv4_kernel' -s 8 -c 4 -f -o gpurun_out/prof_main \
    python tools/profile_workload.py 96 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python tests/tools/accuracy_check.py > gpurun_out/accuracy.txt 2>&1; tail -1 gpurun_out/accuracy.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.txt 2>&1; tail -c 300 gpurun_out/bench_ref.txt
timeout 600 python tools/tail_accuracy.py 12 256 > gpurun_out/tail_accuracy.txt 2>&1; tail -1 gpurun_out/tail_accuracy.txt
timeout 300 python tests/tools/accuracy_check.py 200 >> gpurun_out/accuracy.txt 2>&1; tail -1 gpurun_out/accuracy.txt
timeout 300 python tools/latency_long_reads.py 1 8 64 > gpurun_out/long_reads.txt 2>&1; tail -3 gpurun_out/long_reads.txt
