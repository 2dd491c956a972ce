#!/bin/bash
# Round-end measurement set (run under gpurun on one B200): GPU tests, the bench line (all legs), the reference arm,
# the ncu launch list of a shortened bench, one --set full capture of the conv stack + the three GRU-layer launches,
# quick metric captures of the RNN-only / ResNet-only configurations, accuracy at three sizes, long-read latency.
# Every ncu command runs only after the same command has exited 0 without ncu; numbers printed under ncu are never
# bench values.
set -u
R=${1:-r2}
mkdir -p gpurun_out
python -c "import bench; print(bench.source_fingerprint())" > gpurun_out/${R}_source_sha1.txt      # which kernel sources the captures are of
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.txt
cat gpurun_out/${R}_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_full.txt 2> gpurun_out/${R}_bench_full.err
tail -c 400 gpurun_out/${R}_bench_full.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.txt 2>&1; tail -c 300 gpurun_out/${R}_bench_ref.txt
timeout 600 python bench.py --steps 2 --warmup 3 --reads-per-step 128 --legs main > gpurun_out/${R}_bench_short.txt 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 300 --csv --log-file gpurun_out/${R}_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --reads-per-step 128 --legs main > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 300 python tools/profile_workload.py 96 2 > gpurun_out/${R}_prof_workload.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_gru_fused2_kernel|tc_conv4_kernel' -s 8 -c 4 -f -o gpurun_out/${R}_prof_main \
    python tools/profile_workload.py 96 2 > gpurun_out/${R}_ncu_full.log 2>&1
tail -2 gpurun_out/${R}_ncu_full.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max"
timeout 300 python tools/profile_workload.py 200 1 auto RNN > gpurun_out/${R}_prof_rnn.txt 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none -k regex:'tc_gru_fused2_kernel' -c 3 --csv --log-file gpurun_out/${R}_ncu_rnn_only.csv \
    python tools/profile_workload.py 200 1 auto RNN > gpurun_out/${R}_ncu_rnn.log 2>&1
timeout 300 python tools/profile_workload.py 200 1 auto ResNet > gpurun_out/${R}_prof_resnet.txt 2>&1 && \
timeout 600 ncu --metrics $M --clock-control none -k regex:'tc_conv4_kernel|tc_head_conv_kernel' -c 2 --csv --log-file gpurun_out/${R}_ncu_resnet_only.csv \
    python tools/profile_workload.py 200 1 auto ResNet > gpurun_out/${R}_ncu_resnet.log 2>&1
timeout 300 python tests/tools/accuracy_check.py > gpurun_out/${R}_accuracy.txt 2>&1; tail -1 gpurun_out/${R}_accuracy.txt
timeout 300 python tests/tools/accuracy_check.py 200 >> gpurun_out/${R}_accuracy.txt 2>&1; tail -1 gpurun_out/${R}_accuracy.txt
timeout 600 python tools/tail_accuracy.py 12 256 > gpurun_out/${R}_tail_accuracy.txt 2>&1; tail -1 gpurun_out/${R}_tail_accuracy.txt
timeout 300 python tools/latency_long_reads.py 1 8 64 > gpurun_out/${R}_long_reads.txt 2>&1; tail -3 gpurun_out/${R}_long_reads.txt
timeout 600 python tools/soak.py > gpurun_out/${R}_soak.txt 2>&1; tail -2 gpurun_out/${R}_soak.txt
tools/mma_rate.bin > gpurun_out/${R}_mma_rate.txt 2>&1
nvidia-smi -q -d POWER,CLOCK > gpurun_out/${R}_nvidia_smi_power.txt 2>&1
