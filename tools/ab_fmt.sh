#!/bin/bash
# A/B of the operand formats on one box: tests, accuracy, bench with CF_TC_FMT=0 (bf16x3) vs default (f16e5)
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), 'k2', round(k['k2_conv_stack']['ms']/d['steps'],2), 'k4', round(k['k4_gru_recurrence']['ms']/d['steps'],2), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['roofline']['frac'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "accuracy f16e5"; timeout 300 python tests/tools/accuracy_check.py 2>&1 | tail -1
echo "accuracy bf16x3"; CF_TC_FMT=0 timeout 300 python tests/tools/accuracy_check.py 2>&1 | tail -1
for r in 1 2; do
  echo "f16e5"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show
  echo "bf16x3"; CF_TC_FMT=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show
done
