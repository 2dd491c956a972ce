#!/bin/bash
# A/B several builds of the library on the same box: tools/ab_libs.sh <variant.so>... ("-" = the default library)
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), 'k2', round(k['k2_conv_stack']['ms']/d['steps'],2), 'k4', round(k['k4_gru_recurrence']['ms']/d['steps'],2), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"; }
for r in 1 2; do
  for v in "$@"; do
    echo "$v"
    if [ "$v" = "-" ]; then timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --legs main 2>&1 | tail -1 | show
    else CF_LIB_PATH=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --legs main 2>&1 | tail -1 | show; fi
  done
done
