show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; print(round(d['value']/1e6,1), round(d['ms_per_step'],2), 'k2', round(k['k2_conv_stack']['ms']/d['steps'],2), 'k4', round(k['k4_gru_recurrence']['ms']/d['steps'],2), d['clocks'])"; }
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 4 3 2; do echo "CF_TC_CONV=$v"; CF_TC_CONV=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | show; done
