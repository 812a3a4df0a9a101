#!/bin/bash
# same-box A/B of one environment switch on the quick bench line: VAR=name A=value B=value [REPS=2]
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -k "group_norm or resnet or decoder" > gpurun_out/r2_pytest_ab.log 2>&1; tail -n 2 gpurun_out/r2_pytest_ab.log
for i in $(seq 1 ${REPS:-2}); do
  for v in "$A" "$B"; do
    env $VAR=$v timeout 300 python bench.py --steps 30 --warmup 5 --quick 2> gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['kernels_ms_per_step']
print('$VAR=$v', 'ms/step', d['ms_per_step'], 'img/s', d['value'], 'gn_bwd', k.get('gn_bwd'), 'gn_apply', k.get('gn_apply'), 'clk', d['clocks']['sm_mhz'])"
  done
done
