#!/bin/bash
# parity + bench + microbench, no ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python scripts/microbench.py --iters 10 ${MICRO_ARGS} > gpurun_out/micro.jsonl 2> gpurun_out/micro.err
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roof',d['roofline']['achieved'],d['roofline']['frac'],'wgrad',d['wgrad'])
    for k,v in d['kernels_ms_per_step'].items(): print('  ',k,v)
except Exception as e: print('bench parse failed',e)
for l in open('gpurun_out/micro.jsonl'):
    d=json.loads(l); print(' '.join(f'{k}={v}' for k,v in d.items() if k!='algo_bytes'))
PY
if [ "${PROFILE_STEP:-0}" = "1" ]; then
  timeout 600 python scripts/profile_step.py > gpurun_out/step_profile.txt 2>&1
  head -75 gpurun_out/step_profile.txt | cut -c1-200
fi
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -20
