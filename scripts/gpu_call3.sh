#!/bin/bash
# round-2 GPU pass: parity tests, conv / upsample microbench, bench line with a per-launch dump
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -n 8 gpurun_out/r2_pytest_gpu.log
grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest_gpu.log | head -30
timeout 900 python scripts/microbench.py --only convtc,upconv --iters 10 > gpurun_out/r2_micro_conv.jsonl 2> gpurun_out/r2_micro.err
tail -n 3 gpurun_out/r2_micro.err
BENCH_DUMP_LAUNCHES=gpurun_out/r2_launch_dump.json timeout 1500 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_bench.err
tail -n 5 gpurun_out/r2_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_1gpu.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
    for k,v in d['kernels_ms_per_step'].items(): print('  ',k,v)
    for k in ('dmd_stage','loss_parity'):
        print(k, json.dumps(d.get(k))[:1500])
except Exception as e: print('bench parse failed',e)
for l in open('gpurun_out/r2_micro_conv.jsonl'):
    d=json.loads(l); print(' '.join(f'{k}={v}' for k,v in d.items()))
PY
