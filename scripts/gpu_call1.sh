#!/bin/bash
# round-2 first GPU pass: parity tests, smoke, full bench line (all legs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_bench.err
tail -n 6 gpurun_out/r2_pytest_gpu.log; tail -n 2 gpurun_out/r2_smoke.log; tail -n 5 gpurun_out/r2_bench.err
grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest_gpu.log | head -20
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_1gpu.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roof',d['roofline']['achieved'],d['roofline']['frac'])
    for k in ('dmd_stage','roofline_hbm','gpu_baseline','loss_parity','cpu_baseline'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e: print('bench parse failed',e)
PY
