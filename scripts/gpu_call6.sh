#!/bin/bash
# round-2 1-GPU pass: all GPU tests, smoke, full bench line, the other workloads, reference arm, upsample microbench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -n 6 gpurun_out/r2_pytest_gpu.log
grep -E "^FAILED|^ERROR" gpurun_out/r2_pytest_gpu.log | head -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log; tail -n 2 gpurun_out/r2_smoke.log
BENCH_DUMP_LAUNCHES=gpurun_out/r2_launch_dump.json timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_bench.err; tail -n 3 gpurun_out/r2_bench.err
timeout 600 python bench.py --workload gan --steps 10 --quick > gpurun_out/r2_bench_gan.json 2> gpurun_out/r2_bench_gan.err; echo "gan rc=$?"
timeout 600 python bench.py --workload dmd --steps 10 --quick > gpurun_out/r2_bench_dmd.json 2> gpurun_out/r2_bench_dmd.err; echo "dmd rc=$?"
timeout 600 python bench.py --workload stress512 --steps 10 --quick > gpurun_out/r2_bench_stress512.json 2> gpurun_out/r2_bench_stress512.err; echo "stress512 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
timeout 600 python scripts/microbench.py --only upconv --iters 10 > gpurun_out/r2_micro_upconv.jsonl 2> gpurun_out/r2_micro_up.err; tail -n 2 gpurun_out/r2_micro_up.err
python - <<'PY'
import json
for f in ('r2_bench_1gpu','r2_bench_gan','r2_bench_dmd','r2_bench_stress512','r2_bench_reference'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f,'value',d.get('value'),'e2e',(d.get('e2e') or {}).get('value'),'ms/step',d.get('ms_per_step'),'roof',(d.get('roofline') or {}).get('achieved'),(d.get('roofline') or {}).get('frac'), d.get('run'))
    except Exception as e: print(f,'parse failed',e)
d=json.load(open('gpurun_out/r2_bench_1gpu.json'))
for k,v in d['kernels_ms_per_step'].items(): print('  ',k,v)
for k in ('dmd_stage','roofline_hbm','gpu_baseline','cpu_baseline'): print(k, json.dumps(d.get(k))[:700])
lp=d.get('loss_parity') or {}
print('loss_parity', json.dumps({k:lp.get(k) for k in ('same_weights_per_step','free_running_trajectories','verdict','seconds')})[:2500])
for l in open('gpurun_out/r2_micro_upconv.jsonl'): print(l.strip())
PY
