#!/bin/bash
# N-GPU bench line with every leg (tight timeouts)
N=${NGPU:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --parity-steps ${PARITY_STEPS:-20} ${BENCH_ARGS} > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$? after ${SECONDS}s" | tee -a gpurun_out/r2_bench_${N}gpu.err
grep -E "teardown|Error|error" gpurun_out/r2_bench_${N}gpu.err | tail -5
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'run',d['run'])
    ds=d.get('dmd_stage') or {}; print('dmd_stage', ds.get('value'), ds.get('ms_per_iteration'), ds.get('exchange'))
    lp=d.get('loss_parity') or {}; print('loss_parity', json.dumps(lp.get('same_weights_per_step'))[:600], lp.get('verdict'))
except Exception as e: print('bench parse failed',e)
PY
SECONDS=0
timeout 300 $TR bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2_bench_${N}gpu_ref.json 2> gpurun_out/r2_bench_${N}gpu_ref.err
echo "reference arm rc=$? after ${SECONDS}s"; cut -c1-200 gpurun_out/r2_bench_${N}gpu_ref.json
