#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): exchange consistency eager + graph replay, bench line with all legs
N=${NGPU:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR scripts/check_ddp_sync.py > gpurun_out/r2_ddp_sync_n$N.txt 2>&1
echo "ddp_sync rc=$?" >> gpurun_out/r2_ddp_sync_n$N.txt
grep -E "world|OK|Error|error|rc=" gpurun_out/r2_ddp_sync_n$N.txt | tail -12
NCCL_DEBUG=WARN timeout 1800 $TR bench.py --gpus $N --steps 20 --warmup 5 --parity-steps ${PARITY_STEPS:-30} > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?" >> gpurun_out/r2_bench_${N}gpu.err
tail -n 6 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'run',d['run'])
    for k in ('dmd_stage','loss_parity'):
        print(k, json.dumps(d.get(k))[:2500])
except Exception as e: print('bench parse failed',e)
PY
