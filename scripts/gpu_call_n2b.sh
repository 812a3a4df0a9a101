#!/bin/bash
# short multi-GPU validation of the distributed teardown (tight timeouts)
N=${NGPU:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 240 $TR scripts/check_ddp_sync.py > gpurun_out/r2_ddp_sync_n$N.txt 2>&1
echo "ddp_sync rc=$? after ${SECONDS}s" | tee -a gpurun_out/r2_ddp_sync_n$N.txt
grep -E "world|OK|teardown" gpurun_out/r2_ddp_sync_n$N.txt | tail -6
SECONDS=0
timeout 420 $TR bench.py --gpus $N --steps 20 --warmup 5 --parity-steps ${PARITY_STEPS:-0} --no-dmd-stage > gpurun_out/r2_bench_${N}gpu_quick.json 2> gpurun_out/r2_bench_${N}gpu_quick.err
echo "bench rc=$? after ${SECONDS}s" | tee -a gpurun_out/r2_bench_${N}gpu_quick.err
grep -E "teardown|Error|error" gpurun_out/r2_bench_${N}gpu_quick.err | tail -5
cut -c1-600 gpurun_out/r2_bench_${N}gpu_quick.json
