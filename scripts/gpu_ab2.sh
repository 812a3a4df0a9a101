#!/bin/bash
# full GPU suite, then a same-box A/B of the per-step launch trimming (zero pool + batched dgrad pack) on the quick bench line,
# replayed from the CUDA graph and launched eagerly
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; tail -n 3 gpurun_out/r2_pytest_gpu.log
for mode in "--no-cuda-graph" ""; do
  for v in 0 1 0 1; do
    env DMVAE_ZERO_POOL=$v DMVAE_BATCHED_DGRAD_PACK=$v timeout 300 python bench.py --steps 20 --warmup 4 --quick $mode 2> gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('mode[$mode] pool/pack=$v', 'ms/step', d['ms_per_step'], 'img/s', d['value'], 'clk', d['clocks']['sm_mhz'])"
  done
done
tail -n 5 gpurun_out/r2_ab.err
