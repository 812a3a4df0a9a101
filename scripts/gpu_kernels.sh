#!/bin/bash
# First GPU pass: kernel-level parity. Risky tcgen05 tests run in their own processes under timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "not tcgen05 and not conv_tc" > gpurun_out/k_basic.log 2>&1
echo "basic rc=$?" >> gpurun_out/k_basic.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "tcgen05" > gpurun_out/k_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/k_tc.log
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "conv_tc_many" > gpurun_out/k_tc_big.log 2>&1
echo "tcbig rc=$?" >> gpurun_out/k_tc_big.log
for f in k_basic k_tc k_tc_big; do tail -n 3 gpurun_out/$f.log; done
