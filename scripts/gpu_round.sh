#!/bin/bash
# Full GPU pass: parity tests, smoke, bench (1 GPU), ncu launch list + one full capture of the top kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 40 -c 3 -o gpurun_out/prof_conv_tc -f \
      python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
for f in pytest_gpu.log smoke.log bench.err; do echo "== $f"; tail -n 4 gpurun_out/$f; done
head -c 1500 gpurun_out/bench.json
