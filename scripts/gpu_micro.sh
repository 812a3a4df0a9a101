#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/microbench.py --iters 10 > gpurun_out/micro.jsonl 2> gpurun_out/micro.err
echo "rc=$?" >> gpurun_out/micro.err
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_|bias_grad" -c 14 -o gpurun_out/prof_gn -f \
   python scripts/microbench.py --iters 1 --only gn_stats,gn_apply,gn_bwd,bias_grad > gpurun_out/ncu_gn.log 2>&1
fi
cat gpurun_out/micro.jsonl | cut -c1-200; tail -n 3 gpurun_out/micro.err
