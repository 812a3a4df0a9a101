#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the main kernels
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py > gpurun_out/step_profile.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_tc_kernel|conv_tc_wgrad" -s 60 -c 6 -o gpurun_out/prof_conv -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dmd_loss|dmd_mix|l1l2|reparam" -c 8 -o gpurun_out/prof_loss -f \
    python scripts/microbench.py --iters 1 --only dmd,l1l2,reparam > gpurun_out/ncu_loss.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_bwd|gn_apply|gn_stats|lpips" -c 10 -o gpurun_out/prof_gn2 -f \
    python scripts/microbench.py --iters 1 --only gn_bwd,gn_apply,gn_stats,lpips > gpurun_out/ncu_gn2.log 2>&1
ls -la gpurun_out/*.ncu-rep
