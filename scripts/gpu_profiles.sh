#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the main kernels
# PROFILE_PARTS selects what to run (default: everything): step launches conv loss gn
PARTS=${PROFILE_PARTS:-"step launches conv loss gn"}
mkdir -p gpurun_out
has() { [[ " $PARTS " == *" $1 "* ]]; }
if has step; then
timeout 600 python scripts/profile_step.py > gpurun_out/step_profile.txt 2>&1
fi
if has launches; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
fi
if has conv; then
# halo-resident CTA-pair tiles (N = 256 and N = 128 instantiations), transposed 128-channel tiles, per-tap pair / single tiles, wgrad
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2h_kernel" -s 284 -c 3 -o gpurun_out/prof_conv_halo -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tcT_kernel" -s 51 -c 2 -o gpurun_out/prof_conv_t -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_wgrad" -s 100 -c 5 -o gpurun_out/prof_conv -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_conv.log 2>&1
fi
if has loss; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dmd_loss|dmd_mix|l1l2|reparam" -c 8 -o gpurun_out/prof_loss -f \
    python scripts/microbench.py --iters 1 --only dmd,l1l2,reparam > gpurun_out/ncu_loss.log 2>&1
fi
if has gn; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gn_bwd|gn_apply|gn_stats|lpips" -c 10 -o gpurun_out/prof_gn2 -f \
    python scripts/microbench.py --iters 1 --only gn_bwd,gn_apply,gn_stats,lpips > gpurun_out/ncu_gn2.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
