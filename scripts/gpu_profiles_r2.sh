#!/bin/bash
# round-2 ncu evidence for profiles/: launch list of one bench step, full captures of the dominant kernels, torch.profiler table.
# PROFILE_PARTS selects: step launches conv up gn dmd
PARTS=${PROFILE_PARTS:-"step launches conv up gn dmd"}
mkdir -p gpurun_out
has() { [[ " $PARTS " == *" $1 "* ]]; }
B="python bench.py --steps 1 --warmup 3 --quick --no-cuda-graph"
if has step; then
timeout 600 python scripts/profile_step.py > gpurun_out/r2_step_profile.txt 2>&1
fi
if has launches; then
# one whole eager step somewhere after the warm-up (7 steps x ~1000 launches run in total)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1100 --csv --log-file gpurun_out/r2_launches.csv \
    $B > gpurun_out/r2_ncu_bench.log 2>&1
fi
if has conv; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2h_kernel" -s 250 -c 4 -o gpurun_out/r2_prof_conv_halo -f \
    $B > gpurun_out/r2_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tcT_kernel" -s 60 -c 3 -o gpurun_out/r2_prof_conv_t -f \
    $B >> gpurun_out/r2_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_wgrad" -s 110 -c 5 -o gpurun_out/r2_prof_wgrad -f \
    $B >> gpurun_out/r2_ncu_conv.log 2>&1
fi
if has up; then
# sub-pixel Upsample: two-kernel form vs sub-pixel form, forward + backward, 3 layer shapes (microbench launches 4 per variant)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2h_kernel|conv_tc_wgrad2_kernel" -c 40 -o gpurun_out/r2_prof_upconv -f \
    python scripts/microbench.py --only upconv --iters 1 > gpurun_out/r2_ncu_up.log 2>&1
fi
if has gn; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gn_bwd|gn_apply|lpips_dist|pool_tap|maxpool|adamw_ema" -s 100 -c 14 -o gpurun_out/r2_prof_gn -f \
    $B > gpurun_out/r2_ncu_gn.log 2>&1
fi
if has dmd; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dmd_loss|dmd_mix" -c 6 -o gpurun_out/r2_prof_dmd -f \
    python scripts/microbench.py --iters 1 --only dmd > gpurun_out/r2_ncu_dmd.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep 2>/dev/null
