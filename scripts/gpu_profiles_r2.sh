#!/bin/bash
# round-2 ncu evidence for profiles/: launch list of one bench step, full captures of the dominant kernels (summarised ON the box:
# only the small .md / .csv summaries travel back, gpurun_out/ is capped at 64 MiB), torch.profiler table.
# PROFILE_PARTS selects: step launches conv up gn dmd
PARTS=${PROFILE_PARTS:-"step launches conv up gn dmd"}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
has() { [[ " $PARTS " == *" $1 "* ]]; }
B="python bench.py --steps 1 --warmup 3 --quick --no-cuda-graph"
summ() { python scripts/summarize_ncu.py full gpurun_out/$1.ncu-rep gpurun_out/$1.md && rm -f gpurun_out/$1.ncu-rep; }
if has step; then
timeout 600 python scripts/profile_step.py > gpurun_out/r2_step_profile.txt 2>&1
fi
if has launches; then
# one whole eager step somewhere after the warm-up (7 steps x ~1000 launches run in total)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 4200 -c 1100 --csv --log-file gpurun_out/r2_launches.csv \
    $B > gpurun_out/r2_ncu_bench.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/r2_launches.csv gpurun_out/r2_launches_bench.md
fi
if has conv; then
# a window of consecutive conv launches (halo pair / transposed / per-tap tiles, weight gradients) out of the third step
timeout 1200 ncu --set full --clock-control none -k regex:"conv_tc2h_kernel|conv_tcT_kernel|conv_tc_wgrad|conv_tc2_kernel" -s 420 -c 48 -o gpurun_out/r2_conv_kernels_full -f \
    $B > gpurun_out/r2_ncu_conv.log 2>&1
summ r2_conv_kernels_full
fi
if has up; then
# sub-pixel Upsample: two-kernel form vs sub-pixel form, forward + backward, 3 layer shapes (microbench launches 4 per variant)
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc2h_kernel|conv_tc_wgrad2_kernel" -c 40 -o gpurun_out/r2_upconv_kernels_full -f \
    python scripts/microbench.py --only upconv --iters 1 > gpurun_out/r2_ncu_up.log 2>&1
summ r2_upconv_kernels_full
fi
if has gn; then
timeout 900 ncu --set full --clock-control none -k regex:"gn_bwd|gn_apply|lpips_dist|pool_tap|maxpool|adamw_ema|pack_dgrad|subpixel" -s 120 -c 16 -o gpurun_out/r2_hbm_kernels_full -f \
    $B > gpurun_out/r2_ncu_gn.log 2>&1
summ r2_hbm_kernels_full
fi
if has dmd; then
timeout 600 ncu --set full --clock-control none -k regex:"dmd_loss|dmd_mix" -c 6 -o gpurun_out/r2_dmd_kernels_full -f \
    python scripts/microbench.py --iters 1 --only dmd > gpurun_out/r2_ncu_dmd.log 2>&1
summ r2_dmd_kernels_full
fi
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out; ls -la gpurun_out/r2_*full.md gpurun_out/r2_launches_bench.md 2>/dev/null
if [ -n "$ALSO_TEST" ]; then
timeout 600 python -m pytest tests -q -m gpu -k "lpips or gan or tokenizer or stress or conv_relu" > gpurun_out/r2_pytest_subset.log 2>&1; tail -n 3 gpurun_out/r2_pytest_subset.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; cut -c1-200 gpurun_out/r2_bench_quick.json
timeout 600 python bench.py --workload stress512 --steps 10 --quick > gpurun_out/r2_bench_stress512.json 2> gpurun_out/r2_bench_stress512.err; cut -c1-200 gpurun_out/r2_bench_stress512.json
timeout 600 python bench.py --workload dmd --steps 10 --quick > gpurun_out/r2_bench_dmd.json 2> gpurun_out/r2_bench_dmd.err; cut -c1-200 gpurun_out/r2_bench_dmd.json
fi
