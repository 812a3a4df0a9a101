#!/bin/bash
# one pass of end-of-round evidence: bench lines (ours, reference arm, stress512, dmd), conv microbench, ncu launch list + captures
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ev_bench_1gpu.json 2> gpurun_out/ev_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ev_bench_reference.json 2>> gpurun_out/ev_bench.err
timeout 300 python bench.py --workload stress512 --steps 10 --no-cpu-baseline > gpurun_out/ev_bench_stress512.json 2>> gpurun_out/ev_bench.err
timeout 300 python bench.py --workload dmd --steps 10 --no-cpu-baseline > gpurun_out/ev_bench_dmd.json 2>> gpurun_out/ev_bench.err
timeout 300 python scripts/microbench.py --only convtc --iters 10 > gpurun_out/ev_micro_conv.jsonl 2> gpurun_out/ev_micro.err
PROFILE_PARTS="${PROFILE_PARTS:-launches conv}" bash scripts/gpu_profiles.sh
tail -n 3 gpurun_out/ev_bench.err
for f in gpurun_out/ev_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("roofline") or {}).get("achieved"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(sys.argv[1], "parse failed", e)
PY
done
