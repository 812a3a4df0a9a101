#!/usr/bin/env python
"""Turns ncu outputs under gpurun_out/ into the small text summaries kept in profiles/.
   python scripts/summarize_ncu.py launches gpurun_out/launches.csv  profiles/r01_launches.md
   python scripts/summarize_ncu.py full     gpurun_out/prof.ncu-rep  profiles/r01_prof.md"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(<[^(]{0,60}>)?)", name)
    return (m.group(1) if m else name)[:90]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        a = agg[short(row["Kernel Name"])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: {src}; {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms total\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"| {v[1]:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k}` |\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({src})\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{short(d.get('Kernel Name', '?'))}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"- {k} = {d[k]} {units[hdr.index(k)]}\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
