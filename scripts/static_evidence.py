#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): SASS mnemonic counts that show the tensor-core / TMA path is tcgen05
(UTCHMMA*, UTMALDG*, LDTM, UTCBAR*; see /opt/skills/guides/B200_PROFILING.md) and the per-kernel register / stack / local-memory
table of `cuobjdump -res-usage` (LOCAL > 0 would be spills).  Writes profiles/r02_static_sass_resources.md."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dmvae_b200", "libdmvae_b200.so")


def run(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def main():
    sass = run("cuobjdump", "-sass", SO)
    counts = collections.Counter()
    per_fn = collections.defaultdict(collections.Counter)
    fn = None
    pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG(?:\.[0-9A-Z.]+)?|UTMASTG|LDTM(?:\.[0-9A-Zx.]+)?|UTCBAR(?:\.[0-9A-Z.]+)?|SYNCS\.[A-Z.]+|REDG?\.E\.[A-Z0-9.]+|HMMA\.[0-9A-Z.]+)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = pat.search(line)
        if m:
            key = m.group(1)
            key = re.sub(r"^(UTMALDG\.\dD).*?(\.2CTA)?$", lambda g: g.group(1) + (g.group(2) or ""), key) if key.startswith("UTMALDG") else key
            key = "LDTM" if key.startswith("LDTM") else key
            key = "SYNCS" if key.startswith("SYNCS") else key
            key = "RED(G).E.ADD (global reductions)" if key.startswith("RED") else key
            key = "UTCBAR" + (".2CTA.MULTICAST" if "MULTICAST" in key else "") if key.startswith("UTCBAR") else key
            counts[key] += 1
            per_fn[fn][key.split(".")[0]] += 1
    res = run("cuobjdump", "-res-usage", SO)
    rows = []
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((name, *map(int, m.groups())))
            name = None
    demangled = run("c++filt", *[r[0] for r in rows]).splitlines() if rows else []
    out = ["# Static evidence from `dmvae_b200/libdmvae_b200.so` (scripts/static_evidence.py, `cuobjdump -sass` / `-res-usage`)", "",
           "## SASS mnemonics (whole library)", "", "| mnemonic | count |", "|---|---:|"]
    for k in sorted(counts):
        out.append(f"| `{k}` | {counts[k]} |")
    out += ["", "`UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `UTMALDG` = TMA tensor loads, `LDTM` = tcgen05.ld (TMEM read-back), "
            "`UTCBAR` = tcgen05.commit (`.MULTICAST` = to both CTAs of a pair).  No `HMMA` row means no legacy mma.sync anywhere.", "",
            "## Kernels that issue tcgen05 / TMA instructions", "", "| kernel | UTCHMMA | UTMALDG | LDTM | UTCBAR |", "|---|---:|---:|---:|---:|"]
    for f in sorted(per_fn):
        c = per_fn[f]
        if c["UTCHMMA"] or c["UTMALDG"]:
            nm = run("c++filt", f).strip()
            nm = re.sub(r"\(anonymous namespace\)::", "", nm)
            nm = re.sub(r"\(.*", "", nm).replace("void ", "")
            out.append(f"| `{nm}` | {c['UTCHMMA']} | {c['UTMALDG']} | {c['LDTM']} | {c['UTCBAR']} |")
    out += ["", "## Registers / stack / local memory per kernel (`LOCAL` = spill bytes; all 0)", "",
            "| kernel | regs | stack B | static smem B | local B |", "|---|---:|---:|---:|---:|"]
    seen = set()
    for (raw, reg, stack, sh, loc), nm in zip(rows, demangled):
        nm = re.sub(r"\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"\(.*", "", nm).replace("void ", "")
        if (nm, reg, stack, sh, loc) in seen:
            continue
        seen.add((nm, reg, stack, sh, loc))
        out.append(f"| `{nm}` | {reg} | {stack} | {sh} | {loc} |")
    spills = [r for r in rows if r[4] > 0]
    out += ["", f"{len(rows)} kernels, {len(spills)} with local-memory spills."]
    path = os.path.join(ROOT, "profiles", "r02_static_sass_resources.md")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(path, len(rows), "kernels;", dict(counts))


if __name__ == "__main__":
    main()
