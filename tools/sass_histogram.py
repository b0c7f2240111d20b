#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libgsb.so (cuobjdump -sass): the evidence behind statements such as "the
rasterisers issue packed FFMA2", "records are staged by LDGSTS / UBLKCP", "gradients leave through RED".

    python tools/sass_histogram.py [--lib path/to/libgsb.so] [--out profiles/r2_sass_opcodes.md]

Counts are STATIC instruction counts per kernel (all paths of all loops once), not executed instructions; the
per-evaluation figures quoted in DESIGN.md come from the innermost loop bodies listed at the end of the report.
"""
from __future__ import annotations

import argparse
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "FMNMX", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "UBLKRED",
        "REDG", "ATOMG", "ATOMS", "LDGMC", "STGMC", "BAR", "SYNCS", "SHFL", "VOTE", "MATCH", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=str(ROOT / "gaussiansplattingmlx_b200" / "libgsb.so"))
    ap.add_argument("--out", default=None)
    ap.add_argument("--filter", default=r"gsb::k_", help="regex on the demangled kernel name")
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    hist = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            hist[cur][op] += 1
            hist[cur]["*"] += 1
            if op == "MUFU" and mods:
                hist[cur]["MUFU." + mods.split(".")[1]] += 1
    names = demangle(list(hist))
    rows = []
    for k, c in hist.items():
        dn = names.get(k, k)
        if not re.search(args.filter, dn):
            continue
        short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("gsb::", "")
        rows.append((short, c))
    rows.sort(key=lambda r: -r[1]["*"])
    lines = [f"# SASS opcode histogram of `{Path(args.lib).name}` ({', '.join(arch)})", "",
             "Static instruction counts per kernel from `cuobjdump -sass` (tools/sass_histogram.py). FFMA2 / FMUL2 / FADD2 = packed",
             "f32x2 arithmetic; LDGSTS = 16-byte `cp.async` global->shared; UBLKCP = TMA 1-D bulk copy (`cp.async.bulk`); UBLKRED = bulk",
             "reduce-add to global; REDG = `red.global.add`; LDGMC / STGMC = `multimem.ld_reduce` / `multimem.st` (NVLS); SYNCS = mbarrier operations; MATCH = `match.any` (onesweep ranking).", "",
             "| kernel | total | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]
    for short, c in rows:
        lines.append(f"| `{short}` | {c['*']} | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")
    mufu = collections.Counter()
    for short, c in rows:
        for k, v in c.items():
            if k.startswith("MUFU."):
                mufu[(short, k)] += v
    if mufu:
        lines += ["", "MUFU by function: " + ", ".join(f"`{s}` {k} x{v}" for (s, k), v in sorted(mufu.items()))]
    text = "\n".join(lines) + "\n"
    if args.out:
        Path(args.out).write_text(text)
        print(f"wrote {args.out} ({len(rows)} kernels)")
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main()
