#!/usr/bin/env python3
"""Static check of the compiled rasterisers: the dense Gaussian loops of k_raster_bwd must not carry register shuffles.

The backward sits at its 128-register cap (16 one-warp CTAs per SM).  Small source changes around the loops - a conditional
update of one half of a packed f32x2 state register, a second definition point of the loop-carried state - make ptxas
keep the pairs split across the loop: +12 ... 45 MOV / IMAD.MOV per Gaussian and 0.74 -> 0.87 ms at C3 (DESIGN.md
section 7).  The symptom is visible without a GPU, in the SASS: this script lists every loop of a kernel with its
instruction mix and fails if a loop dominated by packed arithmetic holds more than a handful of MOVs.

usage: python tools/check_sass.py [libgsb.so]      (exit code 1 on a regression)"""
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
MOV_LIMIT = 16         # per loop (both clamp variants of a dense loop together, ~430 instructions; healthy builds: 7-10, split pairs: 23-33)


_dump = {}


def kernel_sass(so: Path, name: str):
    if so not in _dump:
        _dump[so] = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout.split("\n")
    out = _dump[so]
    ins, on = [], False
    for line in out:
        if "Function :" in line:
            on = name in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def loops(ins, min_len=60, max_len=1000):
    """(first, last, histogram) of every backward branch spanning min_len ... max_len instructions"""
    index = {a: i for i, (a, _) in enumerate(ins)}
    found = []
    for i, (_, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)* (?:\w+, )?0x([0-9a-f]+)", text)
        if not (m and text.startswith("@")):
            continue
        target = int(m.group(1), 16)
        if target in index and min_len < i - index[target] < max_len:
            hist = {}
            for _, t in ins[index[target]:i + 1]:
                parts = t.split()
                op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
                hist[op] = hist.get(op, 0) + 1
            found.append((index[target], i, hist))
    return found


def check(so: Path, verbose=True):
    bad = []
    for kernel in ("k_raster_bwdILb0", "k_raster_bwdILb1"):
        for first, last, hist in loops(kernel_sass(so, kernel)):
            packed = hist.get("FFMA2", 0) + hist.get("FMUL2", 0) + hist.get("FADD2", 0)
            movs = hist.get("MOV", 0) + hist.get("IMAD", 0)      # IMAD.MOV shows up as IMAD
            dense = packed >= 60
            if verbose and dense:
                print(f"{kernel}: dense loop {first}-{last}: {last - first + 1} instructions, {packed} packed, "
                      f"{hist.get('MOV', 0)} MOV, {hist.get('IMAD', 0)} IMAD, {hist.get('MUFU', 0)} MUFU")
            if dense and hist.get("MOV", 0) > MOV_LIMIT:
                bad.append((kernel, first, last, hist.get("MOV", 0), movs))
    return bad


def main():
    so = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gaussiansplattingmlx_b200" / "libgsb.so"
    if not shutil.which("cuobjdump"):
        print("cuobjdump not found")
        return 0
    bad = check(so)
    for kernel, first, last, mov, _ in bad:
        print(f"REGRESSION {kernel}: loop {first}-{last} carries {mov} MOV (limit {MOV_LIMIT}): packed state registers are split")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
