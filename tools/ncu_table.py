#!/usr/bin/env python3
"""One markdown row per kernel from `ncu --set full` reports: duration, DRAM bytes and bandwidth against the measured
copy peak (MEASURED_PEAKS.json, else the 6 549 GB/s this pool measured), issue-slot / FMA / ALU pipe utilisation, L2 hit rate,
occupancy.  For kernels captured several times (sort passes, scans of different sizes) the LARGEST launch is listed.
usage: python tools/ncu_table.py report.ncu-rep [more.ncu-rep ...] > profiles/<round>_ncu_all_kernels.md"""
import csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
try:
    PEAK = json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    PEAK = 6549.4
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
         "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
COLS = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs")]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    for r in rows[2:]:
        def val(k):
            return float(r[ix[k]].replace(",", "")) * SCALE.get(units[ix[k]], 1.0) if k in ix and r[ix[k]] not in ("", "n/a") else float("nan")
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("gsb::", "")
        t = val("gpu__time_duration.sum")
        b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        yield name, dict(t=t, bytes=b, grid=r[ix["Grid Size"]], block=r[ix["Block Size"]], extra=[val(k) for k, _ in COLS])


def main():
    best = {}
    for rep in sys.argv[1:]:
        for name, d in rows_of(rep):
            if name not in best or d["t"] > best[name]["t"]:
                best[name] = d
    print(f"| kernel | grid x block | time us | DRAM MB | DRAM GB/s | % of {PEAK:.0f} GB/s | " + " | ".join(h for _, h in COLS) + " |")
    print("|---|---|---|---|---|---|" + "---|" * len(COLS))
    for name, d in sorted(best.items(), key=lambda kv: -kv[1]["t"]):
        gbs = d["bytes"] / d["t"] / 1e9
        ex = " | ".join(f"{x:.1f}" if x == x else "-" for x in d["extra"])
        print(f"| `{name}` | {d['grid']} x {d['block']} | {d['t'] * 1e6:.1f} | {d['bytes'] / 1e6:.1f} | {gbs:.0f} | {100 * gbs / PEAK:.1f} | {ex} |")


if __name__ == "__main__":
    main()
