#!/usr/bin/env python3
"""Summarise an .ncu-rep (one `--set full` capture) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"), ("launch__waves_per_multiprocessor", "waves"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
]

def traffic_json(reps, out_path):
    """--traffic OUT.json rep...: average DRAM bytes (read + write) per launch of every kernel captured."""
    import json
    acc = {}
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("gsb::", "").split("<")[0]
            b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            t = float(r[hdr.index("gpu__time_duration.sum")])
            acc.setdefault(name, []).append((b, t, units[hdr.index("gpu__time_duration.sum")], r[hdr.index("Grid Size")]))
    res = {}
    for name, v in acc.items():
        big = max(x[0] for x in v)
        sel = [x for x in v if x[0] >= 0.5 * big]          # ignore the tiny launches of a multi-size kernel (sort passes)
        res[name] = {"dram_bytes_per_launch": sum(x[0] for x in sel) / len(sel), "launches_averaged": len(sel),
                     "source": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum"}
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--traffic":
        return traffic_json(sys.argv[3:], sys.argv[2])
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"{rep}: empty"); continue
        hdr, units = rows[0], rows[1]
        print(f"== {rep}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0][-60:]
            print(f"-- {name}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
            parts = []
            for key, short in WANT:
                if key in hdr:
                    v = r[hdr.index(key)]
                    try:
                        v = f"{float(v):.4g}"
                    except ValueError:
                        pass
                    parts.append(f"{short}={v}{units[hdr.index(key)] if short in ('time','dram_rd','dram_wr') else ''}")
            print("   " + "  ".join(parts))

if __name__ == "__main__":
    main()
