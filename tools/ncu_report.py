#!/usr/bin/env python
"""Key numbers of `ncu --set full` reports (read on the CPU box): python tools/ncu_report.py gpurun_out/a.ncu-rep … > profiles/x.md"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
        ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall lg_throttle %"),
        ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short_scoreboard %"),
        ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "stall mio_throttle %"),
        ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math_pipe_throttle %")]


def main(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            print(f"### `{r[col['Kernel Name']].split('(')[0]}` — {p}\n")
            print("| metric | value |")
            print("|---|---|")
            for key, label in WANT:
                if key in col:
                    print(f"| {label} (`{key}`) | {r[col[key]]} {units[col[key]]} |")
            try:
                t = float(r[col["gpu__time_duration.sum"]]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[col["gpu__time_duration.sum"]]]
                f = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
                b = float(r[col["dram__bytes_read.sum"]]) * f[units[col["dram__bytes_read.sum"]]] + \
                    float(r[col["dram__bytes_write.sum"]]) * f[units[col["dram__bytes_write.sum"]]]
                print(f"| **DRAM traffic / duration** | {b / 1e9:.3f} GB / {t * 1e3:.3f} ms = {b / t / 1e9:.0f} GB/s |")
            except Exception:
                pass
            print()


if __name__ == "__main__":
    main(sys.argv[1:])
