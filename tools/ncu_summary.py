#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel duration, DRAM traffic, FP64 pipe, occupancy, stalls.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % busy"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]] for r in data]
    print("| metric | " + " | ".join(n.split("(")[0].replace("void ", "")[:44] for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in idx:
            continue
        i = idx[key]
        vals = []
        for r in data:
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
