#!/usr/bin/env python
"""Summarise an .ncu-rep (one line block per distinct kernel) into markdown for profiles/."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of ncu peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    seen = set()
    for v in rows[2:]:
        name = v[h.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        print(f"### `{name[:140]}`\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in h:
                i = h.index(k)
                print(f"| {label} (`{k}`) | {v[i]} {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
