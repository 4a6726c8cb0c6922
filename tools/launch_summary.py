#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (share of GPU time per kernel)."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        tot[r[ki]] += v
        cnt[r[ki]] += 1
    total = sum(tot.values())
    print("| share of GPU time | launches | mean per launch (us) | kernel |")
    print("|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v / total < 0.005:
            continue
        print(f"| {100 * v / total:.1f}% | {cnt[k]} | {v / cnt[k]:.1f} | `{k[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
