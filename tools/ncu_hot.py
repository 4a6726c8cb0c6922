#!/usr/bin/env python
"""Per-instruction view of one captured launch: opcode mix by executed warp instructions and the top stall sites.
   python tools/ncu_hot.py <rep> <launch index>"""
import csv
import subprocess
import sys
from collections import Counter


def main(rep, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines())]
    print(rows[0][1][:120])
    h = rows[1]
    ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    data = [(i, r[ia].strip(), int(r[ie] or 0), int(r[iss] or 0)) for i, r in enumerate(rows[2:]) if len(r) > max(ia, ie, iss) and (r[ie] or '0').isdigit() and r[ia].strip()]
    tot = sum(d[2] for d in data)
    print("warp instructions", tot)
    c = Counter()
    for _, s, e, _ in data:
        t = s.split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.split(".")[0]] += e
    print("opcode mix %:", [(k, round(v / tot * 100, 1)) for k, v in c.most_common(24)])
    samples = sum(d[3] for d in data)
    print("top stall sites (line, % of samples, executed, instruction):")
    for d in sorted(data, key=lambda d: -d[3])[:16]:
        print(f"  {d[0]:5d} {100 * d[3] / max(samples, 1):5.1f}% {d[2]:9d}  {d[1][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
