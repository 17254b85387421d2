#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X).
Usage: python tools/launch_summary.py profiles/launches_r1_final.csv [first_id] > profiles/..._summary.txt
Only launches with ID >= first_id are counted (skip warm-up); shares, not absolute times, are the
evidence: ncu serialises launches and runs them cold-cache."""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if int(r[ii]) < first:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        tot[name] += float(r[vi].replace(",", "")) / 1e3
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{path}: launches with ID >= {first}; total {total:.1f} us")
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{name[:72]:72s} n={cnt[name]:3d} sum={t:10.1f}us share={100 * t / total:5.1f}% avg={t / cnt[name]:9.1f}us")


if __name__ == "__main__":
    main()
