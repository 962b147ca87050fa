#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of device time)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name).replace("(anonymous namespace)::", "")
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        agg[name][0] += 1; agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {us / tot * 100:.1f} % |")
    print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100 % |")


if __name__ == "__main__":
    main(sys.argv[1])
