#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>10s} {'share':>6s}")
    for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:70s} {c:5d} {v / 1e3:10.1f} {100 * v / tot:5.1f}%")
    print(f"total {tot / 1e6:.3f} ms over {sum(c for c, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
