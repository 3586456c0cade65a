#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean time and share.  Usage: summarize_launches.py launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        t_us = t / 1e3 if unit in ("nsecond", "ns") else t * ({"usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3}.get(unit, 1))
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        rows.append((int(r["ID"]), name, t_us, r["Grid Size"], r["Block Size"]))
    rows = [r for r in rows if r[0] >= skip]
    agg = OrderedDict()
    for _, name, t, g, b in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    print("# %s  (launch ids >= %d, %d launches, %.1f us total; cold-cache, serialised: compare shares)" % (path, skip, len(rows), total))
    print("%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "mean_us", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %10.1f %6.1f%%" % (name[:60], n, t, t / n, 100 * t / total))


if __name__ == "__main__":
    main()
