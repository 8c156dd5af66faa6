#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total time and share.
Usage: python tools/ncu_launches.py launches.csv [skip_first_n_launches]   (skip = warm-up launches to leave out)"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("eigb200::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.3f ms (cold-cache, serialised under ncu)" % (len(rows), tot))
print("%-60s %8s %10s %7s" % ("kernel", "launches", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %8d %10.3f %6.1f%%" % (k[:60], v[0], v[1], 100 * v[1] / tot))
