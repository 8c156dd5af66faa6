#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions by stall samples, with the dominant stall reason.
Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
si = ci["Warp Stall Sampling (All Samples)"]
ie = ci["Instructions Executed"]
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(float(r[si] or 0) for r in body)
tot_inst = sum(float(r[ie] or 0) for r in body)
print("kernel:", rows[0][1] if rows[0] else "?")
print("total samples %.0f, warp instructions executed %.0f, SASS lines %d" % (tot, tot_inst, len(body)))
agg = {}
for h, i in stall_cols:
    agg[h] = sum(float(r[i] or 0) for r in body)
print("stall totals:", ", ".join("%s=%.0f" % (k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0))
order = sorted(range(len(body)), key=lambda k: -float(body[k][si] or 0))
print("%5s %7s %6s %9s  %-14s %s" % ("line", "samples", "%", "executed", "top stall", "SASS"))
for k in order[:top]:
    r = body[k]
    v = float(r[si] or 0)
    if v == 0:
        break
    st = max(stall_cols, key=lambda hi: float(r[hi[1]] or 0))
    print("%5d %7.0f %5.1f%% %9s  %-14s %s" % (k, v, 100 * v / tot, r[ie], st[0][6:], r[ci["Source"]].strip()[:100]))
