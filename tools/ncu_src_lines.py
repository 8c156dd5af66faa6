"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line: share of executed instructions and of stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_src_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; agg = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if cur is None or hdr is None or len(r) <= ii or r[2] != "-": continue
    try: agg.append((int(r[ii]), int(r[si]), cur, r[0], r[1].strip()[:120]))
    except ValueError: pass
tot = sum(a[0] for a in agg); tots = sum(a[1] for a in agg)
print("total warp instructions %d, samples %d" % (tot, tots))
for n, s, f, l, src in sorted(agg, key=lambda x: -x[0])[:top]:
    print("%5.1f%% inst %5.1f%% samp  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / max(tots, 1), f, l, src))
