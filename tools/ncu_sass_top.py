"""Top SASS instructions (by executed count and by stall samples) per kernel from `ncu --page source --csv --print-source sass`.
usage: python tools/ncu_sass_top.py sass.csv [kernel-index] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"fn": r[1], "rows": []}; secs.append(cur); continue
    if r and r[0] == "Address": hdr = r; continue
    if cur is not None: cur["rows"].append(r)
ix = {h: i for i, h in enumerate(hdr)}
s = secs[which]
ins = []
for r in s["rows"]:
    if len(r) < len(hdr) or not r[0].startswith("0x"): continue
    try: ins.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[1].strip()))
    except ValueError: pass
tot = sum(i[0] for i in ins); tots = sum(i[1] for i in ins)
print(s["fn"][:70], "| SASS", len(ins), "| executed", tot, "| samples", tots)
ops = {}
for n, sm, src in ins:
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    op = op.split(".")[0]
    a = ops.setdefault(op, [0, 0]); a[0] += n; a[1] += sm
print("-- by opcode")
for op, (n, sm) in sorted(ops.items(), key=lambda x: -x[1][0])[:top]:
    print("%6.2f%% inst %6.2f%% samp  %s" % (100.0 * n / tot, 100.0 * sm / max(tots, 1), op))
