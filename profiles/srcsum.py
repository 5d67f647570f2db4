#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: instruction and stall-sample share per source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
files = {}
cur = None
hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = hdr.index("Instructions Executed"); sx = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ix: continue
    if r[2] != "-": continue  # SASS rows carry an address; source rows have '-'
    try: n = int(r[ix] or 0); s = int(r[sx] or 0)
    except ValueError: continue
    agg[(cur, int(r[0]))] = (n, s, r[1].strip()[:100])
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp-instr {tot}  samples {ts}")
for (f, l), (n, s, t) in sorted(agg.items()):
    if n > tot * thr or s > ts * thr:
        print(f"{f}:{l:4d} inst {100*n/tot:5.1f}%  stall-smp {100*s/ts:5.1f}%  {t}")
