#!/usr/bin/env python
"""Per-source-line memory pressure from an `ncu --page source --csv --print-source cuda,sass` dump:
instructions, stall samples, L1 tag requests (global), shared wavefronts, L2 sectors."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
cur = None; hdr = None; agg = {}
cols = ["Instructions Executed", "# Samples", "L1 Tag Requests Global", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Global"]
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = [hdr.index(c) for c in cols]; continue
    if hdr is None or len(r) <= max(ix) or r[2] != "-": continue
    try: vals = [int(float(r[i] or 0)) for i in ix]
    except ValueError: continue
    agg[(cur, int(r[0]))] = (vals, r[1].strip()[:90])
tot = [sum(v[0][k] for v in agg.values()) or 1 for k in range(len(cols))]
print("totals", dict(zip(cols, tot)))
for (f, l), (v, t) in sorted(agg.items()):
    if any(v[k] > tot[k] * thr for k in range(len(cols))):
        print(f"{f}:{l:4d} inst {100*v[0]/tot[0]:4.1f}% smp {100*v[1]/tot[1]:4.1f}% tagG {100*v[2]/tot[2]:4.1f}% shW {100*v[3]/tot[3]:4.1f}% (ideal {100*v[4]/tot[3]:4.1f}%) L2sec {100*v[5]/tot[5]:4.1f}%  {t}")
