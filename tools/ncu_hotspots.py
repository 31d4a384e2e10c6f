#!/usr/bin/env python
"""Segment the SASS of one kernel from `ncu -i rep --page source --csv` by execution count and stall samples.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hotspots.py src.csv [min_share]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    try: data.append((int(r[ie]), r[isrc], int(r[ismp] or 0)))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data); smp = sum(d[2] for d in data)
print(f"{len(data)} SASS instructions, {tot:.3e} warp-instructions executed, {smp} stall samples")
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(data[i][0] - data[start][0]) > 0.2 * max(data[start][0], 1):
        s = sum(d[0] for d in data[start:i]); q = sum(d[2] for d in data[start:i])
        if s > thr * tot or q > thr * smp:
            print(f"instr {start:5d}-{i-1:5d}  n={i-start:4d}  exec/instr={data[start][0]:11d}  share={s/tot:6.1%}  samples={q/max(smp,1):6.1%}  {data[start][1][:50]}")
        start = i
