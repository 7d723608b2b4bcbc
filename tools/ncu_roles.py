#!/usr/bin/env python3
"""Attribute `ncu --page source --csv` stall samples to code segments (roles of a warp-specialised kernel):
segments are split where the per-instruction execution count changes.  usage: ncu_roles.py src.csv [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]; si = hdr.index("Warp Stall Sampling (All Samples)"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
data = [(int(r[si] or 0), r[src].strip(), int(r[ie] or 0)) for r in rows[hi + 1:] if len(r) > si and r[si].strip().isdigit()]
segs, start = [], 0
for i in range(1, len(data)):
    a, b = data[i - 1][2], data[i][2]
    if a and b and (b > 1.5 * a or a > 1.5 * b):
        segs.append((start, i - 1)); start = i
segs.append((start, len(data) - 1))
print("total samples", sum(d[0] for d in data), "instructions", len(data))
KEYS = ['LDTM', 'STTM', 'UTCHMMA', 'UTMALDG', 'LDS', 'STS', 'TRYWAIT', 'F2FP', 'FADD', 'BAR', 'STG']
for s, e in segs:
    tot = sum(d[0] for d in data[s:e + 1])
    if tot < thr:
        continue
    kinds = {k: sum(1 for d in data[s:e + 1] if k in d[1]) for k in KEYS}
    print("%5d-%5d n=%4d exec=%8d samples=%6d  %s" % (s, e, e - s + 1, data[s][2], tot, {k: v for k, v in kinds.items() if v}))
    for i in sorted(range(s, e + 1), key=lambda i: -data[i][0])[:5]:
        print("        %5d %6d %s" % (i, data[i][0], data[i][1][:90]))
