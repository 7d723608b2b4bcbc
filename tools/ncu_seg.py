#!/usr/bin/env python3
"""Print per-instruction stall samples and dominant stall reasons for a SASS index range of `ncu --page source --csv`.
usage: ncu_seg.py src.csv first last [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
a, b = int(sys.argv[2]), int(sys.argv[3]); thr = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]
st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = rows[hi + 1:]
agg = {}
tot = 0
for n in range(a, b + 1):
    r = data[n]
    s = int(r[2] or 0); tot += s
    reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in st), reverse=True)
    for v, h in reasons: agg[h] = agg.get(h, 0) + v
    if s >= thr:
        print("%5d %5d ex=%7s  %-70s %s" % (n, s, r[5], r[1].strip()[:70], " ".join("%s=%d" % (h, v) for v, h in reasons[:3] if v)))
print("total", tot, sorted(agg.items(), key=lambda x: -x[1])[:8])
