#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions and stall-reason totals.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
hdr = rows[hi]
si, src, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ridx = {h: hdr.index(h) for h in reasons}
data = []
tot_r = {h: 0 for h in reasons}
for r in rows[hi + 1:]:
    if len(r) <= max(si, ie) or not r[si].strip().isdigit():
        continue
    d = (int(r[si] or 0), r[src].strip(), int(r[ie] or 0), {h: int(r[ridx[h]] or 0) for h in reasons if len(r) > ridx[h] and r[ridx[h]].isdigit()})
    data.append(d)
    for h, v in d[3].items():
        tot_r[h] += v
tot = sum(d[0] for d in data) or 1
print("total samples", tot, " instructions", len(data), " executed warp-instr", sum(d[2] for d in data))
print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100 * v / tot) for h, v in sorted(tot_r.items(), key=lambda x: -x[1]) if v))
for i, (s, t, n, rs) in sorted(enumerate(data), key=lambda x: -x[1][0])[:N]:
    top = max(rs.items(), key=lambda x: x[1])[0][6:] if rs else ""
    print("%5d %7d %5.1f%% exec=%8d %-14s %s" % (i, s, 100 * s / tot, n, top, t[:100]))
