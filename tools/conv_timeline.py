"""Development aid: reads the event dump of a -DSRCNN_TL build of conv_tc.cu (SRCNN_TL_DUMP=file) and prints, per image row,
when each pipeline event happened (SM clocks, relative), the barrier completion times implied by the LAST arriving warp, and a
model of the in-order tensor pipe (isolated MMA rates) against the observed completions.
usage: python tools/conv_timeline.py dump.txt [first_row nrows]"""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.int64)
rows = a[:, 0]
ev = a[:, 1:].reshape(len(rows), 32, 4)
t0 = ev[ev > 0].min()
ev = np.where(ev > 0, ev - t0, -1)
r0 = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = int(sys.argv[3]) if len(sys.argv) > 3 else 16
mx = lambda x: x.max(axis=-1)
mn = lambda x: np.where(x >= 0, x, 1 << 60).min(axis=-1)
E = lambda e: ev[:, e, :]
I = lambda e: ev[:, e, 0]
cols = [("im.fet", mn(E(0))), ("im.M1ok", mn(E(1))), ("im.arr-", mn(E(2))), ("im.arr+", mx(E(2))),
        ("E1.dr+", None), ("M1.rdy", I(22)), ("M1.go", I(3)), ("M1.iss", I(4)),
        ("E1.wk-", mn(E(5))), ("E1.wk+", mx(E(5))), ("E1.ld+", mx(E(6))), ("E1.M2ok", mx(E(7))), ("E1.arr+", mx(E(8))),
        ("M2.rdy", I(23)), ("M2.go", I(9)), ("M2.iss", I(10)),
        ("E2.wk-", mn(E(11))), ("E2.wk+", mx(E(11))), ("E2.ld+", mx(E(12))), ("E2.st+", mx(E(13))), ("E2.arr+", mx(E(21))),
        ("M3.rdy", I(24)), ("M3.a3ok", I(14)), ("M3.go", I(15)), ("M3.iss", I(16)),
        ("G.wk-", mn(E(17))), ("G.wk+", mx(E(17))), ("G.ld+", mx(E(18))), ("G.end+", mx(E(20)))]
# D1 drained for row r = E1 ld done of row r-2
e1ld = mx(E(6)); dr = np.full(len(rows), -1, dtype=np.int64); dr[2:] = e1ld[:-2]
cols[4] = ("D1.free", dr)
print("row  " + " ".join("%8s" % c[0] for c in cols))
for i in range(r0, r0 + n):
    print("%4d " % rows[i] + " ".join("%8d" % (c[1][i] if c[1][i] < (1 << 59) else -1) for c in cols))
print("row period (M1.iss):", float(np.diff(I(4)[8:120]).mean()))

# in-order tensor pipe model with the isolated rates (tools/probe/mma_rate.cu): batches sorted by issue-start time
B = []
for i in range(len(rows)):
    if I(3)[i] > 0: B.append((I(3)[i], I(4)[i], 192.0, "M1", rows[i], mn(E(5))[i]))
    if I(9)[i] > 0: B.append((I(9)[i], I(10)[i], 80.0, "M2", rows[i], mn(E(11))[i]))
    if I(15)[i] > 0: B.append((I(15)[i], I(16)[i], 110.0, "M3", rows[i], mn(E(17))[i]))
B.sort()
t = 0.0; idle = 0.0; work = 0.0; first = None
print("\nbatch  row   issue-start issue-end   model-start model-done  observed-wake  (wake - model-done)   pipe idle before")
out = []
for (s, e, w, nm, r, wake) in B:
    st = max(t, float(s)); gap = st - t if t > 0 else 0.0
    t = st + w
    out.append((nm, r, s, e, st, t, wake, wake - t, gap))
    if first is None: first = st
    idle += gap; work += w
for o in out[3 * r0 // 2: 3 * r0 // 2 + 3 * n]:
    print("%-5s %4d %10d %10d %12.0f %10.0f %12d %12.0f %18.0f" % o)
print("model: work %.0f  idle %.0f  -> pipe busy %.1f %% if MMAs ran at their isolated rates" % (work, idle, 100 * work / (work + idle)))
