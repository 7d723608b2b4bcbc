"""Development aid: print the per-row event timeline written by a -DSRCNN_TRACE build of conv_tc.cu (gpurun_out/trace.bin)."""
import sys, numpy as np
t = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.bin", dtype=np.int64).reshape(-1, 20)
a, b = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100, 124)
t0 = t[a, 0]
names = ["im.start", "im.fetchd", "im.go", "A1full", "m1.go", "m1.issd", "e1.go", "D1empt", "e1.A2ok", "A2full", "-", "m2.go", "e2.go", "D2empt", "A3full", "-", "m3.go", "g.go", "D3empt", "g.done"]
print("row " + " ".join("%8s" % n for n in names[:17]))
for r in range(a, b):
    print("%3d " % r + " ".join("%8d" % (t[r, e] - t0) if t[r, e] else "       ." for e in range(17)))
d = np.diff(t[a:b, 3]); print("A1full row-to-row period: mean %.0f" % d.mean(), d)
