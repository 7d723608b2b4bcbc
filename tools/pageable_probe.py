"""Development aid: the host-buffer matrix of the end-to-end call (pinned / pageable source x pinned / pageable destination),
with and without pinned staging of pageable results (SRCNN_B200_STAGE), and the drop-in ProcessSRCNN call.  GPU box only.
Measured (r1): pin->pin 0.59, page->pin 0.70, pin->page 0.78, page->page 0.89 ms; without staging 1.9 / 2.3 ms for pageable
results.  Staging a pageable SOURCE through the worker threads was slower (0.92 ms) than the driver's own path: not done."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import libsrcnn_b200 as S, oracle_lib as O
h, w, d, m, f = 1080, 1920, 3, 2.0, 2
img = O.synth("uniform", h, w, d, 42); ow, oh = S.output_dims(w, h, m)
def t(fn, n=20):
    fn(); fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
for stage in ("1", "0", "1", "0"):
    os.environ["SRCNN_B200_STAGE"] = stage
    ctx = S.Context(0)
    hs = S.pinned_empty((h, w, d)); hs[...] = img; ho = S.pinned_empty((oh, ow, d)); po = np.empty((oh, ow, d), np.uint8)
    a = t(lambda: ctx.pass_host(hs, m, f, False, out=ho)); b = t(lambda: ctx.pass_host(img, m, f, False, out=ho)); c = t(lambda: ctx.pass_host(hs, m, f, False, out=po)); e = t(lambda: ctx.pass_host(img, m, f, False, out=po))
    print("stage=%s pin->pin %.3f  page->pin %.3f  pin->page %.3f  page->page %.3f ms  ok=%s" % (stage, a, b, c, e, np.array_equal(ho, po)), flush=True)
    ctx.close()
os.environ["SRCNN_B200_STAGE"] = "1"
print("drop-in ProcessSRCNN %.3f ms" % t(lambda: S.process_srcnn(img, m, f, False, False, copy=False), 10))
