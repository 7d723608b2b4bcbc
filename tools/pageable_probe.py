"""Development aid: the host-buffer matrix of the end-to-end call (pinned / pageable source x pinned / pageable destination)
for the staging switches of the host pipeline (SRCNN_B200_STAGE_IN, SRCNN_B200_STAGE_ALL, SRCNN_B200_STRIP_MB), the drop-in
ProcessSRCNN call, and what a fresh 25 MB new[] costs on this box.  GPU box only:  python tools/pageable_probe.py"""
import ctypes as C, os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import libsrcnn_b200 as S, oracle_lib as O
h, w, d, m, f = 1080, 1920, 3, 2.0, 2
img = O.synth("uniform", h, w, d, 42); ow, oh = S.output_dims(w, h, m)


def t(fn, n=20):
    fn(); fn(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


libc = C.CDLL("libc.so.6"); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
libc.memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]; libc.memset.restype = C.c_void_p


def fresh():
    p = libc.malloc(ow * oh * d); libc.memset(p, 1, ow * oh * d); libc.free(p)


print("malloc + first touch + free of %d MB: %.3f ms (single thread)" % (ow * oh * d >> 20, t(fresh)), flush=True)
for stage_in, stage_all, mb in (("1", "1", ""), ("0", "1", ""), ("1", "0", ""), ("0", "0", ""), ("1", "1", "2"), ("1", "1", "8"), ("0", "1", "2")):
    os.environ["SRCNN_B200_STAGE_IN"] = stage_in; os.environ["SRCNN_B200_STAGE_ALL"] = stage_all
    if mb:
        os.environ["SRCNN_B200_STRIP_MB"] = mb
    else:
        os.environ.pop("SRCNN_B200_STRIP_MB", None)
    ctx = S.Context(0)
    hs = S.pinned_empty((h, w, d)); hs[...] = img; ho = S.pinned_empty((oh, ow, d)); po = np.empty((oh, ow, d), np.uint8)
    a = t(lambda: ctx.pass_host(hs, m, f, False, out=ho)); b = t(lambda: ctx.pass_host(img, m, f, False, out=ho))
    c = t(lambda: ctx.pass_host(hs, m, f, False, out=po)); e = t(lambda: ctx.pass_host(img, m, f, False, out=po))
    print("stage_in=%s stage_all=%s strip_mb=%-2s  pin->pin %.3f  page->pin %.3f  pin->page %.3f  page->page %.3f ms  ok=%s"
          % (stage_in, stage_all, mb or "-", a, b, c, e, np.array_equal(ho, po)), flush=True)
    ctx.close()
for k in ("SRCNN_B200_STAGE_IN", "SRCNN_B200_STAGE_ALL", "SRCNN_B200_STRIP_MB"):
    os.environ.pop(k, None)
# a FRESH pageable result buffer per call (what ProcessSRCNN's new[] is), with and without the background first touch
for pf in ("1", "0"):
    os.environ["SRCNN_B200_PREFAULT"] = pf
    ctx = S.Context(0)
    def fresh_call():
        p = libc.malloc(ow * oh * d)
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), (oh, ow, d))
        ctx.pass_host(img, m, f, False, out=buf)
        libc.free(p)
    print("prefault=%s  page->FRESH malloc'd buffer %.3f ms" % (pf, t(fresh_call)), flush=True)
    ctx.close()
os.environ.pop("SRCNN_B200_PREFAULT", None)
print("drop-in ProcessSRCNN (defaults) %.3f ms" % t(lambda: S.process_srcnn(img, m, f, False, False, copy=False), 12), flush=True)
# would keeping freed blocks in the heap help?  (process-wide malloc policy: not something a library should set silently)
libc.mallopt.argtypes = [C.c_int, C.c_int]
libc.mallopt(-3, 1 << 30); libc.mallopt(-1, 1 << 30)        # M_MMAP_THRESHOLD, M_TRIM_THRESHOLD
print("after mallopt(M_MMAP_THRESHOLD / M_TRIM_THRESHOLD = 1 GB): malloc + first touch + free %.3f ms, ProcessSRCNN %.3f ms"
      % (t(fresh), t(lambda: S.process_srcnn(img, m, f, False, False, copy=False), 12)), flush=True)
os.environ["SRCNN_B200_TRACE"] = "1"
c2 = S.Context(0); po = np.empty((oh, ow, d), np.uint8)
for _ in range(4):
    c2.pass_host(img, m, f, False, out=po)
