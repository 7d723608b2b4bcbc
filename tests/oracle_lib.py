"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE):

  port()  -> oracle/_build/libsrcnn_oracle.so  plain-C restatement (oracle/srcnn_oracle.c)
  ref()   -> oracle/_ref/libsrcnn_ref.so       the unmodified reference compiled in place (None if absent)

Both expose the same stage-level surface through the small wrapper class below.
"""
import ctypes as C
import pathlib, subprocess
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
u8p, f32p, f64p, i32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int))


def _ptr(a, t):
    return a.ctypes.data_as(t)


def out_dims(w, h, m):
    """rs_w = (unsigned)((float)w * multiply), reference src/libsrcnn.cpp:662-663"""
    return int(np.float32(w) * np.float32(m)), int(np.float32(h) * np.float32(m))


class Oracle:
    def __init__(self, lib, prefix, kind):
        self.lib, self.p, self.kind = lib, prefix, kind
        g = lambda n: getattr(lib, prefix + n)
        self._color_in = g("color_in"); self._color_in.restype = None
        self._resize = g("resize"); self._resize.restype = C.c_int
        self._table = g("contrib_table"); self._table.restype = C.c_int
        self._c1, self._c2, self._c3 = g("conv1"), g("conv2"), g("conv3")
        self._color_out = g("color_out"); self._color_out.restype = None
        self._process = g("process"); self._process.restype = C.c_int
        self._free = g("free"); self._free.restype = None; self._free.argtypes = [C.c_void_p]
        self.threads = (lib.ref_omp_threads if kind == "reference" else lib.orc_threads)()

    def color_in(self, img):
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        pl = [np.empty((h, w), np.float32) for _ in range(d)]
        self._color_in(_ptr(img, u8p), w, h, d, _ptr(pl[0], f32p), _ptr(pl[1], f32p), _ptr(pl[2], f32p),
                       _ptr(pl[3], f32p) if d == 4 else None)
        return pl

    def resize(self, plane, dw, dh, filt):
        sh, sw = plane.shape
        plane = np.ascontiguousarray(plane, np.float32)
        o = np.empty((dh, dw), np.float32)
        self._resize(_ptr(plane, f32p), sw, sh, dw, dh, int(filt), _ptr(o, f32p))
        return o

    def table(self, filt, dst, src, maxwin=64):
        L = np.zeros(dst, np.int32); R = np.zeros(dst, np.int32); W = np.zeros((dst, maxwin), np.float64)
        self._table(int(filt), dst, src, _ptr(L, i32p), _ptr(R, i32p), _ptr(W, f64p), maxwin)
        return L, R, W

    def conv1(self, y):
        h, w = y.shape; y = np.ascontiguousarray(y, np.float32)
        o = np.empty((64, h, w), np.float32); self._c1(_ptr(y, f32p), w, h, _ptr(o, f32p)); return o

    def conv2(self, l1):
        _, h, w = l1.shape; l1 = np.ascontiguousarray(l1, np.float32)
        o = np.empty((32, h, w), np.float32); self._c2(_ptr(l1, f32p), w, h, _ptr(o, f32p)); return o

    def conv3(self, l2):
        _, h, w = l2.shape; l2 = np.ascontiguousarray(l2, np.float32)
        o = np.empty((h, w), np.float32); self._c3(_ptr(l2, f32p), w, h, _ptr(o, f32p)); return o

    def convs(self, y):
        return self.conv3(self.conv2(self.conv1(y)))

    def color_out(self, planes):
        d = len(planes); h, w = planes[0].shape
        planes = [np.ascontiguousarray(p, np.float32) for p in planes]
        arr = (f32p * d)(*[_ptr(p, f32p) for p in planes])
        o = np.empty((h, w, d), np.uint8)
        self._color_out(arr, w, h, d, _ptr(o, u8p))
        return o

    def process(self, img, multiply, filt=2, stepscale=False, want_conv=True):
        """-> (rc, out[h,w,d] | None, conv[h,w] | None)"""
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        out = C.c_void_p(); osz = C.c_uint(0); cv = C.c_void_p(); csz = C.c_uint(0)
        rc = self._process(_ptr(img, u8p), w, h, d, C.c_float(multiply), int(filt), int(bool(stepscale)),
                           C.byref(out), C.byref(osz),
                           C.byref(cv) if want_conv else None, C.byref(csz) if want_conv else None)
        o = c = None
        if out.value:
            o = np.ctypeslib.as_array(C.cast(out, u8p), (osz.value,)).copy(); self._free(out)
        if cv.value:
            c = np.ctypeslib.as_array(C.cast(cv, u8p), (csz.value,)).copy(); self._free(cv)
        return rc, o, c


def build():
    subprocess.run(["make", "-s", "-C", str(ORACLE_DIR), "all"], check=True)


_cache = {}


def port():
    if "port" not in _cache:
        p = ORACLE_DIR / "_build" / "libsrcnn_oracle.so"
        if not p.exists():
            build()
        _cache["port"] = Oracle(C.CDLL(str(p)), "orc_", "port")
    return _cache["port"]


def ref():
    if "ref" not in _cache:
        p = ORACLE_DIR / "_ref" / "libsrcnn_ref.so"
        if not p.exists():
            try:
                build()
            except Exception:
                pass
        _cache["ref"] = Oracle(C.CDLL(str(p)), "ref_", "reference") if p.exists() else None
    return _cache["ref"]


def best():
    """the strongest checker available: compiled reference if present, else the port"""
    return ref() or port()


# ---- synthetic inputs ----
def synth(kind, h, w, d, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, size=(h, w, d), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, size=(h, w, d), dtype=np.uint8) * 255).astype(np.uint8)
    if kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        v = ((((yy // 2) + (xx // 2)) & 1) * 255).astype(np.uint8)
        return np.repeat(v[:, :, None], d, axis=2).copy()
    if kind == "smooth":
        yy, xx = np.mgrid[0:h, 0:w]; x = xx / max(w, 1); y = yy / max(h, 1)
        chans = []
        for c in range(d):
            phi = 0.17 * c
            f = 127.5 + 100 * np.sin(2 * np.pi * (3 * x + 2 * y + phi)) * np.cos(2 * np.pi * (5 * y - x))
            chans.append(np.clip(f + rng.normal(0, 4, size=(h, w)), 0, 255))
        return np.stack(chans, axis=2).astype(np.uint8)
    if kind == "const0":
        return np.zeros((h, w, d), np.uint8)
    if kind == "const255":
        return np.full((h, w, d), 255, np.uint8)
    raise ValueError(kind)


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
