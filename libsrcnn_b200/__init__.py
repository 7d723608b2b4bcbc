"""libsrcnn_b200 -- Python (ctypes) binding of the B200-native SRCNN library.

The product is the shared library ``libsrcnn_b200/lib/libsrcnn.so`` (C++ drop-in API ``ProcessSRCNN`` /
``ConfigureFilterSRCNN`` + the ``extern "C"`` layer of ``include/srcnn_b200.h``).  This module is only the
thin binding the tests and the benchmark use; it contains no image arithmetic and no fallback: if the library
is not built, or there is no B200, every entry point raises.
"""
import ctypes as C
import os
import pathlib

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
# SRCNN_B200_LIB: development aid (A/B timing of two builds of the same library); never a different implementation
LIB_PATH = pathlib.Path(os.environ.get("SRCNN_B200_LIB", ROOT / "lib" / "libsrcnn.so"))

NEAREST, BILINEAR, BICUBIC, LANCZOS3, BSPLINE = range(5)
CONV_TENSOR, CONV_EXACT = 0, 1
OK, EINVAL, ESIZE, ENODEV, ECUDA, ENOTHING = 0, -1, -2, -20, -21, -100

_u8p, _f32p, _f64p, _i32p = (C.POINTER(C.c_ubyte), C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int))
_lib = None


class SrcnnError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__("libsrcnn-b200 error %d %s" % (code, msg))
        self.code = code


def build():
    """compile the library in-tree (nvcc, sm_100a)"""
    import subprocess
    subprocess.run(["make", "-s", "-C", str(ROOT / "csrc")], check=True)


def lib():
    """the loaded shared library; raises if it has not been built (no fallback of any kind)"""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SrcnnError(ENODEV, "%s missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(str(LIB_PATH))
        vp, u, f, i = C.c_void_p, C.c_uint, C.c_float, C.c_int
        L.srcnn_b200_create.argtypes = [i, C.POINTER(vp)]
        L.srcnn_b200_destroy.argtypes = [vp]; L.srcnn_b200_destroy.restype = None
        L.srcnn_b200_last_error.argtypes = [vp]; L.srcnn_b200_last_error.restype = C.c_char_p
        L.srcnn_b200_set_conv_mode.argtypes = [vp, i]
        L.srcnn_b200_get_conv_mode.argtypes = [vp]
        L.srcnn_b200_output_dims.argtypes = [u, u, f, C.POINTER(u), C.POINTER(u)]; L.srcnn_b200_output_dims.restype = None
        L.srcnn_b200_process_dims.argtypes = [u, u, f, i, C.POINTER(u), C.POINTER(u)]
        L.srcnn_b200_process_has_conv.argtypes = [u, f, i]
        L.srcnn_b200_strip_geometry.argtypes = [u, u, u, f, i, u, u] + [C.POINTER(u)] * 4
        L.srcnn_b200_pass_host.argtypes = [vp, vp, u, u, u, f, i, vp, vp]
        L.srcnn_b200_pass_host_rows.argtypes = [vp, vp, u, u, u, f, i, u, u, vp, vp]
        L.srcnn_b200_pass_device.argtypes = [vp, vp, u, u, u, f, i, u, u, vp, vp]
        L.srcnn_b200_pass_host_async.argtypes = [vp, vp, u, u, u, f, i, vp, vp]
        L.srcnn_b200_wait.argtypes = [vp]
        L.srcnn_b200_set_copy_threads.argtypes = [vp, u]
        L.srcnn_b200_step_factors.argtypes = [u, f, _f32p, i]
        L.srcnn_b200_multi_create.argtypes = [_i32p, i, C.POINTER(vp)]
        L.srcnn_b200_multi_destroy.argtypes = [vp]; L.srcnn_b200_multi_destroy.restype = None
        L.srcnn_b200_multi_device_count.argtypes = [vp]
        L.srcnn_b200_multi_last_error.argtypes = [vp]; L.srcnn_b200_multi_last_error.restype = C.c_char_p
        L.srcnn_b200_multi_set_conv_mode.argtypes = [vp, i]
        L.srcnn_b200_multi_pass_host.argtypes = [vp, vp, u, u, u, f, i, vp, vp]
        L.srcnn_b200_multi_process_host.argtypes = [vp, vp, u, u, u, f, i, i, vp, vp]
        L.srcnn_b200_multi_batch_host.argtypes = [vp, vp, u, u, u, u, f, i, vp, vp]
        L.srcnn_b200_set_graphs.argtypes = [vp, i]
        L.srcnn_b200_graph_replays.argtypes = [vp]; L.srcnn_b200_graph_replays.restype = C.c_ulonglong
        L.srcnn_b200_sync.argtypes = [vp]
        L.srcnn_b200_stream.argtypes = [vp]; L.srcnn_b200_stream.restype = vp
        L.srcnn_b200_process_host.argtypes = [vp, vp, u, u, u, f, i, i, vp, vp]
        L.srcnn_b200_batch_host.argtypes = [vp, vp, u, u, u, u, f, i, vp, vp]
        L.srcnn_b200_batch_device.argtypes = [vp, vp, u, u, u, u, f, i, vp, vp]
        L.srcnn_b200_host_alloc.argtypes = [C.c_size_t]; L.srcnn_b200_host_alloc.restype = vp
        L.srcnn_b200_host_free.argtypes = [vp]; L.srcnn_b200_host_free.restype = None
        L.srcnn_b200_get_plane.argtypes = [vp, i, vp, C.c_size_t, C.POINTER(u), C.POINTER(u), C.POINTER(u)]
        L.srcnn_b200_set_timing.argtypes = [vp, i]
        L.srcnn_b200_set_keep_planes.argtypes = [vp, i]
        L.srcnn_b200_last_timings.argtypes = [vp, _f32p]
        L.srcnn_b200_launch_count.argtypes = [vp]; L.srcnn_b200_launch_count.restype = C.c_ulonglong
        L.srcnn_b200_last_group.argtypes = [vp]; L.srcnn_b200_last_group.restype = C.c_uint
        L.srcnn_b200_stage_color_in.argtypes = [vp, vp, u, u, u, C.POINTER(vp)]
        L.srcnn_b200_stage_resize.argtypes = [vp, vp, u, u, u, u, i, vp]
        L.srcnn_b200_stage_conv.argtypes = [vp, vp, u, u, i, vp]
        L.srcnn_b200_stage_color_out.argtypes = [vp, C.POINTER(vp), u, u, u, vp, vp]
        L.srcnn_b200_contrib_table.argtypes = [i, u, u, _i32p, _i32p, _f64p, i]
        L.srcnn_b200_contrib_table_compact.argtypes = [i, u, u, _i32p, _f64p, i, _i32p]
        # the C++ drop-in API, by its (reference-identical) mangled names
        L.ProcessSRCNN = getattr(L, "_Z12ProcessSRCNNPKhjjjfRPhRjPS1_Pj")
        L.ProcessSRCNN.restype = i
        L.ProcessSRCNN.argtypes = [vp, u, u, u, f, C.POINTER(vp), C.POINTER(u), C.POINTER(vp), C.POINTER(u)]
        L.ConfigureFilterSRCNN = getattr(L, "_Z20ConfigureFilterSRCNN15SRCNNFilterTypeb")
        L.ConfigureFilterSRCNN.restype = None
        L.ConfigureFilterSRCNN.argtypes = [i, C.c_bool]
        _lib = L
    return _lib


def output_dims(w, h, multiply):
    ow, oh = C.c_uint(), C.c_uint()
    lib().srcnn_b200_output_dims(w, h, multiply, C.byref(ow), C.byref(oh))
    return ow.value, oh.value


def process_dims(w, h, multiply, stepscale):
    ow, oh = C.c_uint(), C.c_uint()
    n = lib().srcnn_b200_process_dims(w, h, multiply, int(bool(stepscale)), C.byref(ow), C.byref(oh))
    return n, ow.value, oh.value


def strip_rows(total_rows, parts, index):
    """rows [a, b) of strip `index` when an image of `total_rows` output rows is cut into `parts` row strips
    (the sharding rule of the multi-GPU drivers: contiguous, balanced, disjoint, covering)"""
    return (index * total_rows) // parts, ((index + 1) * total_rows) // parts


def strip_geometry(w, h, d, multiply, filt, row0, row1):
    """-> (halo_rows (a, b) of the interpolated planes, source rows (a, b) read) for the strip [row0, row1)"""
    v = [C.c_uint() for _ in range(4)]
    rc = lib().srcnn_b200_strip_geometry(w, h, d, multiply, int(filt), row0, row1, *[C.byref(x) for x in v])
    if rc != 0:
        raise SrcnnError(rc, "strip_geometry")
    return (v[0].value, v[1].value), (v[2].value, v[3].value)


def contrib_table(filt, dst, src, maxwin=64):
    """host-side contribution table (left, right, weights[dst, maxwin]); needs no GPU"""
    L = np.zeros(dst, np.int32); R = np.zeros(dst, np.int32); W = np.zeros((dst, maxwin), np.float64)
    win = lib().srcnn_b200_contrib_table(int(filt), dst, src, L.ctypes.data_as(_i32p), R.ctypes.data_as(_i32p),
                                         W.ctypes.data_as(_f64p), maxwin)
    if win < 0:
        raise SrcnnError(win, "contrib_table")
    return L, R, W


def contrib_table_compact(filt, dst, src, maxwin=16, flags=False):
    """compact form read by the fused front-end kernel -> (cwin, cleft, weights[dst, maxwin], exact29); cwin 0 = none.
    flags=True returns the raw bit set instead of exact29 (bit 0 exact29, bit 1 pow2pair)"""
    L = np.zeros(dst, np.int32); W = np.zeros((dst, maxwin), np.float64); ex = C.c_int(0)
    win = lib().srcnn_b200_contrib_table_compact(int(filt), dst, src, L.ctypes.data_as(_i32p), W.ctypes.data_as(_f64p),
                                                 maxwin, C.byref(ex))
    if win < 0:
        raise SrcnnError(win, "contrib_table_compact")
    return win, L, W, (ex.value if flags else bool(ex.value & 1))


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


_dele = None


def process_srcnn(img, multiply, filt=BICUBIC, stepscale=False, want_conv=True, copy=True):
    """The drop-in C++ API exactly as an application calls it: ConfigureFilterSRCNN + ProcessSRCNN.
    -> (rc, out | None, conv | None); the new[] buffers are copied and released via libstdc++'s operator delete[].
    copy=False (timing): the buffers are released without being copied into numpy arrays -> (rc, size, conv size)."""
    global _dele
    L = lib()
    h, w, d = img.shape
    img = np.ascontiguousarray(img, np.uint8)
    L.ConfigureFilterSRCNN(int(filt), bool(stepscale))
    out, osz, cv, csz = C.c_void_p(), C.c_uint(0), C.c_void_p(), C.c_uint(0)
    rc = L.ProcessSRCNN(_vp(img), w, h, d, C.c_float(multiply), C.byref(out), C.byref(osz),
                        C.byref(cv) if want_conv else None, C.byref(csz) if want_conv else None)
    o = c = None
    if _dele is None:
        _dele = C.CDLL("libstdc++.so.6")._ZdaPv     # operator delete[](void*)
        _dele.argtypes = [C.c_void_p]; _dele.restype = None
    dele = _dele
    if not copy:
        if out.value:
            dele(out)
        if cv.value:
            dele(cv)
        return rc, osz.value, csz.value
    if out.value:
        o = np.ctypeslib.as_array(C.cast(out, _u8p), (osz.value,)).copy(); dele(out)
    if cv.value:
        c = np.ctypeslib.as_array(C.cast(cv, _u8p), (csz.value,)).copy(); dele(cv)
    return rc, o, c


class Context:
    """one GPU + one stream (include/srcnn_b200.h: srcnn_b200_ctx)"""

    def __init__(self, device=0, conv_mode=CONV_TENSOR):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.srcnn_b200_create(device, C.byref(h))
        if rc != 0:
            raise SrcnnError(rc, "srcnn_b200_create(device=%d): no usable B200 (there is no CPU fallback)" % device)
        self.h = h
        self.set_conv_mode(conv_mode)

    def close(self):
        if getattr(self, "h", None):
            self.L.srcnn_b200_destroy(self.h); self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise SrcnnError(rc, "%s: %s" % (what, (self.L.srcnn_b200_last_error(self.h) or b"").decode()))

    def set_conv_mode(self, mode):
        self._ck(self.L.srcnn_b200_set_conv_mode(self.h, mode), "set_conv_mode")

    # ---- passes -----------------------------------------------------------------------------------------
    def pass_host(self, img, multiply, filt=BICUBIC, want_conv=True, rows=None, out=None, conv=None):
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        ow, oh = output_dims(w, h, multiply)
        r0, r1 = rows if rows is not None else (0, oh)
        if out is None:
            out = np.empty((max(r1 - r0, 0), ow, d), np.uint8)
        if conv is None and want_conv:
            conv = np.empty((max(r1 - r0, 0), ow), np.uint8)
        rc = self.L.srcnn_b200_pass_host_rows(self.h, _vp(img), w, h, d, multiply, int(filt), r0, r1, _vp(out),
                                              _vp(conv) if want_conv else None)
        self._ck(rc, "pass_host_rows")
        return out, (conv if want_conv else None)

    def pass_host_async(self, img, multiply, filt, out, conv=None):
        """non-blocking pass: img / out / conv should be pinned (pinned_empty) and must stay alive until wait()"""
        h, w, d = img.shape
        assert img.flags["C_CONTIGUOUS"] and img.dtype == np.uint8
        self._ck(self.L.srcnn_b200_pass_host_async(self.h, _vp(img), w, h, d, multiply, int(filt), _vp(out), _vp(conv)),
                 "pass_host_async")

    def wait(self):
        self._ck(self.L.srcnn_b200_wait(self.h), "wait")

    def pass_device(self, d_src, w, h, d, multiply, filt, d_out, d_conv=0, rows=None):
        """device pointers (ints); asynchronous on the context's stream"""
        ow, oh = output_dims(w, h, multiply)
        r0, r1 = rows if rows is not None else (0, oh)
        self._ck(self.L.srcnn_b200_pass_device(self.h, d_src, w, h, d, multiply, int(filt), r0, r1, d_out,
                                               d_conv or None), "pass_device")

    def set_graphs(self, on=True):
        self._ck(self.L.srcnn_b200_set_graphs(self.h, int(on)), "set_graphs")

    @property
    def graph_replays(self):
        return int(self.L.srcnn_b200_graph_replays(self.h))

    def sync(self):
        self._ck(self.L.srcnn_b200_sync(self.h), "sync")

    @property
    def stream(self):
        return self.L.srcnn_b200_stream(self.h)

    def process(self, img, multiply, filt=BICUBIC, stepscale=False, want_conv=True):
        """whole ProcessSRCNN semantics through the C layer -> (rc, out|None, conv|None)"""
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        n, ow, oh = process_dims(w, h, multiply, stepscale)
        if n == 0:
            return ENOTHING, None, None
        out = np.empty((oh, ow, d), np.uint8)
        want_conv = want_conv and bool(self.L.srcnn_b200_process_has_conv(w, multiply, int(bool(stepscale))))
        conv = np.empty((oh, ow), np.uint8) if want_conv else None
        rc = self.L.srcnn_b200_process_host(self.h, _vp(img), w, h, d, multiply, int(filt), int(bool(stepscale)),
                                            _vp(out), _vp(conv))
        if rc != 0:
            return rc, None, None
        return rc, out, conv

    def batch(self, frames, multiply, filt=BICUBIC, want_conv=False, out=None, conv=None):
        n, h, w, d = frames.shape
        frames = np.ascontiguousarray(frames, np.uint8)
        ow, oh = output_dims(w, h, multiply)
        if out is None:
            out = np.empty((n, oh, ow, d), np.uint8)
        if not want_conv:
            conv = None
        elif conv is None:
            conv = np.empty((n, oh, ow), np.uint8)
        self._ck(self.L.srcnn_b200_batch_host(self.h, _vp(frames), n, w, h, d, multiply, int(filt), _vp(out), _vp(conv)),
                 "batch_host")
        return out, conv

    def batch_device(self, d_frames, n, w, h, d, multiply, filt, d_out, d_conv=0):
        """device pointers (ints) of n contiguous frames / results; asynchronous on the context's stream"""
        self._ck(self.L.srcnn_b200_batch_device(self.h, d_frames, n, w, h, d, multiply, int(filt), d_out, d_conv or None),
                 "batch_device")

    # ---- introspection ----------------------------------------------------------------------------------
    def get_plane(self, which, shape_hint):
        oh, ow = shape_hint
        buf = np.empty(((oh + 12) * ow,), np.float32)
        r0, nr, wd = C.c_uint(), C.c_uint(), C.c_uint()
        self._ck(self.L.srcnn_b200_get_plane(self.h, which, _vp(buf), buf.size, C.byref(r0), C.byref(nr), C.byref(wd)),
                 "get_plane")
        return r0.value, buf[: nr.value * wd.value].reshape(nr.value, wd.value).copy()

    def set_keep_planes(self, on=True):
        self._ck(self.L.srcnn_b200_set_keep_planes(self.h, int(on)), "set_keep_planes")

    def set_timing(self, on=True):
        self._ck(self.L.srcnn_b200_set_timing(self.h, int(on)), "set_timing")

    def timings(self):
        ms = (C.c_float * 7)()
        self._ck(self.L.srcnn_b200_last_timings(self.h, ms), "last_timings")
        return dict(zip(["color_in", "resize1", "resize2", "pack", "conv", "color_out", "total"], list(ms)))

    @property
    def launches(self):
        return int(self.L.srcnn_b200_launch_count(self.h))

    @property
    def last_group(self):
        return int(self.L.srcnn_b200_last_group(self.h))

    # ---- single stages ----------------------------------------------------------------------------------
    def color_in(self, img):
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        pl = [np.empty((h, w), np.float32) for _ in range(d)]
        arr = (C.c_void_p * d)(*[p.ctypes.data for p in pl])
        self._ck(self.L.srcnn_b200_stage_color_in(self.h, _vp(img), w, h, d, arr), "stage_color_in")
        return pl

    def resize(self, plane, dw, dh, filt):
        sh, sw = plane.shape
        plane = np.ascontiguousarray(plane, np.float32)
        o = np.empty((dh, dw), np.float32)
        self._ck(self.L.srcnn_b200_stage_resize(self.h, _vp(plane), sw, sh, dw, dh, int(filt), _vp(o)), "stage_resize")
        return o

    def conv(self, y, mode=None):
        h, w = y.shape
        y = np.ascontiguousarray(y, np.float32)
        o = np.empty((h, w), np.float32)
        if mode is None:
            mode = self.L.srcnn_b200_get_conv_mode(self.h)
        self._ck(self.L.srcnn_b200_stage_conv(self.h, _vp(y), w, h, int(mode), _vp(o)), "stage_conv")
        return o

    def color_out(self, planes, want_conv=False):
        d = len(planes); h, w = planes[0].shape
        planes = [np.ascontiguousarray(p, np.float32) for p in planes]
        arr = (C.c_void_p * d)(*[p.ctypes.data for p in planes])
        o = np.empty((h, w, d), np.uint8)
        cv = np.empty((h, w), np.uint8) if want_conv else None
        self._ck(self.L.srcnn_b200_stage_color_out(self.h, arr, w, h, d, _vp(o), _vp(cv)), "stage_color_out")
        return (o, cv) if want_conv else o


class Multi:
    """one host call over several GPUs of the box, single process (include/srcnn_b200.h: srcnn_b200_multi)"""

    def __init__(self, devices=None, conv_mode=CONV_TENSOR):
        self.L = lib()
        h = C.c_void_p()
        if devices is None:
            rc = self.L.srcnn_b200_multi_create(None, 0, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.L.srcnn_b200_multi_create(arr, len(devices), C.byref(h))
        if rc != 0:
            raise SrcnnError(rc, "srcnn_b200_multi_create(%r): no usable B200 (there is no CPU fallback)" % (devices,))
        self.h = h
        self._ck(self.L.srcnn_b200_multi_set_conv_mode(self.h, conv_mode), "set_conv_mode")

    def close(self):
        if getattr(self, "h", None):
            self.L.srcnn_b200_multi_destroy(self.h); self.h = None

    __del__ = close

    def _ck(self, rc, what):
        if rc != 0:
            raise SrcnnError(rc, "%s: %s" % (what, (self.L.srcnn_b200_multi_last_error(self.h) or b"").decode()))

    @property
    def ndev(self):
        return int(self.L.srcnn_b200_multi_device_count(self.h))

    def set_conv_mode(self, mode):
        self._ck(self.L.srcnn_b200_multi_set_conv_mode(self.h, mode), "set_conv_mode")

    def pass_host(self, img, multiply, filt=BICUBIC, want_conv=True, out=None, conv=None):
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        ow, oh = output_dims(w, h, multiply)
        if out is None:
            out = np.empty((oh, ow, d), np.uint8)
        if conv is None and want_conv:
            conv = np.empty((oh, ow), np.uint8)
        self._ck(self.L.srcnn_b200_multi_pass_host(self.h, _vp(img), w, h, d, multiply, int(filt), _vp(out),
                                                   _vp(conv) if want_conv else None), "multi_pass_host")
        return out, (conv if want_conv else None)

    def process(self, img, multiply, filt=BICUBIC, stepscale=False, want_conv=True):
        h, w, d = img.shape
        img = np.ascontiguousarray(img, np.uint8)
        n, ow, oh = process_dims(w, h, multiply, stepscale)
        if n == 0:
            return ENOTHING, None, None
        out = np.empty((oh, ow, d), np.uint8)
        want_conv = want_conv and bool(self.L.srcnn_b200_process_has_conv(w, multiply, int(bool(stepscale))))
        conv = np.empty((oh, ow), np.uint8) if want_conv else None
        rc = self.L.srcnn_b200_multi_process_host(self.h, _vp(img), w, h, d, multiply, int(filt), int(bool(stepscale)),
                                                  _vp(out), _vp(conv))
        if rc != 0:
            return rc, None, None
        return rc, out, conv

    def batch(self, frames, multiply, filt=BICUBIC, want_conv=False, out=None, conv=None):
        n, h, w, d = frames.shape
        frames = np.ascontiguousarray(frames, np.uint8)
        ow, oh = output_dims(w, h, multiply)
        if out is None:
            out = np.empty((n, oh, ow, d), np.uint8)
        if not want_conv:
            conv = None
        elif conv is None:
            conv = np.empty((n, oh, ow), np.uint8)
        self._ck(self.L.srcnn_b200_multi_batch_host(self.h, _vp(frames), n, w, h, d, multiply, int(filt), _vp(out), _vp(conv)),
                 "multi_batch_host")
        return out, conv


def device_count():
    return int(lib().srcnn_b200_device_count())


def pinned_empty(shape, dtype=np.uint8):
    """numpy array in CUDA pinned host memory (srcnn_b200_host_alloc); keep the returned array alive"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().srcnn_b200_host_alloc(n)
    if not p:
        raise SrcnnError(ECUDA, "host_alloc(%d)" % n)
    buf = (C.c_ubyte * n).from_address(p)
    a = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return a
