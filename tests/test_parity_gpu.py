"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI / the C++ drop-in API, against the oracle.

Two arithmetic modes of the fused convolution are checked:
  exact  (SRCNN_B200_CONV_EXACT)  : FP32 in the reference's summation order  -> every stage and the final image BIT-EXACT
  tensor (SRCNN_B200_CONV_TENSOR) : tcgen05 FP16 operands / FP32 accumulate  -> north-star tolerance:
         final 8-bit output within 1 LSB on every channel and PSNR >= 50 dB, conv plane within 1 LSB,
         network output |dY'| < 1.0; the interpolated Y plane is bit-exact in both modes (tolerance 1e-5 relative).
"""
import numpy as np
import pytest
from PIL import Image

import libsrcnn_b200 as S
import oracle_lib as O

pytestmark = pytest.mark.gpu
G = O.ROOT / "tests" / "golden"
MODES = [("exact", S.CONV_EXACT), ("tensor", S.CONV_TENSOR)]
MODE_IDS = [m[0] for m in MODES]


@pytest.fixture(scope="module")
def ctx():
    c = S.Context(0)
    yield c
    c.close()


def _check_u8(name, got, want, exact):
    got = np.asarray(got).reshape(want.shape)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if exact:
        assert diff.max() == 0, "%s: max diff %d, %d px differ" % (name, diff.max(), int((diff > 0).sum()))
    else:
        assert diff.max() <= 1, "%s: max diff %d LSB" % (name, diff.max())
        assert O.psnr(got, want) >= 50.0, "%s: PSNR %.2f dB" % (name, O.psnr(got, want))


# ---------------------------------------------------------------------------------------------- single stages
@pytest.mark.parametrize("d", [3, 4])
@pytest.mark.parametrize("hw", [(37, 53), (64, 64), (1, 1), (5, 3)])
def test_stage_color_in_bit_exact(ctx, d, hw):
    img = O.synth("uniform", hw[0], hw[1], d, 11)
    for a, b in zip(ctx.color_in(img), O.best().color_in(img)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("shape", [((40, 56), (80, 112)), ((31, 45), (46, 67)), ((64, 64), (40, 30)),
                                   ((20, 30), (20, 60)), ((20, 30), (45, 30)), ((9, 7), (27, 29)), ((50, 40), (75, 20))])
def test_stage_resize_bit_exact(ctx, filt, shape):
    (sh, sw), (dh, dw) = shape
    p = np.random.default_rng(3).uniform(-20, 280, size=(sh, sw)).astype(np.float32)
    assert np.array_equal(ctx.resize(p, dw, dh, filt), O.best().resize(p, dw, dh, filt))


@pytest.mark.parametrize("hw", [(40, 56), (3, 7), (17, 130), (33, 16), (1, 1), (129, 125)])
def test_stage_conv_exact_bit_exact(ctx, hw):
    y = np.random.default_rng(4).uniform(-20, 280, size=hw).astype(np.float32)
    assert np.array_equal(ctx.conv(y, S.CONV_EXACT), O.best().convs(y))


@pytest.mark.parametrize("hw", [(40, 56), (3, 7), (17, 130), (64, 124), (65, 125), (1, 1), (130, 250)])
@pytest.mark.parametrize("kind", ["uniform", "binary"])
def test_stage_conv_tensor_within_tolerance(ctx, hw, kind):
    rng = np.random.default_rng(5)
    if kind == "uniform":
        y = rng.uniform(-20, 280, size=hw).astype(np.float32)
    else:                                   # overshoot range seen with Lanczos-3 on 0/255 noise
        y = rng.choice(np.array([-117.0, 0.0, 255.0, 372.0], np.float32), size=hw)
    want = O.best().convs(y)
    got = ctx.conv(y, S.CONV_TENSOR)
    assert np.abs(got - want).max() < 1.0


@pytest.mark.parametrize("d", [3, 4])
def test_stage_color_out_bit_exact(ctx, d):
    rng = np.random.default_rng(6)
    pl = [rng.uniform(-30, 290, size=(19, 27)).astype(np.float32) for _ in range(d)]
    pl[0] = np.clip(pl[0], 0, 255)          # the conv III plane is already clamped (reference :519-522)
    out, conv = ctx.color_out(pl, want_conv=True)
    assert np.array_equal(out, O.best().color_out(pl))
    assert np.array_equal(conv, pl[0].astype(np.uint8))


# ---------------------------------------------------------------------------------------------- whole path
@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_c1_butterfly_golden(ctx, mode):
    """BASELINE config 1 against the reference's own golden PNGs (reference Pictures/butterfly_srcnn*.png)"""
    ctx.set_conv_mode(mode[1])
    src = np.asarray(Image.open(G / "butterfly.png").convert("RGB"))
    gold = np.asarray(Image.open(G / "butterfly_srcnn.png").convert("RGB"))
    gconv = np.asarray(Image.open(G / "butterfly_srcnn_convolution.png").convert("L"))
    rc, out, conv = ctx.process(src, 2.0, S.BICUBIC, False, True)
    assert rc == 0
    _check_u8("rgb", out, gold, mode[0] == "exact")
    _check_u8("conv", conv, gconv, mode[0] == "exact")


CASES = [
    dict(id="c2-like", kind="uniform", h=135, w=240, d=3, m=2.0, f=S.BICUBIC, s=False),
    dict(id="c3-like-rgba", kind="uniform", h=90, w=160, d=4, m=2.0, f=S.BICUBIC, s=False),
    dict(id="c4-like-lanczos", kind="uniform", h=128, w=128, d=3, m=2.0, f=S.LANCZOS3, s=False),
    dict(id="c5-like-bspline-step", kind="uniform", h=72, w=128, d=3, m=1.5, f=S.BSPLINE, s=True),
    dict(id="binary-lanczos-stress", kind="binary", h=96, w=100, d=3, m=2.0, f=S.LANCZOS3, s=False),
    dict(id="checker-nearest", kind="checker", h=50, w=70, d=3, m=2.0, f=S.NEAREST, s=False),
    dict(id="smooth-bilinear-x2.5", kind="smooth", h=41, w=67, d=4, m=2.5, f=S.BILINEAR, s=False),
    dict(id="step-x3", kind="smooth", h=30, w=44, d=3, m=3.0, f=S.BICUBIC, s=True),
    dict(id="step-x4-rgba", kind="uniform", h=20, w=28, d=4, m=4.0, f=S.BILINEAR, s=True),
    dict(id="step-x6", kind="uniform", h=14, w=18, d=3, m=6.0, f=S.BICUBIC, s=True),
    # x8 = x2 x2 x2, then the remaining factor is 1: the loop breaks before the pass that would receive convbuff, so the
    # reference produces NO convolution plane (src/libsrcnn.cpp:1003-1028)
    dict(id="step-x8-no-conv-plane", kind="smooth", h=10, w=12, d=3, m=8.0, f=S.BICUBIC, s=True),
    dict(id="const0", kind="const0", h=24, w=24, d=3, m=2.0, f=S.BICUBIC, s=False),
    dict(id="const255", kind="const255", h=24, w=24, d=4, m=2.0, f=S.BICUBIC, s=False),
    dict(id="tiny-1x1", kind="uniform", h=1, w=1, d=3, m=2.0, f=S.BICUBIC, s=False),
    dict(id="thin-3x200", kind="uniform", h=3, w=200, d=3, m=2.0, f=S.BICUBIC, s=False),
    dict(id="downscale-x0.5", kind="smooth", h=64, w=96, d=3, m=0.5, f=S.BICUBIC, s=False),
    dict(id="bad-filter-id", kind="uniform", h=16, w=16, d=3, m=2.0, f=9, s=False),
    # magnifications in (1, 1.5): the separate colour-in / resize / pack kernels (the fused front end starts at 1.5)
    dict(id="x1.25-bicubic", kind="uniform", h=64, w=88, d=3, m=1.25, f=S.BICUBIC, s=False),
    dict(id="x1.4-lanczos-rgba", kind="smooth", h=50, w=70, d=4, m=1.4, f=S.LANCZOS3, s=False),
    dict(id="x1.1-bspline", kind="uniform", h=40, w=60, d=3, m=1.1, f=S.BSPLINE, s=False),
    dict(id="x1.5-bicubic-rgba", kind="uniform", h=66, w=90, d=4, m=1.5, f=S.BICUBIC, s=False),
    dict(id="x2-wide-lanczos", kind="uniform", h=40, w=700, d=3, m=2.0, f=S.LANCZOS3, s=False),
]


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_process_matches_oracle(ctx, case, mode):
    ctx.set_conv_mode(mode[1])
    img = O.synth(case["kind"], case["h"], case["w"], case["d"], 21)
    rc_o, out_o, conv_o = O.best().process(img, case["m"], case["f"], case["s"])
    rc, out, conv = ctx.process(img, case["m"], case["f"], case["s"], True)
    assert rc == rc_o
    _check_u8("rgb", out, out_o, mode[0] == "exact")
    assert (conv is None) == (conv_o is None)
    if conv_o is not None:
        _check_u8("conv", conv, conv_o, mode[0] == "exact")


@pytest.mark.parametrize("filt", [S.BICUBIC, S.LANCZOS3, S.BSPLINE])
def test_interpolated_plane_and_network_output(ctx, filt):
    """the FP32 planes the public API never exposes: interpolated Y within 1e-5 relative (expected 0),
    tensor-mode network output within 1.0 of the oracle's FP32 plane"""
    img = O.synth("binary" if filt == S.LANCZOS3 else "uniform", 60, 84, 3, 33)
    o = O.best()
    y = o.resize(o.color_in(img)[0], 168, 120, filt)
    yprime = o.convs(y)
    ctx.set_conv_mode(S.CONV_TENSOR)
    ctx.set_keep_planes(True)           # the tensor path does not write the FP32 Y plane unless asked to
    ctx.pass_host(img, 2.0, filt)
    ctx.set_keep_planes(False)
    r0, gy = ctx.get_plane(0, (120, 168))
    assert r0 == 0
    rel = np.abs(gy - y) / np.maximum(np.abs(y), 1.0)
    assert rel.max() <= 1e-5
    assert np.array_equal(gy, y)
    _, gyp = ctx.get_plane(4, (120, 168))
    assert np.abs(gyp - yprime).max() < 1.0
    ctx.set_conv_mode(S.CONV_EXACT)
    ctx.pass_host(img, 2.0, filt)
    _, gyp = ctx.get_plane(4, (120, 168))
    assert np.array_equal(gyp, yprime)
    _, gy = ctx.get_plane(0, (120, 168))
    assert np.array_equal(gy, y)                  # bit-exact mode: the product route's kernel with separately rounded mul / add
    # the PRODUCT route's own front end (tensor mode, x2: one double FMA per tap): north-star gate 1e-5 relative on the plane;
    # the fused form differs from the reference by < 2^-52 relative before the rounding to FP32, so nearly every sample is equal
    ctx.set_conv_mode(S.CONV_TENSOR)
    ctx.set_keep_planes(2)
    ctx.pass_host(img, 2.0, filt)
    r0, gy = ctx.get_plane(0, (120, 168))
    ctx.set_keep_planes(False)
    assert r0 == 0
    rel = np.abs(gy - y) / np.maximum(np.abs(y), 1.0)
    assert rel.max() <= 1e-5
    assert (gy != y).mean() < 1e-3
    with pytest.raises(S.SrcnnError):
        ctx.get_plane(1, (120, 168))              # the product route has no output-size chroma planes


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("d", [3, 4])
def test_equal_size_pass(ctx, mode, d):
    """multiply == 1 without step-scale: no axis is resampled.  DELIBERATE deviation from the reference, whose equal-size
    shortcut copies only HALF of each plane (sizeof(unsigned short) instead of sizeof(float), src/frawscale.cpp:185-193)
    and leaves the rest uninitialised heap.  Defined behaviour here: the whole plane is copied, i.e. the result is
    colour-in -> conv I/II/III on the un-resized Y -> colour-out, composed below from the oracle's own stages.  The rows
    the reference does compute from valid data (the top half minus the conv halo) must also equal the reference's own
    ProcessSRCNN output."""
    h, w = 46, 62
    img = O.synth("uniform", h, w, d, 17)
    o = O.best()
    pl = o.color_in(img)
    yprime = o.convs(pl[0])
    want = o.color_out([yprime] + pl[1:])
    want_conv = yprime.astype(np.uint8)
    ctx.set_conv_mode(mode[1])
    for call in ("pass", "process", "dropin"):
        if call == "pass":
            out, conv = ctx.pass_host(img, 1.0, S.BICUBIC, True)
        elif call == "process":
            rc, out, conv = ctx.process(img, 1.0, S.BICUBIC, False, True)
            assert rc == 0
        else:
            if mode[0] == "exact":
                continue                                   # the process-wide context of the C++ API runs tensor mode
            rc, out, conv = S.process_srcnn(img, 1.0, S.BICUBIC, False, True)
            assert rc == 0
        _check_u8(call + " rgb", out, want, mode[0] == "exact")
        _check_u8(call + " conv", conv, want_conv, mode[0] == "exact")
    # chroma / alpha come out exactly (no arithmetic besides the colour conversion) in both modes
    if d == 4:
        assert np.array_equal(np.asarray(out).reshape(h, w, d)[..., 3], img[..., 3])
    # the part of the reference's own output that is computed from initialised memory: rows < h/2 - 6
    rc_r, out_r, _ = o.process(img, 1.0, S.BICUBIC, False, False)
    assert rc_r == 0
    valid = h // 2 - 6
    _check_u8("reference top rows", np.asarray(out).reshape(h, w, d)[:valid], out_r.reshape(h, w, d)[:valid], mode[0] == "exact")
    # step mode with an effective factor of 1 produces nothing (src/libsrcnn.cpp:1007-1008)
    assert ctx.process(img, 1.0, S.BICUBIC, True, True)[0] == S.ENOTHING


@pytest.mark.parametrize("d,hw,filt", [(3, (61, 83), S.BICUBIC), (4, (47, 130), S.LANCZOS3), (3, (200, 333), S.BSPLINE),
                                       (4, (9, 5), S.BILINEAR), (3, (33, 64), S.NEAREST)])
def test_chroma_planes_fp32_path_bit_exact(ctx, d, hw, filt):
    """Cb / Cr / A at x2 go through the front end's FP32 two-tap path (weights 0.25 / 0.75 / 1: fmaf(wA, vA, wB*vB) is the
    correctly rounded exact sum, the same value the reference's double accumulation yields).  Checked against the
    oracle's FP64 resize and against the kernel's own FP64 chroma path (SRCNN_B200_PREP_F64=1), border tiles, odd
    sizes and strips included."""
    import os
    h, w = hw
    img = O.synth("uniform", h, w, d, 77)
    o = O.best()
    planes = o.color_in(img)
    fc = S.NEAREST if filt == S.NEAREST else S.BILINEAR
    want = [o.resize(planes[k], 2 * w, 2 * h, fc) for k in range(1, d)]
    ctx.set_conv_mode(S.CONV_TENSOR)
    ctx.set_keep_planes(True)        # otherwise the output-size chroma planes are never materialised (x2 colour-out path)
    got = {}
    for mode in ("f32", "f64"):
        if mode == "f64":
            os.environ["SRCNN_B200_PREP_F64"] = "1"
        try:
            ctx.pass_host(img, 2.0, filt)
            got[mode] = [ctx.get_plane(k, (2 * h, 2 * w))[1] for k in range(1, d)]
        finally:
            os.environ.pop("SRCNN_B200_PREP_F64", None)
    for k in range(d - 1):
        assert np.array_equal(got["f32"][k], want[k]), "plane %d: FP32 path differs from the oracle" % (k + 1)
        assert np.array_equal(got["f64"][k], want[k]), "plane %d: FP64 path differs from the oracle" % (k + 1)
    # a strip in the middle: rows [r0-6, r1+6) of the planes, same values
    r0, r1 = (2 * h) // 3, (2 * h) // 3 + max(1, (2 * h) // 4)
    ctx.pass_host(img, 2.0, filt, rows=(r0, r1))
    a, pl = ctx.get_plane(1, (2 * h, 2 * w))
    assert np.array_equal(pl, want[0][a:a + pl.shape[0]])
    ctx.set_keep_planes(False)


@pytest.mark.parametrize("d,hw,filt", [(3, (62, 84), S.BICUBIC), (4, (47, 130), S.LANCZOS3), (3, (9, 6), S.BILINEAR),
                                       (4, (33, 64), S.NEAREST), (3, (120, 250), S.BSPLINE)])
def test_chroma_resampled_inside_colour_out(d, hw, filt):
    """x2 with an output width that is a multiple of 4: Cb / Cr / A are handed over at source resolution and resampled
    inside colour-out (color_out_up_kernel).  Exact mode must stay bit-identical to the oracle, equal to the route through
    output-size planes (SRCNN_B200_CHROMA_UP=0), for whole images and strips; the chroma plane getters report that the
    planes do not exist."""
    import os
    h, w = hw
    img = O.synth("uniform", h, w, d, 91)
    rc, want, want_conv = O.best().process(img, 2.0, filt, False, True)
    assert rc == 0
    outs = {}
    for up in ("1", "0"):
        os.environ["SRCNN_B200_CHROMA_UP"] = up
        try:
            c = S.Context(0, S.CONV_EXACT)
        finally:
            os.environ.pop("SRCNN_B200_CHROMA_UP", None)
        out, conv = c.pass_host(img, 2.0, filt, True)
        assert np.array_equal(out.reshape(-1), want) and np.array_equal(conv.reshape(-1), want_conv), "CHROMA_UP=" + up
        if up == "1":
            with pytest.raises(S.SrcnnError):
                c.get_plane(1, (2 * h, 2 * w))
        r0, r1 = (2 * h) // 3, (2 * h) // 3 + max(1, h // 2)
        so, sc = c.pass_host(img, 2.0, filt, True, rows=(r0, r1))
        assert np.array_equal(so, out[r0:r1]) and np.array_equal(sc, conv[r0:r1])
        c.set_conv_mode(S.CONV_TENSOR)
        outs[up], _ = c.pass_host(img, 2.0, filt, False)
        c.close()
    assert np.array_equal(outs["1"], outs["0"])          # the conv kernel is the same; only the chroma route differs


def test_dropin_cpp_api_matches_oracle():
    """ConfigureFilterSRCNN + ProcessSRCNN by their mangled names, new[] ownership, conv plane optional"""
    img = O.synth("uniform", 48, 64, 4, 8)
    rc_o, out_o, conv_o = O.best().process(img, 2.0, S.LANCZOS3, False)
    rc, out, conv = S.process_srcnn(img, 2.0, S.LANCZOS3, False, True)
    assert rc == rc_o == 0
    _check_u8("rgb", out, out_o, False)
    _check_u8("conv", conv, conv_o, False)
    rc, out2, conv2 = S.process_srcnn(img, 2.0, S.LANCZOS3, False, False)
    assert rc == 0 and conv2 is None and np.array_equal(out2, out)
    rc, out3, _ = S.process_srcnn(img, 3.0, S.BICUBIC, True, True)
    rc_o, out_o3, _ = O.best().process(img, 3.0, S.BICUBIC, True)
    assert rc == rc_o == 0
    _check_u8("rgb-step3", out3, out_o3, False)
    S.lib().ConfigureFilterSRCNN(S.BICUBIC, False)


# ---------------------------------------------------------------------------------------------- sharding
@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("nstrips", [2, 3, 8])
def test_row_strips_equal_full_image(ctx, mode, nstrips):
    """row-strip shards (6-row halo recomputed) reproduce the single-call result bit for bit"""
    ctx.set_conv_mode(mode[1])
    img = O.synth("uniform", 101, 150, 4, 50)
    full, fconv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
    oh = full.shape[0]
    cuts = [round(i * oh / nstrips) for i in range(nstrips + 1)]
    parts, cparts = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        o, c = ctx.pass_host(img, 2.0, S.BICUBIC, True, rows=(a, b))
        parts.append(o); cparts.append(c)
    assert np.array_equal(np.concatenate(parts, 0), full)
    assert np.array_equal(np.concatenate(cparts, 0), fconv)


def test_single_row_strips(ctx):
    ctx.set_conv_mode(S.CONV_TENSOR)
    img = O.synth("smooth", 20, 70, 3, 51)
    full, _ = ctx.pass_host(img, 2.0, S.BSPLINE, False)
    for r in [0, 1, 5, 20, 38, 39]:
        o, _ = ctx.pass_host(img, 2.0, S.BSPLINE, False, rows=(r, r + 1))
        assert np.array_equal(o[0], full[r])


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_batch_equals_per_frame(ctx, mode):
    ctx.set_conv_mode(mode[1])
    frames = np.stack([O.synth("uniform", 48, 48, 3, 60 + i) for i in range(5)])
    out, conv = ctx.batch(frames, 2.0, S.LANCZOS3, want_conv=True)
    for i in range(5):
        o, c = ctx.pass_host(frames[i], 2.0, S.LANCZOS3, True)
        assert np.array_equal(out[i], o) and np.array_equal(conv[i], c)


@pytest.mark.parametrize("d,hw,filt,n", [(3, (64, 64), S.LANCZOS3, 7), (4, (40, 72), S.BICUBIC, 5), (3, (33, 50), S.BICUBIC, 3)])
def test_frame_groups_equal_single_frames(ctx, d, hw, filt, n):
    """batches of small frames run the fused convolution as ONE launch per group of frames (srcnn_b200_batch_host /
    _batch_device): every frame must come out bit-identical to its own single pass (tensor mode included -- the FP32
    sums do not depend on how rows are cut into work units), group sizes that do not divide the batch included"""
    import os
    import torch
    h, w = hw
    frames = np.stack([O.synth("uniform", h, w, d, 300 + i) for i in range(n)])
    ctx.set_conv_mode(S.CONV_TENSOR)
    singles = [ctx.pass_host(frames[i], 2.0, filt, True) for i in range(n)]
    out, conv = ctx.batch(frames, 2.0, filt, want_conv=True)
    for i in range(n):
        assert np.array_equal(out[i], singles[i][0]) and np.array_equal(conv[i], singles[i][1]), "host batch, frame %d" % i
    d_src = torch.from_numpy(frames).cuda()
    d_out = torch.zeros((n, 2 * h, 2 * w, d), dtype=torch.uint8, device="cuda")
    d_conv = torch.zeros((n, 2 * h, 2 * w), dtype=torch.uint8, device="cuda")
    ctx.batch_device(d_src.data_ptr(), n, w, h, d, 2.0, filt, d_out.data_ptr(), d_conv.data_ptr())
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy(), out) and np.array_equal(d_conv.cpu().numpy(), conv)
    # and with grouping switched off (a context created with SRCNN_B200_GROUP=0)
    os.environ["SRCNN_B200_GROUP"] = "0"
    try:
        c2 = S.Context(0, S.CONV_TENSOR)
    finally:
        os.environ.pop("SRCNN_B200_GROUP", None)
    o2, cv2 = c2.batch(frames, 2.0, filt, want_conv=True)
    c2.close()
    assert np.array_equal(o2, out) and np.array_equal(cv2, conv)


def test_device_resident_pass_equals_host_pass(ctx):
    import torch
    ctx.set_conv_mode(S.CONV_TENSOR)
    img = O.synth("uniform", 90, 120, 3, 70)
    want, wconv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
    d_src = torch.from_numpy(img).cuda()
    d_out = torch.empty((180, 240, 3), dtype=torch.uint8, device="cuda")
    d_conv = torch.empty((180, 240), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.pass_device(d_src.data_ptr(), 120, 90, 3, 2.0, S.BICUBIC, d_out.data_ptr(), d_conv.data_ptr())
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy(), want) and np.array_equal(d_conv.cpu().numpy(), wconv)


# ---------------------------------------------------------------------------------------------- full sizes
def test_c2_full_size_against_oracle(ctx):
    """BASELINE config 2 at full size: 1920x1080 RGB x2 bicubic -> 3840x2160, against the oracle (CPU, ~10-30 s)"""
    img = O.synth("uniform", 1080, 1920, 3, 42)
    rc_o, out_o, conv_o = O.best().process(img, 2.0, S.BICUBIC, False)
    assert rc_o == 0
    out_o = out_o.reshape(2160, 3840, 3)
    ctx.set_conv_mode(S.CONV_EXACT)
    out, conv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
    _check_u8("exact rgb", out, out_o, True)
    _check_u8("exact conv", conv, conv_o.reshape(2160, 3840), True)
    ctx.set_conv_mode(S.CONV_TENSOR)
    out, conv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
    _check_u8("tensor rgb", out, out_o, False)
    _check_u8("tensor conv", conv, conv_o.reshape(2160, 3840), False)


FULL = [
    dict(id="c3-7680x4320-rgba", h=4320, w=7680, d=4, m=2.0, f=S.BICUBIC, seed=43, strips=8),
    dict(id="c4-frame-512-lanczos", h=512, w=512, d=3, m=2.0, f=S.LANCZOS3, seed=44, strips=2),
    dict(id="c5-3840x2160-x1.5-bspline", h=2160, w=3840, d=3, m=1.5, f=S.BSPLINE, seed=45, strips=4),
]


@pytest.mark.parametrize("case", FULL, ids=[c["id"] for c in FULL])
def test_full_size_properties(ctx, case):
    """BASELINE configs 3-5 at full size through size-independent properties:
    (a) tensor mode within 1 LSB / 50 dB of the bit-exact mode (itself pinned to the oracle at small sizes and at C2);
    (b) row-strip shards concatenate to the single-call result bit for bit;
    (c) an oracle run on a crop of the source agrees with the interior of the same region (all stages are local)."""
    img = O.synth("uniform", case["h"], case["w"], case["d"], case["seed"])
    ow, oh = S.output_dims(case["w"], case["h"], case["m"])
    ctx.set_conv_mode(S.CONV_TENSOR)
    full, fconv = ctx.pass_host(img, case["m"], case["f"], True)
    # (b)
    cuts = [round(i * oh / case["strips"]) for i in range(case["strips"] + 1)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        o, c = ctx.pass_host(img, case["m"], case["f"], True, rows=(a, b))
        assert np.array_equal(o, full[a:b]) and np.array_equal(c, fconv[a:b])
    # (a)
    ctx.set_conv_mode(S.CONV_EXACT)
    ex, exconv = ctx.pass_host(img, case["m"], case["f"], True)
    _check_u8("tensor-vs-exact rgb", full, ex, False)
    _check_u8("tensor-vs-exact conv", fconv, exconv, False)
    # (c) top-left crop: source 96x128 -> compare the region at least 16 output px away from the crop's cut edges
    ch, cw = min(96, case["h"]), min(128, case["w"])
    rc, o_crop, _ = O.best().process(img[:ch, :cw], case["m"], case["f"], False)
    cow, coh = S.output_dims(cw, ch, case["m"])
    o_crop = o_crop.reshape(coh, cow, case["d"])
    ih = coh - 16 if ch < case["h"] else coh
    iw = cow - 16 if cw < case["w"] else cow
    assert np.array_equal(ex[:ih, :iw], o_crop[:ih, :iw])


def _oracle_rows(img, m, filt, step, r0, r1, margin_src=24):
    """oracle output rows [r0, r1) of the full result, computed from just the source rows they depend on: the resampling
    maps are shift-invariant for a crop that starts on a whole source row when m is an integer, so a crop with `margin_src`
    extra source rows on each cut side (conv halo 6 output rows + the widest resampling window) reproduces the interior
    bit for bit.  Returns (rgb rows, conv rows)."""
    h = img.shape[0]
    mi = int(m)
    assert mi == m
    a = max(0, r0 // mi - margin_src)
    b = min(h, (r1 + mi - 1) // mi + margin_src)
    rc, o, c = O.best().process(img[a:b], m, filt, step, True)
    assert rc == 0
    ow = img.shape[1] * mi
    o = o.reshape((b - a) * mi, ow, img.shape[2]); c = c.reshape((b - a) * mi, ow)
    return o[r0 - a * mi: r1 - a * mi], c[r0 - a * mi: r1 - a * mi]


def test_c3_strips_directly_against_oracle(ctx):
    """BASELINE config 3 (7680x4320 RGBA x2 -> 15360x8640, 8 row strips): both border strips and an interior strip of the
    TENSOR path directly against the oracle run on the source rows those strips depend on."""
    img = O.synth("uniform", 4320, 7680, 4, 43)
    oh = 8640
    ctx.set_conv_mode(S.CONV_TENSOR)
    for k in (0, 3, 7):
        r0, r1 = S.strip_rows(oh, 8, k)
        got, gconv = ctx.pass_host(img, 2.0, S.BICUBIC, True, rows=(r0, r1))
        want, wconv = _oracle_rows(img, 2.0, S.BICUBIC, False, r0, r1)
        _check_u8("strip %d rgba" % k, got, want, False)
        _check_u8("strip %d conv" % k, gconv, wconv, False)
        if k == 3:                                        # and bit for bit in exact mode, one strip
            ctx.set_conv_mode(S.CONV_EXACT)
            got, gconv = ctx.pass_host(img, 2.0, S.BICUBIC, True, rows=(r0, r1))
            ctx.set_conv_mode(S.CONV_TENSOR)
            _check_u8("strip %d rgba exact" % k, got, want, True)
            _check_u8("strip %d conv exact" % k, gconv, wconv, True)


def test_c4_frames_directly_against_oracle(ctx):
    """BASELINE config 4 (512x512 RGB x2 Lanczos-3 frames, batched): every frame of a batch that spans two frame groups,
    TENSOR path, directly against the oracle"""
    n = 18
    frames = np.stack([O.synth("uniform", 512, 512, 3, 44 + i) for i in range(n)])
    ctx.set_conv_mode(S.CONV_TENSOR)
    out, conv = ctx.batch(frames, 2.0, S.LANCZOS3, want_conv=True)
    for i in (0, 1, 7, 15, 16, 17):
        rc, want, wconv = O.best().process(frames[i], 2.0, S.LANCZOS3, False, True)
        assert rc == 0
        _check_u8("frame %d rgb" % i, out[i], want.reshape(1024, 1024, 3), False)
        _check_u8("frame %d conv" % i, conv[i], wconv.reshape(1024, 1024), False)


def test_c5_full_size_directly_against_oracle(ctx):
    """BASELINE config 5 at full size (3840x2160 RGB x1.5 B-spline, step-scale ON, convolution output ON): the whole
    frame, TENSOR path through the ProcessSRCNN-level entry point, directly against the oracle"""
    img = O.synth("uniform", 2160, 3840, 3, 45)
    rc_o, want, wconv = O.best().process(img, 1.5, S.BSPLINE, True, True)
    assert rc_o == 0
    ctx.set_conv_mode(S.CONV_TENSOR)
    rc, out, conv = ctx.process(img, 1.5, S.BSPLINE, True, True)
    assert rc == 0 and out.shape == (3240, 5760, 3)
    _check_u8("c5 rgb", out, want.reshape(3240, 5760, 3), False)
    _check_u8("c5 conv", conv, wconv.reshape(3240, 5760), False)


# ---------------------------------------------------------------------------------------------- host-side drivers
def test_async_passes_equal_blocking_passes(ctx):
    """srcnn_b200_pass_host_async: several frames in flight on one context (pinned buffers), then one wait -- every result equals
    the blocking call's; a blocking call issued while asynchronous work is pending waits for it first"""
    ctx.set_conv_mode(S.CONV_TENSOR)
    frames = [O.synth("uniform", 120, 200, 3, 400 + i) for i in range(5)]
    want = [ctx.pass_host(f, 2.0, S.BICUBIC, True) for f in frames]
    h_in = [S.pinned_empty(f.shape) for f in frames]
    h_out = [S.pinned_empty((240, 400, 3)) for _ in frames]
    h_conv = [S.pinned_empty((240, 400)) for _ in frames]
    for a, f in zip(h_in, frames):
        a[...] = f
    for o in h_out:
        o[...] = 0
    for i in range(5):
        ctx.pass_host_async(h_in[i], 2.0, S.BICUBIC, h_out[i], h_conv[i])
    ctx.wait()
    for i in range(5):
        assert np.array_equal(h_out[i], want[i][0]) and np.array_equal(h_conv[i], want[i][1]), "frame %d" % i
    # pending asynchronous work + a blocking call on the same context
    ctx.pass_host_async(h_in[0], 2.0, S.BICUBIC, h_out[0], None)
    ctx.pass_host_async(h_in[1], 2.0, S.BICUBIC, h_out[1], None)
    o, _ = ctx.pass_host(frames[2], 2.0, S.BICUBIC, False)
    assert np.array_equal(o, want[2][0]) and np.array_equal(h_out[0], want[0][0]) and np.array_equal(h_out[1], want[1][0])
    for a in h_in + h_out + h_conv:
        S.lib().srcnn_b200_host_free(a.ctypes.data)


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
def test_single_process_multi_gpu_driver(ctx, mode, monkeypatch):
    """srcnn_b200_multi: ONE host call cut into row strips / frame blocks over several per-device contexts and threads.  On a
    one-GPU box the workers are three contexts on device 0 (the strip arithmetic, the threading and the per-context state are
    the same); with more GPUs the first three devices.  Results must be bit-identical to the single-context call."""
    monkeypatch.setenv("SRCNN_B200_MULTI_MIN_ROWS", "16")
    n = S.device_count()
    devs = [0, 1 % n, 2 % n]
    m = S.Multi(devs, mode[1])
    assert m.ndev == 3
    ctx.set_conv_mode(mode[1])
    try:
        img = O.synth("uniform", 101, 150, 4, 50)
        want, wconv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
        got, gconv = m.pass_host(img, 2.0, S.BICUBIC, True)
        assert np.array_equal(got, want) and np.array_equal(gconv, wconv)
        # the whole ProcessSRCNN semantics, step-scale chain included (every pass on all workers)
        for mult, step in ((2.0, False), (3.0, True), (4.0, True), (1.5, True)):
            rc1, o1, c1 = ctx.process(img[:40, :60], mult, S.LANCZOS3, step, True)
            rc2, o2, c2 = m.process(img[:40, :60], mult, S.LANCZOS3, step, True)
            assert rc1 == rc2 == 0 and np.array_equal(o1, o2) and np.array_equal(c1, c2), (mult, step)
        assert m.process(img, 1.0, S.BICUBIC, True)[0] == S.ENOTHING
        # frame batches: contiguous blocks of frames per worker
        frames = np.stack([O.synth("uniform", 48, 64, 3, 60 + i) for i in range(7)])
        bo, bc = ctx.batch(frames, 2.0, S.LANCZOS3, want_conv=True)
        mo, mc = m.batch(frames, 2.0, S.LANCZOS3, want_conv=True)
        assert np.array_equal(bo, mo) and np.array_equal(bc, mc)
    finally:
        m.close()


def test_two_contexts_on_two_devices():
    """one process, one context per GPU (function attributes are per device) -- needs two GPUs"""
    if S.device_count() < 2:
        pytest.skip("needs two GPUs")
    img = O.synth("uniform", 200, 300, 3, 9)
    a, b = S.Context(0), S.Context(1)
    try:
        for mode in (S.CONV_TENSOR, S.CONV_EXACT):
            a.set_conv_mode(mode); b.set_conv_mode(mode)
            for filt, mult in ((S.BICUBIC, 2.0), (S.LANCZOS3, 2.0), (S.BSPLINE, 1.5)):
                oa, ca = a.pass_host(img, mult, filt, True)
                ob, cb = b.pass_host(img, mult, filt, True)
                assert np.array_equal(oa, ob) and np.array_equal(ca, cb)
    finally:
        a.close(); b.close()


def test_dropin_is_reentrant():
    """concurrent ProcessSRCNN calls from several threads (the reference is re-entrant apart from its two settings,
    src/libsrcnn.cpp:91-92): every call returns the result of a lone call"""
    import threading
    imgs = [O.synth("uniform", 90 + 10 * i, 130, 3 + (i & 1), 700 + i) for i in range(6)]
    want = [S.process_srcnn(im, 2.0, S.BICUBIC, False, True) for im in imgs]
    got = [None] * len(imgs)

    def work(i):
        for _ in range(3):
            got[i] = S.process_srcnn(imgs[i], 2.0, S.BICUBIC, False, True)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(imgs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i in range(len(imgs)):
        assert got[i][0] == want[i][0] == 0
        assert np.array_equal(got[i][1], want[i][1]) and np.array_equal(got[i][2], want[i][2]), "image %d" % i


def test_dropin_over_several_devices():
    """SRCNN_B200_DEVICES: the C++ drop-in call itself cut into row strips over the listed devices (a fresh process: the setting
    is read once)"""
    import subprocess, sys
    n = S.device_count()
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, libsrcnn_b200 as S, oracle_lib as O\n"
        "img = O.synth('uniform', 300, 200, 4, 5)\n"
        "rc, out, conv = S.process_srcnn(img, 2.0, S.BICUBIC, False, True)\n"
        "c = S.Context(0); w, wc = c.pass_host(img, 2.0, S.BICUBIC, True)\n"
        "assert rc == 0 and np.array_equal(out, w.reshape(-1)) and np.array_equal(conv, wc.reshape(-1))\n"
        "rc, out, _ = S.process_srcnn(img[:50, :60], 3.0, S.BICUBIC, True, False)\n"
        "rc2, w2, _ = c.process(img[:50, :60], 3.0, S.BICUBIC, True, False)\n"
        "assert rc == rc2 == 0 and np.array_equal(out, w2.reshape(-1))\n"
        "print('OK')\n" % (str(O.ROOT), str(O.ROOT / "tests")))
    env = dict(__import__("os").environ, SRCNN_B200_DEVICES="0,%d" % (1 % n), SRCNN_B200_MULTI_MIN_ROWS="32")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_cuda_graph_replay_equals_plain_launches(ctx):
    """srcnn_b200_set_graphs: the repeated device-resident pass is captured (second call) and replayed (from the third on); results
    equal the plain launch sequence, a change of geometry or pointers falls back to a fresh capture"""
    import torch
    ctx.set_conv_mode(S.CONV_TENSOR)
    img = O.synth("uniform", 150, 200, 3, 71)
    want, wconv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
    d_src = torch.from_numpy(img).cuda()
    d_out = torch.zeros((300, 400, 3), dtype=torch.uint8, device="cuda")
    d_conv = torch.zeros((300, 400), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.set_graphs(True)
    try:
        r0 = ctx.graph_replays
        for i in range(5):
            d_out.zero_(); d_conv.zero_(); torch.cuda.synchronize()
            ctx.pass_device(d_src.data_ptr(), 200, 150, 3, 2.0, S.BICUBIC, d_out.data_ptr(), d_conv.data_ptr())
            ctx.sync()
            assert np.array_equal(d_out.cpu().numpy(), want) and np.array_equal(d_conv.cpu().numpy(), wconv), "call %d" % i
        assert ctx.graph_replays - r0 >= 3
        # another frame through the same pointers: the graph reads the new data
        img2 = O.synth("uniform", 150, 200, 3, 72)
        d_src.copy_(torch.from_numpy(img2)); torch.cuda.synchronize()
        ctx.pass_device(d_src.data_ptr(), 200, 150, 3, 2.0, S.BICUBIC, d_out.data_ptr(), d_conv.data_ptr()); ctx.sync()
        ctx.set_graphs(False)
        want2, _ = ctx.pass_host(img2, 2.0, S.BICUBIC, False)
        assert np.array_equal(d_out.cpu().numpy(), want2)
        ctx.set_graphs(True)
        # a strip of the same image (different geometry): its own capture
        for i in range(3):
            ctx.pass_device(d_src.data_ptr(), 200, 150, 3, 2.0, S.BICUBIC, d_out[100:].data_ptr(), 0, rows=(100, 300)); ctx.sync()
            assert np.array_equal(d_out[100:].cpu().numpy(), want2[100:])
    finally:
        ctx.set_graphs(False)


@pytest.mark.parametrize("m,filt,d,hw", [(1.5, S.BSPLINE, 3, (120, 200)), (2.5, S.BICUBIC, 4, (64, 96)), (3.0, S.LANCZOS3, 3, (60, 80)),
                                         (1.75, S.BILINEAR, 3, (96, 128))])
def test_chroma_handover_at_other_magnifications(m, filt, d, hw):
    """tensor mode, magnification >= 1.5 other than x2: Cb / Cr / A are handed over at source resolution as well and resampled inside
    colour-out with FP32 two-tap arithmetic (approximate there -- the weights are not powers of two -- and gated by the final
    1 LSB / 50 dB); SRCNN_B200_CHROMA_EXACT=1 and the bit-exact conv mode keep the reference's double arithmetic on output-size planes"""
    import os
    h, w = hw
    img = O.synth("uniform", h, w, d, 97)
    rc, want, want_conv = O.best().process(img, m, filt, False, True)
    assert rc == 0
    ow, oh = S.output_dims(w, h, m)
    c = S.Context(0, S.CONV_TENSOR)
    out, conv = c.pass_host(img, m, filt, True)
    _check_u8("rgb", out, want.reshape(oh, ow, d), False)
    _check_u8("conv", conv, want_conv.reshape(oh, ow), False)
    if ow % 4 == 0:
        with pytest.raises(S.SrcnnError):
            c.get_plane(1, (oh, ow))                      # product route: no output-size chroma planes
    # strips of the product route equal the full image
    r0, r1 = oh // 3, oh // 3 + max(1, oh // 4)
    so, sc = c.pass_host(img, m, filt, True, rows=(r0, r1))
    assert np.array_equal(so, out[r0:r1]) and np.array_equal(sc, conv[r0:r1])
    # bit-exact mode never takes the approximate route
    c.set_conv_mode(S.CONV_EXACT)
    ex, exconv = c.pass_host(img, m, filt, True)
    _check_u8("exact rgb", ex, want.reshape(oh, ow, d), True)
    c.close()
    os.environ["SRCNN_B200_CHROMA_EXACT"] = "1"
    try:
        c2 = S.Context(0, S.CONV_TENSOR)
        out2, _ = c2.pass_host(img, m, filt, False)
        c2.get_plane(1, (oh, ow))                         # the planes exist on this route
        c2.close()
    finally:
        os.environ.pop("SRCNN_B200_CHROMA_EXACT", None)
    _check_u8("rgb (exact chroma route)", out2, want.reshape(oh, ow, d), False)
    # the two routes differ only where an FP32-rounded chroma sample crosses a truncation boundary
    assert (np.abs(out.astype(int) - out2.astype(int)) > 1).sum() == 0
    assert (out != out2).mean() < 0.01
