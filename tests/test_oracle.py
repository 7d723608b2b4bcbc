"""CPU tests (-m "not gpu"): pin the oracle.

1. the compiled reference (oracle/_ref) and the plain-C port reproduce the reference's golden vectors
   (tests/golden/butterfly_srcnn*.png, from the reference's Pictures/) bit-exactly;
2. the port equals the compiled reference stage by stage and end to end (max |diff| == 0).
"""
import numpy as np
import pytest
from PIL import Image
import oracle_lib as O

G = O.ROOT / "tests" / "golden"


def _butterfly():
    return (np.asarray(Image.open(G / "butterfly.png").convert("RGB")),
            np.asarray(Image.open(G / "butterfly_srcnn.png").convert("RGB")),
            np.asarray(Image.open(G / "butterfly_srcnn_convolution.png").convert("L")))


def _oracles():
    xs = [O.port()]
    if O.ref() is not None:
        xs.append(O.ref())
    return xs


@pytest.mark.parametrize("which", ["port", "reference"])
def test_golden_butterfly(which):
    o = O.port() if which == "port" else O.ref()
    if o is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    src, gold, gconv = _butterfly()
    rc, out, conv = o.process(src, 2.0, filt=2, stepscale=False)
    assert rc == 0
    assert np.array_equal(out.reshape(512, 512, 3), gold)
    assert np.array_equal(conv.reshape(512, 512), gconv)


needs_ref = pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("d", [3, 4])
def test_color_in_matches_reference(d):
    img = O.synth("uniform", 37, 53, d, 1)
    a, b = O.port().color_in(img), O.ref().color_in(img)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@needs_ref
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("dst,src", [(512, 256), (97, 64), (100, 100 // 2 + 7), (33, 77), (300, 299), (8, 3)])
def test_contrib_table_matches_reference(filt, dst, src):
    La, Ra, Wa = O.port().table(filt, dst, src)
    Lb, Rb, Wb = O.ref().table(filt, dst, src)
    assert np.array_equal(La, Lb) and np.array_equal(Ra, Rb)
    assert np.array_equal(Wa, Wb)


@needs_ref
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("shape", [((40, 56), (80, 112)), ((31, 45), (46, 67)), ((64, 64), (40, 30)), ((20, 30), (20, 60))])
def test_resize_matches_reference(filt, shape):
    (sh, sw), (dh, dw) = shape
    p = np.random.default_rng(3).uniform(-20, 280, size=(sh, sw)).astype(np.float32)
    assert np.array_equal(O.port().resize(p, dw, dh, filt), O.ref().resize(p, dw, dh, filt))


@needs_ref
def test_convs_match_reference():
    y = np.random.default_rng(4).uniform(-20, 280, size=(23, 31)).astype(np.float32)
    a1, b1 = O.port().conv1(y), O.ref().conv1(y)
    assert np.array_equal(a1, b1)
    a2, b2 = O.port().conv2(a1), O.ref().conv2(b1)
    assert np.array_equal(a2, b2)
    assert np.array_equal(O.port().conv3(a2), O.ref().conv3(b2))


@needs_ref
def test_color_out_matches_reference():
    rng = np.random.default_rng(5)
    pl = [rng.uniform(-30, 290, size=(19, 27)).astype(np.float32) for _ in range(4)]
    assert np.array_equal(O.port().color_out(pl), O.ref().color_out(pl))
    assert np.array_equal(O.port().color_out(pl[:3]), O.ref().color_out(pl[:3]))


@needs_ref
@pytest.mark.parametrize("case", [
    dict(kind="uniform", h=24, w=40, d=3, m=2.0, f=2, s=False),
    dict(kind="uniform", h=24, w=40, d=4, m=2.0, f=3, s=False),
    dict(kind="binary", h=20, w=28, d=3, m=1.5, f=4, s=True),
    dict(kind="smooth", h=18, w=22, d=3, m=3.0, f=2, s=True),     # x2 then x1.5
    dict(kind="uniform", h=12, w=16, d=4, m=4.0, f=1, s=True),    # x2 then x2
    dict(kind="checker", h=21, w=33, d=3, m=2.5, f=0, s=False),
    dict(kind="uniform", h=10, w=14, d=3, m=1.0, f=2, s=True),    # step mode, factor 1 -> -100, no output
])
def test_process_matches_reference(case):
    img = O.synth(case["kind"], case["h"], case["w"], case["d"], 7)
    ra = O.port().process(img, case["m"], case["f"], case["s"])
    rb = O.ref().process(img, case["m"], case["f"], case["s"])
    assert ra[0] == rb[0]
    for x, y in zip(ra[1:], rb[1:]):
        assert (x is None) == (y is None)
        if x is not None:
            assert np.array_equal(x, y)


def test_error_codes_port():
    img = O.synth("uniform", 8, 8, 3, 0)
    assert O.port().process(img, 0.0)[0] == -2
    assert O.port().process(img, -1.0)[0] == -2
