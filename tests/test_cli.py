"""The command-line harness (libsrcnn_b200/csrc/srcnn_cli.cpp, SURVEY.md 8(f)-1: the counterpart of the reference's
src/test.cpp) and its image codec (imgio.cpp).

CPU tests: the PNG / BMP / PNM codec against PIL, the command-line conventions of the reference (default output
names, option parsing, src/test.cpp:290-448), and -- with the harness linked against the COMPILED REFERENCE
(oracle/_ref/srcnn_ref_cli, test infrastructure) -- that it reproduces the reference's golden images bit for bit.
GPU tests: the same harness linked against the B200 library reproduces the goldens (exact mode: bit-identical,
tensor mode: within 1 LSB / PSNR >= 50 dB).
"""
import os
import pathlib
import shutil
import subprocess

import numpy as np
import pytest
from PIL import Image

import oracle_lib as O

ROOT = O.ROOT
G = ROOT / "tests" / "golden"
CLI = ROOT / "libsrcnn_b200" / "bin" / "srcnn"
REF_CLI = ROOT / "oracle" / "_ref" / "srcnn_ref_cli"


def _run(exe, *args, env=None, cwd=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, env=e, cwd=cwd, timeout=300)


def _need_cli():
    if not CLI.exists():
        subprocess.run(["make", "-C", str(ROOT / "libsrcnn_b200" / "csrc")], check=True, capture_output=True)
    assert CLI.exists()


def _rng_img(h, w, d, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, d) if d > 1 else (h, w), dtype=np.uint8)


PNG_CASES = {
    "rgb": lambda: Image.fromarray(_rng_img(37, 53, 3)),
    "rgba": lambda: Image.fromarray(_rng_img(19, 64, 4, 1)),
    "gray": lambda: Image.fromarray(_rng_img(40, 31, 1, 2)),
    "gray_alpha": lambda: Image.fromarray(_rng_img(12, 17, 4, 3)).convert("LA"),
    "palette": lambda: Image.fromarray(_rng_img(33, 29, 3, 4)).quantize(64),
    "bilevel": lambda: Image.fromarray(_rng_img(21, 45, 1, 5) > 127),
    "rgb16": None,   # written by hand below (PIL cannot write 16-bit RGB)
    "smooth_big": lambda: Image.fromarray(O.synth("smooth", 300, 411, 3, 6)),
}


@pytest.mark.parametrize("name", sorted(k for k, v in PNG_CASES.items() if v))
def test_png_decode_encode_matches_pil(tmp_path, name):
    """decode (all filter types / colour types PIL produces) -> RGB(A) -> encode; compared with PIL's view of both files"""
    _need_cli()
    src = tmp_path / (name + ".png")
    im = PNG_CASES[name]()
    im.save(src, optimize=(name == "smooth_big"))
    dst = tmp_path / "out.png"
    r = _run(CLI, src, "--decode-only", dst)
    assert r.returncode == 0, r.stdout + r.stderr
    ref = Image.open(src)
    want = np.asarray(ref.convert("RGBA" if ("A" in ref.mode or "transparency" in ref.info) else "RGB"))
    got = np.asarray(Image.open(dst))
    assert got.shape == want.shape and np.array_equal(got, want)


def test_png_16bit_keeps_high_byte(tmp_path):
    import struct, zlib
    h, w = 5, 7
    a = np.random.default_rng(7).integers(0, 65536, (h, w, 3), dtype=np.uint16)
    raw = b"".join(b"\x00" + a[y].astype(">u2").tobytes() for y in range(h))
    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b))
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)) + \
          chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    src = tmp_path / "rgb16.png"; src.write_bytes(png)
    _need_cli()
    r = _run(CLI, src, "--decode-only", tmp_path / "o.png")
    assert r.returncode == 0, r.stdout
    assert np.array_equal(np.asarray(Image.open(tmp_path / "o.png")), (a >> 8).astype(np.uint8))


@pytest.mark.parametrize("fmt,d", [("bmp", 3), ("ppm", 3), ("pgm", 1)])
def test_bmp_and_pnm_decode(tmp_path, fmt, d):
    _need_cli()
    a = _rng_img(23, 35, d, 11)
    src = tmp_path / ("in." + fmt)
    Image.fromarray(a).save(src)
    r = _run(CLI, src, "--decode-only", tmp_path / "o.png")
    assert r.returncode == 0, r.stdout
    want = a if d == 3 else np.repeat(a[:, :, None], 3, axis=2)            # gray -> RGB like the reference's convImage
    assert np.array_equal(np.asarray(Image.open(tmp_path / "o.png")), want)


def test_rejects_garbage_and_progressive_jpeg(tmp_path):
    _need_cli()
    (tmp_path / "x.png").write_bytes(b"not an image at all")
    assert _run(CLI, tmp_path / "x.png", "--decode-only").returncode != 0
    Image.fromarray(_rng_img(8, 8, 3)).save(tmp_path / "x.jpg", progressive=True)
    r = _run(CLI, tmp_path / "x.jpg", "--decode-only")
    assert r.returncode != 0 and "progressive" in r.stdout                   # baseline only: refused with a message
    (tmp_path / "t.jpg").write_bytes((tmp_path / "x.jpg").read_bytes()[:60])
    assert _run(CLI, tmp_path / "t.jpg", "--decode-only").returncode != 0    # truncated
    assert _run(CLI).returncode != 0                                         # no arguments: usage, non-zero


@pytest.mark.parametrize("name,kw", [("444", dict(subsampling=0)), ("422", dict(subsampling=1)), ("420", dict(subsampling=2)),
                                     ("q30", dict(subsampling=2, quality=30)), ("gray", None), ("odd", dict(subsampling=2))])
def test_baseline_jpeg_decode_matches_a_libjpeg_loader(tmp_path, name, kw):
    """the reference CLI loads JPEG through FLTK / libjpeg (src/test.cpp:145-148, 563-566; its sample Pictures/castle.jpg is a
    baseline 4:2:0 file): the built-in baseline decoder (jpegdec.cpp: libjpeg's fancy upsampling and colour tables, a plain
    double-precision IDCT) must stay within 3 LSB of PIL's libjpeg result, with a mean difference below 0.1 LSB"""
    _need_cli()
    h, w = (77, 101) if name != "odd" else (33, 47)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 9.0 + yy / 17.0), 127 + 100 * np.cos(xx / 5.0), 60 + yy * 2.0], 2).clip(0, 255).astype(np.uint8)
    f = tmp_path / ("t_%s.jpg" % name)
    if kw is None:
        Image.fromarray(img[..., 0]).save(f, quality=90)
    else:
        Image.fromarray(img).save(f, quality=kw.get("quality", 92), subsampling=kw["subsampling"])
    r = _run(CLI, f, "--decode-only", tmp_path / "o.png")
    assert r.returncode == 0 and "JPEG" in r.stdout, r.stdout
    got = np.asarray(Image.open(tmp_path / "o.png")).astype(int)
    want = np.asarray(Image.open(f).convert("RGB")).astype(int)
    assert got.shape == want.shape
    diff = np.abs(got - want)
    assert diff.max() <= 3 and diff.mean() < 0.1, (diff.max(), diff.mean())


def test_corrupt_and_truncated_png_are_rejected_cleanly(tmp_path):
    """damaged files must produce an error message and a non-zero exit code, never a crash or a bogus image"""
    _need_cli()
    good = tmp_path / "good.png"
    Image.fromarray(_rng_img(40, 60, 3, 9)).save(good)
    raw = good.read_bytes()
    cases = {
        "cut_in_idat": raw[: len(raw) // 2],
        "cut_after_sig": raw[:8],
        "bad_zlib": raw[:60] + bytes(b ^ 0x5A for b in raw[60:120]) + raw[120:],
        "huge_dims": raw[:16] + (0x7FFFFFFF).to_bytes(4, "big") + (0x7FFFFFFF).to_bytes(4, "big") + raw[24:],
        "empty": b"",
    }
    import struct, zlib
    ihdr = struct.pack(">IIBBBBB", 0x7FFFFFFF, 0x7FFFFFFF, 8, 2, 0, 0, 0)      # absurd size, VALID chunk CRCs
    def chunk(t, b):
        return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b))
    cases["huge_dims_valid_crc"] = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    for name, data in cases.items():
        p = tmp_path / (name + ".png")
        p.write_bytes(data)
        r = _run(CLI, p, "--decode-only", tmp_path / "o.png")
        assert r.returncode not in (0, -11, 139), "%s: rc %s\n%s" % (name, r.returncode, r.stdout + r.stderr)
        assert "Failed to load image" in r.stdout, name


def test_default_output_names_follow_the_reference(tmp_path):
    """<src>_resized.png next to the source when no output name is given (src/test.cpp:385-411)"""
    _need_cli()
    shutil.copy(G / "butterfly.png", tmp_path / "pic.png")
    r = _run(CLI, tmp_path / "pic.png", "--decode-only")
    assert r.returncode == 0 and (tmp_path / "pic_resized.png").exists(), r.stdout


def test_without_gpu_the_cli_reports_minus_20(tmp_path):
    """no CPU fallback: on a machine without a B200 ProcessSRCNN returns -20 and the harness says so"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _need_cli()
    r = _run(CLI, G / "butterfly.png", tmp_path / "o.png")
    assert r.returncode != 0 and "error code = -20" in r.stdout, r.stdout


@pytest.mark.skipif(not REF_CLI.exists(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("args,gold", [((), "butterfly_srcnn.png"), (("--scale=2.0", "--filter=2"), "butterfly_srcnn.png")])
def test_harness_on_the_reference_library_reproduces_goldens(tmp_path, args, gold):
    """the harness only uses the public API, so linked against the compiled reference it must reproduce the reference's
    own Pictures/butterfly_srcnn*.png exactly -- this pins the harness and the PNG codec, not the GPU path"""
    shutil.copy(G / "butterfly.png", tmp_path / "butterfly.png")
    r = _run(REF_CLI, tmp_path / "butterfly.png", *args)
    assert r.returncode == 0 and "Test Ok" in r.stdout, r.stdout + r.stderr
    out = np.asarray(Image.open(tmp_path / "butterfly_resized.png"))
    conv = np.asarray(Image.open(tmp_path / "butterfly_convolution.png"))
    assert np.array_equal(out, np.asarray(Image.open(G / gold).convert("RGB")))
    assert np.array_equal(conv, np.asarray(Image.open(G / "butterfly_srcnn_convolution.png").convert("L")))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "tensor"])
def test_cli_on_b200_reproduces_goldens(tmp_path, mode):
    _need_cli()
    shutil.copy(G / "butterfly.png", tmp_path / "butterfly.png")
    r = _run(CLI, tmp_path / "butterfly.png", env={"SRCNN_B200_CONV": mode})
    assert r.returncode == 0 and "Test Ok" in r.stdout, r.stdout + r.stderr
    out = np.asarray(Image.open(tmp_path / "butterfly_resized.png")).astype(int)
    conv = np.asarray(Image.open(tmp_path / "butterfly_convolution.png")).astype(int)
    gold = np.asarray(Image.open(G / "butterfly_srcnn.png").convert("RGB")).astype(int)
    gconv = np.asarray(Image.open(G / "butterfly_srcnn_convolution.png").convert("L")).astype(int)
    lim = 0 if mode == "exact" else 1
    assert np.abs(out - gold).max() <= lim and np.abs(conv - gconv).max() <= lim
    assert O.psnr(out, gold) >= 50.0


@pytest.mark.gpu
def test_cli_step_scale_and_filters_match_oracle(tmp_path):
    """--step --scale=3 --filter=3 on an RGBA PNG against the oracle's ProcessSRCNN (step chain, Lanczos-3, alpha path)"""
    _need_cli()
    img = O.synth("smooth", 40, 52, 4, 21)
    Image.fromarray(img).save(tmp_path / "a.png")
    r = _run(CLI, tmp_path / "a.png", "--scale=3", "--step", "--filter=3", tmp_path / "o.png", env={"SRCNN_B200_CONV": "exact"})
    assert r.returncode == 0, r.stdout + r.stderr
    rc, want, want_conv = O.best().process(img, 3.0, 3, True, True)
    assert rc == 0
    got = np.asarray(Image.open(tmp_path / "o.png"))
    assert got.shape[2] == 4 and np.array_equal(got.reshape(-1), want)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "a_convolution.png")).reshape(-1), want_conv)
