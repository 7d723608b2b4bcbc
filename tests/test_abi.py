"""CPU tests (-m "not gpu") of the boundary: the shared library loads, exports every symbol the headers declare,
its host-side logic (size arithmetic, step-scale schedule, contribution tables, argument validation) matches the
oracle, and -- with no GPU in the box -- the product refuses to run instead of falling back to a CPU path."""
import ctypes as C
import re
import numpy as np
import pytest

import libsrcnn_b200 as S
import oracle_lib as O

HDR = (O.ROOT / "include" / "srcnn_b200.h").read_text()


def _no_gpu():
    return S.lib().srcnn_b200_device_count() == 0


def test_library_exports_all_declared_symbols():
    L = S.lib()
    names = sorted(set(re.findall(r"\b(srcnn_b200_[a-z0-9_]+)\s*\(", HDR)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    # the C++ drop-in API keeps the reference's mangled names (reference src/libsrcnn.h:46-54)
    assert hasattr(L, "_Z12ProcessSRCNNPKhjjjfRPhRjPS1_Pj")
    assert hasattr(L, "_Z20ConfigureFilterSRCNN15SRCNNFilterTypeb")
    assert L.srcnn_b200_abi_version() == 1


def test_public_header_is_source_compatible(tmp_path):
    """a caller written against the reference header compiles and links against include/libsrcnn.h + our .so"""
    import subprocess
    src = tmp_path / "caller.cpp"
    src.write_text('#include "libsrcnn.h"\n#include <cstdio>\n'
                   'int main(){ unsigned char px[12]={0}; unsigned char* o=nullptr; unsigned osz=0;\n'
                   ' unsigned char* c=nullptr; unsigned csz=0; ConfigureFilterSRCNN(SRCNNF_Lanczos3, true);\n'
                   ' ConfigureFilterSRCNN(SRCNNF_Bicubic);\n'
                   ' int rc = ProcessSRCNN(nullptr,2,2,3,2.0f,o,osz,&c,&csz); std::printf("%d %x\\n", rc, LIBSRCNN_VERSION);\n'
                   ' rc = ProcessSRCNN(px,2,2,3,-1.0f,o,osz,nullptr,nullptr); std::printf("%d\\n", rc); return 0; }\n')
    exe = tmp_path / "caller"
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-I", str(O.ROOT / "include"), str(src), "-o", str(exe),
                    "-L", str(S.LIB_PATH.parent), "-lsrcnn", "-Wl,-rpath," + str(S.LIB_PATH.parent)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["-1", "10a28", "-2"]


@pytest.mark.parametrize("w,h,m", [(256, 256, 2.0), (1920, 1080, 2.0), (3840, 2160, 1.5), (7, 5, 1.3), (100, 37, 2.7),
                                    (333, 77, 0.5), (512, 512, 3.999)])
def test_output_dims_match_oracle(w, h, m):
    assert S.output_dims(w, h, m) == O.out_dims(w, h, m)


@pytest.mark.parametrize("m", [1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 5.0, 0.5, 2.0000002, 7.3])
def test_step_schedule_matches_oracle(m):
    """final size and 'nothing produced' of the step-scale chain equal the oracle's (reference :980-1061)"""
    w, h = 20, 12
    img = O.synth("uniform", h, w, 3, 0)
    rc, out, _ = O.best().process(img, m, filt=1, stepscale=True, want_conv=False)
    n, ow, oh = S.process_dims(w, h, m, True)
    if rc == -100:
        assert n == 0
    else:
        assert rc == 0 and n >= 1 and out.size == ow * oh * 3


@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("dst,src", [(512, 256), (3840, 1920), (97, 64), (1024, 512), (33, 77), (300, 299), (8, 3), (5760, 3840)])
def test_contrib_table_matches_oracle(filt, dst, src):
    La, Ra, Wa = S.contrib_table(filt, dst, src)
    Lb, Rb, Wb = O.best().table(filt, dst, src)
    assert np.array_equal(La, Lb) and np.array_equal(Ra, Rb) and np.array_equal(Wa, Wb)


@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("dst,src", [(512, 256), (3840, 1920), (97, 64), (1024, 512), (300, 299), (8, 3), (5760, 3840), (1024, 256)])
def test_compact_table_is_the_reference_table_minus_exact_zeros(filt, dst, src):
    """the table the fused front-end kernel reads: same weights at the same source indices as the oracle's table,
    only taps whose weight is exactly 0.0 are missing (they add +-0 to the double accumulator: no effect)"""
    cwin, cl, cw, exact29 = S.contrib_table_compact(filt, dst, src)
    Lb, Rb, Wb = O.best().table(filt, dst, src)
    if cwin == 0:
        assert src < 8
        return
    assert cwin % 2 == 0 and cwin <= Wb.shape[1]
    assert (cl >= 0).all() and (cl + cwin <= src).all()
    for u in range(dst):
        full = np.zeros(src); full[Lb[u]:Rb[u] + 1] = Wb[u, :Rb[u] - Lb[u] + 1]
        comp = np.zeros(src); comp[cl[u]:cl[u] + cwin] = cw[u, :cwin]
        assert np.array_equal(full, comp)
    # <= 29 significant bits <=> the low 24 mantissa bits of every weight are clear
    bits = cw[:, :cwin].copy().view(np.uint64)
    assert exact29 == bool(((bits & np.uint64((1 << 24) - 1)) == 0).all())
    if filt == S.BILINEAR and dst == 2 * src:
        assert exact29          # 0.25 / 0.75: the fused multiply-add shortcut applies to the chroma planes at x2
    # pow2pair (FP32 chroma path of the fused front end): two taps, FP32-exact, one of each pair 0 or a power of two
    flags = S.contrib_table_compact(filt, dst, src, flags=True)[3]
    w2 = cw[:, :cwin]
    man = lambda a: (a.copy().view(np.uint64) & np.uint64((1 << 52) - 1)) == 0          # 0 or a power of two
    want = cwin == 2 and bool((w2.astype(np.float32).astype(np.float64) == w2).all()) and bool((w2 >= 0).all()) \
        and bool((man(w2[:, 0]) | man(w2[:, 1])).all())
    assert bool(flags & 2) == want
    if filt in (S.BILINEAR, S.NEAREST) and dst == 2 * src:
        assert flags & 2
    if filt == S.BILINEAR and (dst, src) == (5760, 3840):
        assert not flags & 2    # x1.5: thirds are not FP32-exact


def test_argument_validation_needs_no_gpu():
    img = O.synth("uniform", 8, 8, 3, 0)
    assert S.process_srcnn(img, 0.0)[0] == -2
    assert S.process_srcnn(img, -2.0)[0] == -2
    assert S.process_srcnn(img, 1.0, stepscale=True)[0] == -100          # reference :1007-1008
    L = S.lib()
    o, osz = C.c_void_p(), C.c_uint()
    assert L.ProcessSRCNN(None, 8, 8, 3, C.c_float(2.0), C.byref(o), C.byref(osz), None, None) == -1
    assert L.ProcessSRCNN(C.c_void_p(img.ctypes.data), 0, 8, 3, C.c_float(2.0), C.byref(o), C.byref(osz), None, None) == -1
    assert L.ProcessSRCNN(C.c_void_p(img.ctypes.data), 8, 8, 2, C.c_float(2.0), C.byref(o), C.byref(osz), None, None) == -1
    S.lib().ConfigureFilterSRCNN(S.BICUBIC, False)


def test_no_cpu_fallback_without_gpu():
    if not _no_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(S.SrcnnError) as e:
        S.Context(0)
    assert e.value.code == S.ENODEV
    img = O.synth("uniform", 8, 8, 3, 0)
    rc, out, conv = S.process_srcnn(img, 2.0)
    assert rc == S.ENODEV and out is None and conv is None
