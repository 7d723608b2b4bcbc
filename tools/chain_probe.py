#!/usr/bin/env python3
"""Step-scale chains (factor > 2, reference src/libsrcnn.cpp:980-1061): does the chained result stay within 1 LSB / 50 dB of the
reference when the INTERMEDIATE x2 passes also run on the tcgen05 kernel (SRCNN_B200_CHAIN=tensor) instead of the bit-exact
CUDA-core kernel?  Prints max |diff|, the share of differing bytes and the PSNR per case for both settings, plus the time of a
x4 chain to 4K (960x540 -> 3840x2160).  Run on a B200:  python tools/chain_probe.py"""
import os, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import libsrcnn_b200 as S
import oracle_lib as O

CASES = [("uniform", 96, 128, 3, 3.0, S.BICUBIC), ("uniform", 96, 128, 4, 4.0, S.BICUBIC), ("binary", 80, 100, 3, 4.0, S.LANCZOS3),
         ("smooth", 120, 160, 3, 3.0, S.BSPLINE), ("uniform", 64, 64, 3, 6.0, S.BICUBIC), ("checker", 64, 80, 3, 4.0, S.BILINEAR),
         ("uniform", 270, 480, 3, 4.0, S.BICUBIC), ("uniform", 128, 128, 3, 8.0, S.LANCZOS3)]


def main():
    for chain in ("exact", "tensor"):
        os.environ["SRCNN_B200_CHAIN"] = chain
        ctx = S.Context(0, S.CONV_TENSOR)
        os.environ.pop("SRCNN_B200_CHAIN")
        for kind, h, w, d, m, f in CASES:
            img = O.synth(kind, h, w, d, 5)
            rc_o, want, wconv = O.best().process(img, m, f, True, True)
            rc, got, gconv = ctx.process(img, m, f, True, True)
            assert rc == rc_o == 0
            diff = np.abs(got.reshape(-1).astype(int) - want.astype(int))
            assert (gconv is None) == (wconv is None)
            dc = np.abs(gconv.reshape(-1).astype(int) - wconv.astype(int)) if wconv is not None else np.zeros(1, int)
            print("chain=%-6s %-8s %3dx%-3d d=%d x%.0f f=%d : max|diff| %d LSB (%.4f %% of bytes differ, %.5f %% by > 1), PSNR %.2f dB, conv max %d"
                  % (chain, kind, w, h, d, m, f, diff.max(), 100.0 * (diff > 0).mean(), 100.0 * (diff > 1).mean(),
                     O.psnr(got.reshape(-1), want), dc.max()), flush=True)
        img = O.synth("uniform", 540, 960, 3, 6)
        ctx.process(img, 4.0, S.BICUBIC, True, False)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.process(img, 4.0, S.BICUBIC, True, False)
        dt = (time.perf_counter() - t0) / 5
        print("chain=%-6s x4 chain 960x540 -> 3840x2160 (pageable host buffers): %.2f ms per call = %.1f output MP/s"
              % (chain, dt * 1e3, 3840 * 2160 / dt / 1e6), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
