#!/usr/bin/env python3
"""Sweep of the host-pipeline knobs (sub-strip size, staging, copy threads) for the end-to-end call; GPU box only.
    python tools/e2e_probe.py [c2|c5]
Prints one line per setting: ms per call through srcnn_b200_pass_host with pinned buffers, with pageable buffers,
and through the C++ drop-in ProcessSRCNN (new[] outputs)."""
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import libsrcnn_b200 as S      # noqa: E402
import oracle_lib as O         # noqa: E402

WL = {"c2": (1080, 1920, 3, 2.0, 2), "c5": (2160, 3840, 3, 1.5, 4), "c4": (512, 512, 3, 2.0, 3)}


def timeit(fn, n=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    h, w, d, m, f = WL[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    img = O.synth("uniform", h, w, d, 42)
    ow, oh = S.output_dims(w, h, m)
    ref = None
    for strip_mb, stage, threads in [(1000, 1, 4), (8, 1, 4), (4, 1, 4), (3, 1, 4), (2, 1, 4), (1.5, 1, 4), (1, 1, 4),
                                     (3, 0, 1), (3, 1, 1), (3, 1, 2), (3, 1, 6), (3, 1, 8)]:
        os.environ["SRCNN_B200_STRIP_MB"] = str(strip_mb)
        os.environ["SRCNN_B200_STAGE"] = str(stage)
        os.environ["SRCNN_B200_COPY_THREADS"] = str(threads)
        ctx = S.Context(0)
        h_src = S.pinned_empty((h, w, d)); h_src[...] = img
        h_out = S.pinned_empty((oh, ow, d))
        p_out = np.empty((oh, ow, d), np.uint8)
        t_pin = timeit(lambda: ctx.pass_host(h_src, m, f, False, out=h_out))
        t_page = timeit(lambda: ctx.pass_host(img, m, f, False, out=p_out))
        if ref is None:
            ref = h_out.copy()
        ok = np.array_equal(h_out, ref) and np.array_equal(p_out, ref)
        ctx.close()
        print("strip_mb=%-6s stage=%d threads=%d : pinned %.3f ms (%.0f MP/s)  pageable %.3f ms (%.0f MP/s)  identical=%s"
              % (strip_mb, stage, threads, t_pin, ow * oh / t_pin / 1e3, t_page, ow * oh / t_page / 1e3, ok), flush=True)
    import subprocess
    for strip_mb in (1000, 3):
        env = dict(os.environ, SRCNN_B200_STRIP_MB=str(strip_mb), SRCNN_B200_STAGE="1", SRCNN_B200_COPY_THREADS="4")
        subprocess.run([sys.executable, __file__, sys.argv[1] if len(sys.argv) > 1 else "c2", "dropin"], env=env)


def dropin():
    h, w, d, m, f = WL[sys.argv[1]]
    img = O.synth("uniform", h, w, d, 42)
    ow, oh = S.output_dims(w, h, m)
    t = timeit(lambda: S.process_srcnn(img, m, f, False, False, copy=False), 10)
    print("drop-in ProcessSRCNN strip_mb=%s : %.3f ms (%.0f MP/s)" % (os.environ.get("SRCNN_B200_STRIP_MB"), t, ow * oh / t / 1e3), flush=True)


if __name__ == "__main__":
    dropin() if len(sys.argv) > 2 and sys.argv[2] == "dropin" else main()
