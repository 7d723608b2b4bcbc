"""Development aid: where do row-strip results differ from the full-image result? (run on the GPU box)"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import libsrcnn_b200 as S
import oracle_lib as O
ctx = S.Context(0, S.CONV_TENSOR)
img = O.synth("uniform", 101, 150, 4, 50)
full, fconv = ctx.pass_host(img, 2.0, S.BICUBIC, True)
full2, fconv2 = ctx.pass_host(img, 2.0, S.BICUBIC, True)
print("repeatable:", np.array_equal(full, full2), np.array_equal(fconv, fconv2))
oh = full.shape[0]
for cuts in ([0, 101, 202], [0, 100, 202], [0, 67, 135, 202]):
    for a, b in zip(cuts[:-1], cuts[1:]):
        o, c = ctx.pass_host(img, 2.0, S.BICUBIC, True, rows=(a, b))
        d = (c.astype(int) - fconv[a:b].astype(int))
        rows = np.unique(np.nonzero(d)[0]) + a
        cols = np.unique(np.nonzero(d)[1])
        print("strip", a, b, "conv diffs:", int((d != 0).sum()), "max", int(np.abs(d).max()), "rows", rows[:20], "cols", cols[:20])
