"""Development aid: front-end / conv / colour-out kernel times of the C2 and C4-group workloads for the library named by SRCNN_B200_LIB.
usage: [SRCNN_B200_LIB=path] python tools/prep_time.py   (GPU box)"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import libsrcnn_b200 as S
import oracle_lib as O
ctx = S.Context(0, S.CONV_TENSOR)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
for name, (h, w, d, f, nfr) in {"c2": (1080, 1920, 3, S.BICUBIC, 1), "c3": (4320, 7680, 4, S.BICUBIC, 1), "c4x16": (512, 512, 3, S.LANCZOS3, 16),
                                "c5": (2160, 3840, 3, S.BSPLINE, 1)}.items():
    mult = 1.5 if name == "c5" else 2.0
    ow, oh = S.output_dims(w, h, mult)
    d_src = torch.from_numpy(np.stack([O.synth("uniform", h, w, d, 42 + i) for i in range(nfr)])).cuda()
    d_out = torch.empty((nfr, oh, ow, d), dtype=torch.uint8, device="cuda")
    def step():
        if nfr > 1: ctx.batch_device(d_src.data_ptr(), nfr, w, h, d, mult, f, d_out.data_ptr(), 0)
        else: ctx.pass_device(d_src.data_ptr(), w, h, d, mult, f, d_out.data_ptr(), 0)
    for _ in range(4): step()
    ctx.sync(); ctx.set_timing(True); ts = []
    for i in range(20):
        with torch.cuda.stream(stream): flush.fill_(i & 0xff)
        step(); ctx.sync(); ts.append(ctx.timings())
    ctx.set_timing(False)
    m = {k: float(np.mean([t[k] for t in ts])) * 1e3 for k in ("color_in", "conv", "color_out", "total")}
    print("%-28s %-6s front end %7.1f us  conv %7.1f us  colour-out %7.1f us  total %7.1f us  (checksum %d)"
          % (str(S.LIB_PATH)[-28:], name, m["color_in"], m["conv"], m["color_out"], m["total"], int(d_out.sum().item())), flush=True)
