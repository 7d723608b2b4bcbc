import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import libsrcnn_b200 as S
import oracle_lib as O
h, w, d, m = 540, 960, 3, 2.0
ow, oh = S.output_dims(w, h, m)
ctx = S.Context(0, S.CONV_EXACT)
img = O.synth("uniform", h, w, d, 42)
d_src = torch.from_numpy(img).cuda()
d_out = torch.empty((oh, ow, d), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.pass_device(d_src.data_ptr(), w, h, d, m, S.BICUBIC, d_out.data_ptr(), 0)
ctx.sync(); ctx.set_timing(True)
ts = []
for i in range(5):
    ctx.pass_device(d_src.data_ptr(), w, h, d, m, S.BICUBIC, d_out.data_ptr(), 0)
    ctx.sync(); ts.append(ctx.timings()["conv"])
print("%s exact conv 1920x1080: %.3f ms (checksum %d)" % (str(S.LIB_PATH)[-24:], float(np.mean(ts)), int(d_out.sum().item())))
