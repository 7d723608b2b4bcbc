"""Development aid: conv-kernel time of the C2 workload (4K frame) for the library named by SRCNN_B200_LIB.
usage: [SRCNN_B200_LIB=path] python tools/conv_time.py [steps]   (run on the GPU box)"""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch
import libsrcnn_b200 as S
import oracle_lib as O
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
h, w, d, m = 1080, 1920, 3, 2.0
ow, oh = S.output_dims(w, h, m)
ctx = S.Context(0, S.CONV_TENSOR)
img = O.synth("uniform", h, w, d, 42)
d_src = torch.from_numpy(img).cuda()
d_out = torch.empty((oh, ow, d), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
for _ in range(5):
    ctx.pass_device(d_src.data_ptr(), w, h, d, m, S.BICUBIC, d_out.data_ptr(), 0)
ctx.sync(); ctx.set_timing(True)
ts = []
for i in range(steps):
    with torch.cuda.stream(stream):
        flush.fill_(i & 0xff)
    ctx.pass_device(d_src.data_ptr(), w, h, d, m, S.BICUBIC, d_out.data_ptr(), 0)
    ctx.sync(); ts.append(ctx.timings()["conv"])
ts = np.array(ts)
print("%-40s conv ms: mean %.4f  min %.4f  max %.4f  (checksum %d)" % (str(S.LIB_PATH)[-40:], ts.mean(), ts.min(), ts.max(), int(d_out.sum().item())))
