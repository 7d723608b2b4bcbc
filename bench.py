#!/usr/bin/env python3
"""bench.py -- throughput of the SRCNN hot path (ProcessSRCNN: colour-in -> resize -> fused conv I/II/III -> colour-out).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A step = one pass of the hot path over one synthetic frame per GPU (BASELINE.json configs[1]: 1920x1080 RGB x2.0
bicubic -> 3840x2160).  Frames are independent units, so N GPUs process N frames per step with no collective
(weak scaling); torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the timing.
Other workloads (not the headline; BASELINE configs 3-5): c3 = ONE 7680x4320 RGBA image cut into row strips with a
6-row halo, one strip per rank (strong scaling); c4 = 512x512 Lanczos-3 frames, a batch per rank; c5 = x1.5 B-spline
with the convolution plane.

Printed JSON (one line, rank 0):
  value     whole-job output megapixels/s, inputs resident in HBM, CUDA-event timed per step, L2 flushed between steps
            (the per-KERNEL events behind `roofline` / `stage_ms` are recorded on every 4th timed step only: they are
            instrumentation, and each one costs a few microseconds of GPU idle time between two kernels)
  e2e       same metric through the C-ABI call srcnn_b200_pass_host with pinned HOST buffers (H2D + kernels + D2H)
  roofline  the fused-conv kernel: algorithmic 16 064 flop per output pixel / its CUDA-event duration, against the
            measured dense bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU code (oracle/_ref, OpenMP, all host cores) on a bounded sample of the workload
`--impl reference` times that CPU implementation alone and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

NCU_CONV_DRAM_BYTES = 17243904   # conv_tc_kernel at C2: 17.172 MB read + 0.072 MB written (profiles/r1_ncu_summary.md)
FLOP_PER_PX = 16064            # conv I 81*64 + conv II 64*32 + conv III 25*32 MACs, x2
WORKLOADS = {
    # name: (h, w, d, multiply, filter, description)
    "c2": (1080, 1920, 3, 2.0, 2, "synthetic 1920x1080 RGB x2.0 bicubic -> 3840x2160 (BASELINE configs[1])"),
    "c3": (4320, 7680, 4, 2.0, 2, "synthetic 7680x4320 RGBA x2.0 bicubic -> 15360x8640, row strips with 6-row halos (BASELINE configs[2])"),
    "c4": (512, 512, 3, 2.0, 3, "synthetic 512x512 RGB x2.0 Lanczos-3 -> 1024x1024 frames, batch per GPU (BASELINE configs[3])"),
    "c5": (2160, 3840, 3, 1.5, 4, "synthetic 3840x2160 RGB x1.5 B-spline -> 5760x3240, conv-output on (BASELINE configs[4])"),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    fb = dict(tf=1590.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")
    try:
        j = json.loads(p.read_text())
        # the conv kernel is timed alone between L2 flushes (a burst, not a long step): the burst figure applies
        tf = next(float(j[k]) for k in ("bf16_tflops", "bf16_tflops_burst", "bf16_tflops_sustained") if k in j)
        hbm = next((float(j[k]) for k in ("hbm_gbs", "hbm_gbs_burst") if k in j), fb["hbm"])
        if tf < 100.0:                       # a file in PFLOP/s rather than TFLOP/s
            tf *= 1000.0
        return dict(tf=tf, hbm=hbm, src="measured (MEASURED_PEAKS.json, burst bf16)")
    except Exception:                        # absent or unreadable: the documented fallback
        return fb


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms (the timed
    region of the default run lasts only tens of milliseconds, far below nvidia-smi's own sampling period)"""

    def __init__(self, index):
        self.rows, self.ok, self.run = [], False, True
        try:
            import pynvml as N
            N.nvmlInit()
            self.N, self.h = N, N.nvmlDeviceGetHandleByIndex(index)
            self.max = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()
        except Exception as e:                                              # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        N = self.N
        while self.run:
            try:
                self.rows.append((time.time(), float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)),
                                  int(N.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:                                               # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "?")]}
        self.run = False; self.t.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": ["no samples"]}
        N = self.N
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        names = [("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max,
                "reasons": [n for n, b in names if bits & b], "samples": len(rows), "source": "NVML, 2 ms period, timed region only"}


def cpu_reference_run(h, w, d, m, filt, conv_on, budget_s=12.0, check_ctx=None):
    """the reference's CPU implementation (oracle/_ref; the plain-C port if the reference was not compiled) on a
    bounded sample: a horizontal crop of the workload frame sized for a few seconds of CPU work"""
    import oracle_lib as O
    orc = O.best()
    sh = min(h, max(64, (270 * 1920) // max(w, 1)))            # ~0.5 M source px
    img = O.synth("uniform", sh, w, d, 42)
    ow, oh = O.out_dims(w, sh, m)
    t0 = time.perf_counter(); rc, out, _ = orc.process(img, m, filt, False, conv_on); dt = time.perf_counter() - t0
    assert rc == 0
    runs, best = 1, dt
    while runs < 3 and (time.perf_counter() - t0) + best < budget_s:
        t1 = time.perf_counter(); orc.process(img, m, filt, False, conv_on); best = min(best, time.perf_counter() - t1); runs += 1
    res = dict(value=ow * oh / best / 1e6, unit="output MP/s", cores=orc.threads,
               kind="reference" if orc.kind == "reference" else "port",
               sample="%dx%dx%d crop of the workload frame -> %dx%d, best of %d run(s), %.2f s each, OpenMP threads=%d"
                      % (w, sh, d, ow, oh, runs, best, orc.threads))
    if check_ctx is not None:
        # the parity gate of SURVEY 8(d), on the same sample: the GPU result of the crop against the CPU result just computed
        got, _ = check_ctx.pass_host(img, m, filt, conv_on)
        diff = np.abs(got.reshape(-1).astype(np.int16) - out.astype(np.int16))
        res["parity"] = {"max_abs_lsb": int(diff.max()), "psnr_db": round(float(O.psnr(got.reshape(-1), out)), 2),
                         "gate": "<= 1 LSB and >= 50 dB against the CPU result of the same sample",
                         "ok": bool(diff.max() <= 1 and O.psnr(got.reshape(-1), out) >= 50.0)}
    return res, best, ow * oh


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is the OpenMP CPU code with ALL host threads.
    # (Set before the OpenMP runtime is loaded: nothing below has imported torch or the oracle library yet.)
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    h, w, d, m, filt, desc = wl
    conv_on = args.workload == "c5"
    vals = []
    for i in range(args.warmup + args.steps):
        cb, dt, px = cpu_reference_run(h, w, d, m, filt, conv_on, budget_s=0.0)
        if i >= args.warmup:
            vals.append((cb, dt))
    cb = vals[0][0]
    tot = sum(dt for _, dt in vals)
    v = px * len(vals) / tot / 1e6
    cb["value"] = v
    emit({
        "impl": "reference", "metric": "output MP/s, SRCNN x2 to 4K (ProcessSRCNN hot path)", "value": v, "unit": "output MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(vals) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "output MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def emit(line):
    """the ONE JSON line goes to the process's original stdout; everything else that libraries write to fd 1 (NCCL prints
    its version banner there) has been diverted to stderr by quiet_stdout()"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--conv", default="tensor", choices=["tensor", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU and step (default 1; c4: 32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import libsrcnn_b200 as S
    import oracle_lib as O

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    h, w, d, m, filt, desc = wl
    conv_on = args.workload == "c5"
    strips = args.workload == "c3"                                  # one image, one row strip per rank (strong scaling)
    nfr = args.frames or (32 if args.workload == "c4" else 1)       # frames per rank and step (c4: a batch)
    if strips:
        nfr = 1
    ow, oh = S.output_dims(w, h, m)
    r0, r1 = (oh * rank // world, oh * (rank + 1) // world) if strips else (0, oh)
    out_px = (r1 - r0) * ow * nfr                                   # this rank's output pixels per step
    job_px = (ow * oh) if strips else ow * oh * nfr * world         # whole-job output pixels per step
    ctx = S.Context(local, S.CONV_TENSOR if args.conv == "tensor" else S.CONV_EXACT)
    imgs = np.stack([O.synth("uniform", h, w, d, (43 if strips else 42 + rank * nfr + i)) for i in range(nfr)])
    d_src = torch.from_numpy(imgs).cuda()
    d_out = torch.empty((nfr, r1 - r0, ow, d), dtype=torch.uint8, device="cuda")
    d_conv = torch.empty((nfr, r1 - r0, ow), dtype=torch.uint8, device="cuda") if conv_on else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    def step():
        if nfr > 1:          # a batch of frames: srcnn_b200_batch_device (small frames share one fused-conv launch per group)
            ctx.batch_device(d_src.data_ptr(), nfr, w, h, d, m, filt, d_out.data_ptr(), d_conv.data_ptr() if conv_on else 0)
            return
        for i in range(nfr):
            ctx.pass_device(d_src[i].data_ptr(), w, h, d, m, filt, d_out[i].data_ptr(),
                            d_conv[i].data_ptr() if conv_on else 0, rows=(r0, r1))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    ctx.sync()

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps
    # The per-kernel events (ctx timing) are instrumentation the product path does not have, and every event recorded between
    # two kernels costs a few microseconds of GPU idle time: they are switched on for every 4th timed step only (>= 5 samples);
    # the kernel durations reported below are the averages over those steps of the SAME timed region.
    probe_every = max(1, min(4, args.steps // 5))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    conv_ms, stage_ms = [], []
    launches0 = ctx.launches
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(i & 0xff)                                          # evict the previous step's lines from L2
        probed = (i % probe_every) == 0
        ctx.set_timing(probed)
        ev[i][0].record(stream)
        step()
        ev[i][1].record(stream)
        ctx.sync()                                                         # also collects the per-kernel events
        if probed:
            t = ctx.timings(); conv_ms.append(t["conv"]); stage_ms.append(t)   # (of the step's last frame / frame group)
            conv_frames = ctx.last_group
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = ctx.launches - launches0
    ctx.set_timing(False)
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = job_px * args.steps / (total_ms_max * 1e-3) / 1e6

    # ---- end to end: the C-ABI call with pinned host buffers (H2D + kernels + D2H inside the timed region)
    h_src = S.pinned_empty((nfr, h, w, d)); h_src[...] = imgs
    h_out = S.pinned_empty((nfr, r1 - r0, ow, d)); h_conv = S.pinned_empty((nfr, r1 - r0, ow)) if conv_on else None

    def e2e_step():
        if nfr > 1:
            ctx.batch(h_src, m, filt, conv_on, out=h_out, conv=h_conv)
        else:
            ctx.pass_host(h_src[0], m, filt, conv_on, rows=(r0, r1), out=h_out[0], conv=h_conv[0] if conv_on else None)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = job_px * args.steps / float(te.item()) / 1e6
    # the drop-in C++ call itself (pageable input, new[] output released with delete[]), rank 0, for reference
    dropin = None
    if rank == 0 and not strips:
        S.process_srcnn(imgs[0], m, filt, False, conv_on, copy=False)
        t0 = time.perf_counter(); n = max(3, args.steps // 4)
        for _ in range(n):
            rc, _, _ = S.process_srcnn(imgs[0], m, filt, False, conv_on, copy=False)
        dropin = ow * oh * n / (time.perf_counter() - t0) / 1e6 if rc == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    if strips:                                                     # source rows each rank uploads (strip + halo), summed
        h2d_bytes = sum((lambda g: (g[1][1] - g[1][0]) * w * d)(S.strip_geometry(w, h, d, m, filt, oh * k // world, oh * (k + 1) // world))
                        for k in range(world))
    else:
        h2d_bytes = int(imgs.nbytes)
    conv_avg_ms = float(np.mean(conv_ms))
    conv_px = (r1 - r0) * ow                                       # one launch = one frame (or strip) of this rank
    conv_px *= conv_frames                                         # batches: the step's last conv launch covered a group of frames
    achieved_tf = conv_px * FLOP_PER_PX / (conv_avg_ms * 1e-3) / 1e12
    stage_avg = {k: float(np.mean([s[k] for s in stage_ms])) for k in stage_ms[0]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_reference_run(h, w, d, m, filt, conv_on, check_ctx=ctx)
    line = {
        "metric": "output MP/s, SRCNN x2 to 4K (ProcessSRCNN hot path)", "value": value, "unit": "output MP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
        "higher_is_better": True, "scaling": "strong" if strips else "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (conv), f64 (resize), f32 (colour)" if args.conv == "tensor" else "f32 (conv exact mode), f64 (resize)",
        "data": "synthetic",
        "config": {"workload": desc, "frames_per_step": 1 if strips else world * nfr,
                   "sharding": ("row strips [oh*g/N, oh*(g+1)/N) with a recomputed 6-row halo, no collective" if strips
                                else "%d frame(s) per GPU and step, no collective" % nfr),
                   "conv_mode": args.conv, "l2": "256 MiB buffer written between timed steps (L2 flush)",
                   "kernel_events": "per-kernel CUDA events on every %s timed step (%d samples)" % (
                       {1: "", 2: "2nd", 3: "3rd", 4: "4th"}[probe_every], len(conv_ms)),
                   "conv_output": conv_on},
        "e2e": {"value": e2e_val, "unit": "output MP/s",
                "h2d_bytes_per_step": h2d_bytes * (1 if strips else world),
                "d2h_bytes_per_step": job_px * d + (job_px if conv_on else 0),
                "api": "srcnn_b200_%s (C-ABI, pinned host buffers)" % ("batch_host" if nfr > 1 else "pass_host_rows" if strips else "pass_host"),
                "dropin_ProcessSRCNN_pageable": dropin},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "fused conv I+II+III (%s)" % args.conv, "bound": "tensor", "achieved": achieved_tf,
                     "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf"],
                     "traffic": NCU_CONV_DRAM_BYTES if (args.workload == "c2" and args.conv == "tensor") else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture (profiles/r1_ncu_summary.md); the 33 MB Y' write stays in L2",
                     "peak_source": pk["src"], "kernel_ms": conv_avg_ms,
                     "algorithmic_flop_per_launch": conv_px * FLOP_PER_PX},
        "stage_ms": stage_avg,
        "cpu_baseline": cpu,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
