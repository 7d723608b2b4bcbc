#!/usr/bin/env python3
"""bench.py -- throughput of the SRCNN hot path (ProcessSRCNN: colour-in -> resize -> fused conv I/II/III -> colour-out).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4|c5] [--no-sub]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1], "c2"): a step = one pass of the hot path over one synthetic 1920x1080 RGB
frame per GPU, x2.0 bicubic -> 3840x2160.  Frames are independent units, so N GPUs process N frames per step with no
collective (weak scaling); torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the timing.

Printed JSON (ONE line, rank 0):
  value          whole-job output megapixels/s, inputs resident in HBM, CUDA events around every step on the launching stream,
                 L2 flushed between steps, max over ranks
  e2e            same metric through the C-ABI call srcnn_b200_pass_host with pinned HOST buffers (H2D + kernels + D2H timed);
                 e2e.dropin_ProcessSRCNN_pageable = the C++ drop-in call itself (pageable input, new[] output)
  roofline       the fused-conv kernel: algorithmic 16 064 flop per output pixel / its CUDA-event duration, against the
                 measured dense bf16 peak in MEASURED_PEAKS.json
  roofline_mem   the two memory-side kernels of the step (front end, colour-out): algorithmic bytes / CUDA-event duration
                 against the measured HBM copy bandwidth
  cpu_baseline   the reference's own CPU code (oracle/_ref, OpenMP, all host cores) on the WHOLE workload frame, with the
                 parity gate (GPU result vs that CPU result)
  sub            the sharded workloads of BASELINE.json, measured the same way in the same run (not the headline):
                   c3 = ONE 7680x4320 RGBA image as N row strips with 6-row halos (strong scaling)
                   c4 = 4096 frames 512x512 RGB Lanczos-3 in total, 4096/N per GPU (batch-sharded)
`--impl reference` times the reference CPU implementation alone (rank 0) on the same config and prints the same line
with "impl": "reference".
"""
import argparse
import json
import os
import pathlib
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FLOP_PER_PX = 16064            # conv I 81*64 + conv II 64*32 + conv III 25*32 MACs, x2
METRIC = "output MP/s, SRCNN x2 to 4K (ProcessSRCNN hot path)"
WORKLOADS = {
    # name: (h, w, d, multiply, filter, description)
    "c2": (1080, 1920, 3, 2.0, 2, "synthetic 1920x1080 RGB x2.0 bicubic -> 3840x2160 (BASELINE configs[1])"),
    "c3": (4320, 7680, 4, 2.0, 2, "synthetic 7680x4320 RGBA x2.0 bicubic -> 15360x8640, row strips with 6-row halos (BASELINE configs[2])"),
    "c4": (512, 512, 3, 2.0, 3, "synthetic 512x512 RGB x2.0 Lanczos-3 -> 1024x1024 frames, batch-sharded (BASELINE configs[3])"),
    "c5": (2160, 3840, 3, 1.5, 4, "synthetic 3840x2160 RGB x1.5 B-spline -> 5760x3240, step-scale and conv-output on (BASELINE configs[4])"),
}
C4_TOTAL_FRAMES = 4096         # north_star: "batch of 4096 synthetic 512x512 RGB frames", split over the ranks
C4_DISTINCT = 64               # distinct synthetic frames generated per rank (tiled on the device up to the batch size)
C4_E2E_CHUNK = 256             # frames per srcnn_b200_batch_host call of the c4 end-to-end leg (one pinned buffer set)


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
        return dict(tf=tf, hbm=hbm, src="measured (MEASURED_PEAKS.json: burst bf16 / copy bandwidth)")
    except Exception:                        # absent or unreadable: the documented fallback
        return fb


def config_of(name, world, nfr):
    """the `config` object of the JSON line -- identical for both arms (--impl b200 / reference) of one command line"""
    h, w, d, m, filt, desc = WORKLOADS[name]
    strips = name == "c3"
    return {"workload": desc, "frames_per_step": 1 if strips else world * nfr,
            "sharding": ("row strips [oh*g/N, oh*(g+1)/N) with a recomputed 6-row halo, no collective" if strips
                         else "%d frame(s) per GPU and step, no collective" % nfr),
            "conv_output": name == "c5",
            "l2": "256 MiB buffer written between timed steps (L2 flush)"}


def frames_per_rank(name, world, arg_frames):
    if name == "c3":
        return 1
    if arg_frames:
        return arg_frames
    return max(1, C4_TOTAL_FRAMES // world) if name == "c4" else 1


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


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_frames(name, count):
    import oracle_lib as O
    h, w, d, m, filt, _ = WORKLOADS[name]
    seed = {"c2": 42, "c3": 43, "c4": 44, "c5": 45}[name]
    return [O.synth("uniform", h, w, d, seed + i) for i in range(count)]


def cpu_run_once(name, img):
    """ONE whole workload unit (a full frame / the full image) through the reference's CPU ProcessSRCNN -> (seconds, out, conv)"""
    import oracle_lib as O
    h, w, d, m, filt, _ = WORKLOADS[name]
    step, conv_on = name == "c5", name == "c5"
    t0 = time.perf_counter()
    rc, out, conv = O.best().process(img, m, filt, step, conv_on)
    dt = time.perf_counter() - t0
    assert rc == 0
    return dt, out, conv


def cpu_baseline_leg(name, ctx, budget_s=25.0):
    """the CPU baseline beside the GPU number: the reference's implementation on the WHOLE workload frame (1 warm-up, then
    best of up to 3 runs inside the time budget), plus the parity gate of SURVEY 8(d) on that full frame"""
    import oracle_lib as O
    import libsrcnn_b200 as S
    h, w, d, m, filt, _ = WORKLOADS[name]
    orc = O.best()
    img = cpu_frames(name, 1)[0]
    ow, oh = O.out_dims(w, h, m)
    t_start = time.perf_counter()
    dt, out, conv = cpu_run_once(name, img)                       # warm-up (page faults of the 64 + 32 planes) -- also the checker
    runs, best = 0, None
    while runs < 3 and (runs == 0 or (time.perf_counter() - t_start) + best < budget_s):
        dt, _, _ = cpu_run_once(name, img)
        best = dt if best is None else min(best, dt); runs += 1
    res = dict(value=ow * oh / best / 1e6, unit="output MP/s", cores=orc.threads,
               kind="reference" if orc.kind == "reference" else "port",
               sample="the whole workload frame %dx%dx%d -> %dx%d (same config as the GPU arm), 1 warm-up + best of %d run(s), "
                      "%.2f s each, OpenMP threads=%d" % (w, h, d, ow, oh, runs, best, orc.threads))
    if ctx is not None:
        rc, got, gconv = ctx.process(img, m, filt, name == "c5", name == "c5")
        diff = np.abs(got.reshape(-1).astype(np.int16) - out.astype(np.int16))
        psnr = float(O.psnr(got.reshape(-1), out))
        ok = bool(rc == 0 and diff.max() <= 1 and psnr >= 50.0)
        par = {"max_abs_lsb": int(diff.max()), "psnr_db": round(psnr, 2)}
        if conv is not None:
            dc = np.abs(gconv.reshape(-1).astype(np.int16) - conv.astype(np.int16))
            par["conv_max_abs_lsb"] = int(dc.max()); ok = ok and bool(dc.max() <= 1)
        # the interpolated Y plane of the product route (FP32 copy written on request) against the reference's resize: 1e-5 relative
        try:
            import torch                                     # (one device-resident pass over the whole frame: a host pass is cut
            d_src = torch.from_numpy(img).cuda()             #  into sub-strips, and the plane getters describe the last one)
            d_out = torch.empty((oh, ow, d), dtype=torch.uint8, device="cuda")
            ctx.set_keep_planes(2)
            ctx.pass_device(d_src.data_ptr(), w, h, d, m, filt, d_out.data_ptr(), 0)
            ctx.sync()
            r0, gy = ctx.get_plane(0, (oh, ow))
            ctx.set_keep_planes(False)
            if r0 != 0 or gy.shape != (oh, ow):
                raise RuntimeError("plane getter returned rows from %d, shape %r" % (r0, gy.shape))
            ry = orc.resize(orc.color_in(img)[0], ow, oh, filt)
            rel = float((np.abs(gy - ry) / np.maximum(np.abs(ry), 1.0)).max())
            par["y_plane_rel_max"] = rel; par["y_plane_differing_samples"] = int((gy != ry).sum())
            ok = ok and rel <= 1e-5
        except Exception as e:                                              # noqa: BLE001
            par["y_plane_rel_max"] = None; par["y_plane_note"] = repr(e)
        par.update(gate="final image (and conv plane) <= 1 LSB and >= 50 dB, interpolated Y plane <= 1e-5 relative: GPU (tensor path) "
                        "against the CPU result of the same full frame", ok=ok)
        res["parity"] = par
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is the OpenMP CPU code with ALL host threads.
    # (Set before the OpenMP runtime is loaded: nothing below has imported torch or the oracle library yet.)
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import oracle_lib as O
    name = args.workload
    h, w, d, m, filt, desc = WORKLOADS[name]
    nfr = frames_per_rank(name, world, args.frames)
    ow, oh = O.out_dims(w, h, m)
    # a step of the CPU arm = ONE whole unit of the workload (a full frame; c3: the full image; c4: one frame of the batch --
    # its frames are identical in cost, the throughput per frame is what scales), every step on the whole unit, never a crop
    imgs = cpu_frames(name, 1)
    budget_s = 170.0                                              # the arm must end within a few minutes
    t_begin = time.perf_counter()
    times = []
    for i in range(args.warmup + args.steps):
        dt, _, _ = cpu_run_once(name, imgs[0])
        if i >= args.warmup:
            times.append(dt)
        # stop early rather than overrun the budget (slow hosts), but never without a timed step
        if times and (time.perf_counter() - t_begin) + dt > budget_s:
            break
        if not times and (time.perf_counter() - t_begin) + 2 * dt > budget_s:
            times.append(cpu_run_once(name, imgs[0])[0])
            break
    tot = sum(times)
    v = ow * oh * len(times) / tot / 1e6
    threads = O.best().threads
    cb = dict(value=v, unit="output MP/s", cores=threads, kind="reference" if O.best().kind == "reference" else "port",
              sample="the whole workload frame %dx%dx%d -> %dx%d per step (same config as the GPU arm), %d timed step(s)%s, "
                     "%.2f s each, OpenMP threads=%d" % (w, h, d, ow, oh, len(times),
                                                          "" if len(times) == args.steps else " (of %d requested: time budget %d s)" % (args.steps, budget_s),
                                                          tot / len(times), threads))
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "output MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3,
        "steps_timed": len(times),
        "higher_is_better": True, "scaling": "strong" if name == "c3" else "weak", "vs_baseline": None, "dtype": "f32 (conv), f64 (resize)",
        "data": "synthetic",
        "config": config_of(name, world, nfr),
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


# ------------------------------------------------------------------------------------------------- GPU arm
class Env:
    """what every measurement shares: rank / world, the context, the L2-flush buffer, barrier and max-over-ranks"""

    def __init__(self, conv_mode):
        import torch
        import libsrcnn_b200 as S
        self.torch, self.S = torch, S
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.ctx = S.Context(self.local, S.CONV_TENSOR if conv_mode == "tensor" else S.CONV_EXACT)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")        # > 126 MB L2
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def measure(env, name, steps, warmup, arg_frames=0, want_e2e=True, want_dropin=False, sample_clocks=False):
    """one workload, device-resident (CUDA events per step, L2 flushed between steps) and end to end (host buffers)"""
    import oracle_lib as O
    torch, S, ctx = env.torch, env.S, env.ctx
    rank, world = env.rank, env.world
    h, w, d, m, filt, desc = WORKLOADS[name]
    conv_on = name == "c5"
    strips = name == "c3"                                           # one image, one row strip per rank (strong scaling)
    nfr = frames_per_rank(name, world, arg_frames)                  # frames per rank and step (c4: its share of the batch)
    ow, oh = S.output_dims(w, h, m)
    r0, r1 = (oh * rank // world, oh * (rank + 1) // world) if strips else (0, oh)
    job_px = (ow * oh) if strips else ow * oh * nfr * world         # whole-job output pixels per step
    seed0 = {"c2": 42, "c3": 43, "c4": 44, "c5": 45}[name]
    ndist = min(nfr, C4_DISTINCT)
    imgs = np.stack([O.synth("uniform", h, w, d, (seed0 if strips else seed0 + rank * ndist + i)) for i in range(ndist)])
    d_src = torch.from_numpy(imgs).cuda()
    if nfr > ndist:                                                 # big batches: the distinct frames tiled on the device
        d_src = d_src.repeat((nfr + ndist - 1) // ndist, 1, 1, 1)[:nfr].contiguous()
    d_out = torch.empty((nfr, r1 - r0, ow, d), dtype=torch.uint8, device="cuda")
    d_conv = torch.empty((nfr, r1 - r0, ow), dtype=torch.uint8, device="cuda") if conv_on else None

    def step():
        if nfr > 1:          # a batch of frames: srcnn_b200_batch_device (small frames share one fused-conv launch per group)
            ctx.batch_device(d_src.data_ptr(), nfr, w, h, d, m, filt, d_out.data_ptr(), d_conv.data_ptr() if conv_on else 0)
            return
        ctx.pass_device(d_src[0].data_ptr(), w, h, d, m, filt, d_out[0].data_ptr(),
                        d_conv[0].data_ptr() if conv_on else 0, rows=(r0, r1))

    for _ in range(warmup):
        step()
    ctx.sync()

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between steps.
    # The per-kernel events (ctx timing) are instrumentation the product path does not have, and every event recorded between
    # two kernels costs a few microseconds of GPU idle time: they are switched on for every 4th timed step only (>= 5 samples
    # at the default step count); the kernel durations reported are the averages over those steps of the SAME timed region.
    probe_every = max(1, min(4, steps // 5))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    stage_ms = []
    conv_frames = 1
    launches0 = ctx.launches
    sampler = ClockSampler(env.local) if (rank == 0 and sample_clocks) else None
    env.barrier()
    t_wall0 = time.time()
    for i in range(steps):
        with torch.cuda.stream(env.stream):
            env.flush.fill_(i & 0xff)                                      # evict the previous step's lines from L2
        probed = (i % probe_every) == 0
        ctx.set_timing(probed)
        ev[i][0].record(env.stream)
        step()
        ev[i][1].record(env.stream)
        ctx.sync()                                                         # also collects the per-kernel events
        if probed:
            stage_ms.append(ctx.timings())                                 # (of the step's last frame / frame group)
            conv_frames = ctx.last_group
    env.barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = ctx.launches - launches0
    ctx.set_timing(False)
    total_ms_max = env.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    value = job_px * steps / (total_ms_max * 1e-3) / 1e6

    res = {"value": value, "ms_per_step": total_ms_max / steps, "steps": steps, "launches": launches, "clocks": clocks,
           "job_px": job_px, "nfr": nfr, "probe_every": probe_every}

    # ---- kernel rooflines of this rank's launches (rank 0 reports its own)
    stage_avg = {k: float(np.mean([s[k] for s in stage_ms])) for k in stage_ms[0]}
    unit_px = (r1 - r0) * ow * conv_frames                         # output pixels of one conv / colour-out launch
    src_rows = (S.strip_geometry(w, h, d, m, filt, r0, r1)[1] if strips else (0, h))
    src_px = (src_rows[1] - src_rows[0]) * w * conv_frames         # source pixels one front-end launch reads
    halo_px = ((min(oh, r1 + 6) - max(0, r0 - 6)) * ow) * conv_frames
    res["stage_ms"] = stage_avg
    res["conv_px_per_launch"] = unit_px
    res["conv_ms"] = stage_avg["conv"]
    # algorithmic bytes: front end = u8 source in, FP16 conv-input plane + source-size FP32 chroma / alpha out;
    # colour-out = FP32 Y' + source-size FP32 chroma / alpha in, d bytes per pixel (+1 with the conv plane) out
    fe_ms = stage_avg["color_in"] + stage_avg["resize1"] + stage_avg["resize2"] + stage_avg["pack"]
    fe_bytes = src_px * d + halo_px * 2 + src_px * 4 * (d - 1)
    co_bytes = unit_px * 4 + src_px * 4 * (d - 1) + unit_px * (d + (1 if conv_on else 0))
    if m < 1.5:                                                    # no hand-over below x1.5: chroma / alpha planes at output size
        fe_bytes = src_px * d + halo_px * 2 + halo_px * 4 * (d - 1)
        co_bytes = unit_px * 4 * d + unit_px * (d + (1 if conv_on else 0))
    res["mem_kernels"] = [("front end (colour-in + Y resize + conv-input pack + chroma hand-over)", fe_bytes, fe_ms),
                          ("colour-out (+ x2 chroma / alpha resampling)", co_bytes, stage_avg["color_out"])]

    if not want_e2e:
        return res
    # ---- end to end: the C-ABI call with pinned host buffers (H2D + kernels + D2H inside the timed region)
    chunk = min(nfr, C4_E2E_CHUNK)                                  # c4: the rank's batch streams through one pinned buffer set
    ncalls = (nfr + chunk - 1) // chunk
    h_src = S.pinned_empty((chunk, h, w, d))
    for i in range(chunk):
        h_src[i] = imgs[i % ndist]
    h_out = S.pinned_empty((chunk, r1 - r0, ow, d)); h_conv = S.pinned_empty((chunk, r1 - r0, ow)) if conv_on else None

    def e2e_step():
        if nfr > 1:
            for c in range(ncalls):
                k = min(chunk, nfr - c * chunk)
                ctx.batch(h_src[:k], m, filt, conv_on, out=h_out[:k], conv=h_conv[:k] if conv_on else None)
        else:
            ctx.pass_host(h_src[0], m, filt, conv_on, rows=(r0, r1), out=h_out[0], conv=h_conv[0] if conv_on else None)

    e2e_steps = steps if nfr <= 64 else max(2, min(steps, 3))
    for _ in range(2 if nfr <= 64 else 1):
        e2e_step()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    if strips:                                                     # source rows each rank uploads (strip + halo), summed
        h2d = sum((lambda g: (g[1][1] - g[1][0]) * w * d)(S.strip_geometry(w, h, d, m, filt, oh * k // world, oh * (k + 1) // world))
                  for k in range(world))
    else:
        h2d = h * w * d * nfr * world
    res["e2e"] = {"value": job_px * e2e_steps / e2e_s / 1e6, "unit": "output MP/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": job_px * d + (job_px if conv_on else 0), "steps": e2e_steps,
                  "api": "srcnn_b200_%s (C-ABI, pinned host buffers)" % (
                      "batch_host, %d call(s) of <= %d frames per step" % (ncalls, chunk) if nfr > 1
                      else "pass_host_rows" if strips else "pass_host")}
    # the stream-ordered non-blocking API (srcnn_b200_pass_host_async, two calls in flight per context): same copies per step,
    # consecutive steps overlap -- what a streaming caller gets; reported next to the blocking call, not instead of it
    if nfr == 1 and not strips:
        h_out2 = S.pinned_empty((r1 - r0, ow, d)); h_conv2 = S.pinned_empty((r1 - r0, ow)) if conv_on else None
        outs = [(h_out[0], h_conv[0] if conv_on else None), (h_out2, h_conv2)]
        for k in range(4):
            ctx.pass_host_async(h_src[0], m, filt, outs[k & 1][0], outs[k & 1][1])
        ctx.wait()
        env.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            ctx.pass_host_async(h_src[0], m, filt, outs[k & 1][0], outs[k & 1][1])
        ctx.wait()
        env.barrier()
        a_s = env.max_over_ranks(time.perf_counter() - t0)
        res["e2e"]["pipelined_async"] = {"value": job_px * steps / a_s / 1e6, "unit": "output MP/s",
                                         "api": "srcnn_b200_pass_host_async + srcnn_b200_wait (two calls in flight, pinned buffers)"}
        S.lib().srcnn_b200_host_free(h_out2.ctypes.data)
        if h_conv2 is not None:
            S.lib().srcnn_b200_host_free(h_conv2.ctypes.data)
    for a in (h_src, h_out, h_conv):
        if a is not None:
            S.lib().srcnn_b200_host_free(a.ctypes.data)
    # the drop-in C++ call itself (pageable input, new[] output released with delete[]), rank 0
    if want_dropin and rank == 0 and not strips and nfr == 1:
        for _ in range(3):                                         # (its own context from the library's pool: buffers, tables)
            S.process_srcnn(imgs[0], m, filt, name == "c5", conv_on, copy=False)
        n = max(10, steps); ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            rc, _, _ = S.process_srcnn(imgs[0], m, filt, name == "c5", conv_on, copy=False)
            ts.append(time.perf_counter() - t0)
        if rc == 0:
            res["e2e"]["dropin_ProcessSRCNN_pageable"] = ow * oh / float(np.median(ts)) / 1e6
            res["e2e"]["dropin_note"] = "ProcessSRCNN: pageable input, result in a fresh new[] buffer, median of %d calls" % n
    return res


def measure_one_call(ngpu, steps):
    """ONE host call over all GPUs of the box, single process: srcnn_b200_multi_pass_host on the C3 image (row strips, one per
    GPU, 6-row halos, no exchange) with pinned host buffers -- the in-library counterpart of the reference spreading one
    ProcessSRCNN call over all host cores.  Runs on rank 0 BEFORE the process group exists (the other ranks are still waiting
    in the rendezvous and leave their GPUs idle)."""
    import libsrcnn_b200 as S
    import oracle_lib as O
    h, w, d, m, filt, desc = WORKLOADS["c3"]
    ow, oh = S.output_dims(w, h, m)
    img = O.synth("uniform", h, w, d, 43)
    h_src = S.pinned_empty((h, w, d)); h_src[...] = img
    h_out = S.pinned_empty((oh, ow, d))
    res = {"config": config_of("c3", ngpu, 1), "unit": "output MP/s", "scaling": "strong",
           "api": "srcnn_b200_multi_pass_host (one call, one process, %d GPUs; pinned host buffers)" % ngpu}
    mg = S.Multi(list(range(ngpu)))
    try:
        for _ in range(2):
            mg.pass_host(h_src, m, filt, False, out=h_out)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            mg.pass_host(h_src, m, filt, False, out=h_out)
            ts.append(time.perf_counter() - t0)
        res.update(value=ow * oh / (sum(ts) / len(ts)) / 1e6, ms_per_call=sum(ts) / len(ts) * 1e3, steps=steps,
                   h2d_bytes_per_step=int(img.nbytes), d2h_bytes_per_step=ow * oh * d)
        # the same call on ONE GPU of the same handle type, for the ratio
        m1 = S.Multi([0])
        m1.pass_host(h_src, m, filt, False, out=h_out)
        t0 = time.perf_counter()
        for _ in range(max(2, steps // 2)):
            m1.pass_host(h_src, m, filt, False, out=h_out)
        res["value_1gpu"] = ow * oh * max(2, steps // 2) / (time.perf_counter() - t0) / 1e6
        m1.close()
    finally:
        mg.close()
        for a in (h_src, h_out):
            S.lib().srcnn_b200_host_free(a.ctypes.data)
    return res


def measure_chain(env, calls=7):
    """step-scale chain (reference src/libsrcnn.cpp:980-1061): x4 = x2 . x2 from 960x540 to 3840x2160 through the C-ABI call with host
    buffers; the 8-bit intermediate stays in HBM, the intermediate pass runs on the bit-exact FP32 kernel"""
    import oracle_lib as O
    img = O.synth("uniform", 540, 960, 3, 6)
    rc, out, _ = env.ctx.process(img, 4.0, env.S.BICUBIC, True, False)
    if rc != 0 or out is None or out.shape != (2160, 3840, 3):
        raise RuntimeError("chain_x4: rc %d" % rc)
    ts = []
    for _ in range(calls):
        t0 = time.perf_counter()
        env.ctx.process(img, 4.0, env.S.BICUBIC, True, False)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return {"config": {"workload": "synthetic 960x540 RGB x4.0 bicubic, precision-step on (two x2 passes) -> 3840x2160"},
            "value": 3840 * 2160 / dt / 1e6, "unit": "output MP/s", "ms_per_call": dt * 1e3, "calls": calls,
            "api": "srcnn_b200_process_host (C-ABI, pageable host buffers, blocking), median"}


def mem_rooflines(res, hbm):
    return [{"kernel": k, "bytes": int(b), "ms": ms, "gbs": (b / (ms * 1e-3) / 1e9 if ms > 0 else None),
             "frac": (b / (ms * 1e-3) / 1e9 / hbm if ms > 0 else None), "peak_gbs": hbm} for k, b, ms in res["mem_kernels"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--conv", default="tensor", choices=["tensor", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the c3 / c4 sub-results")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU and step (default 1; c4: 4096 / N)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    one_call = None
    if (int(os.environ.get("WORLD_SIZE", "1")) > 1 and int(os.environ.get("RANK", "0")) == 0 and args.workload == "c2"
            and args.conv == "tensor" and not args.no_sub):
        try:
            one_call = measure_one_call(int(os.environ["WORLD_SIZE"]), min(args.steps, 8))
        except Exception as e:                                              # noqa: BLE001
            one_call = {"error": repr(e)}
    env = Env(args.conv)
    name = args.workload
    h, w, d, m, filt, desc = WORKLOADS[name]
    main_res = measure(env, name, args.steps, args.warmup, args.frames, want_e2e=True, want_dropin=True, sample_clocks=True)
    pk = peaks()

    sub = {}
    if not args.no_sub and name == "c2" and args.conv == "tensor":
        for sname, ssteps in (("c3", min(args.steps, 10)), ("c4", min(args.steps, 3))):
            try:
                r = measure(env, sname, ssteps, 3, 0, want_e2e=True)
                tf = r["conv_px_per_launch"] * FLOP_PER_PX / (r["conv_ms"] * 1e-3) / 1e12
                sub[sname] = {"config": config_of(sname, env.world, r["nfr"]), "value": r["value"], "unit": "output MP/s",
                              "scaling": "strong" if sname == "c3" else "weak (4096 frames in total: strong in the batch, fixed per-frame work)",
                              "ms_per_step": r["ms_per_step"], "steps": r["steps"], "e2e": r["e2e"], "gpu_launches": r["launches"],
                              "roofline": {"kernel": "fused conv I+II+III (tensor)", "achieved": tf, "peak": pk["tf"], "unit": "TFLOP/s",
                                           "frac": tf / pk["tf"], "kernel_ms": r["conv_ms"]},
                              "roofline_mem": mem_rooflines(r, pk["hbm"]), "stage_ms": r["stage_ms"]}
            except Exception as e:                                          # noqa: BLE001  (a sub-result must not cost the headline)
                sub[sname] = {"error": repr(e)}

    if not args.no_sub and name == "c2" and args.conv == "tensor" and env.world == 1:
        try:
            sub["chain_x4"] = measure_chain(env)
        except Exception as e:                                              # noqa: BLE001
            sub["chain_x4"] = {"error": repr(e)}

    if env.rank != 0:
        env.close()
        return

    conv_ms = main_res["conv_ms"]
    achieved_tf = main_res["conv_px_per_launch"] * FLOP_PER_PX / (conv_ms * 1e-3) / 1e12
    cpu = None
    if env.world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(name, env.ctx)
    traffic = ncu_traffic(name, args.conv)
    line = {
        "metric": METRIC, "value": main_res["value"], "unit": "output MP/s",
        "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong" if name == "c3" else "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (conv), f32 (Y resize, colour)" if args.conv == "tensor" else "f32 (conv exact mode), f64 (resize)",
        "data": "synthetic",
        "config": config_of(name, env.world, main_res["nfr"]),
        "method": {"conv_mode": args.conv,
                   "kernel_events": "per-kernel CUDA events on every %s timed step (%d samples)" % (
                       {1: "", 2: "2nd", 3: "3rd", 4: "4th"}[main_res["probe_every"]], (args.steps + main_res["probe_every"] - 1) // main_res["probe_every"])},
        "e2e": main_res["e2e"],
        "gpu_launches": main_res["launches"],
        "clocks": main_res["clocks"],
        "roofline": {"kernel": "fused conv I+II+III (%s)" % args.conv, "bound": "tensor", "achieved": achieved_tf,
                     "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved_tf / pk["tf"],
                     "traffic": traffic[0], "traffic_source": traffic[1],
                     "peak_source": pk["src"], "kernel_ms": conv_ms,
                     "algorithmic_flop_per_launch": main_res["conv_px_per_launch"] * FLOP_PER_PX,
                     "achieved_conv_I_II_only": achieved_tf * 14464.0 / FLOP_PER_PX,
                     "note": "achieved counts conv I + II + III (16 064 flop per output pixel); conv I + II alone are 14 464 (SURVEY.md 8(d) asks for both)"},
        "roofline_mem": mem_rooflines(main_res, pk["hbm"]),
        "stage_ms": main_res["stage_ms"],
        "cpu_baseline": cpu,
    }
    if one_call is not None:
        sub["c3_one_call"] = one_call
    if sub:
        line["sub"] = sub
    emit(line)
    env.close()


def ncu_traffic(name, conv):
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv kernel from the committed ncu --set full capture (per launch)"""
    p = ROOT / "profiles" / "ncu_conv_traffic.json"
    try:
        j = json.loads(p.read_text())
        e = j[name + ":" + conv]
        return int(e["bytes"]), e["source"]
    except Exception:                                                       # noqa: BLE001
        return None, "no ncu capture committed for this workload"


if __name__ == "__main__":
    main()
