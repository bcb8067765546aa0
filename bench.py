#!/usr/bin/env python
"""bench.py — 1080p P-frames/sec of the DVC P-frame hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one closed-loop GOP (GOP=10: one I-frame taken as given + 9 P-frames) of synthetic
1088x1920 frames per rank (BASELINE.json configs[1]; configs[2] shards GOPs over ranks: weak
scaling, no collective on the data path, one all-reduce of the statistics at the end).
`value` = whole-job P-frames/s with frames resident in HBM; `e2e` = the same through the
host-buffer entry point (pinned host frames, H2D inside the timed region, D2H of the metrics).
`--impl reference` times the reference's algorithm on the host CPU cores (the oracle port of the
PyTorch fp32 path; the reference itself is Python and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, GOP, BATCH = 1088, 1920, 10, 1
FLOP_PER_PX = 1397809 + 5376          # conv 2*MAC + GDN 1x1, per pixel per P-frame
METRIC = "1080p P-frames/sec (whole box)"
# --config: the headline workload (configs[1]/[2]) or one of the other BASELINE.json configs (lines for profiles/)
CONFIGS = {"hd": (1088, 1920, 1, "DVC P-frame forward 1088x1920, GOP=10 (9 P-frames/step/rank), B=1, configs[1]; "
                                 "GOPs sharded by rank (configs[2])"),
           "4k": (2176, 3840, 1, "DVC P-frame forward 3840x2160 padded to 2176 rows, GOP=10, B=1, configs[3]"),
           "multiview": (768, 1280, 8, "DVC P-frame forward, 8 camera views of 1280x720 padded to 768 rows folded "
                                       "into the batch (B=8), GOP=10, configs[4]")}


def hbm_kernel_bytes(tag, h, w, b):
    """Algorithmic HBM bytes of one launch of a memory-bound kernel, fp32 tensors as the reference
    holds them + the engine's 4 B/channel hi|lo records; `tag` is the profile name '@kernel[:variant]'."""
    px = float(h * w * b)
    name, _, var = tag[1:].partition(":")
    if name == "k_mc_prep":                       # ref gather 12 + mv 8 in; warpframe 12 + [warp,ref] record 32 out
        return 64 * px
    if name == "k_mc_finish":                     # warpnet res 12 + warpframe 12 + cur 12 in; prediction 12 + record 32 out
        return 80 * px
    if name == "k_spynet_prep":                   # im1 12 + im2 gather 12 + coarse flow 2 in; record 32 + flow_up 8 out
        lvl = int(var[1:])
        return 66 * px / 4 ** (3 - lvl)
    if name == "k_recon_losses":                  # cur, pred, warp 36 + res 12 in; clipped 12 out
        return 60 * px
    if name == "k_upadd_act":                     # 64-ch records (256 B): low/4 + skip in; x and relu(x) out
        return (256 / 4 + 256 + 512) * px / (1 if var == "full" else 4)
    if name == "k_pool_act":                      # 4 records in, x and relu(x) out, per output pixel
        return (4 * 256 + 512) * px / (4 if var == "full" else 16)
    if name == "k_quant_bits_factorized":         # 8 B/element
        return 8 * (128 * px / 256 if var == "mv" else 64 * px / 4096)
    if name == "k_quant_bits_laplace":            # 12 B/element
        return 12 * 96 * px / 256
    if name == "k_avg_pool2_planar":
        return None
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1383.4), d.get("hbm_gbs", 6548.2), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def run_reference(args, rank, world):
    """Reference arm: the reference algorithm on the host CPU (oracle port), all host threads."""
    if rank != 0:
        return
    import torch
    from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
    from oracle import dvc_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = init_state_dict(0)
    frames = synthetic_gop(H, W, gop=2, gop_id=0)[:, 0]
    with torch.no_grad():
        for _ in range(args.warmup):
            dvc_oracle.pframe_forward(sd, frames[1:2], frames[0:1])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dvc_oracle.pframe_forward(sd, frames[1:2], frames[0:1])
        dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = "1 P-frame at %dx%d per step (bounded sample of the GOP workload), fp32, torch CPU" % (H, W)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "P-frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIGS[args.config][3]},
            "cpu_baseline": {"value": v, "unit": "P-frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "P-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    import torch
    from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
    from oracle import dvc_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = init_state_dict(0)
    small = synthetic_gop(256, 256, gop=2, gop_id=0)[:, 0]
    frames = synthetic_gop(H, W, gop=2, gop_id=0)[:, 0]
    with torch.no_grad():
        dvc_oracle.pframe_forward(sd, small[1:2], small[0:1])  # thread-pool warm-up
        t0 = time.perf_counter()
        n = 0
        while n < 2 and time.perf_counter() - t0 < 25.0:
            dvc_oracle.pframe_forward(sd, frames[1:2], frames[0:1])
            n += 1
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "P-frames/s", "cores": cores, "kind": "port",
            "sample": "%d P-frame(s) at %dx%d, fp32, oracle port of the reference on torch CPU" % (n, H, W)}


def gpu_baseline_sample(dev, n_frames=3):
    """Like-for-like GPU baseline (SURVEY 8d last row): the reference's PyTorch algorithm (the oracle port: F.conv2d
    -> cuDNN, grid_sample-equivalent gathers, ...) on the SAME B200 in fp32, once with TF32 disabled (the
    parity-equivalent baseline) and once with PyTorch's default cuDNN TF32 (what a user of the reference gets).
    Outside every timed region of our arm; a reported baseline like cpu_baseline, never on the product path."""
    import torch
    from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop
    from oracle import dvc_oracle
    sd = {k: v.to(dev) for k, v in init_state_dict(0).items()}
    frames = synthetic_gop(H, W, gop=2, gop_id=0, batch=BATCH).to(dev)
    out = {"kind": "port", "what": "oracle port of DVC/net.py:70-220 on torch CUDA (cuDNN) fp32, %d P-frames at "
                                   "%dx%d B=%d after 1 warm-up" % (n_frames, H, W, BATCH), "unit": "P-frames/s"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        with torch.no_grad(), torch.device(dev):
            for key, tf32 in (("fp32_tf32_off", False), ("fp32_tf32_on", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                dvc_oracle.pframe_forward(sd, frames[1], frames[0])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n_frames):
                    dvc_oracle.pframe_forward(sd, frames[1], frames[0])
                e1.record()
                torch.cuda.synchronize()
                out[key] = BATCH * n_frames / (e0.elapsed_time(e1) * 1e-3)
    except Exception as ex:  # the baseline must never take the bench down
        out["error"] = "%s: %s" % (type(ex).__name__, str(ex)[:200])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()
    return out


def parity_against_reference_golden(rows_first_gop):
    """First timed GOP (rank 0: gop_id 0) against the UNMODIFIED reference's own closed-loop rows for the same GOP
    (tests/golden/hd_gop10.npz, oracle/gen_golden_r2.py): north-star gates bpp 0.5 %, PSNR 0.02 dB on the GOP means."""
    import math
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "hd_gop10.npz")
    if not os.path.exists(p):
        return {"ok": None, "why": "tests/golden/hd_gop10.npz missing"}
    with np.load(p) as z:
        ref = z["rows"]                         # [9, 8]: mse warploss interloss bpp_f bpp_z bpp_mv bpp psnr
    got = rows_first_gop.double().numpy()       # [9, 7]
    bpp, bpp_ref = got[:, 6].mean(), ref[:, 6].mean()
    psnr = float(np.mean([10.0 * math.log10(1.0 / m) for m in got[:, 0]]))
    psnr_ref = float(ref[:, 7].mean())
    frame_bpp_rel = float(np.max(np.abs(got[:, 6] - ref[:, 6]) / ref[:, 6]))
    frame_psnr = float(np.max(np.abs(np.array([10.0 * math.log10(1.0 / m) for m in got[:, 0]]) - ref[:, 7])))
    bpp_rel = float(abs(bpp - bpp_ref) / bpp_ref)
    psnr_db = abs(psnr - psnr_ref)
    return {"against": "unmodified reference, closed-loop GOP-10 rows (tests/golden/hd_gop10.npz)",
            "bpp": float(bpp), "bpp_ref": float(bpp_ref), "bpp_rel": bpp_rel, "psnr": psnr, "psnr_ref": psnr_ref,
            "psnr_db": psnr_db, "max_frame_bpp_rel": frame_bpp_rel, "max_frame_psnr_db": frame_psnr,
            "gates": {"bpp_rel": 0.005, "psnr_db": 0.02}, "ok": bool(bpp_rel <= 0.005 and psnr_db <= 0.02)}


def fast_mode_sample(args, dev, dev_gops, host_gops):
    """Second figure on the same line (SURVEY 7.2-1 "ship two modes and report both"): precision='fast' = ONE fp16 MMA
    per product instead of three on hi/lo pairs.  Same workload, same timing rules (W warm-up GOPs, K timed GOPs, CUDA
    events); parity is metric-level only: the GOP means against the unmodified reference's rows."""
    import torch
    from fastvideocodec_b200 import VideoCompressor
    from fastvideocodec_b200._lib import check, lib, ptr, stream_ptr
    from fastvideocodec_b200.synthetic import init_state_dict
    m = VideoCompressor(precision="fast")
    m.load_state_dict(init_state_dict(0))
    m = m.to(dev).eval()
    ctx = m._context(BATCH, H, W, dev)
    rec = torch.empty((2, BATCH, 3, H, W), device=dev)
    scal = torch.empty((GOP - 1, 7), device=dev)

    def gop(step):
        frames = dev_gops[step % len(dev_gops)]
        prev = frames[0]
        for i in range(1, GOP):
            out = rec[i & 1]
            check(lib().fvc_pframe_forward(ctx.handle, ptr(frames[i]), ptr(prev), ptr(out), ptr(scal[i - 1]),
                                           stream_ptr()), "fvc_pframe_forward")
            prev = out

    for s in range(args.warmup):
        gop(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        gop(args.warmup + s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    for s in range(2):
        m.gop_forward_host(host_gops[s % len(host_gops)], want_recon=False)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    rows = []
    for s in range(args.steps):
        rows.append(m.gop_forward_host(host_gops[s % len(host_gops)], want_recon=False)[1].clone())
    e3.record()
    torch.cuda.synchronize()
    n = args.steps * (GOP - 1) * BATCH
    out = {"precision": "fast", "dtype": "f16 x1 MMA (fp32 accumulate)", "value": n / (ms * 1e-3), "unit": "P-frames/s",
           "ms_per_step": ms / args.steps, "e2e": {"value": n / (e2.elapsed_time(e3) * 1e-3), "unit": "P-frames/s"},
           "roofline_frac": FLOP_PER_PX * H * W * BATCH * n / (ms * 1e-3) / 1e12 / _peaks()[0],
           "parity": parity_against_reference_golden(rows[0]) if (H, W, BATCH) == (1088, 1920, 1) else None,
           "note": "metric-level parity only (bpp 0.5 %, PSNR 0.02 dB); quantised latents are NOT bit-exact in this mode"}
    m.release()
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from fastvideocodec_b200 import VideoCompressor, reduce_stats, stats_vector, summarize
    from fastvideocodec_b200._lib import check, lib, ptr, stream_ptr
    from fastvideocodec_b200.synthetic import init_state_dict, synthetic_gop

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    model = VideoCompressor(precision=args.precision)
    model.load_state_dict(init_state_dict(0))
    model = model.to(dev).eval()
    impl_name = {0: "simt", 1: "tc", 2: "tc-fast"}[model.impl]

    # each rank codes its own GOPs (GOP g seeded 1234+g; rank r owns g = r mod world): weak scaling.
    # --total-gops T (configs[2]: "64 GOPs x 10 frames"): the T GOPs are dealt round-robin (gop.shard_gops), every rank
    # times its share, the job time is the slowest rank's: strong scaling.
    n_local = 2
    strong = args.total_gops > 0
    if strong:
        from fastvideocodec_b200 import shard_gops
        args.steps = len(shard_gops(args.total_gops, rank, world))
    host_gops = [synthetic_gop(H, W, gop=GOP, gop_id=rank + i * world, batch=BATCH).contiguous().pin_memory()
                 for i in range(n_local)]                      # [G,B,3,H,W] each
    dev_gops = [g.to(dev) for g in host_gops]
    ctx = model._context(BATCH, H, W, dev)
    rec = torch.empty((2, BATCH, 3, H, W), device=dev)
    scal = torch.empty((args.steps + args.warmup, GOP - 1, 7), device=dev)

    def gop_resident(step):
        frames = dev_gops[step % n_local]
        prev = frames[0]
        for i in range(1, GOP):
            out = rec[i & 1]
            check(lib().fvc_pframe_forward(ctx.handle, ptr(frames[i]), ptr(prev), ptr(out), ptr(scal[step, i - 1]),
                                           stream_ptr()), "fvc_pframe_forward")
            prev = out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident (value) ------------------------------------------------------------------
    for s in range(args.warmup):
        gop_resident(s)
    barrier()
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for s in range(args.steps):
            gop_resident(args.warmup + s)
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t)
    frames_total = (args.total_gops if strong else world * args.steps) * (GOP - 1) * BATCH
    value = frames_total / (ms_max * 1e-3)

    # ---- end to end: host frames -> metrics on host -----------------------------------------
    for s in range(min(args.warmup, 3)):
        model.gop_forward_host(host_gops[s % n_local], want_recon=False)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    rows = []
    for s in range(args.steps):
        _, sc = model.gop_forward_host(host_gops[s % n_local], want_recon=False)
        rows.append(sc.clone())
    e3.record()
    barrier()
    t2 = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = frames_total / (float(t2) * 1e-3)
    h2d = GOP * BATCH * 3 * H * W * 4
    d2h = (GOP - 1) * 7 * 4

    # the same call fed with the frames as the reference's loader holds them before ToTensor (uint8 HWC, dataset.py:68-75):
    # a quarter of the upload, ToTensor on the device (extra information; `e2e` above stays the float32 contract)
    u8_gops = [(g * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous().pin_memory()
               for g in host_gops[:max(1, min(n_local, args.steps))]]
    model.gop_forward_host(u8_gops[0], want_recon=False)
    barrier()
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e4.record()
    for s in range(args.steps):
        model.gop_forward_host(u8_gops[s % len(u8_gops)], want_recon=False)
    e5.record()
    barrier()
    t3 = torch.tensor([e4.elapsed_time(e5)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    e2e_u8 = {"value": frames_total / (float(t3) * 1e-3), "unit": "P-frames/s", "h2d_bytes_per_step": GOP * BATCH * 3 * H * W,
              "d2h_bytes_per_step": d2h, "input": "uint8 [G,B,H,W,3] host frames, ToTensor on the device (fvc_gop_forward_host_u8)"}

    # ---- statistics: the one collective of the path ---------------------------------------------
    stats = summarize(reduce_stats(stats_vector(torch.cat(rows, 0))))

    # ---- roofline of the dominant kernels (convolution engine), rank 0 only ---------------------
    roof, cpu, gpu_base, parity, fast = None, None, None, None, None
    if rank == 0:
        tf_peak, hbm_peak, how = _peaks()
        os.environ["FVC_PROFILE"] = "1"
        pm = VideoCompressor()
        pm.load_state_dict(init_state_dict(0))
        pm = pm.to(dev).eval()
        fr = dev_gops[0]
        conv_s, texts = [], []
        with torch.no_grad():
            for i in range(1, 6):
                pm(fr[i], fr[i - 1])
                conv_s.append(lib().fvc_ctx_last_conv_seconds(pm._last_ctx.handle))
                texts.append(lib().fvc_ctx_profile_text(pm._last_ctx.handle).decode())
        os.environ["FVC_PROFILE"] = "0"
        conv_t = sorted(conv_s[1:])[len(conv_s[1:]) // 2]
        flops = FLOP_PER_PX * H * W * BATCH
        ach = flops / conv_t / 1e12
        # memory-bound kernels: achieved HBM GB/s per launch from the same CUDA-event pass (median of 4 frames)
        per = {}
        for t in texts[1:]:
            seen = {}
            for ln in t.splitlines():
                name, ms = ln.rsplit(" ", 1)
                if name.startswith("@"):
                    seen.setdefault(name, []).append(float(ms))
            for name, v in seen.items():
                per.setdefault(name, []).append(sum(v) / len(v))
        other_ms = sum(float(ln.rsplit(" ", 1)[1]) for ln in texts[-1].splitlines() if ln.startswith("@"))
        hbm_kernels = {}
        for name, v in sorted(per.items()):
            ms = sorted(v)[len(v) // 2]
            nbytes = hbm_kernel_bytes(name, H, W, BATCH)
            if nbytes:
                gbs = nbytes / (ms * 1e-3) / 1e9
                hbm_kernels[name[1:]] = {"ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / hbm_peak, 3)}
        traffic, traffic_src = None, None   # DRAM bytes of the convolution launches of one frame
        tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tp) and (H, W, BATCH) == (1088, 1920, 1):
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("dram_bytes_per_frame")
            traffic_src = "profiles/conv_traffic.json (ncu dram__bytes_read+write over the conv launches of one frame; " \
                          "%s); not measured in this run" % tj.get("source", "committed ncu pass")
        roof = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "convolution engine (%s), all conv launches of one P-frame" % impl_name,
                "algorithmic_flop_per_frame": flops, "conv_seconds_per_frame": conv_t,
                "conv_share_of_frame": conv_t / (ms_max * 1e-3 / (args.steps * (GOP - 1))), "peak_source": how,
                "hbm_kernels": hbm_kernels, "hbm_peak_gbs": hbm_peak,
                "non_conv_ms_per_frame_serialised": round(other_ms, 4)}
        pm.release()
        if (H, W, BATCH) == (1088, 1920, 1):
            parity = parity_against_reference_golden(rows[0])
        if world == 1:
            model.release()
            torch.cuda.empty_cache()
            if args.precision != "fast" and impl_name == "tc":
                fast = fast_mode_sample(args, dev, dev_gops, host_gops)
            gpu_base = gpu_baseline_sample(dev)
            cpu = cpu_baseline_sample()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "P-frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / max(1, args.steps), "higher_is_better": True,
                "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": {"tc": "f16 hi/lo pairs x3 MMAs (fp32 accumulate)", "tc-fast": "f16 x1 MMA (fp32 accumulate)",
                                                  "simt": "f32"}[impl_name],
                "data": "synthetic",
                "config": {"workload": CONFIGS[args.config][3] + (" (%d views per rank)" % BATCH if args.config == "multiview" else ""),
                           "engine": impl_name, "l2": "per-frame working set (GBs of activations) exceeds the 126 MB L2",
                           "parallelism": "gop-sharded x%d, no data-path collective" % world,
                           **({"total_gops": args.total_gops, "steps_are": "GOPs of rank 0's share"} if strong else {})},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": "P-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "d2h_what": "the 7 scalars per P-frame only (eval.py reads metrics, not frames; "
                                    "models.py:376-383)"},
                "e2e_u8_ingest": e2e_u8,
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "gpu_baseline": gpu_base,
                "parity": parity, "parity_stats": stats, "fast_mode": fast}
        print(json.dumps(line), flush=True)
    model.release()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="hd", choices=sorted(CONFIGS))
    ap.add_argument("--total-gops", type=int, default=0,
                    help="strong-scaling form (configs[2]): this many GOPs in total, dealt round-robin to the ranks; "
                         "overrides --steps")
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"],
                    help="exact: fp16 hi/lo pairs, 3 MMAs per product (element-level parity; the headline); "
                         "fast: 1 fp16 MMA per product (metric-level parity)")
    args = ap.parse_args()
    global H, W, BATCH
    H, W, BATCH = CONFIGS[args.config][:3]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config == "multiview" and world > 1 and BATCH % world == 0:
        BATCH //= world      # configs[4] "one view per GPU": the 8 camera views are split over the ranks (8 GPUs: B = 1 each)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        with torch.no_grad():
            run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
