#!/usr/bin/env python
"""bench.py — BASELINE.json's headline: VAE encode+decode frames/s at 720x1280x129 frames (bf16 model, tiled).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one synthetic clip: enable_tiling(); encode(1x3x129x720x1280) ->
mode() -> decode() (BASELINE config 4; 84 encoder + 84 decoder sub-model calls, 6467.8 conv TFLOP).  At N > 1
the SAME clip is partitioned by tile over the ranks (strong scaling, hunyuanvideo_efficiency_b200/vae/tile_parallel.py).
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/) on the
host cores on a bounded sample of the same workload (BASELINE config 1, one canonical tile), scaled by conv FLOPs.

Timed region: K steps with the clip resident in HBM, device timed (CUDA events, max over ranks), tiles dealt over the
model's `tile_streams` CUDA streams, no instrumentation.  `e2e`: the same K steps from a pinned fp32 host clip with the
H2D copy, bf16 cast and D2H of the reconstruction inside the timed region.  `roofline` / `kernel_ms_per_step_rank0`:
CUDA events around every C-ABI launch (hyvae_profile_begin/end) on extra steps right after the timed region with all
kernels serialised on ONE stream, because event pairs of kernels that overlap across streams do not measure a kernel.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vae_encode_decode_frames_per_sec_720p_129f"
UNIT = "frames/s"
FRAMES, HEIGHT, WIDTH = 129, 720, 1280
CPU_SAMPLE_SHAPE = (1, 3, 17, 256, 256)   # bounded sample the CPU arm times (scaled to the workload by conv FLOPs)
WORKLOAD = "config4: enable_tiling(); encode(1x3x129x720x1280) -> mode() -> decode(); 84+84 sub-model calls"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel from the committed `ncu --set full`
# capture (profiles/r02_ncu_conv_wino_128_v2.txt): conv_wino_kernel<half>, 128 -> 128 channels, 17x256x256 voxels with residual
# and fused GroupNorm statistics.  Algorithmic bytes of that launch: Winograd-T planes 33 x 258 x 258 x 128 x 2 B = 562 MB +
# residual 285 MB + y 285 MB = 1133 MB.
NCU_TRAFFIC = {"bytes": 1116.4e6, "note": "one 128->128 17x256x256 launch of conv_wino_kernel (ncu --set full, "
                                          "profiles/r02_ncu_conv_wino_128_v2.txt): 849.5 MB read + 266.9 MB written vs 1133 MB algorithmic "
                                          "(33 Winograd-T planes with halo + residual + y); tensor pipe active 90.6 % of the cycles"}


def _peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p.update(json.load(f))
            p["_source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.proc = index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop_evt.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_sample(threads: int):
    """One bounded CPU sample of the workload through the oracle port of the reference: fp32 encode+decode of a
    1x3x17x256x256 clip with the HY config (untiled: BASELINE config 1, one canonical tile of the tiled workload; about
    10-20 s on 16 host threads).  Returns (seconds, conv_flops_of_sample)."""
    import torch
    from oracle import flops as FL
    from oracle import vae_oracle as O
    from oracle import weights as W
    torch.set_num_threads(threads)
    cfg = W.HY_VAE_CONFIG
    if not hasattr(cpu_sample, "_sd"):
        cpu_sample._sd = W.make_state_dict(cfg)
    shape = CPU_SAMPLE_SHAPE
    x = W.make_video(shape)
    tl = O.Tiling.from_cfg(cfg)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.forward(cpu_sample._sd, cfg, x, tl)
    dt = time.perf_counter() - t0
    fe, (t, h, w) = FL.encoder_tile_flops(cfg, *shape[:1], *shape[2:])
    fd, _ = FL.decoder_tile_flops(cfg, 1, t, h, w)
    return dt, fe + fd


def full_workload_flops(workload: str = "config4", frames: int = FRAMES, height: int = HEIGHT, width: int = WIDTH):
    """SURVEY 8d conv_flops of one step of the workload: every nn.Conv3d the reference executes in its own tile
    decomposition, 2*Cout*Cin*k^3*voxels with the TRUE channel counts (3, not the 8 / 16 the tensor-core kernels store) and
    without the attention's Linear projections.  6467.8 TFLOP for config 4 (oracle/flops.py is bookkeeping, no VAE arithmetic)."""
    from oracle import flops as FL
    from oracle import vae_oracle as O
    from oracle import weights as W
    cfg = W.HY_VAE_CONFIG
    tl = O.Tiling.from_cfg(cfg, True, True)
    fe = FL.path_flops(cfg, (1, 3, frames, height, width), tl, "encode")[0] if workload != "config2" else 0.0
    fd = FL.path_flops(cfg, (1, 16, (frames - 1) // 4 + 1, height // 8, width // 8), tl, "decode")[0] if workload != "config3" else 0.0
    return fe + fd


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    full = full_workload_flops()
    for _ in range(args.warmup):
        cpu_sample(threads)
    ts = []
    for _ in range(args.steps):
        dt, fl = cpu_sample(threads)
        ts.append(dt)
    sec_per_sample = sum(ts) / len(ts)
    sec_full = sec_per_sample * full / fl  # scale the sample to the whole clip by conv FLOPs
    value = FRAMES / sec_full
    sample = (f"oracle port of the reference, fp32, {threads} host threads: encode+decode of {'x'.join(map(str, CPU_SAMPLE_SHAPE))} (HY config, "
              f"{fl / 1e12:.2f} conv TFLOP, {sec_per_sample:.1f} s), scaled to the {full / 1e12:.1f} TFLOP of the full tiled workload")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sampled": True},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ the real bar
def torch_gpu_baseline(dev, host_video, steps: int = 1, warmup: int = 1):
    """The reference's algorithm through PyTorch + cuDNN on the SAME B200 ("the bar on the B200 box is the
    unmodified reference running through PyTorch 2.11 + cuDNN"; unet_causal_3d_blocks.py:68-75,359-363,661).  /root/reference
    does not exist on the GPU box and carries no installable package, so this is the oracle port (a line-cited functional
    restatement, pinned to the unmodified reference by tests/test_oracle_golden.py) moved to cuda in bf16: F.pad(replicate) +
    cuDNN conv3d (NCDHW), ATen group_norm / silu / add, nearest upsample by repeat_interleave, F.scaled_dot_product_attention
    with the dense additive mask, the reference's tile loops and Python blend loops.  Favourable to the reference in one
    respect: the frame-causal mask is built vectorised, not by the reference's 17 408-iteration Python loop per call
    (unet_causal_3d_blocks.py:38-46).  Same workload as the timed region: BASELINE config 4 in full."""
    import torch
    from oracle import vae_oracle as O
    from oracle import weights as W
    cfg = W.HY_VAE_CONFIG
    sd = {k: v.to(dev, torch.bfloat16) for k, v in W.make_state_dict(cfg).items()}
    tl = O.Tiling.from_cfg(cfg, True, True)
    x = host_video.to(dev, torch.bfloat16)
    saved = O.ATTN_IMPL
    O.ATTN_IMPL = "sdpa"
    try:
        with torch.no_grad():
            for _ in range(warmup):
                O.forward(sd, cfg, x, tl)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                dec, mean, _ = O.forward(sd, cfg, x, tl)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": x.shape[2] / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup, "kind": "port",
                "what": "oracle port of the reference on cuda:0 in bf16 through PyTorch "
                        f"{torch.__version__} / cuDNN {torch.backends.cudnn.version()}: cuDNN conv3d, ATen group_norm, "
                        "F.scaled_dot_product_attention with the dense mask (built vectorised), reference tile + blend loops; "
                        "full config 4, inputs resident in HBM, CUDA events",
                "cudnn_benchmark": bool(torch.backends.cudnn.benchmark)}
    finally:
        O.ATTN_IMPL = saved
        del sd, x
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hunyuanvideo_efficiency_b200 import _native as N
    from hunyuanvideo_efficiency_b200.vae import AutoencoderKLCausal3D
    from hunyuanvideo_efficiency_b200.vae import tile_parallel as TP

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = _peaks()

    # synthetic weights of the HY architecture (no checkpoint offline); bf16 model as in BASELINE configs 2-4
    from hunyuanvideo_efficiency_b200.synthetic import HY_VAE_CONFIG, make_state_dict, make_video
    vae = AutoencoderKLCausal3D.from_config(HY_VAE_CONFIG)
    vae.load_state_dict(make_state_dict(HY_VAE_CONFIG))
    vae = vae.to(torch.bfloat16).to(dev).eval().requires_grad_(False)
    vae.enable_tiling()
    if args.tile_streams is not None:
        vae.tile_streams = args.tile_streams
    frames, height, width = args.frames, args.height, args.width
    metric, scaling, frames_per_step = METRIC, "strong", frames
    runner = TP.TileParallelVAE(vae, rank, world) if world > 1 else None
    if args.workload == "config2":
        # BASELINE config 2: tiled decode of 16x33x90x160 latents to a 129-frame 720x1280 video (tiles over ranks at N > 1)
        from hunyuanvideo_efficiency_b200.synthetic import make_latent
        host_video = make_latent((1, 16, (frames - 1) // 4 + 1, height // 8, width // 8)).pin_memory()
        metric = "vae_tiled_decode_frames_per_sec_720p_129f"
        workload = f"config2: enable_tiling(); decode(1x16x{(frames - 1) // 4 + 1}x{height // 8}x{width // 8}) -> {frames}x{height}x{width}"

        def step(x):
            return runner.decode(x) if runner is not None else vae.decode(x).sample
    elif args.workload == "config3":
        # BASELINE config 3: batched dataset encode of 65-frame 544x960 clips, one clip per rank and step, no collective
        frames, height, width = 65, 544, 960
        host_video = make_video((1, 3, frames, height, width), seed=1234 + rank).pin_memory()
        metric, scaling, frames_per_step = "vae_batched_encode_frames_per_sec_544x960_65f", "weak", frames * world
        workload = f"config3: enable_tiling(); encode(1x3x65x544x960) per rank (30 encoder sub-model calls per clip), clip-parallel"
        runner = None

        def step(x):
            return vae.encode(x).latent_dist.parameters
    else:
        host_video = make_video((1, 3, frames, height, width)).pin_memory()     # fp32 [-1,1], the dataset's .pt format
        workload = WORKLOAD if (frames, height, width) == (FRAMES, HEIGHT, WIDTH) else f"encode+decode 1x3x{frames}x{height}x{width}, tiled"

        def step(x):
            if runner is not None:
                return runner.roundtrip(x)
            post = vae.encode(x).latent_dist
            return vae.decode(post.mode()).sample
    video = host_video.to(dev, torch.bfloat16)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            out = step(video)
        # ---- timed region: K steps, device timed, inputs resident in HBM
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = N.launch_count()
        # Tiles run on `tile_streams` CUDA streams (vae/model.py run_tiles): per-kernel CUDA events then overlap across
        # streams and no longer measure a kernel on its own.  The timed region therefore runs uninstrumented (unless
        # --profile-in-region), and the roofline leg is measured right after it on the same inputs, serialised on ONE
        # stream, where a kernel's event pair brackets only that kernel.
        in_region = args.profile_in_region or vae.tile_streams <= 1
        if in_region:
            N.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = step(video)
        e1.record()
        barrier()
        launches_timed = N.launch_count() - launches0
        if in_region:
            prof = N.profile_end()
            prof_pass = {"where": "timed region", "tile_streams": vae.tile_streams, "steps": args.steps, "ms_per_step": None}
        else:
            n_prof = 1 if args.no_profile else min(args.steps, 2)
            saved_streams, vae.tile_streams = vae.tile_streams, 1
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            N.profile_begin()
            p0.record()
            for _ in range(n_prof):
                step(video)
            p1.record()
            barrier()
            prof = N.profile_end()
            vae.tile_streams = saved_streams
            for v in prof.values():   # scale to the timed region's step count: the JSON reports per-step figures
                for k in ("ms", "work", "executed"):
                    if k in v:
                        v[k] *= args.steps / n_prof
                v["launches"] *= args.steps / n_prof
            prof_pass = {"where": "extra steps right after the timed region, same inputs, all kernels serialised on one stream",
                         "tile_streams": 1, "steps": n_prof, "ms_per_step": p0.elapsed_time(p1) / n_prof}
        launches = launches_timed
        clocks = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_per_step = ms.item() / args.steps

        # ---- end to end: host fp32 clip (pinned) -> H2D -> encode+decode -> D2H of the reconstruction, every step
        host_out = torch.empty(tuple(out.shape), dtype=torch.bfloat16).pin_memory() if out is not None else None
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        # Every step copies ITS clip from pinned host memory and its result back; like a dataset loop (infer.py) the H2D of
        # step i+1 and the D2H of step i-1 run on side streams while step i computes.  All of it is inside t0..t1.
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)

        def fetch():
            with torch.cuda.stream(h2d):
                return host_video.to(dev, non_blocking=True).to(torch.bfloat16)

        n_e2e = 0 if args.no_e2e else args.steps
        t0.record()
        h2d.wait_stream(cur)
        x_next = fetch() if n_e2e else None
        for i in range(n_e2e):
            cur.wait_stream(h2d)
            x, x_next = x_next, None
            x.record_stream(cur)
            if i + 1 < n_e2e:
                x_next = fetch()
            o = step(x)
            if host_out is not None:
                d2h.wait_stream(cur)
                with torch.cuda.stream(d2h):
                    host_out.copy_(o, non_blocking=True)
                o.record_stream(d2h)
        cur.wait_stream(d2h)
        t1.record()
        barrier()
        ms2 = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_ms = ms2.item() / args.steps

    # the library books conv work with the STORED channel counts (3 -> 8 / 16 zero padded): rescale rank 0's share so that the
    # job total equals SURVEY 8d's count exactly (the attention projections are booked under their own class, attn_proj)
    lib_total = torch.tensor([prof["conv_tc"]["work"] + prof["conv_direct"]["work"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lib_total)
    survey_total = full_workload_flops(args.workload, frames, height, width) * args.steps * (world if args.workload == "config3" else 1)
    survey_ratio = survey_total / lib_total.item() if lib_total.item() > 0 else 1.0
    if rank == 0:
        tc = prof["conv_tc"]
        tc["library_work"] = tc["work"]
        tc["work"] *= survey_ratio
        prof["conv_direct"]["work"] *= survey_ratio
        tc_tflops = tc["work"] / (tc["ms"] * 1e9) if tc["ms"] > 0 else 0.0
        exec_tflops = tc.get("executed", tc["work"]) / (tc["ms"] * 1e9) if tc["ms"] > 0 else 0.0
        conv_flops = (prof["conv_tc"]["work"] + prof["conv_direct"]["work"]) / args.steps
        peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])  # kernel timed inside a long step
        shares = {k: round(v["ms"] / args.steps, 3) for k, v in prof.items()}
        line = {
            "metric": metric, "value": frames_per_step / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "bf16 model and I/O; fp16 tensor-core operands (bf16 weights convert exactly; Winograd-T / phase tap sums rounded once to fp16), fp32 accumulate",
            "data": "synthetic",
            "config": {"workload": workload,
                       "weights": "random-init, HY VAE config [128,256,512,512], 16 latent channels",
                       "l2": "inputs and activations far larger than the 126 MB L2 (713 MB clip); no flush needed",
                       "tile_streams": vae.tile_streams,
                       "partition": ("clips over ranks" if args.workload == "config3" else "tiles over ranks") if world > 1 else "single GPU",
                       "tile_exchange": (getattr(runner, "last_exchange", None) if runner is not None else None)},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "e2e": None if args.no_e2e else {"value": frames_per_step / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": host_video.numel() * 4, "d2h_bytes_per_step": int(host_out.numel() * 2)},
            "roofline": {"kernel": "tcgen05 implicit-GEMM CausalConv3d (conv_wino_kernel = Winograd F(2,3) along T for the stride-1 3x3x3 layers; "
                                   "conv_tc2_kernel sub-pixel phases / strided / k=1; conv_halo_kernel conv_in; conv_stack_kernel conv_out)",
                         "bound": "tensor", "achieved": tc_tflops, "peak": peak, "unit": "TFLOP/s", "frac": tc_tflops / peak,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['_source']}): the kernels run inside a multi-second step under the 1 kW cap",
                         "frac_of_burst_peak": tc_tflops / peaks["bf16_tflops"],
                         "achieved_definition": "reference conv FLOPs (SURVEY 8d: 2*Cout*Cin*k^3*voxels of every nn.Conv3d in the reference's tile "
                                                "decomposition) / sum of CUDA-event times of the conv launches in the timed region",
                         "work_definition_check": {"survey_8d_conv_tflop_per_step": survey_total / args.steps / 1e12,
                                                   "library_counted_tflop_per_step_all_ranks": lib_total.item() / args.steps / 1e12,
                                                   "note": "library counters use stored (zero-padded) channel counts; achieved uses SURVEY 8d's"},
                         "executed_tflops": exec_tflops,
                         "executed_frac_of_sustained_peak": exec_tflops / peak, "executed_frac_of_burst_peak": exec_tflops / peaks["bf16_tflops"],
                         "tensor_pipe_active_ncu": {"conv_wino_kernel 128->128": 0.906, "conv_wino_kernel 256->256": 0.936,
                                                    "source": "profiles/r02_ncu_conv_wino_{128,256}_v2.txt (sm__pipe_tensor_cycles_active, one launch each)"},
                         "executed_note": "the stride-1 3x3x3 convs run as Winograd F(2,3) along T (2T-1 instead of 3T plane GEMMs) and the "
                                          "post-upsample convs as sub-pixel phases over the low-res tensor (8/27 or 12/27 of the reference MACs), "
                                          "so executed < algorithmic and achieved exceeds the cuBLAS-measured peak",
                         "traffic": NCU_TRAFFIC["bytes"], "traffic_note": NCU_TRAFFIC["note"],
                         "launches_per_step": tc["launches"] / args.steps, "ms_per_step": tc["ms"] / args.steps,
                         "measured": prof_pass, "rank": 0},
            "conv_path": {"conv_tflop_per_step_rank0": conv_flops / 1e12,
                          "executed_conv_tflop_per_step_rank0": tc.get("executed", tc["work"]) / args.steps / 1e12,
                          "path_util_vs_sustained_peak": (conv_flops * (world if world > 1 else 1) / 1e12) / (ms_per_step / 1e3) / (peak * world)},
            "attention": {"kernel": "attn_fused_kernel (flash-style tcgen05: S, P in TMEM / shared memory; frame-causal block skipping)",
                          "ms_per_step": prof["attn"]["ms"] / args.steps, "launches_per_step": prof["attn"]["launches"] / args.steps,
                          "dense_tflop_per_step": prof["attn"]["work"] / args.steps / 1e12,
                          "achieved_dense_tflops": prof["attn"]["work"] / (prof["attn"]["ms"] * 1e9) if prof["attn"]["ms"] > 0 else 0.0,
                          "note": "SURVEY 8d attn_flops definition (4*L^2*D per call, dense, not halved for causality); the q/k/v/out "
                                  "projections are k=1 launches of the conv class"},
            "kernel_ms_per_step_rank0": shares,
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "config4":
            threads = os.cpu_count() or 1
            dt, fl = cpu_sample(threads)
            full = full_workload_flops()
            line["cpu_baseline"] = {
                "value": FRAMES / (dt * full / fl), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"oracle port (fp32): encode+decode {'x'.join(map(str, CPU_SAMPLE_SHAPE))}, {fl / 1e12:.2f} conv TFLOP in {dt:.1f} s, "
                          f"scaled by conv FLOPs to the {full / 1e12:.1f} TFLOP workload"}
        if world == 1 and not args.no_torch_baseline and args.workload == "config4":
            del video, out
            torch.cuda.empty_cache()
            try:
                tb = torch_gpu_baseline(dev, host_video)
                tb["ours_over_torch_cudnn"] = line["value"] / tb["value"]
            except Exception as e:  # reported, never fatal for the headline line
                tb = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            line["torch_gpu_baseline"] = tb
        print(json.dumps(line), flush=True)
    if world > 1:
        if runner is not None:
            runner.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the PyTorch/cuDNN leg (the reference's algorithm on the same GPU)")
    ap.add_argument("--workload", default="config4", choices=["config4", "config2", "config3"],
                    help="config4 (default, the headline): 720p x 129f encode+decode; config2: tiled decode only; config3: batched 544x960x65f encode")
    ap.add_argument("--no-profile", action="store_true", help="shorten the serialised roofline pass to one step")
    ap.add_argument("--profile-in-region", action="store_true",
                    help="record the per-launch CUDA events inside the timed region even with tile_streams > 1 (they then overlap)")
    ap.add_argument("--tile-streams", type=int, default=None, help="CUDA streams the tile sub-model calls are dealt over (default: the model's, 2)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (used for the ncu launch-list pass only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
