#!/usr/bin/env python
"""Benchmark of the B200-native Marigold-DC guided denoising loop (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference algorithm's CPU path (oracle port) on the host cores

A "step" is ONE GUIDED DDIM STEP of the hot path (marigold_dc.py:801-904): UNet forward, x0 prediction, VAE decode,
masked L1+L2 loss, backward through decoder and UNet, grad-norm rescale, Adam on latent/scale/shift, DDIM update.
Workload (config b of BASELINE.json): synthetic 480x640 RGB + 500 sparse points, resolution 768 (latent 72x96), bf16,
random-init SD2 UNet / VAE.  `value` = guided steps/s over all ranks with every input resident in HBM; `e2e` = the same
metric through the drop-in pipeline call with pinned HOST inputs and a host copy of the dense output, i.e. whole frames
(prologue + 50 steps + final decode + copies).  Frames are independent, so ranks shard frames (weak scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs.  "b" is the configuration the metric is quoted on and the default; the
# others are selected with --config and reported under their own workload name.
WORKLOADS = {
    "b": dict(H=480, W=640, resolution=768, n_points=500, frame_steps=50, max_depth=10.0, kind="nyu", batch=1,
              name="config b: NYUv2-shaped 480x640 RGB + 500 sparse points, resolution 768"),
    "b640": dict(H=480, W=640, resolution=640, n_points=500, frame_steps=50, max_depth=10.0, kind="nyu", batch=1,
                 name="config b at resolution 640: 480x640 RGB + 500 sparse points"),
    "c": dict(H=352, W=1216, resolution=1216, n_points=500, frame_steps=50, max_depth=80.0, kind="kitti", batch=1,
              name="config c: KITTI-shaped 352x1216 RGB + 64-line LiDAR-pattern sparsity (~5 %), resolution 1216 (native; SURVEY G9)"),
    "d": dict(H=480, W=640, resolution=768, n_points=500, frame_steps=50, max_depth=10.0, kind="nyu", batch=1, sequence=64,
              name="config d: 64-frame synthetic sequence (480x640, 500 points per frame, resolution 768) sharded over the ranks"),
    "e": dict(H=768, W=1024, resolution=1024, n_points=500, frame_steps=50, max_depth=10.0, kind="nyu", batch=4,
              name="config e: 768x1024 RGB + 500 sparse points, resolution 1024, batch of 4 frames per GPU"),
}
WORKLOAD = WORKLOADS["b"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=2, help="frames per rank for the e2e measurement")
    ap.add_argument("--tiny", action="store_true", help="narrow UNet/VAE on a 96x128 frame (debugging only; invalid as a bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--vae", default="original", choices=["original", "light"],
                    help="original = SD2 AutoencoderKL (BASELINE.json's configs); light = AutoencoderTiny, the reference CLI's default "
                         "(predict.py:44-52) -- a different workload, reported under its own name")
    ap.add_argument("--no-batch2", action="store_true", help="skip the informational two-frames-in-flight measurement")
    ap.add_argument("--config", default="b", choices=sorted(WORKLOADS), help="BASELINE.json workload (default b, the one the metric is quoted on)")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the informational torch-bf16 GPU baseline (oracle modules, cuDNN / cuBLAS / SDPA)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (nvidia-smi clocks line of B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ models / inputs
def make_models(device, tiny: bool, dtype=torch.bfloat16, seed=1234, vae_kind="original"):
    """Random-init SD2-derived UNet (in_channels 8) and SD2 VAE (or, vae_kind "light", the AutoencoderTiny the reference
    CLI uses by default); weights created directly on `device`."""
    from oracle.sd2_modules import (AutoencoderKL, UNet2DConditionModel, UNetConfig, VAEConfig, tiny_unet_config,
                                    tiny_vae_config)
    from oracle.marigold_dc import make_empty_text_embedding

    ucfg, vcfg = (tiny_unet_config(), tiny_vae_config()) if tiny else (UNetConfig(), VAEConfig())
    torch.manual_seed(seed)
    with torch.device(device):
        unet, vae = UNet2DConditionModel(ucfg), AutoencoderKL(vcfg)
        if vae_kind == "light":
            from oracle.taesd import AutoencoderTiny
            vae = AutoencoderTiny()
    unet, vae = unet.to(dtype).requires_grad_(False), vae.to(dtype).requires_grad_(False)
    ctx = make_empty_text_embedding(ucfg.cross_attention_dim, device=device, dtype=dtype)
    return unet, vae, ctx


def make_product_models(device, tiny: bool, vae_kind="original", seed=1234):
    """Random-init weights of the same architectures for OUR arm, generated by the product package itself
    (depth_completion_b200.synthetic.random_init_modules): nothing under oracle/ is imported on that arm."""
    from depth_completion_b200.config import UNetConfig, VAEConfig
    from depth_completion_b200.synthetic import random_init_modules

    if tiny:
        ucfg = UNetConfig(block_out_channels=(64, 128, 128, 128), attention_heads=(1, 2, 2, 2), cross_attention_dim=64)
        vcfg = VAEConfig(block_out_channels=(64, 64, 128, 128))
    else:
        ucfg, vcfg = UNetConfig(), VAEConfig()
    if vae_kind == "light":
        vcfg = VAEConfig(block_out_channels=(64, 64, 64, 64), layers_per_block=0, norm_num_groups=0, scaling_factor=1.0, kind="tiny")
    return random_init_modules(ucfg, vcfg, device, torch.bfloat16, seed)


def workload(tiny: bool, config: str = "b"):
    w = dict(WORKLOADS[config])
    if tiny:
        w.update(H=96, W=128, resolution=128, n_points=100)
        if "sequence" in w:
            w["sequence"] = 8
    return w


def config_dict(w, latent=None, vae_kind="original"):
    """Identical for both arms (the reference arm runs on our arm's config)."""
    lat = f" (latent {latent[0]}x{latent[1]})" if latent else ""
    if vae_kind == "light":
        return {"workload": f"{w['H']}x{w['W']} RGB + {w['n_points']} sparse points, resolution {w['resolution']}{lat}, "
                            f"{w['frame_steps']}-step guided completion, 1 frame in flight per GPU, random-init SD2 UNet (866M) / "
                            "AutoencoderTiny (2.4M; the reference CLI's default --vae light, NOT the BASELINE.json config)",
                "l2": "per-step working set >> 126 MB L2, no flush needed",
                "frames_sharding": "independent frames per rank, no collective inside the step"}
    return {"workload": f"{w['name']}{lat}, {w['frame_steps']}-step guided completion, {w.get('batch', 1)} frame(s) in flight per GPU, "
                        "random-init SD2 UNet (866M) / VAE decoder (49.5M)",
            "l2": "per-step working set (activations + 1.8 GB weights) >> 126 MB L2, no flush needed",
            "frames_sharding": "independent frames per rank, no collective inside the step"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU baseline
def _frames_for(w, n, seed=0):
    from depth_completion_b200.synthetic import make_batch

    return make_batch(n, H=w["H"], W=w["W"], kind=w.get("kind", "nyu"), n_points=w["n_points"], max_depth=w["max_depth"],
                      min_field=1.0 if w.get("kind") == "kitti" else 0.5, seed=seed)


def cpu_reference_steps(n_steps: int, warmup: int, config: str, tiny: bool, vae_kind: str = "original"):
    """Times guided steps of the reference algorithm (oracle port, fp32) on the host cores ON THE STATED CONFIGURATION:
    the same frame geometry, resolution and batch as our arm -- a bounded sample in the number of steps only."""
    from depth_completion_b200.synthetic import make_frame  # noqa: F401  (inputs come from the product's generator)
    from oracle.marigold_dc import OraclePipeline

    w = workload(tiny, config)
    # all host threads: torchrun exports OMP_NUM_THREADS=1 for nproc > 1, which would time a single core
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    torch.set_num_threads(max(1, ncpu))
    unet, vae, ctx = make_models("cpu", tiny, dtype=torch.float32, vae_kind=vae_kind)
    pipe = OraclePipeline(unet, vae, ctx)
    fr = _frames_for(w, w.get("batch", 1))
    times = []

    def tr(rec):
        times.append(time.perf_counter())

    t0 = time.perf_counter()
    pipe(fr["img"], fr["sparse"], fr["max_depth"], steps=w["frame_steps"], resolution=w["resolution"], trace=tr,
         max_steps=warmup + n_steps)
    stamps = [t0] + times
    dt = stamps[-1] - stamps[warmup]
    sps = n_steps * w.get("batch", 1) / dt  # a batch of B frames advances B guided frame-steps per iteration
    return dict(value=sps, cores=torch.get_num_threads(), seconds=dt,
                sample=f"{n_steps} guided step(s) after {warmup} warm-up (incl. the prologue) of the stated workload itself "
                       f"({w['H']}x{w['W']}, resolution {w['resolution']}, batch {w.get('batch', 1)}), fp32, oracle port of "
                       f"marigold_dc.py:801-904 with frozen weights, {dt:.1f} s; no scaling applied")


# ------------------------------------------------------------------------------------------------ main arms
REF_MAX_TIMED, REF_MAX_WARM = 6, 1  # a full-size CPU step is ~12 s on 16 cores: bound the run to a couple of minutes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    w = workload(args.tiny, args.config)
    k, wu = max(1, min(args.steps, REF_MAX_TIMED)), max(0, min(args.warmup, REF_MAX_WARM))
    r = cpu_reference_steps(k, wu, args.config, args.tiny, args.vae)
    from depth_completion_b200.config import processed_geometry
    ph, pw, pad_h, pad_w = processed_geometry(w["H"], w["W"], w["resolution"])
    line = {
        "impl": "reference", "metric": "guided_steps_per_sec", "value": r["value"], "unit": "steps/s",
        "n_gpus": args.gpus, "steps": k, "warmup": wu, "ms_per_step": 1e3 / r["value"],
        "requested": {"steps": args.steps, "warmup": args.warmup,
                      "note": f"full-size CPU steps are bounded to {REF_MAX_TIMED} timed + {REF_MAX_WARM} warm-up so the run ends within minutes"},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, vae_kind=args.vae, latent=((ph + pad_h) // 8, (pw + pad_w) // 8)),
        "cpu_baseline": {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def torch_gpu_baseline(dev, w, tiny, vae_kind, n_steps=6, warmup=2):
    """Informational: the SAME algorithm executed by PyTorch's own kernels on this GPU (oracle modules in bf16, frozen
    weights, cuDNN benchmark mode as predict.py:21 sets it, cuBLAS, SDPA) -- the kernel-class bar SURVEY.md section 2.1
    names.  Warm steps are excluded; CUDA-event timing."""
    from oracle.marigold_dc import OraclePipeline

    torch.backends.cudnn.benchmark = True
    unet, vae, ctx = make_models(dev, tiny, dtype=torch.bfloat16, vae_kind=vae_kind)
    pipe = OraclePipeline(unet, vae, ctx)
    fr = _frames_for(w, w.get("batch", 1))
    ev = []

    def tr(rec):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)

    pipe(fr["img"].to(dev), fr["sparse"].to(dev), fr["max_depth"], steps=w["frame_steps"], resolution=w["resolution"], trace=tr,
         max_steps=warmup + n_steps)
    torch.cuda.synchronize()
    ms = ev[warmup - 1].elapsed_time(ev[-1]) / n_steps
    del pipe, unet, vae
    torch.cuda.empty_cache()
    return {"value": 1e3 / ms * w.get("batch", 1), "unit": "steps/s", "ms_per_step": ms, "steps": n_steps, "warmup": warmup,
            "what": "oracle pipeline (restated diffusers modules) in torch bf16 on this GPU: cuDNN (benchmark=True) / cuBLAS / SDPA "
                    "kernels + autograd with frozen weights + torch.optim.Adam, python loop as in marigold_dc.py:801-904"}


def run_ours(args):
    import torch.distributed as dist

    from depth_completion_b200.flops import step_flops
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline, shard_frames

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    w = workload(args.tiny, args.config)
    B = w.get("batch", 1)                       # frames per pipeline call
    seq = w.get("sequence")                      # config d: a fixed frame set sharded over the ranks (strong scaling)
    unet, vae, ctx = make_product_models(dev, args.tiny, args.vae)   # no oracle code on this arm
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    H, W, res, fs = w["H"], w["W"], w["resolution"], w["frame_steps"]
    if seq:
        # every rank generates the same sequence and takes its contiguous shard
        mine = shard_frames(seq, rank, world)
        fr_all = _frames_for(w, seq)
        fr = {k: (v[mine.start:mine.stop] if torch.is_tensor(v) else v) for k, v in fr_all.items()}
        n_calls = len(mine)
    else:
        # every rank owns its own frames (seed offset by rank): frames are independent
        n_calls = max(1, args.frames)
        fr = _frames_for(w, n_calls * B, seed=100 * rank)
    imgs_h, sparses_h = fr["img"].pin_memory(), fr["sparse"].pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- one full call: builds the engine for these shapes, packs weights, warms every kernel up
    dense, _ = pipe(imgs_h[:B].to(dev), sparses_h[:B].to(dev), w["max_depth"], steps=fs, resolution=res)
    assert torch.isfinite(dense).all()
    eng = list(pipe._engines.values())[-1]

    # ---- device-resident timing of K guided steps (inputs already in HBM): re-begin a frame, W warm-up, K timed
    def begin_frame():
        # the pipeline's own prologue, then the engine is left at step 0 with everything resident
        pipe(imgs_h[:B].to(dev), sparses_h[:B].to(dev), w["max_depth"], steps=fs, resolution=res, _begin_only=True)

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin_frame()
    done, left_w = 0, args.warmup
    while left_w > 0:  # warm-up steps
        n = min(left_w, fs - done)
        eng.run(n)
        done += n
        left_w -= n
        if done == fs:
            begin_frame()
            done = 0
    total_ms, left = 0.0, args.steps
    launches0 = eng.launch_count()
    barrier()
    with ClockSampler(local) as clk:
        while left > 0:
            n = min(left, fs - done)
            e0.record()
            eng.run(n)
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            done += n
            left -= n
            if done == fs and left > 0:
                begin_frame()
                done = 0
    barrier()
    launches = eng.launch_count() - launches0
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    # a step of a B-frame call advances B frames by one guided step each: `value` counts frame-steps, as at B = 1
    steps_per_s = world * B * args.steps / (total_ms * 1e-3)

    accuracy = None
    # ---- end to end: pinned host inputs -> pipeline call (H2D, prologue, 50 steps, final decode) -> host dense
    e2e = None
    if not args.no_e2e:
        out_h = torch.empty(n_calls * B, 1, H, W, dtype=torch.float32).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for i in range(n_calls):
            sl = slice(i * B, (i + 1) * B)
            img_d = imgs_h[sl].to(dev, non_blocking=True)
            sp_d = sparses_h[sl].to(dev, non_blocking=True)
            d, _ = pipe(img_d, sp_d, w["max_depth"], steps=fs, resolution=res)
            out_h[sl].copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:  # gather the dense maps on the device (the only collective, outside the step)
            mine_d = out_h.to(dev)
            if seq:    # shards may differ by one frame: pad to the longest
                longest = (seq + world - 1) // world
                pad = torch.zeros(longest, 1, H, W, device=dev)
                pad[: mine_d.shape[0]] = mine_d
                mine_d = pad
            gl = [torch.empty_like(mine_d) for _ in range(world)]
            dist.all_gather(gl, mine_d)
            torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = dt.item()
        frames = seq if seq else world * n_calls * B
        e2e = {"value": frames * fs / dt, "unit": "steps/s", "frames_per_sec": frames / dt, "frames": frames,
               "sec_per_frame": dt / (n_calls * B), "sec_per_call": dt / n_calls, "seconds": dt,
               "h2d_bytes_per_step": (imgs_h[0].numel() + sparses_h[0].numel() * 4) / fs,
               "d2h_bytes_per_step": H * W * 4 / fs}
        accuracy = holdout_accuracy(out_h, fr, n_calls * B)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM / implicit-GEMM conv): FLOPs of its launches in one step
    #      / the sum of their in-situ durations (CUDA events around every launch of an instrumented step)
    burst, sustained, hbm, how = peaks()
    prof = eng.profile_gemm_step()
    sf = step_flops(pipe.unet_cfg, pipe.vae_cfg, H, W, res)
    # DRAM bytes per launch of that kernel from the committed ncu pass over one guided step of this workload
    # (profiles/summarize_launches.py); only valid for the full-size default workload
    traffic, tnote = None, None
    for tp in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tp)
        if not args.tiny and args.vae == "original" and args.config == "b" and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            tnote = f"dram__bytes_read+write per launch, mean over the step's launches (profiles/{tp})"
            break
    step_ms = total_ms / args.steps
    roof = {"bound": "tensor", "kernel": "umma_gemm_pair_kernel + umma_gemm_kernel (gemm.cuh, cta_group::2 / ::1)",
            "achieved": prof["tflops"], "peak": sustained,
            "unit": "TFLOP/s", "frac": prof["tflops"] / sustained, "traffic": traffic, "traffic_note": tnote,
            "peak_source": how + " sustained",
            "launches_per_step": prof["launches"], "kernel_ms_per_step": prof["ms"], "kernel_flops_per_step": prof["flops"],
            "step_algorithmic_tflops": B * sf["step"] / 1e12,
            "step_frac_of_peak": B * sf["step"] / 1e12 / (step_ms * 1e-3) / sustained,
            "step_frac_of_burst_peak": B * sf["step"] / 1e12 / (step_ms * 1e-3) / burst}

    dev_mem_gb = eng.device_bytes() / 2 ** 30
    # ---- informational: the same call with two frames in flight per GPU (the reference's --batch-size 2): the UNet's
    #      low-resolution layers are weight-streaming / launch-latency bound at one frame, so a second frame is cheap
    two = None
    if not args.no_e2e and not args.no_batch2 and B == 1 and not seq and n_calls >= 2:
        barrier()
        pipe(imgs_h[:2].to(dev), sparses_h[:2].to(dev), w["max_depth"], steps=fs, resolution=res, _begin_only=True)  # builds the N=2 engine
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2, _ = pipe(imgs_h[:2].to(dev, non_blocking=True), sparses_h[:2].to(dev, non_blocking=True), w["max_depth"], steps=fs,
                     resolution=res)
        out_h[:2].copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt2 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
        two = {"frames_in_flight_per_gpu": 2, "frames_per_sec": 2 * world / dt2.item(), "sec_per_frame_pair": dt2.item()}

    # ---- informational: two batch-1 calls in flight on two engines / two CUDA streams (video.complete_sequence): frame
    #      k+1's prologue and loop overlap frame k's loop; same frames, same arithmetic, one weight bank
    inflight = None
    if not args.no_e2e and not args.no_batch2 and B == 1 and not seq and n_calls >= 2:
        from depth_completion_b200.video import complete_sequence

        reps = 2  # 2 x n_calls frames
        imgs_d = torch.cat([imgs_h[:n_calls].to(dev)] * reps)
        sps_d = torch.cat([sparses_h[:n_calls].to(dev)] * reps)
        complete_sequence(pipe, imgs_d[:2], sps_d[:2], w["max_depth"], frames_in_flight=2, steps=fs, resolution=res)  # builds both engines
        barrier()
        t0 = time.perf_counter()
        complete_sequence(pipe, imgs_d, sps_d, w["max_depth"], frames_in_flight=2, steps=fs, resolution=res)
        torch.cuda.synchronize()
        dt3 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt3, op=dist.ReduceOp.MAX)
        inflight = {"frames_in_flight_per_gpu": 2, "engines": 2, "frames": int(imgs_d.shape[0]) * world,
                    "frames_per_sec": int(imgs_d.shape[0]) * world / dt3.item(), "seconds": dt3.item()}

    torch_base = None
    if rank == 0 and world == 1 and not args.no_torch_baseline:
        pipe._invalidate()  # free our workspace first: the two arms never share the GPU
        try:
            torch_base = torch_gpu_baseline(dev, w, args.tiny, args.vae)
        except Exception as e:  # noqa: BLE001 -- informational only, never fails the bench line
            torch_base = {"error": repr(e)[:300]}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            r = cpu_reference_steps(2, 1, args.config, args.tiny, args.vae)
            cpu = {"value": r["value"], "unit": "steps/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {
            "metric": "guided_steps_per_sec", "value": steps_per_s, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if seq else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(w, (eng.lh, eng.lw), args.vae),
            "frames_per_sec_device": steps_per_s / fs,
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "torch_gpu_baseline": torch_base,
            "e2e_two_frames_in_flight": two,
            "e2e_two_calls_in_flight": inflight,
            "accuracy": accuracy,
            "device_mem_gb": dev_mem_gb,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def holdout_accuracy(dense_h, fr, n_frames):
    """MAE / RMSE (utils.py:692-739) of rank 0's end-to-end dense maps in metres: at the hold-out points the loop never
    saw and at the guidance points it fitted.  With random-init weights these are parity numbers (the GPU tests compare
    them with the oracle's), not quality numbers.  Never raises: the bench line must not depend on it."""
    try:
        gt = fr["gt"][:n_frames].float()
        err = dense_h[:n_frames].float() - gt
        hold, seen = fr["holdout"][:n_frames].bool(), fr["sparse"][:n_frames] > 0
        return {"holdout_mae_m": err[hold].abs().mean().item(), "holdout_rmse_m": err[hold].pow(2).mean().sqrt().item(),
                "guided_mae_m": err[seen].abs().mean().item(), "holdout_points": int(hold.sum()), "guided_points": int(seen.sum()),
                "note": "random-init weights: parity numbers, not quality"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
