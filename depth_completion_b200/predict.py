"""`python -m depth_completion_b200.predict SRC_ROOT DST_ROOT [options]`: the command line of `/root/reference/predict.py`
on the B200 pipeline.

Same arguments, option names, defaults and fall-backs as `predict.py:25-343` (option table) and `:385-455` (checks), the
same outputs (`<dst>/<dataset>/dense/<rel>.{bl2,npz,npy}` and `<dst>/<dataset>/vis/<rel stem>_vis.jpg`, `:717-764`).  What
differs, because the hot path behind it differs:

* `--precision fp32`, `--model lcm`, `--train-method per-input` and `--closed-form True` together with `edge` / `smooth`
  name arithmetic this library does not have: they stop with a CRITICAL message instead of silently
  running the default arm.  Everything else of the option table runs in the library, including `--interp-mode nearest`,
  `--train-latents False` (plain DDIM sampling + closed-form scale / shift), the three optimisers, projections and norms.
* `--compile-graph` / `--compile-mode` are accepted and ignored: the guided step is always one captured CUDA graph.
* `--use-segmask`: the reference loads the masks and never hands them to the pipeline call (`predict.py:659-694`); here
  the flag only checks that the files exist, like the reference's pairing step does (`:520-571`).
* Extensions (not in the reference): `--weights synthetic` builds random-init modules of the SD2 architecture so the
  command runs on an air-gapped box (no checkpoint download); `--marigold-ckpt` / `--vae-ckpt` point `from_pretrained`
  at local directories; `--overlap-prologue` encodes batch k+1 under batch k's guided steps; under `torchrun` the
  independent frames of every dataset are sharded across the ranks.

`run()` is the body without click and without model construction, so tests can drive it with any pipeline object.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import click
import torch

from . import dataset_io as dio
from . import vis as visual
from .cli_common import LOG_LEVELS, CommaSeparated, get_logger
from .pipeline import SUPPORTED_LOSS_FUNCS

MARIGOLD_CKPT_ORIGINAL = "prs-eth/marigold-v1-0"   # marigold_dc.py:16
MARIGOLD_CKPT_LCM = "prs-eth/marigold-lcm-v1-0"    # marigold_dc.py:17
VAE_CKPT_LIGHT = "madebyollin/taesd"               # marigold_dc.py:18
VIEWS = ("image", "sparse", "dense")


class OptionError(Exception):
    """An option combination the reference exits on (`logger.critical` + `sys.exit(1)`)."""


def resolve_options(o: dict, log) -> dict:
    """The checks and fall-backs of predict.py:396-455 on the parsed options, in the reference's order."""
    o = dict(o)
    if o["vis"]:
        order = []
        for view in o["vis_order"]:
            if view not in VIEWS:
                log.error(f"Invalid order (skipped): {view}")
                continue
            order.append(view)
        if not order:
            raise OptionError("No valid visualization order specified")
        o["vis_order"] = order
    kept = []
    for f in o["loss_funcs"]:
        if f not in SUPPORTED_LOSS_FUNCS:
            log.error(f"Invalid loss function (skipped): {f}")
        else:
            kept.append(f)
    o["loss_funcs"] = kept
    if o["use_prev_latent"] and o["batch_size"] > 1:
        log.warning("Currently, batch_size is forced to 1 when use_prev_latent=True. This will be fixed in the future")
        o["batch_size"] = 1
    if (o["projection"] in ("log", "log10") or o["inv"]) and o["norm"] == "const":
        log.error("norm=const is not allowed when projection=log or log10. Falling back to norm=minmax")
        o["norm"] = "minmax"
    if o["model"] == "lcm" and o["train_latents"]:
        log.error("LCM-based Marigold model does not support trainable latents (train_latents=True). "
                  "Falling back to train_latents=False")
        o["train_latents"] = False
    if not o["train_latents"] and not o["closed_form"]:
        log.error("When trainable latentes are not used, closed-form solution must be enabled. "
                  "Falling back to closed_form=True")
        o["closed_form"] = True
    # what the B200 path does not compute: refuse, never substitute
    if o["precision"] != "bf16":
        raise OptionError("--precision fp32: the B200 engine computes in bfloat16 with fp32 accumulation only")
    if o["model"] == "lcm":
        raise OptionError("--model lcm: the LCM scheduler arm is outside the B200 hot path (DDIM, v-prediction)")
    if o["train_latents"] and o["train_method"] != "per-step":
        raise OptionError("--train-method per-input is outside the B200 hot path (and fails upstream: SURVEY.md section 2)")
    if o["train_latents"] and o["closed_form"] and any(f in ("edge", "smooth") for f in o["loss_funcs"]):
        raise OptionError("--closed-form True with edge / smooth losses is not built (gradient of the dense terms through the fit)")
    if o["compile_graph"]:
        log.warning("--compile-graph is ignored: the guided step always runs as one captured CUDA graph")
    if o["compress"] is None:
        o["compress"] = "bl2" if dio.have_blosc2() else "npz"
        if o["compress"] == "npz" and o["save_dense"]:
            log.warning("blosc2 is not installed: dense maps are written as .npz instead of the default .bl2")
    elif o["compress"] == "bl2" and o["save_dense"] and not dio.have_blosc2():
        raise OptionError("--compress bl2 needs the blosc2 package, which is not installed")
    return o


def pipe_kwargs_of(o: dict) -> dict:
    """The keyword arguments of the pipeline call, as predict.py:669-694 passes them."""
    return dict(min_depth=o["min_depth"], projection=o["projection"], inv=o["inv"], norm=o["norm"],
                percentile=tuple(o["percentile"]), steps=o["steps"], resolution=o["res"], interp_mode=o["interp_mode"],
                loss_funcs=o["loss_funcs"], opt=o["opt"], lr=(o["lr_latent"], o["lr_scaling"]), kld=o["kld"],
                kld_mode=o["kld_mode"], kld_weight=o["kld_weight"], closed_form=o["closed_form"],
                train_latents=o["train_latents"], train_method=o["train_method"], train_steps=o["train_steps"])


def check_segmasks(dataset_dirs, log) -> None:
    """`--use-segmask True`: the existence checks of predict.py:520-571 (the masks themselves never reach the call)."""
    for ds in dataset_dirs:
        seg = ds / "segmask"
        if not seg.exists():
            log.error(f"No segmentation directory found at {seg}. Segmentation masks will not used for {ds.name}")
        elif not (seg / "map.csv").exists():
            log.error(f"No segmentation mapping file found at {seg / 'map.csv'}. "
                      f"Segmentation masks will not be used for {ds.name}")


def run(pipe, src_root: Path, dst_root: Path, o: dict, log, rank: int = 0, world: int = 1, progress: bool = True) -> dict:
    """Everything after model construction (predict.py:500-770) for already-resolved options `o`."""
    src_root, dst_root = Path(src_root), Path(dst_root)
    dataset_dirs = dio.find_dataset_dirs(src_root)
    if not dataset_dirs:
        raise OptionError(f"No dataset directories found at {src_root}")
    log.info(f"Found {len(dataset_dirs):,} dataset directories")
    if o["use_segmask"]:
        check_segmasks(dataset_dirs, log)
    for ds in dataset_dirs:
        n = len(dio.find_pairs(ds))
        if n == 0:
            raise OptionError("No valid input pairs found")
        log.info(f"Found {n:,} input pairs for {ds.name}")
    dst_root.mkdir(parents=True, exist_ok=True)

    bars, t_vis = {}, [0.0]

    def bar_of(ds):
        if ds not in bars and progress:
            import tqdm

            idx = dataset_dirs.index(ds)
            for b in bars.values():
                b.close()
            bars[ds] = tqdm.tqdm(total=len(dio.find_pairs(ds)), dynamic_ncols=True,
                                 desc=f"{idx + 1}/{len(dataset_dirs)} - {ds.name}", disable=rank != 0)
        return bars.get(ds)

    def on_batch(ds, n_frames, seconds):
        b = bar_of(ds)
        if b is not None:
            b.set_postfix({"time/infer": seconds, "time/vis": t_vis[0]})
            b.update(n_frames * world)
        t_vis[0] = 0.0

    def on_frame(ds, img_rel, sparse_rel, img, sparse, dense):
        if not o["vis"]:
            return
        t0 = time.time()
        panels = visual.frame_panels(img, sparse, dense, o["vis_order"], max_depth=o["max_depth"], min_depth=o["min_depth"])
        grid = visual.make_grid(panels, resize=tuple(o["vis_res"]))
        out_dir = dst_root / (ds.relative_to(src_root) if ds != src_root else Path("."))
        path = (out_dir / "vis" / img_rel).parent / f"{img_rel.stem}_vis.jpg"     # predict.py:759-764
        visual.save_img_tensor(grid, path)
        t_vis[0] += time.time() - t0

    saved = dio.complete_dataset(pipe, src_root, dst_root, o["max_depth"], o["max_sparse_depth"], batch_size=o["batch_size"],
                                 use_prev_latent=o["use_prev_latent"], beta=o["beta"], compress=o["compress"], rank=rank,
                                 world=world, save_dense=o["save_dense"], on_frame=on_frame, on_batch=on_batch,
                                 overlap_prologue=o.get("overlap_prologue", False), **pipe_kwargs_of(o))
    for b in bars.values():
        b.close()
    for ds in dataset_dirs:
        log.success(f"Finished processing {ds.name}")
    log.success(f"Finished processing all {len(dataset_dirs):,} datasets")
    return saved


def build_pipeline(o: dict, log, device="cuda"):
    """predict.py:457-498: Marigold checkpoint, optional AutoencoderTiny, trailing DDIM scheduler.  `--weights synthetic`
    replaces the two checkpoint reads with random-init modules of the same architectures (no network)."""
    from .pipeline import MarigoldDepthCompletionPipeline

    if o["weights"] == "synthetic":
        from .config import UNetConfig, VAEConfig
        from .synthetic import random_init_modules

        vcfg = VAEConfig()
        if o["vae"] == "light":
            vcfg = VAEConfig(block_out_channels=(64, 64, 64, 64), layers_per_block=0, norm_num_groups=0, scaling_factor=1.0,
                             kind="tiny")
        unet, vae, ctx = random_init_modules(UNetConfig(), vcfg, device)
        pipe = MarigoldDepthCompletionPipeline(unet, vae, prediction_type="depth")
        pipe.empty_text_embedding = ctx
        log.warning("--weights synthetic: random-init SD2 UNet / VAE, the dense maps are not depth estimates")
        return pipe
    ckpt = o["marigold_ckpt"] or MARIGOLD_CKPT_ORIGINAL
    pipe = MarigoldDepthCompletionPipeline.from_pretrained(ckpt, prediction_type="depth", torch_dtype=torch.bfloat16).to(device)
    if o["vae"] == "light":
        from diffusers import AutoencoderTiny

        pipe.vae = AutoencoderTiny.from_pretrained(o["vae_ckpt"] or VAE_CKPT_LIGHT, torch_dtype=torch.bfloat16).to(device)
    from diffusers import DDIMScheduler

    pipe.scheduler = DDIMScheduler.from_config(pipe.scheduler.config, timestep_spacing="trailing")
    return pipe


_B = dict(type=bool, show_default=True)


@click.command(help="Predict dense depth maps from sparse depth maps and camera images.")
@click.argument("src_root", type=click.Path(exists=True, path_type=Path, file_okay=False, dir_okay=True))
@click.argument("dst_root", type=click.Path(exists=False, path_type=Path))
@click.option("--model", type=click.Choice(["original", "lcm"]), default="original", show_default=True,
              help="Marigold model (only `original` runs on the B200 path).")
@click.option("--vae", type=click.Choice(["original", "light"]), default="light", show_default=True,
              help=f"original - the Stable Diffusion VAE of Marigold; light - {VAE_CKPT_LIGHT}.")
@click.option("-n", "--steps", type=click.IntRange(min=1), default=50, show_default=True, help="Number of denoising steps.")
@click.option("-r", "--res", type=click.IntRange(min=1), default=768, show_default=True,
              help="Input images are resized so that the longer side has this length.")
@click.option("--norm", type=click.Choice(["const", "minmax", "percentile"]), default="const", show_default=True,
              help="Normalisation of the sparse depth maps.")
@click.option("--percentile", type=CommaSeparated(float), default="0.01,0.99", show_default=True,
              help="min,max percentiles for --norm=percentile.")
@click.option("--max-sparse-depth", type=click.FloatRange(min=0, min_open=True), default=120.0, show_default=True,
              help="Max distance [m] of the input sparse depth maps (decodes the range images).")
@click.option("--max-depth", type=click.FloatRange(min=0, min_open=True), default=120.0, show_default=True,
              help="Max distance [m] of the output dense depth maps.")
@click.option("--min-depth", type=click.FloatRange(min=0), default=0.0, show_default=True,
              help="Min distance [m] of the output dense depth maps.")
@click.option("-v", "--vis", default=True, help="Whether to save visualisations.", **_B)
@click.option("-vr", "--vis-res", type=click.Tuple([int, int]), default=(512, -1), show_default=True,
              help="Resolution (height, width) of the visualisation; -1 keeps the aspect ratio.")
@click.option("-vo", "--vis-order", type=CommaSeparated(str), default="image,sparse,dense", show_default=True,
              help="Panels of the visualisation, left to right.")
@click.option("--save-dense", default=True, help="Whether to save the dense depth maps.", **_B)
@click.option("--log", type=click.Path(path_type=Path), default=None, help="Path to save logs.")
@click.option("--log-level", type=click.Choice(LOG_LEVELS), default="INFO", show_default=True)
@click.option("-p", "--precision", type=click.Choice(["bf16", "fp32"]), default="bf16", show_default=True)
@click.option("-c", "--compress", type=click.Choice(["npz", "bl2", "npy"]), default=None,
              help="File format of the dense maps  [default: bl2; npz where blosc2 is not installed]")
@click.option("--compile-graph", default=False, help="Ignored (the step is always one CUDA graph).", **_B)
@click.option("--compile-mode", type=click.Choice(["max-autotune", "reduce-overhead", "default"]), default="reduce-overhead",
              help="Ignored.")
@click.option("--interp-mode", type=click.Choice(["bilinear", "nearest"]), default="bilinear", show_default=True)
@click.option("--loss-funcs", type=CommaSeparated(str), default="l1,l2", show_default=True,
              help=f"Comma-separated loss functions out of {','.join(SUPPORTED_LOSS_FUNCS)}.")
@click.option("--opt", type=click.Choice(["adam", "sgd", "adagrad"]), default="adam", show_default=True)
@click.option("--lr-latent", type=click.FloatRange(min=0, min_open=True), default=0.05, show_default=True)
@click.option("--lr-scaling", type=click.FloatRange(min=0, min_open=True), default=0.005, show_default=True)
@click.option("--kld", default=False, help="KL-divergence penalty on the latent.", **_B)
@click.option("--kld-mode", type=click.Choice(["simple", "strict"]), default="simple", show_default=True)
@click.option("--kld-weight", type=click.FloatRange(min=0, min_open=True), default=0.1, show_default=True)
@click.option("-bs", "--batch-size", type=click.IntRange(min=1), default=1, show_default=True)
@click.option("--use-prev-latent", default=False, help="Blend the previous frame's latent into the next start.", **_B)
@click.option("--beta", type=click.FloatRange(min=0, min_open=True), default=0.9, show_default=True)
@click.option("--use-segmask", default=False, help="Check segmentation masks (they never reach the call).", **_B)
@click.option("--closed-form", default=False, help="Closed-form scale / shift instead of the learned ones.", **_B)
@click.option("--projection", type=click.Choice(["linear", "log", "log10"]), default="linear", show_default=True)
@click.option("--inv", default=False, help="Work on inverse depth.", **_B)
@click.option("--train-latents", default=True, **_B)
@click.option("--train-method", type=click.Choice(["per-step", "per-input"]), default="per-step", show_default=True)
@click.option("--train-steps", type=click.IntRange(min=1), default=10, show_default=True)
@click.option("--weights", type=click.Choice(["pretrained", "synthetic"]), default="pretrained", show_default=True,
              help="[extension] synthetic = random-init SD2 modules, no checkpoint needed.")
@click.option("--marigold-ckpt", type=str, default=None, help="[extension] local path / hub id of the Marigold checkpoint.")
@click.option("--vae-ckpt", type=str, default=None, help="[extension] local path / hub id of the light VAE checkpoint.")
@click.option("--overlap-prologue", default=False, help="[extension] encode batch k+1 under batch k's guided steps.", **_B)
def main(src_root: Path, dst_root: Path, **o) -> None:
    log = get_logger(o["log_level"], o["log"])
    if o["log"] is not None:
        log.info(f"Saving logs to {o['log']}")
    if not torch.cuda.is_available():
        log.critical("CUDA must be available to run this script.")
        sys.exit(1)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    try:
        o = resolve_options(o, log)
        if world > 1 and o["use_prev_latent"]:
            raise OptionError("--use-prev-latent chains the frames of a dataset serially: run it on one rank")
        pipe = build_pipeline(o, log, device)
        run(pipe, src_root, dst_root, o, log, rank=rank, world=world)
    except OptionError as e:
        log.critical(str(e))
        sys.exit(1)


if __name__ == "__main__":
    main()
