"""Seeded synthetic inputs for the BASELINE.json configs.

There are no datasets or checkpoints offline, so every benchmark and parity test runs on:
  * a uint8 RGB image (uniform noise),
  * a smooth synthetic ground-truth depth field (plane + low-frequency sinusoids),
  * sparse guidance = the field sampled at random pixels (NYU-like) or on LiDAR-like scan lines (KITTI-like),
  * a disjoint hold-out sample of the same field for MAE/RMSE.
"""
from __future__ import annotations

import math

import torch


def depth_field(H: int, W: int, lo: float, hi: float, seed: int, shift_px: float = 0.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    ph = torch.rand(6, generator=g) * 2 * math.pi
    fr = 1.0 + torch.rand(6, generator=g) * 2.0
    ys = torch.linspace(0, 1, H).view(H, 1)
    xs = (torch.linspace(0, 1, W).view(1, W) + shift_px / W)
    f = 0.5 + 0.25 * (ys - 0.5) + 0.15 * (xs - 0.5)
    f = f + 0.08 * torch.sin(2 * math.pi * fr[0] * xs + ph[0]) * torch.cos(2 * math.pi * fr[1] * ys + ph[1])
    f = f + 0.05 * torch.sin(2 * math.pi * fr[2] * (xs + ys) + ph[2])
    f = (f - f.min()) / (f.max() - f.min())
    return (lo + (hi - lo) * f).view(1, 1, H, W)


def random_points_mask(H: int, W: int, n: int, seed: int, exclude: torch.Tensor | None = None) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(H * W, generator=g)
    if exclude is not None:
        perm = perm[~exclude.view(-1)[perm]]
    m = torch.zeros(H * W, dtype=torch.bool)
    m[perm[:n]] = True
    return m.view(1, 1, H, W)


def lidar_mask(H: int, W: int, lines: int, keep: float, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    rows = torch.linspace(H // 3, H - 1, lines).round().long().unique()
    m = torch.zeros(1, 1, H, W, dtype=torch.bool)
    m[0, 0, rows] = torch.rand(len(rows), W, generator=g) < keep
    return m


def make_frame(H=480, W=640, kind="nyu", n_points=500, max_depth=10.0, min_field=0.5, seed=0, shift_px=0.0):
    """Returns dict(img uint8 [1,3,H,W], sparse [1,1,H,W] fp32, gt, holdout_mask, max_depth)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (1, 3, H, W), generator=g, dtype=torch.uint8)
    gt = depth_field(H, W, min_field, max_depth, seed + 1, shift_px)
    if kind == "nyu":
        mask = random_points_mask(H, W, n_points, seed + 2)
    elif kind == "kitti":
        mask = lidar_mask(H, W, 64, 0.275, seed + 2)
    else:
        raise ValueError(kind)
    hold = random_points_mask(H, W, n_points, seed + 3, exclude=mask)
    sparse = torch.where(mask, gt, torch.zeros_like(gt))
    return dict(img=img, sparse=sparse, gt=gt, holdout=hold, max_depth=max_depth)


def make_batch(n_frames: int, **kw):
    """Frames of a synthetic sequence: the depth field translates a few pixels per frame."""
    seed = kw.pop("seed", 0)
    frames = [make_frame(seed=seed + 10 * i, shift_px=3.0 * i, **kw) for i in range(n_frames)]
    out = {k: torch.cat([f[k] for f in frames], 0) for k in ("img", "sparse", "gt", "holdout")}
    out["max_depth"] = frames[0]["max_depth"]
    return out


class ParamBag:
    """Stands in for a diffusers module where only `.config` and `.state_dict()` are read (which is all the drop-in
    pipeline class reads): random-init weights of a given architecture without instantiating any model code."""

    def __init__(self, cfg, state: dict):
        self.config = cfg
        self._state = state

    def state_dict(self):
        return self._state


def random_init_modules(unet_cfg, vae_cfg, device, dtype=torch.bfloat16, seed: int = 1234):
    """(unet, vae, empty_text_embedding) with random weights of the given architecture (BASELINE.json: "random-init SD2
    UNet/VAE").  Parameter names and shapes come from the library itself (`mdc_weight_key` / `mdc_weight_shape` of a
    throw-away handle on a small geometry).  Matrices ~ U(-1, 1) / sqrt(fan_in) (PyTorch's default bound), norm scales
    1, biases small -- enough for finite, well-scaled activations; throughput does not depend on the values."""
    from .engine import StepEngine

    dev = torch.device(device)
    eng = StepEngine(unet_cfg, vae_cfg, 1, 64, 64, 64, 1, dev)
    shapes = eng.weight_shapes()
    eng.close()
    g = torch.Generator(device=dev).manual_seed(seed)
    sds = {"unet": {}, "vae": {}}
    for key in sorted(shapes):
        shp = shapes[key]
        prefix, name = key.split(".", 1)
        if len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = (torch.rand(shp, device=dev, generator=g) * 2 - 1) / fan_in ** 0.5
        elif name.endswith("weight"):   # GroupNorm / LayerNorm scale
            t = torch.ones(shp, device=dev)
        else:
            t = (torch.rand(shp, device=dev, generator=g) * 2 - 1) * 0.02
        sds[prefix][name] = t.to(dtype)
    ctx = torch.randn(1, 2, unet_cfg.cross_attention_dim, device=dev, generator=g).to(dtype)
    return ParamBag(unet_cfg, sds["unet"]), ParamBag(vae_cfg, sds["vae"]), ctx
