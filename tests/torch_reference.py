"""TEST INFRASTRUCTURE (not part of the product package): the per-frame prologue of marigold_dc.py:659-756 -- image
normalise / resize / pad, the VAE *encoder* forward from a state dict, sparse-depth normalisation -- as plain PyTorch ops.
The product does all of this inside libmdc_b200.so (`mdc_begin_frame`); the GPU tests use these functions as the torch
(bf16 / fp32) reference for it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def preprocess_image(image: torch.Tensor, resolution: int, dtype):
    """MarigoldImageProcessor.preprocess semantics."""
    if image.ndim != 4:
        raise ValueError(f"Input image is not 4-dimensional: shape={tuple(image.shape)}")
    if not torch.is_floating_point(image):
        if image.dtype != torch.uint8:
            raise ValueError(f"Image dtype={image.dtype} is not supported.")
        image = image.to(dtype) / 255
    else:
        image = image.to(dtype)
    if image.shape[1] == 1:
        image = image.repeat(1, 3, 1, 1)
    if image.shape[1] != 3:
        raise ValueError(f"Input image is not 1- or 3-channel: {tuple(image.shape)}.")
    if image.min().item() < 0.0 or image.max().item() > 1.0:
        raise ValueError("Input image data is partially outside of the [0,1] range.")
    image = image * 2.0 - 1.0
    h, w = image.shape[-2:]
    m = max(h, w)
    nh, nw = h * resolution // m, w * resolution // m
    if nh == 0 or nw == 0:
        raise ValueError(f"Extreme aspect ratio of the input image: [{w} x {h}]")
    image = F.interpolate(image, (nh, nw), mode="bilinear", antialias=True)
    ph, pw = -nh % 8, -nw % 8
    return F.pad(image, (0, pw, 0, ph), mode="replicate"), (ph, pw)


def _gn(x, sd, key, groups, eps=1e-6):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps)


def _conv(x, sd, key, **kw):
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], **kw)


def _resnet(x, sd, key, groups):
    h = _conv(F.silu(_gn(x, sd, key + ".norm1", groups)), sd, key + ".conv1", padding=1)
    h = _conv(F.silu(_gn(h, sd, key + ".norm2", groups)), sd, key + ".conv2", padding=1)
    if key + ".conv_shortcut.weight" in sd:
        x = _conv(x, sd, key + ".conv_shortcut")
    return x + h


def _mid_attention(x, sd, key, groups):
    n, c, h, w = x.shape
    y = F.group_norm(x.view(n, c, h * w), groups, sd[key + ".group_norm.weight"], sd[key + ".group_norm.bias"], 1e-6)
    y = y.transpose(1, 2)
    q = F.linear(y, sd[key + ".to_q.weight"], sd[key + ".to_q.bias"])[:, None]
    k = F.linear(y, sd[key + ".to_k.weight"], sd[key + ".to_k.bias"])[:, None]
    v = F.linear(y, sd[key + ".to_v.weight"], sd[key + ".to_v.bias"])[:, None]
    o = F.scaled_dot_product_attention(q, k, v)[:, 0]
    o = F.linear(o, sd[key + ".to_out.0.weight"], sd[key + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(n, c, h, w) + x


@torch.no_grad()
def vae_encode_mode(sd: dict, cfg, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode(x).latent_dist.mode(), from the VAE state dict."""
    g = cfg.norm_num_groups
    nb = len(cfg.block_out_channels)
    h = _conv(x, sd, "encoder.conv_in", padding=1)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            h = _resnet(h, sd, f"encoder.down_blocks.{i}.resnets.{j}", g)
        if i != nb - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2)
    h = _resnet(h, sd, "encoder.mid_block.resnets.0", g)
    h = _mid_attention(h, sd, "encoder.mid_block.attentions.0", g)
    h = _resnet(h, sd, "encoder.mid_block.resnets.1", g)
    h = _conv(F.silu(_gn(h, sd, "encoder.conv_norm_out", g)), sd, "encoder.conv_out", padding=1)
    moments = _conv(h, sd, "quant_conv")
    return moments[:, : cfg.latent_channels]


def normalise_sparse(sparses: torch.Tensor, max_depth: float, min_depth: float, norm: str):
    """marigold_dc.py:707-756 (linear projection) as plain PyTorch: returns guide, mask, (lo, hi) of the depth range and
    the masked (min, max) of the guide.  The product path does this inside mdc_begin_frame; tests compare the two."""
    n = sparses.shape[0]
    sparses = sparses.float()
    masks = sparses > 0
    if norm == "minmax":
        lo, hi = masked_minmax(sparses.view(n, -1), masks.view(n, -1))
        lo, hi = lo.view(n, 1, 1, 1), hi.view(n, 1, 1, 1)
    else:
        lo = torch.full((n, 1, 1, 1), float(min_depth), device=sparses.device)
        hi = torch.full((n, 1, 1, 1), float(max_depth), device=sparses.device)
    clamped = sparses.clamp(min=lo, max=hi)
    if norm == "minmax":
        lo, hi = lo.clamp(min=min_depth), hi.clamp(max=max_depth)
    guide = (clamped - lo) / (hi - lo)
    gmin, gmax = masked_minmax(guide.view(n, -1), masks.view(n, -1))
    return guide, masks, (lo.view(n), hi.view(n)), (gmin, gmax)


def masked_minmax(x: torch.Tensor, mask: torch.Tensor):
    """Row-wise masked min / max; ValueError for an empty row (reference: utils.py:89-138)."""
    if x.shape != mask.shape:
        raise ValueError(f"Shape of x {tuple(x.shape)} must be equal to shape of mask {tuple(mask.shape)}")
    inf = torch.tensor(float("inf"), device=x.device, dtype=x.dtype)
    lo = torch.where(mask, x, inf).amin(dim=-1)
    hi = torch.where(mask, x, -inf).amax(dim=-1)
    if torch.isinf(lo).any() or torch.isinf(hi).any():
        raise ValueError("No valid values found in mask for some positions. "
                         "Ensure that mask has at least one True value along the specified dimensions.")
    return lo, hi
