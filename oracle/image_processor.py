"""ORACLE (test infrastructure): MarigoldImageProcessor of diffusers==0.31.0 (vae_scale_factor 8), the
pieces marigold_dc.py:687-692, :367-370 call.  SURVEY.md Appendix A.4.  PARITY UNPINNED (diffusers absent).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

VAE_SCALE_FACTOR = 8


def resize_antialias(image, size, mode="bilinear", is_aa=None):
    antialias = bool(is_aa) and mode in ("bilinear", "bicubic")
    return F.interpolate(image, size, mode=mode, antialias=antialias)


def preprocess(image: torch.Tensor, processing_resolution: int, device, dtype):
    """uint8 [N,3|1,H,W] -> [-1,1] -> resize max edge (bilinear, antialias) -> replicate-pad to x8."""
    if image.ndim != 4:
        raise ValueError(f"Input image is not 4-dimensional: shape={image.shape}")
    orig_res = tuple(image.shape[-2:])
    dtype_max = None
    if not torch.is_floating_point(image):
        if image.dtype != torch.uint8:
            raise ValueError(f"Image dtype={image.dtype} is not supported.")
        dtype_max = 255
    if image.shape[1] == 1:
        image = image.repeat(1, 3, 1, 1)
    if image.shape[1] != 3:
        raise ValueError(f"Input image is not 1- or 3-channel: {image.shape}.")
    image = image.to(device=device, dtype=dtype)
    if dtype_max is not None:
        image = image / dtype_max
    if image.min().item() < 0.0 or image.max().item() > 1.0:
        raise ValueError("Input image data is partially outside of the [0,1] range.")
    image = image * 2.0 - 1.0
    h, w = image.shape[-2:]
    m = max(h, w)
    new_h, new_w = h * processing_resolution // m, w * processing_resolution // m
    if new_h == 0 or new_w == 0:
        raise ValueError(f"Extreme aspect ratio of the input image: [{w} x {h}]")
    image = resize_antialias(image, (new_h, new_w), "bilinear", is_aa=True)
    ph, pw = -new_h % VAE_SCALE_FACTOR, -new_w % VAE_SCALE_FACTOR
    image = F.pad(image, (0, pw, 0, ph), mode="replicate")
    return image, (ph, pw), orig_res


def unpad_image(image, padding):
    ph, pw = padding
    uh = None if ph == 0 else -ph
    uw = None if pw == 0 else -pw
    return image[:, :, :uh, :uw]
