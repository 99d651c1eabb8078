"""Model / geometry configuration of the hot path (host side).

Field meanings follow diffusers' UNet2DConditionModel / AutoencoderKL configs as used by
prs-eth/marigold-v1-0; the reference builds them in predict.py:463-481.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_heads: tuple = (5, 10, 20, 20)  # diffusers calls this attention_head_dim for SD2
    down_attention: tuple = (True, True, True, False)
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    # kind "kl": AutoencoderKL (fields above).  kind "tiny": AutoencoderTiny, the reference CLI's default VAE
    # (predict.py:44-52, 484-488): block_out_channels are the (equal) stage widths, layers_per_block / norm_num_groups
    # are unused.
    kind: str = "kl"
    num_encoder_blocks: tuple = (1, 3, 3, 3)
    num_decoder_blocks: tuple = (3, 3, 3, 1)
    latent_magnitude: float = 3.0


def _get(obj, name, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


def unet_config_from(module) -> UNetConfig:
    """Accepts our dataclass, an object with .cfg, or a diffusers module/config (attribute or dict access)."""
    if isinstance(module, UNetConfig):
        return module
    c = _get(module, "cfg") or _get(module, "config") or module
    if isinstance(c, UNetConfig):
        return c
    boc = tuple(_get(c, "block_out_channels"))
    heads = _get(c, "attention_heads") or _get(c, "attention_head_dim")
    heads = tuple(heads) if isinstance(heads, (list, tuple)) else (heads,) * len(boc)
    down_types = _get(c, "down_block_types")
    if down_types is not None:
        down_attn = tuple("CrossAttn" in t for t in down_types)
    else:
        down_attn = tuple(_get(c, "down_attention", (True,) * (len(boc) - 1) + (False,)))
    return UNetConfig(in_channels=_get(c, "in_channels", 8), out_channels=_get(c, "out_channels", 4),
                      block_out_channels=boc, layers_per_block=_get(c, "layers_per_block", 2), attention_heads=heads,
                      down_attention=down_attn, cross_attention_dim=_get(c, "cross_attention_dim", 1024),
                      norm_num_groups=_get(c, "norm_num_groups", 32))


def vae_config_from(module) -> VAEConfig:
    if isinstance(module, VAEConfig):
        return module
    c = _get(module, "cfg") or _get(module, "config") or module
    if isinstance(c, VAEConfig):
        return c
    if _get(c, "decoder_block_out_channels") is not None:  # AutoencoderTiny (diffusers config names)
        enc, dec = tuple(_get(c, "encoder_block_out_channels")), tuple(_get(c, "decoder_block_out_channels"))
        if enc != dec or len(set(dec)) != 1:
            raise ValueError(f"AutoencoderTiny: encoder / decoder stage widths must all be equal, got {enc} / {dec}")
        if _get(c, "act_fn", "relu") != "relu" or _get(c, "upsampling_scaling_factor", 2) != 2:
            raise ValueError("AutoencoderTiny: only act_fn='relu' and upsampling_scaling_factor=2 are supported")
        return VAEConfig(in_channels=_get(c, "in_channels", 3), out_channels=_get(c, "out_channels", 3),
                         latent_channels=_get(c, "latent_channels", 4), block_out_channels=dec, layers_per_block=0,
                         norm_num_groups=0, scaling_factor=_get(c, "scaling_factor", 1.0), kind="tiny",
                         num_encoder_blocks=tuple(_get(c, "num_encoder_blocks")),
                         num_decoder_blocks=tuple(_get(c, "num_decoder_blocks")),
                         latent_magnitude=float(_get(c, "latent_magnitude", 3)))
    return VAEConfig(in_channels=_get(c, "in_channels", 3), out_channels=_get(c, "out_channels", 3),
                     latent_channels=_get(c, "latent_channels", 4),
                     block_out_channels=tuple(_get(c, "block_out_channels")),
                     layers_per_block=_get(c, "layers_per_block", 2), norm_num_groups=_get(c, "norm_num_groups", 32),
                     scaling_factor=_get(c, "scaling_factor", 0.18215))


def processed_geometry(H: int, W: int, resolution: int):
    """(proc_h, proc_w, pad_h, pad_w): resize-to-max-edge then pad to a multiple of 8 (MarigoldImageProcessor)."""
    m = max(H, W)
    ph, pw = H * resolution // m, W * resolution // m
    if ph == 0 or pw == 0:
        raise ValueError(f"Extreme aspect ratio of the input image: [{W} x {H}]")
    return ph, pw, -ph % 8, -pw % 8
