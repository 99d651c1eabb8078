"""Shared helpers for the GPU parity tests: build the oracle models and a matching StepEngine."""
import torch

from oracle.sd2_modules import AutoencoderKL, UNet2DConditionModel, UNetConfig as OUNetConfig, VAEConfig as OVAEConfig
from oracle.sd2_modules import tiny_unet_config, tiny_vae_config
from oracle.marigold_dc import make_empty_text_embedding


def product_cfgs(ucfg, vcfg):
    from depth_completion_b200.config import UNetConfig, VAEConfig

    u = UNetConfig(in_channels=ucfg.in_channels, out_channels=ucfg.out_channels,
                   block_out_channels=tuple(ucfg.block_out_channels), layers_per_block=ucfg.layers_per_block,
                   attention_heads=tuple(ucfg.attention_heads), down_attention=tuple(ucfg.down_attention),
                   cross_attention_dim=ucfg.cross_attention_dim, norm_num_groups=ucfg.norm_num_groups)
    v = VAEConfig(block_out_channels=tuple(vcfg.block_out_channels), layers_per_block=vcfg.layers_per_block,
                  norm_num_groups=vcfg.norm_num_groups, scaling_factor=vcfg.scaling_factor)
    return u, v


def build_models(device, tiny=True, seed=1234, round_bf16=True):
    """fp32 oracle modules whose weights are bf16-representable (so fp32-oracle vs bf16-engine differ only by
    activation rounding), on `device`."""
    torch.manual_seed(seed)
    ucfg, vcfg = (tiny_unet_config(), tiny_vae_config()) if tiny else (OUNetConfig(), OVAEConfig())
    unet, vae = UNet2DConditionModel(ucfg), AutoencoderKL(vcfg)
    if round_bf16:
        with torch.no_grad():
            for p in list(unet.parameters()) + list(vae.parameters()):
                p.copy_(p.bfloat16().float())
    unet, vae = unet.to(device).requires_grad_(False), vae.to(device).requires_grad_(False)
    ctx = make_empty_text_embedding(ucfg.cross_attention_dim, device=device).bfloat16().float()
    return unet, vae, ctx, ucfg, vcfg


def build_engine(unet, vae, ctx, ucfg, vcfg, n, H, W, resolution, steps, device):
    from depth_completion_b200 import ddim
    from depth_completion_b200.engine import StepEngine

    pu, pv = product_cfgs(ucfg, vcfg)
    eng = StepEngine(pu, pv, n, H, W, resolution, steps, device)
    eng.load_weights(unet.state_dict(), vae.state_dict())
    eng.prepare(ctx, ddim.alphas_cumprod(), ddim.trailing_timesteps(steps))
    return eng


def rel_err(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-12)).item()


def rel_l2(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-12)).item()
