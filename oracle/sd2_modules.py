"""ORACLE (test infrastructure, not product code): PyTorch restatement of the diffusers==0.31.0 modules
the reference's hot path calls -- UNet2DConditionModel (SD2 config, in_channels=8) and AutoencoderKL.

PARITY UNPINNED: diffusers is an un-vendored dependency of /root/reference (requirements.txt:1) and is
not installed in this image, the reference ships no tests or golden vectors, so
this restatement is anchored on the reference's call sites (marigold_dc.py:366, :460-465, :696) and on
structural pins only: parameter counts 865.9 M / 49.5 M and the diffusers state-dict key set
(SURVEY.md Appendix A.1, A.2, A.5).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import math
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- configs
@dataclass
class UNetConfig:
    """UNet2DConditionModel config of prs-eth/marigold-v1-0."""
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    # per-block *head counts* (diffusers' attention_head_dim naming quirk); head_dim = channels // heads
    attention_heads: tuple = (5, 10, 20, 20)
    # which down blocks carry cross-attention transformers (the last SD2 block does not)
    down_attention: tuple = (True, True, True, False)
    cross_attention_dim: int = 1024
    norm_num_groups: int = 32
    norm_eps: float = 1e-5

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


@dataclass
class VAEConfig:
    """AutoencoderKL config of the SD2 VAE."""
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215


def tiny_unet_config() -> UNetConfig:
    """A structurally identical but narrow UNet for fast parity tests (channels stay multiples of 64)."""
    return UNetConfig(block_out_channels=(64, 128, 128, 128), attention_heads=(1, 2, 2, 2), cross_attention_dim=64)


def tiny_vae_config() -> VAEConfig:
    return VAEConfig(block_out_channels=(64, 64, 128, 128))


# ----------------------------------------------------------------------------- shared blocks
class ResnetBlock2D(nn.Module):
    """GN -> SiLU -> conv3x3 -> (+ time_emb_proj(SiLU(temb))) -> GN -> SiLU -> conv3x3, + shortcut (A.1)."""

    def __init__(self, cin: int, cout: int, temb_channels: int | None, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    """diffusers Attention with AttnProcessor2_0 (scaled_dot_product_attention)."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_dim: int | None = None, bias: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        n, t, _ = x.shape
        q = self.to_q(x).view(n, t, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(n, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(n, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(n, t, -1)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, cross_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim // heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim // heads, cross_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    """use_linear_projection=True, one BasicTransformerBlock, GroupNorm eps 1e-6 (A.1)."""

    def __init__(self, dim: int, heads: int, cross_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx):
        n, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(n, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, ctx)
        y = self.proj_out(y)
        return y.reshape(n, h, w, c).permute(0, 3, 1, 2) + res


class Downsample2D(nn.Module):
    def __init__(self, c: int, padding: int):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:  # VAE encoder: asymmetric zero pad (0,1,0,1)
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x, size=None):
        if size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=size, mode="nearest")
        return self.conv(x)


# ----------------------------------------------------------------------------- UNet
class _Block(nn.Module):
    """Generic container matching diffusers' {resnets, attentions, downsamplers|upsamplers} key layout."""

    def __init__(self):
        super().__init__()
        self.resnets = nn.ModuleList()


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin] (A.1)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = UNetConfig()):
        super().__init__()
        self.cfg = cfg
        boc, g, eps, tc = cfg.block_out_channels, cfg.norm_num_groups, cfg.norm_eps, cfg.time_embed_dim
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = nn.Module()
        self.time_embedding.linear_1 = nn.Linear(boc[0], tc)
        self.time_embedding.linear_2 = nn.Linear(tc, tc)

        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, ch in enumerate(boc):
            cin, cout = cout, ch
            blk = _Block()
            if cfg.down_attention[i]:
                blk.attentions = nn.ModuleList()
            for j in range(cfg.layers_per_block):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, tc, g, eps))
                if cfg.down_attention[i]:
                    blk.attentions.append(Transformer2DModel(cout, cfg.attention_heads[i], cfg.cross_attention_dim, g))
            if i != len(boc) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(cout, padding=1)])
            self.down_blocks.append(blk)

        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(boc[-1], boc[-1], tc, g, eps))
        self.mid_block.attentions = nn.ModuleList(
            [Transformer2DModel(boc[-1], cfg.attention_heads[-1], cfg.cross_attention_dim, g)])
        self.mid_block.resnets.append(ResnetBlock2D(boc[-1], boc[-1], tc, g, eps))

        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        rev_heads = list(reversed(cfg.attention_heads))
        rev_attn = list(reversed(cfg.down_attention))
        cout = rev[0]
        for i, ch in enumerate(rev):
            prev_out, cout = cout, ch
            cin = rev[min(i + 1, len(boc) - 1)]
            blk = _Block()
            if rev_attn[i]:
                blk.attentions = nn.ModuleList()
            for j in range(cfg.layers_per_block + 1):
                skip = cin if j == cfg.layers_per_block else cout
                rin = prev_out if j == 0 else cout
                blk.resnets.append(ResnetBlock2D(rin + skip, cout, tc, g, eps))
                if rev_attn[i]:
                    blk.attentions.append(Transformer2DModel(cout, rev_heads[i], cfg.cross_attention_dim, g))
            if i != len(boc) - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(cout)])
            self.up_blocks.append(blk)

        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states):
        n = sample.shape[0]
        n_up = len(self.cfg.block_out_channels) - 1
        forward_upsample_size = any(d % (2 ** n_up) != 0 for d in sample.shape[-2:])
        t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(n)
        temb = timestep_embedding(t, self.cfg.block_out_channels[0]).to(sample.dtype)
        temb = self.time_embedding.linear_2(F.silu(self.time_embedding.linear_1(temb)))

        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(h, temb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, encoder_hidden_states)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)

        h = self.mid_block.resnets[0](h, temb)
        h = self.mid_block.attentions[0](h, encoder_hidden_states)
        h = self.mid_block.resnets[1](h, temb)

        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(torch.cat([h, skips.pop()], dim=1), temb)
                if hasattr(blk, "attentions"):
                    h = blk.attentions[j](h, encoder_hidden_states)
            if hasattr(blk, "upsamplers"):
                size = skips[-1].shape[2:] if forward_upsample_size else None
                h = blk.upsamplers[0](h, size)

        return self.conv_out(F.silu(self.conv_norm_out(h)))


# ----------------------------------------------------------------------------- VAE
class VAEAttention(nn.Module):
    """Single-head attention of the VAE mid block (bias on q/k/v/out, GroupNorm eps 1e-6, residual) (A.2)."""

    def __init__(self, c: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        n, c, h, w = x.shape
        y = self.group_norm(x.view(n, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(y)[:, None], self.to_k(y)[:, None], self.to_v(y)[:, None]
        o = F.scaled_dot_product_attention(q, k, v)[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(n, c, h, w) + x


class _MidBlock(nn.Module):
    def __init__(self, c: int, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, groups, 1e-6), ResnetBlock2D(c, c, None, groups, 1e-6)])
        self.attentions = nn.ModuleList([VAEAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(cfg.latent_channels, rev[0], 3, padding=1)
        self.mid_block = _MidBlock(rev[0], g)
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i, ch in enumerate(rev):
            cin, cout = cout, ch
            blk = _Block()
            for j in range(cfg.layers_per_block + 1):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, None, g, 1e-6))
            if i != len(boc) - 1:
                blk.upsamplers = nn.ModuleList([Upsample2D(cout)])
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            for res in blk.resnets:
                h = res(h)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class Encoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, ch in enumerate(boc):
            cin, cout = cout, ch
            blk = _Block()
            for j in range(cfg.layers_per_block):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, None, g, 1e-6))
            if i != len(boc) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)])
            self.down_blocks.append(blk)
        self.mid_block = _MidBlock(boc[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            for res in blk.resnets:
                h = res(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKL(nn.Module):
    def __init__(self, cfg: VAEConfig = VAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg)
        self.decoder = Decoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)

    def encode_mode(self, x):
        """encode(x).latent_dist.mode(): the mean half of the moments (A.2)."""
        moments = self.quant_conv(self.encoder(x))
        return moments[:, : self.cfg.latent_channels]

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


def count_params(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())
