"""Algorithmic FLOP counts of the hot path (SURVEY.md Appendix B rules): 2 x MAC; conv k x k = 2 Cin Cout k^2 Ho Wo;
linear = 2 Cin Cout tokens; attention = 4 T T_kv d; norms / activations / resampling count zero; the backward is
input-gradient only and counts 1 x forward.  Used for the roofline numerator and for scaling CPU-baseline samples.
"""
from __future__ import annotations

from .config import UNetConfig, VAEConfig, processed_geometry


def _conv(cin, cout, h, w, k=3):
    return 2.0 * cin * cout * k * k * h * w


def _resnet(cin, cout, h, w, temb_dim=0):
    f = _conv(cin, cout, h, w) + _conv(cout, cout, h, w)
    if cin != cout:
        f += _conv(cin, cout, h, w, 1)
    if temb_dim:
        f += 2.0 * temb_dim * cout
    return f


def _transformer(d, heads, h, w, cross_dim):
    t = h * w
    f = 2 * (2.0 * d * d * t)                  # proj_in, proj_out
    f += 4 * (2.0 * d * d * t) + 4.0 * t * t * d   # self attention: q, k, v, out + QK^T, PV
    f += 2 * (2.0 * d * d * t) + 2 * (2.0 * cross_dim * d * 2) + 4.0 * t * 2 * d  # cross attention (2 key tokens)
    f += 2.0 * d * 8 * d * t + 2.0 * 4 * d * d * t  # GEGLU feed-forward
    return f


def unet_forward_flops(cfg: UNetConfig, lh: int, lw: int) -> float:
    boc, L = cfg.block_out_channels, cfg.layers_per_block
    nb, tc = len(boc), boc[0] * 4
    f = _conv(cfg.in_channels, boc[0], lh, lw)
    f += 2.0 * boc[0] * tc + 2.0 * tc * tc
    h, w = lh, lw
    sizes = [(h, w)]
    skips = [boc[0]]
    cout = boc[0]
    for i in range(nb):
        cin, cout = cout, boc[i]
        for j in range(L):
            f += _resnet(cin if j == 0 else cout, cout, h, w, tc)
            if cfg.down_attention[i]:
                f += _transformer(cout, cfg.attention_heads[i], h, w, cfg.cross_attention_dim)
            skips.append(cout)
            sizes.append((h, w))
        if i != nb - 1:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            f += _conv(cout, cout, h, w)
            skips.append(cout)
            sizes.append((h, w))
    c = boc[-1]
    f += 2 * _resnet(c, c, h, w, tc) + _transformer(c, cfg.attention_heads[-1], h, w, cfg.cross_attention_dim)
    cout = boc[-1]
    for i in range(nb):
        prev, cout = cout, boc[nb - 1 - i]
        attn = cfg.down_attention[nb - 1 - i]
        for j in range(L + 1):
            sc = skips.pop()
            h, w = sizes.pop()
            f += _resnet((prev if j == 0 else cout) + sc, cout, h, w, tc)
            if attn:
                f += _transformer(cout, cfg.attention_heads[nb - 1 - i], h, w, cfg.cross_attention_dim)
        if i != nb - 1:
            h, w = sizes[-1]
            f += _conv(cout, cout, h, w)
    f += _conv(boc[0], cfg.out_channels, lh, lw)
    return f


def _tiny_block(c, h, w):
    return 3 * _conv(c, c, h, w)


def tiny_decoder_forward_flops(cfg: VAEConfig, lh: int, lw: int) -> float:
    """AutoencoderTiny decoder: conv, [blocks, nearest x2, conv] per stage, conv to 3 channels."""
    ch, nblk = cfg.block_out_channels, cfg.num_decoder_blocks
    h, w = lh, lw
    f = _conv(cfg.latent_channels, ch[0], h, w)
    for i, n in enumerate(nblk):
        f += n * _tiny_block(ch[i], h, w)
        if i != len(nblk) - 1:
            h, w = 2 * h, 2 * w
            f += _conv(ch[i], ch[i], h, w)
        else:
            f += _conv(ch[i], cfg.out_channels, h, w)
    return f


def tiny_encoder_forward_flops(cfg: VAEConfig, ph: int, pw: int) -> float:
    ch, nblk = cfg.block_out_channels, cfg.num_encoder_blocks
    h, w = ph, pw
    f = 0.0
    for i, n in enumerate(nblk):
        if i == 0:
            f += _conv(cfg.in_channels, ch[0], h, w)
        else:
            h, w = h // 2, w // 2
            f += _conv(ch[i], ch[i], h, w)
        f += n * _tiny_block(ch[i], h, w)
    return f + _conv(ch[-1], cfg.latent_channels, h, w)


def vae_decoder_forward_flops(cfg: VAEConfig, lh: int, lw: int) -> float:
    if cfg.kind == "tiny":
        return tiny_decoder_forward_flops(cfg, lh, lw)
    boc, L = cfg.block_out_channels, cfg.layers_per_block
    nb = len(boc)
    h, w = lh, lw
    c = boc[-1]
    f = _conv(cfg.latent_channels, cfg.latent_channels, h, w, 1) + _conv(cfg.latent_channels, c, h, w)
    t = h * w
    f += 2 * _resnet(c, c, h, w) + 4 * (2.0 * c * c * t) + 4.0 * t * t * c
    cout = c
    for i in range(nb):
        cin, cout = cout, boc[nb - 1 - i]
        for j in range(L + 1):
            f += _resnet(cin if j == 0 else cout, cout, h, w)
        if i != nb - 1:
            h, w = 2 * h, 2 * w
            f += _conv(cout, cout, h, w)
    f += _conv(boc[0], cfg.out_channels, h, w)
    return f


def vae_encoder_forward_flops(cfg: VAEConfig, ph: int, pw: int) -> float:
    if cfg.kind == "tiny":
        return tiny_encoder_forward_flops(cfg, ph, pw)
    boc, L = cfg.block_out_channels, cfg.layers_per_block
    nb = len(boc)
    h, w = ph, pw
    f = _conv(cfg.in_channels, boc[0], h, w)
    cout = boc[0]
    for i in range(nb):
        cin, cout = cout, boc[i]
        for j in range(L):
            f += _resnet(cin if j == 0 else cout, cout, h, w)
        if i != nb - 1:
            h, w = h // 2, w // 2
            f += _conv(cout, cout, h, w)
    c, t = boc[-1], h * w
    f += 2 * _resnet(c, c, h, w) + 4 * (2.0 * c * c * t) + 4.0 * t * t * c
    f += _conv(c, 2 * cfg.latent_channels, h, w) + _conv(2 * cfg.latent_channels, 2 * cfg.latent_channels, h, w, 1)
    return f


def step_flops(ucfg: UNetConfig, vcfg: VAEConfig, H: int, W: int, resolution: int) -> dict:
    """Algorithmic FLOPs of one guided step and of a whole frame (per sample)."""
    ph, pw, pad_h, pad_w = processed_geometry(H, W, resolution)
    lh, lw = (ph + pad_h) // 8, (pw + pad_w) // 8
    u = unet_forward_flops(ucfg, lh, lw)
    d = vae_decoder_forward_flops(vcfg, lh, lw)
    e = vae_encoder_forward_flops(vcfg, ph + pad_h, pw + pad_w)
    return dict(unet_fwd=u, dec_fwd=d, enc_fwd=e, step=2.0 * (u + d), latent=(lh, lw),
                frame=lambda steps: steps * 2.0 * (u + d) + e + d)
