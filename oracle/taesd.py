"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

PyTorch restatement of diffusers 0.31.0 `AutoencoderTiny` ("madebyollin/taesd"), the VAE the reference CLI swaps in
by default (`/root/reference/predict.py:44-52, 484-488`; `marigold_dc.py:18`).  diffusers is absent from this image
and the reference ships no fixture for it, so the module below is restated from the published architecture and pinned
only structurally (parameter count, state-dict key set: tests/test_oracle_formulas.py).

    EncoderTiny : x -> (x + 1) / 2 -> conv3x3(3, 64) -> Block -> [conv3x3 s2 (no bias) -> 3 Blocks] x 3 -> conv3x3(64, 4)
    DecoderTiny : z -> tanh(z / 3) * 3 -> conv3x3(4, 64) -> ReLU -> [3 Blocks -> nearest x2 -> conv3x3 (no bias)] x 3
                  -> Block -> conv3x3(64, 3) -> * 2 - 1
    Block(c)    : relu(conv3(relu(conv3(relu(conv3(x))))) + x)        (64 -> 64: the skip is the identity)

`encode(x).latents` is the encoder output itself (no distribution, no sampling); `scaling_factor` is 1.0, so the
reference's `prepare_latents` / `decode_prediction` (`marigold_dc.py:366, 696`) multiply / divide by one.
State-dict keys follow diffusers: `encoder.layers.<i>.{weight,bias}`, `encoder.layers.<i>.conv.{0,2,4}.{weight,bias}`,
likewise `decoder.layers.<i>...` (activation / upsample modules occupy indices but hold no parameters).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class TinyVAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    encoder_block_out_channels: tuple = (64, 64, 64, 64)
    decoder_block_out_channels: tuple = (64, 64, 64, 64)
    num_encoder_blocks: tuple = (1, 3, 3, 3)
    num_decoder_blocks: tuple = (3, 3, 3, 1)
    latent_magnitude: float = 3.0
    scaling_factor: float = 1.0
    kind: str = "tiny"


class AutoencoderTinyBlock(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(), nn.Conv2d(cout, cout, 3, padding=1), nn.ReLU(),
                                  nn.Conv2d(cout, cout, 3, padding=1))
        self.skip = nn.Conv2d(cin, cout, 1, bias=False) if cin != cout else nn.Identity()
        self.fuse = nn.ReLU()

    def forward(self, x):
        return self.fuse(self.conv(x) + self.skip(x))


class EncoderTiny(nn.Module):
    def __init__(self, cfg: TinyVAEConfig):
        super().__init__()
        layers = []
        for i, nblk in enumerate(cfg.num_encoder_blocks):
            c = cfg.encoder_block_out_channels[i]
            if i == 0:
                layers.append(nn.Conv2d(cfg.in_channels, c, 3, padding=1))
            else:
                layers.append(nn.Conv2d(c, c, 3, padding=1, stride=2, bias=False))
            for _ in range(nblk):
                layers.append(AutoencoderTinyBlock(c, c))
        layers.append(nn.Conv2d(cfg.encoder_block_out_channels[-1], cfg.latent_channels, 3, padding=1))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x.add(1).div(2))


class DecoderTiny(nn.Module):
    def __init__(self, cfg: TinyVAEConfig):
        super().__init__()
        ch = cfg.decoder_block_out_channels
        layers = [nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1), nn.ReLU()]
        n = len(cfg.num_decoder_blocks)
        for i, nblk in enumerate(cfg.num_decoder_blocks):
            last = i == n - 1
            c = ch[i]
            for _ in range(nblk):
                layers.append(AutoencoderTinyBlock(c, c))
            if not last:
                layers.append(nn.Upsample(scale_factor=2))
            layers.append(nn.Conv2d(c, cfg.out_channels if last else c, 3, padding=1, bias=last))
        self.layers = nn.Sequential(*layers)
        self.mag = cfg.latent_magnitude

    def forward(self, z):
        z = torch.tanh(z / self.mag) * self.mag
        return self.layers(z).mul(2).sub(1)


class AutoencoderTiny(nn.Module):
    def __init__(self, cfg: TinyVAEConfig = TinyVAEConfig()):
        super().__init__()
        self.cfg = cfg
        self.encoder = EncoderTiny(cfg)
        self.decoder = DecoderTiny(cfg)

    def encode_mode(self, x):
        """`encode(x).latents` (what retrieve_latents returns for AutoencoderTiny); same name as the KL oracle's method."""
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)
