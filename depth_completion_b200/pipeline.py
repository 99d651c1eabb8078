"""Drop-in for the hot path of tier4/depth_completion: `MarigoldDepthCompletionPipeline.__call__`
(/root/reference/marigold_dc.py:467-985) with the guided denoising loop (:799-909) and the final decode
(:970-985) executed by libmdc_b200.so (hand-written sm_100a kernels) instead of diffusers + autograd.

Same constructor and call signature, same return tuple `(denses [N,1,H,W] fp32 metric, pred_latents [N,4,EH,EW])`,
same ValueError conventions.  Non-default branches the north star leaves out of scope
(per-input training, non-bilinear interpolation, closed_form with edge / smooth) raise NotImplementedError instead of silently doing something else.

The host keeps, in PyTorch, only the argument checks and the seeded initial latent (torch's Philox stream,
marigold_dc.py:661, :677-684); image preprocessing, the VAE encoder and the sparse-depth normalisation run inside the
library (mdc_begin_frame), the per-call options travel through mdc_set_options.
"""
from __future__ import annotations

import torch

from . import ddim, prologue
from ._lib import MdcError
from .config import processed_geometry, unet_config_from, vae_config_from
from .engine import StepEngine

SUPPORTED_LOSS_FUNCS = ["l1", "l2", "edge", "smooth"]  # marigold_dc.py:19
EPSILON = 1e-7  # marigold_dc.py:20
EMPTY_PROMPT_IDS = (49406, 49407)  # CLIP BOS, EOS: tokenizer("", padding="do_not_pad")


class MarigoldDepthCompletionPipeline:
    """Constructor mirrors marigold_dc.py:258-282.

    `unet` / `vae` are anything exposing `.state_dict()` with diffusers key names plus a diffusers-style
    `.config` (or our dataclass `.cfg`).  `scheduler` may be a diffusers DDIMScheduler-like object
    (`alphas_cumprod`, `set_timesteps`, `timesteps`) or None for the built-in trailing DDIM tables.
    `text_encoder` / `tokenizer` are only used to produce the empty-prompt embedding once
    (marigold_dc.py:664-674); alternatively assign `pipe.empty_text_embedding` ([1, 2, cross_dim]).
    """

    def __init__(self, unet, vae, scheduler=None, text_encoder=None, tokenizer=None, prediction_type=None,
                 scale_invariant=True, shift_invariant=True, default_denoising_steps=None,
                 default_processing_resolution=None):
        self._engines: dict = {}
        self._sd_cache = None
        self.unet, self.vae, self.scheduler = unet, vae, scheduler   # property setters below
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        self.prediction_type = prediction_type
        self.scale_invariant, self.shift_invariant = scale_invariant, shift_invariant
        self.default_denoising_steps = default_denoising_steps
        self.default_processing_resolution = default_processing_resolution
        self.empty_text_embedding = None
        p = next(iter(unet.state_dict().values()))
        self.device = p.device
        self.dtype = torch.bfloat16  # the engine computes in bf16 (BASELINE.json configs b-e)

    # The reference swaps modules on a live pipeline (predict.py:484-488 `pipe.vae = AutoencoderTiny...`, :491-494
    # `pipe.scheduler = DDIMScheduler...`): re-derive the config and drop everything built from the old module.
    def _invalidate(self):
        for eng in getattr(self, "_engines", {}).values():
            eng.close()
        self._engines = {}
        self._sd_cache = None

    @property
    def unet(self):
        return self._unet

    @unet.setter
    def unet(self, m):
        self._unet = m
        self.unet_cfg = unet_config_from(m)
        self._invalidate()

    @property
    def vae(self):
        return self._vae

    @vae.setter
    def vae(self, m):
        self._vae = m
        self.vae_cfg = vae_config_from(m)
        self._invalidate()

    @property
    def scheduler(self):
        return self._scheduler

    @scheduler.setter
    def scheduler(self, sch):
        self._scheduler = sch
        self._invalidate()

    # ------------------------------------------------------------------ plumbing
    def to(self, device):
        self.device = torch.device(device)
        for m in (self.unet, self.vae, self.text_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(self.device)
        self._sd_cache = None
        for eng in self._engines.values():  # free the old device's workspace instead of waiting for the collector
            eng.close()
        self._engines.clear()
        return self

    def _state_dicts(self):
        if self._sd_cache is None:
            usd = {k: v.detach().to(self.device) for k, v in self.unet.state_dict().items()}
            vsd = {k: v.detach().to(self.device) for k, v in self.vae.state_dict().items()}
            vsd_enc = {k: v.to(self.dtype) for k, v in vsd.items() if k.startswith(("encoder.", "quant_conv."))}
            self._sd_cache = (usd, vsd, vsd_enc)
        return self._sd_cache

    def _empty_embedding(self):
        if self.empty_text_embedding is None:
            if self.text_encoder is None:
                raise ValueError("no text_encoder given and pipe.empty_text_embedding is not set")
            ids = torch.tensor([list(EMPTY_PROMPT_IDS)], device=self.device)
            if self.tokenizer is not None:
                try:
                    tok = self.tokenizer("", padding="do_not_pad", max_length=self.tokenizer.model_max_length,
                                         truncation=True, return_tensors="pt")
                    if tok.input_ids.shape[-1] == 2:
                        ids = tok.input_ids.to(self.device)
                except Exception:
                    pass
            with torch.no_grad():
                self.empty_text_embedding = self.text_encoder(ids)[0]
        return self.empty_text_embedding

    def _engine(self, N, H, W, resolution, steps) -> StepEngine:
        key = (N, H, W, resolution, steps, str(self.device))
        eng = self._engines.get(key)
        if eng is None:
            if self.device.type != "cuda":
                raise MdcError("MarigoldDepthCompletionPipeline needs a CUDA device: call .to('cuda') (no CPU path)")
            for old in self._engines.values():  # one resident workspace at a time
                old.close()
            self._engines.clear()
            usd, vsd, _ = self._state_dicts()
            eng = StepEngine(self.unet_cfg, self.vae_cfg, N, H, W, resolution, steps, self.device)
            eng.load_weights(usd, vsd)
            ac, ts = ddim.tables_from_scheduler(self.scheduler, steps)
            eng.prepare(self._empty_embedding(), ac, ts)
            self._engines[key] = eng
        return eng

    # ------------------------------------------------------------------ the call (marigold_dc.py:467-493)
    def __call__(self, imgs, sparses, max_depth, min_depth=0.0, projection="linear", inv=False, norm="minmax",
                 percentile=(0.01, 0.99), pred_latents_prev=None, beta=0.9, steps=50, resolution=768,
                 closed_form=None, opt="adam", lr=None, kld=False, kld_weight=0.1, kld_mode="simple",
                 interp_mode="bilinear", loss_funcs=None, seed=2024, train_latents=True, train_method="per-step",
                 train_steps=10, _begin_only=False):
        # --- argument validation, same order and conditions as marigold_dc.py:583-656
        if (imgs.ndim != 4 or sparses.ndim != 4 or imgs.shape[0] != sparses.shape[0]
                or imgs.shape[-2:] != sparses.shape[-2:]):
            raise ValueError("Shape of image must be [N, C, H, W] and shape of sparse must be [N, 1, H, W], but got "
                             f"image.shape: {imgs.shape} and sparse.shape: {sparses.shape}")
        N, _, H, W = imgs.shape
        EH = resolution * H // (8 * max(H, W))
        EW = resolution * W // (8 * max(H, W))
        if pred_latents_prev is not None:
            if pred_latents_prev.ndim != 4 or tuple(pred_latents_prev.shape) != (N, 4, EH, EW):
                raise ValueError(f"Shape of pred_latents_prev must be [N, 4, EH, EW], but got {pred_latents_prev.shape}")
        if closed_form is None:
            closed_form = not train_latents
        elif not closed_form and not train_latents:
            raise ValueError("Closed form solution must be enabled when trainable latents are not used. "
                             "Set closed_form=True when train_latents=False, or just leave closed_form=None")
        if train_method not in ["per-step", "per-input"]:
            raise ValueError(f"Unknown train_method: {train_method}")
        if train_method == "per-input" and train_steps <= 0:
            raise ValueError("train_steps must be > 0 when per-input training is enabled")
        if not (0 < beta < 1):
            raise ValueError(f"beta must be in (0, 1), but got {beta}")
        if norm == "percentile" and not all(0 <= p <= 1 for p in percentile):
            raise ValueError(f"percentile must be in [0, 1], but got {percentile}")
        if projection not in ["linear", "log", "log10"]:
            raise ValueError(f"Unknown projection method: {projection}")
        if (projection in ["log", "log10"] or inv) and min_depth <= EPSILON:
            raise ValueError(f"min_depth must be > {EPSILON} when projection is 'log' or 'log10' or inv is True, "
                             f"but got {min_depth}")
        lr_latent, lr_scaling = (0.05, 0.005) if lr is None else lr
        if loss_funcs is None:
            loss_funcs = ["l1", "l2"]
        else:
            for f in loss_funcs:
                if f not in SUPPORTED_LOSS_FUNCS:
                    raise ValueError(f"Unknown loss function: {f}")
        if norm not in ("minmax", "percentile", "const"):
            raise ValueError(f"Unknown norm method: {norm}")
        if opt not in ("adam", "sgd", "adagrad"):
            raise ValueError(f"Unknown optimizer: {opt}")
        # --- branches outside the hot path this library implements (OUT OF SCOPE rows)
        unsupported = []
        if len(loss_funcs) == 0:
            raise ValueError("loss_funcs must contain at least one loss function")  # compute_loss, marigold_dc.py:171-172
        if kld and kld_mode not in ("simple", "strict"):
            raise ValueError(f"Unknown mode: {kld_mode}")                           # utils.py:78-79
        if train_latents and closed_form and any(f in ("edge", "smooth") for f in loss_funcs):
            unsupported.append("closed_form=True with edge / smooth losses")
        if train_latents and train_method != "per-step":
            unsupported.append("train_method='per-input'")
        if interp_mode not in ("bilinear", "nearest"):  # the two modes the reference CLI offers (predict.py:200-206)
            unsupported.append(f"interp_mode='{interp_mode}'")
        if unsupported:
            raise NotImplementedError("outside the B200 hot path (guided per-step optimisation of the latent): "
                                      + ", ".join(unsupported))
        ph, pw, pad_h, pad_w = processed_geometry(H, W, resolution)
        if ((ph + pad_h) // 8, (pw + pad_w) // 8) != (EH, EW):
            raise ValueError(f"resolution={resolution} gives a {ph}x{pw} processed image whose padded latent "
                             f"{(ph + pad_h) // 8}x{(pw + pad_w) // 8} differs from the pipeline's {EH}x{EW} "
                             "(the reference fails at torch.cat for this combination, marigold_dc.py:459)")

        dev = self.device
        imgs, sparses = imgs.to(dev), sparses.to(dev)
        eng = self._engine(N, H, W, resolution, steps)
        with torch.no_grad():
            # marigold_dc.py:661, :677-684 -- first draw of the seeded generator, in the pipeline dtype
            gen = torch.Generator(device=dev).manual_seed(seed)
            common = torch.randn((1, 4, EH, EW), device=dev, dtype=self.dtype, generator=gen).repeat(N, 1, 1, 1)
            prologue.check_image(imgs)
            x = common if pred_latents_prev is None else beta * common + (1 - beta) * pred_latents_prev.to(dev)
        # marigold_dc.py:687-789 inside libmdc_b200.so (mdc_begin_frame): image preprocess + VAE encoder, sparse-depth
        # normalisation (mask, masked min / max, clamp, guide and its min / max), per-call optimiser state.
        # An empty mask raises ValueError like utils.py:132-136.
        # the reference only looks at `percentile` when norm == "percentile" (marigold_dc.py:628-629, :715-728)
        eng.set_options(projection, inv, opt, loss_funcs, kld, kld_weight, kld_mode,
                        percentile if norm == "percentile" else (0.01, 0.99), closed_form=bool(closed_form and train_latents),
                        interp_mode=interp_mode)
        eng.begin_frame(imgs, sparses, x, max_depth, min_depth, norm, lr_latent, lr_scaling)
        if _begin_only:  # bench.py: leave the engine at step 0 with everything resident in HBM
            return None, None
        if train_latents:
            eng.run(steps)                   # marigold_dc.py:799-904, no host sync inside
            denses = eng.decode_final(closed_form=bool(closed_form))  # marigold_dc.py:970-984
        else:                                # no-grad branch: plain DDIM sampling + closed-form affine (:905-909, :53-128)
            eng.sample(steps)
            denses = eng.decode_final(closed_form=True)
        x_out, scales, shifts, losses = eng.get_state()
        self.last_scales, self.last_shifts, self.last_losses = scales, shifts, losses
        return denses, x_out


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous frame shard of rank `rank`: frames are independent, weights replicated."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
