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

from . import ddim
from ._lib import MdcError
from .config import processed_geometry, unet_config_from, vae_config_from
from .engine import StepEngine

SUPPORTED_LOSS_FUNCS = ["l1", "l2", "edge", "smooth"]  # marigold_dc.py:19
EPSILON = 1e-7  # marigold_dc.py:20
EMPTY_PROMPT_IDS = (49406, 49407)  # CLIP BOS, EOS: tokenizer("", padding="do_not_pad")


def check_image(image: torch.Tensor) -> None:
    """The checks of MarigoldImageProcessor.preprocess that raise (reached from
    marigold_dc.py:687-692)."""
    if image.ndim != 4:
        raise ValueError(f"Input image is not 4-dimensional: shape={tuple(image.shape)}")
    if not torch.is_floating_point(image) and image.dtype != torch.uint8:
        raise ValueError(f"Image dtype={image.dtype} is not supported.")
    if image.shape[1] not in (1, 3):
        raise ValueError(f"Input image is not 1- or 3-channel: {tuple(image.shape)}.")
    if torch.is_floating_point(image) and (image.min().item() < 0.0 or image.max().item() > 1.0):
        raise ValueError("Input image data is partially outside of the [0,1] range.")


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
        # The engine computes in bf16 with fp32 accumulation (BASELINE.json configs b-e, predict.py --dtype bf16).  A
        # pipeline built from fp32 modules (predict.py:463-481 with torch_dtype=float32) would silently get bf16
        # arithmetic, so that request is refused unless the caller opts in (weights are then rounded to bf16).
        self.dtype = torch.bfloat16
        self.allow_fp32_modules = True  # random-init / test modules are fp32 containers of bf16-representable values

    # The reference swaps modules on a live pipeline (predict.py:484-488 `pipe.vae = AutoencoderTiny...`, :491-494
    # `pipe.scheduler = DDIMScheduler...`): re-derive the config and drop everything built from the old module.
    def _invalidate(self):
        for eng in list(getattr(self, "_engines", {}).values()) + [getattr(self, "_keeper", None)]:
            if eng is not None:
                eng.close()
        self._engines = {}
        self._keeper = None
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
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, prediction_type=None, torch_dtype=torch.bfloat16, **kwargs):
        """The constructor call of the reference's only caller (predict.py:463-481): loads the diffusers
        `MarigoldDepthPipeline` checkpoint and wraps its modules.  Needs `diffusers` (the reference's own dependency,
        requirements.txt:1); raises ImportError with that hint when it is not installed.  torch_dtype=float32 is refused:
        this library has no fp32 arithmetic path (it would silently compute in bf16 otherwise)."""
        if torch_dtype not in (torch.bfloat16, None):
            raise NotImplementedError(f"torch_dtype={torch_dtype}: the B200 engine computes in bfloat16 only (predict.py --dtype bf16); "
                                      "an fp32 pipeline would silently run at reduced precision")
        try:
            from diffusers import MarigoldDepthPipeline  # noqa: PLC0415
        except ImportError as e:  # pragma: no cover - diffusers is absent from the build image
            raise ImportError("from_pretrained needs the reference's own dependency `diffusers` (requirements.txt:1) to read "
                              "the checkpoint; construct the class from modules / state dicts instead") from e
        base = MarigoldDepthPipeline.from_pretrained(pretrained_model_name_or_path, prediction_type=prediction_type,
                                                     torch_dtype=torch.bfloat16, **kwargs)
        return cls.from_diffusers(base)

    @classmethod
    def from_diffusers(cls, base):
        """Wraps the modules of a loaded diffusers Marigold pipeline (anything with .unet / .vae / .scheduler /
        .text_encoder / .tokenizer), as MarigoldDepthCompletionPipeline(**base.components) does in the reference."""
        pipe = cls(base.unet, base.vae, getattr(base, "scheduler", None), getattr(base, "text_encoder", None),
                   getattr(base, "tokenizer", None), prediction_type=getattr(base, "prediction_type", None),
                   scale_invariant=getattr(base, "scale_invariant", True), shift_invariant=getattr(base, "shift_invariant", True),
                   default_denoising_steps=getattr(base, "default_denoising_steps", None),
                   default_processing_resolution=getattr(base, "default_processing_resolution", None))
        pipe.allow_fp32_modules = False
        return pipe

    def to(self, device=None, dtype=None):
        if isinstance(device, torch.dtype):
            device, dtype = None, device
        if dtype is not None and dtype != torch.bfloat16:
            raise NotImplementedError(f"dtype={dtype}: the B200 engine computes in bfloat16 only")
        if device is None:
            return self
        self.device = torch.device(device)
        for m in (self.unet, self.vae, self.text_encoder):
            if m is not None and hasattr(m, "to"):
                m.to(self.device)
        self._invalidate()  # free the old device's workspace and weights instead of waiting for the collector
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

    MAX_ENGINES = 4            # resident workspaces (one per call geometry), least recently used evicted first
    MAX_ENGINE_BYTES = 110e9   # ... and never more than this much HBM in workspaces (a 4 x 768x1024 engine is 85 GB)

    def _estimate_bytes(self, N, H, W, resolution) -> float:
        """Workspace of a not-yet-built engine, scaled by pixel count from a resident one (0 when there is none)."""
        if not self._engines:
            return 0.0
        ref = list(self._engines.values())[-1]
        ph, pw, pad_h, pad_w = processed_geometry(H, W, resolution)
        return ref.device_bytes() * (N * (ph + pad_h) * (pw + pad_w)) / max(1, ref.n * ref.lh * ref.lw * 64)

    def _engine(self, N, H, W, resolution, steps, slot: int = 0, concurrent: bool = False) -> StepEngine:
        """One StepEngine per call geometry, kept resident (LRU) and all sharing ONE set of packed weights: a sequence
        whose last batch is short, or alternating resolutions, builds workspace for the new shape but never re-packs the
        parameters (the reference serves every shape from one set of modules, predict.py:585-700)."""
        key = (N, H, W, resolution, steps, str(self.device), slot, bool(concurrent))
        eng = self._engines.pop(key, None)
        if eng is None:
            if self.device.type != "cuda":
                raise MdcError("MarigoldDepthCompletionPipeline needs a CUDA device: call .to('cuda') (no CPU path)")
            if not self.allow_fp32_modules:
                bad = [k for k, v in self.unet.state_dict().items() if v.dtype == torch.float32][:1]
                if bad:
                    raise NotImplementedError("the modules are float32 (torch_dtype=float32, predict.py:463-481): the B200 "
                                              "engine computes in bfloat16 only and will not silently lower the precision")
            # make room first (dict order = least recently used first).  The first engine evicted stays behind as a
            # workspace-less keeper of the packed weights, so eviction never forces a re-pack.
            est = self._estimate_bytes(N, H, W, resolution)
            while self._engines and (len(self._engines) >= self.MAX_ENGINES or
                                     sum(e.device_bytes() for e in self._engines.values()) + est > self.MAX_ENGINE_BYTES):
                old = self._engines.pop(next(iter(self._engines)))
                if self._keeper is None:
                    old.release_workspace()
                    self._keeper = old
                else:
                    old.close()
            donor = self._keeper or (list(self._engines.values())[-1] if self._engines else None)
            eng = StepEngine(self.unet_cfg, self.vae_cfg, N, H, W, resolution, steps, self.device, share_weights_with=donor,
                             concurrent=concurrent)
            if not eng.weights_loaded():
                usd, vsd, _ = self._state_dicts()
                eng.load_weights(usd, vsd, only_missing=donor is not None)
            ac, ts = ddim.tables_from_scheduler(self.scheduler, steps)
            eng.prepare(self._empty_embedding(), ac, ts)
        self._engines[key] = eng  # most recently used last
        return eng

    # ------------------------------------------------------------------ the call (marigold_dc.py:467-493)
    def __call__(self, imgs, sparses, max_depth, min_depth=0.0, projection="linear", inv=False, norm="minmax",
                 percentile=(0.01, 0.99), pred_latents_prev=None, beta=0.9, steps=50, resolution=768,
                 closed_form=None, opt="adam", lr=None, kld=False, kld_weight=0.1, kld_mode="simple",
                 interp_mode="bilinear", loss_funcs=None, seed=2024, train_latents=True, train_method="per-step",
                 train_steps=10, _begin_only=False):
        """The reference call (marigold_dc.py:467-493): `submit` enqueues the prologue and the guided steps on the current
        CUDA stream, `collect` decodes and synchronises."""
        ticket = self.submit(imgs, sparses, max_depth, min_depth, projection, inv, norm, percentile, pred_latents_prev, beta, steps,
                             resolution, closed_form, opt, lr, kld, kld_weight, kld_mode, interp_mode, loss_funcs, seed, train_latents,
                             train_method, train_steps, _begin_only=_begin_only)
        if _begin_only:  # bench.py / tests: leave the engine at step 0 with everything resident in HBM
            return None, None
        return self.collect(ticket)

    def submit(self, imgs, sparses, max_depth, min_depth=0.0, projection="linear", inv=False, norm="minmax",
               percentile=(0.01, 0.99), pred_latents_prev=None, beta=0.9, steps=50, resolution=768,
               closed_form=None, opt="adam", lr=None, kld=False, kld_weight=0.1, kld_mode="simple",
               interp_mode="bilinear", loss_funcs=None, seed=2024, train_latents=True, train_method="per-step",
               train_steps=10, _begin_only=False, _slot=0, _concurrent=False, _img_latents=None):
        """First half of the call: argument validation, per-frame prologue and all guided steps, enqueued on the current
        CUDA stream without waiting for them (only the prologue's empty-mask check synchronises, a few ms in).  `_slot`
        selects one of several engines of the same geometry so that frames on different streams can be in flight at once
        (video.complete_sequence(frames_in_flight=...)); returns a ticket for `collect`."""
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
        eng = self._engine(N, H, W, resolution, steps, _slot, _concurrent)
        with torch.no_grad():
            # marigold_dc.py:661, :677-684 -- first draw of the seeded generator, in the pipeline dtype
            gen = torch.Generator(device=dev).manual_seed(seed)
            common = torch.randn((1, 4, EH, EW), device=dev, dtype=self.dtype, generator=gen).repeat(N, 1, 1, 1)
            check_image(imgs)
            x = common if pred_latents_prev is None else beta * common + (1 - beta) * pred_latents_prev.to(dev)
        # marigold_dc.py:687-789 inside libmdc_b200.so (mdc_begin_frame): image preprocess + VAE encoder, sparse-depth
        # normalisation (mask, masked min / max, clamp, guide and its min / max), per-call optimiser state.
        # An empty mask raises ValueError like utils.py:132-136.
        # the reference only looks at `percentile` when norm == "percentile" (marigold_dc.py:628-629, :715-728)
        eng.set_options(projection, inv, opt, loss_funcs, kld, kld_weight, kld_mode,
                        percentile if norm == "percentile" else (0.01, 0.99), closed_form=bool(closed_form and train_latents),
                        interp_mode=interp_mode)
        eng.begin_frame(imgs, sparses, x, max_depth, min_depth, norm, lr_latent, lr_scaling, img_latents=_img_latents)
        if not _begin_only:
            if train_latents:
                eng.run(steps)               # marigold_dc.py:799-904, asynchronous, no host sync inside
            else:                            # no-grad branch: plain DDIM sampling (:905-909)
                eng.sample(steps)
        return dict(eng=eng, closed_form=bool(closed_form) if train_latents else True)

    def encode_ahead(self, imgs, steps=50, resolution=768):
        """The image half of the per-frame prologue (marigold_dc.py:687-698: preprocess + VAE encoder) on its own, enqueued on
        the current CUDA stream by a SECOND engine of the call geometry (shared weights, no grid-wide barriers), so it can
        run under another frame's guided loop (`video.complete_sequence(overlap_prologue=True)`).
        Returns img_latents [N,4,EH,EW] bf16 for `submit(..., _img_latents=...)`."""
        if imgs.ndim != 4:
            raise ValueError(f"Input image is not 4-dimensional: shape={tuple(imgs.shape)}")
        N, _, H, W = imgs.shape
        imgs = imgs.to(self.device, non_blocking=True)
        check_image(imgs)
        eng = self._engine(N, H, W, resolution, steps, "enc", True)
        return eng.encode(imgs)

    def collect(self, ticket):
        """Second half of the call, on the stream `submit` ran on: final decode + de-normalisation (marigold_dc.py:970-984)
        and the returned latents; synchronises that stream."""
        eng = ticket["eng"]
        denses = eng.decode_final(closed_form=ticket["closed_form"])
        x_out, scales, shifts, losses = eng.get_state()
        self.last_scales, self.last_shifts, self.last_losses = scales, shifts, losses
        return denses, x_out


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous frame shard of rank `rank`: frames are independent, weights replicated."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
