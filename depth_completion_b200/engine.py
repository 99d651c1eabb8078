"""ctypes binding of the C ABI in include/mdc.h: one StepEngine per (device, shapes, model config).

All arithmetic of the guided loop happens in libmdc_b200.so; this class only marshals device pointers of torch
tensors and keeps the handle alive.  There is no fallback: every method raises MdcError if the library is missing
or a call fails.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import MdcError, check, lib, ptr
from .config import UNetConfig, VAEConfig, processed_geometry

MAX_BLOCKS = 8


class MdcConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("n_batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("proc_h", C.c_int), ("proc_w", C.c_int), ("pad_h", C.c_int), ("pad_w", C.c_int), ("steps", C.c_int),
        ("unet_in_ch", C.c_int), ("unet_out_ch", C.c_int), ("unet_nblocks", C.c_int),
        ("unet_layers_per_block", C.c_int), ("unet_groups", C.c_int), ("cross_dim", C.c_int),
        ("unet_block_ch", C.c_int * MAX_BLOCKS), ("unet_heads", C.c_int * MAX_BLOCKS),
        ("unet_down_attn", C.c_int * MAX_BLOCKS),
        ("vae_nblocks", C.c_int), ("vae_layers_per_block", C.c_int), ("vae_groups", C.c_int),
        ("vae_latent_ch", C.c_int), ("vae_block_ch", C.c_int * MAX_BLOCKS), ("vae_scaling", C.c_float),
        ("vae_kind", C.c_int), ("tiny_enc_blocks", C.c_int * MAX_BLOCKS), ("tiny_dec_blocks", C.c_int * MAX_BLOCKS),
        ("tiny_magnitude", C.c_float), ("concurrent", C.c_int),
    ]


def _declare(l):
    if getattr(l, "_mdc_declared", False):
        return
    l.mdc_create.argtypes = [C.POINTER(MdcConfig), C.POINTER(C.c_void_p)]
    l.mdc_create_shared.argtypes = [C.POINTER(MdcConfig), C.c_void_p, C.POINTER(C.c_void_p)]
    l.mdc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    l.mdc_set_weights.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
    l.mdc_weights_loaded.argtypes = [C.c_void_p]
    l.mdc_release_workspace.argtypes = [C.c_void_p]
    l.mdc_weight_is_loaded.argtypes = [C.c_void_p, C.c_int]
    l.mdc_destroy.argtypes = [C.c_void_p]
    l.mdc_destroy.restype = None
    l.mdc_num_weights.argtypes = [C.c_void_p]
    l.mdc_weight_key.argtypes = [C.c_void_p, C.c_int]
    l.mdc_weight_key.restype = C.c_char_p
    l.mdc_weight_shape.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    l.mdc_set_weight.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.c_int]
    l.mdc_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    l.mdc_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_float, C.c_float]
    l.mdc_run.argtypes = [C.c_void_p, C.c_int]
    l.mdc_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.mdc_decode_final.argtypes = [C.c_void_p, C.c_void_p]
    l.mdc_sample.argtypes = [C.c_void_p, C.c_int]
    l.mdc_decode_final_closed_form.argtypes = [C.c_void_p, C.c_void_p]
    l.mdc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    l.mdc_begin_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                  C.c_int, C.c_float, C.c_float]
    l.mdc_begin_frame_encoded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_float, C.c_int, C.c_float, C.c_float]
    l.mdc_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_int, C.c_int]
    l.mdc_launch_count.argtypes = [C.c_void_p]
    l.mdc_launch_count.restype = C.c_longlong
    l.mdc_device_bytes.argtypes = [C.c_void_p]
    l.mdc_device_bytes.restype = C.c_longlong
    l.mdc_dbg_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    l.mdc_dbg_frame_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.mdc_dbg_backward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    l.mdc_dbg_read_tensor.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    l.mdc_dbg_tensor_shape.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.mdc_dbg_num_tensors.argtypes = [C.c_void_p]
    l.mdc_dbg_tensor_name.argtypes = [C.c_void_p, C.c_int]
    l.mdc_dbg_tensor_name.restype = C.c_char_p
    l.mdc_dbg_read_x_adam.argtypes = [C.c_void_p, C.c_void_p]
    l.mdc_dbg_read_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.mdc_dbg_loss.argtypes = [C.c_void_p] * 6
    l.mdc_dbg_update.argtypes = [C.c_void_p] * 4
    l.mdc_dbg_set_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.mdc_dbg_profile_gemm_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.mdc_dbg_profile_ops.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    l.mdc_dbg_time_tapes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    l._mdc_declared = True


class StepEngine:
    """Owns one mdc_handle: the UNet + VAE-decoder tapes and workspace for fixed (N, H, W, resolution, steps)."""

    def __init__(self, unet_cfg: UNetConfig, vae_cfg: VAEConfig, n_batch: int, height: int, width: int,
                 resolution: int, steps: int, device: torch.device | int = 0, share_weights_with: "StepEngine | None" = None,
                 concurrent: bool = False):
        """share_weights_with: another live engine of the same model on the same device whose packed parameters this
        engine reuses (mdc_create_shared) -- a new frame geometry then costs workspace only, no re-packing."""
        self._h = C.c_void_p(0)
        self.lib = lib()
        _declare(self.lib)
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if dev.type != "cuda":
            raise MdcError("the guided loop runs on a CUDA device only (no CPU path exists)")
        self.device = dev
        ph, pw, pad_h, pad_w = processed_geometry(height, width, resolution)
        self.n, self.H, self.W, self.steps = n_batch, height, width, steps
        self.ph, self.pw, self.pad = ph, pw, (pad_h, pad_w)
        self.lh, self.lw = (ph + pad_h) // 8, (pw + pad_w) // 8
        self.unet_cfg, self.vae_cfg = unet_cfg, vae_cfg
        c = MdcConfig()
        c.device = dev.index or 0
        c.n_batch, c.height, c.width = n_batch, height, width
        c.proc_h, c.proc_w, c.pad_h, c.pad_w, c.steps = ph, pw, pad_h, pad_w, steps
        c.unet_in_ch, c.unet_out_ch = unet_cfg.in_channels, unet_cfg.out_channels
        c.unet_nblocks, c.unet_layers_per_block = len(unet_cfg.block_out_channels), unet_cfg.layers_per_block
        c.unet_groups, c.cross_dim = unet_cfg.norm_num_groups, unet_cfg.cross_attention_dim
        for i, v in enumerate(unet_cfg.block_out_channels):
            c.unet_block_ch[i] = v
            c.unet_heads[i] = unet_cfg.attention_heads[i]
            c.unet_down_attn[i] = int(unet_cfg.down_attention[i])
        c.vae_nblocks, c.vae_layers_per_block = len(vae_cfg.block_out_channels), vae_cfg.layers_per_block
        c.vae_groups, c.vae_latent_ch, c.vae_scaling = vae_cfg.norm_num_groups, vae_cfg.latent_channels, vae_cfg.scaling_factor
        for i, v in enumerate(vae_cfg.block_out_channels):
            c.vae_block_ch[i] = v
        c.vae_kind = 1 if vae_cfg.kind == "tiny" else 0
        if c.vae_kind:
            for i, (ne, nd) in enumerate(zip(vae_cfg.num_encoder_blocks, vae_cfg.num_decoder_blocks)):
                c.tiny_enc_blocks[i], c.tiny_dec_blocks[i] = ne, nd
            c.tiny_magnitude = vae_cfg.latent_magnitude
        c.concurrent = int(bool(concurrent))  # several handles in flight on one GPU: no grid-barrier kernels (mdc.h)
        with torch.cuda.device(dev):
            if share_weights_with is not None:
                check(self.lib.mdc_create_shared(C.byref(c), share_weights_with._h, C.byref(self._h)))
            else:
                check(self.lib.mdc_create(C.byref(c), C.byref(self._h)))
        self._stream = None
        self._keep = []

    def _on_stream(self):
        """Hands torch's CURRENT stream on the engine's device to the library (mdc_set_stream) so that the library's
        kernels are ordered with the caller's torch ops exactly as the reference's own torch ops would be."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        if st != self._stream:
            check(self.lib.mdc_set_stream(self._h, C.c_void_p(st)))
            self._stream = st

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.mdc_destroy(self._h)
            self._h = C.c_void_p(0)

    def release_workspace(self):
        """Frees tapes and workspace but keeps the packed weights alive for `share_weights_with` (mdc_release_workspace)."""
        check(self.lib.mdc_release_workspace(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def weight_keys(self):
        n = self.lib.mdc_num_weights(self._h)
        return [self.lib.mdc_weight_key(self._h, i).decode() for i in range(n)]

    def weight_shapes(self) -> dict:
        """{key: shape} of every parameter the tapes need (mdc_weight_shape)."""
        out = {}
        for i in range(self.lib.mdc_num_weights(self._h)):
            shp, nd = (C.c_longlong * 4)(), C.c_int(0)
            check(self.lib.mdc_weight_shape(self._h, i, shp, C.byref(nd)))
            out[self.lib.mdc_weight_key(self._h, i).decode()] = tuple(shp[: nd.value])
        return out

    def weights_loaded(self) -> bool:
        return bool(self.lib.mdc_weights_loaded(self._h))

    def load_weights(self, unet_sd: dict, vae_sd: dict, only_missing: bool = False):
        """Hands every parameter the tapes need to mdc_set_weights in ONE call (keys: 'unet.' / 'vae.' + diffusers name);
        the library re-packs all of them with a single kernel launch.  only_missing: just the parameters some layout of
        which is not packed yet (an engine sharing another's weight bank)."""
        self._on_stream()
        missing, keys, tensors = [], [], []
        for i, key in enumerate(self.weight_keys()):
            if only_missing and self.lib.mdc_weight_is_loaded(self._h, i):
                continue
            prefix, name = key.split(".", 1)
            sd = unet_sd if prefix == "unet" else vae_sd
            if name not in sd:
                missing.append(key)
                continue
            t = sd[name].detach()
            if t.dtype not in (torch.float32, torch.bfloat16):
                t = t.float()
            keys.append(key.encode())
            tensors.append(t.to(self.device).contiguous())
        if missing:
            raise MdcError(f"state dicts lack {len(missing)} parameters, e.g. {missing[:3]}")
        n = len(keys)
        shapes = (C.c_longlong * (4 * n))(*([1] * (4 * n)))
        for i, t in enumerate(tensors):
            if t.ndim > 4:
                raise MdcError(f"parameter {keys[i].decode()} has rank {t.ndim}")
            for j, d in enumerate(t.shape):
                shapes[4 * i + j] = d
        with torch.cuda.device(self.device):
            check(self.lib.mdc_set_weights(
                self._h, n, (C.c_char_p * n)(*keys), (C.c_void_p * n)(*[t.data_ptr() for t in tensors]), shapes,
                (C.c_int * n)(*[t.ndim for t in tensors]), (C.c_int * n)(*[0 if t.dtype == torch.float32 else 1 for t in tensors])))

    def prepare(self, ctx: torch.Tensor, alphas_cumprod: torch.Tensor, timesteps: np.ndarray):
        ctx = ctx.to(self.device, torch.bfloat16).contiguous()
        assert ctx.numel() == 2 * self.unet_cfg.cross_attention_dim, "empty-prompt embedding must be [1, 2, cross_dim]"
        ac = np.ascontiguousarray(alphas_cumprod.detach().float().cpu().numpy())
        ts = np.ascontiguousarray(np.asarray(timesteps, dtype=np.int32))
        if ac.shape[0] != 1000 or ts.shape[0] != self.steps:
            raise ValueError(f"expected 1000 alphas_cumprod and {self.steps} timesteps, got {ac.shape[0]} and {ts.shape[0]}")
        self._on_stream()
        check(self.lib.mdc_prepare(self._h, ptr(ctx), ac.ctypes.data_as(C.c_void_p), ts.ctypes.data_as(C.c_void_p),
                                   int(ts.shape[0])))

    # ------------------------------------------------------------------ per call
    def begin(self, img_latents, x, guide, mask, guide_minmax, depth_minmax, lr_latent=0.05, lr_scaling=0.005):
        self._on_stream()
        il = img_latents.to(self.device, torch.bfloat16).contiguous()
        x = x.to(self.device, torch.bfloat16).contiguous()
        guide = guide.to(self.device, torch.float32).contiguous()
        mask = mask.to(self.device, torch.uint8).contiguous()
        assert tuple(il.shape) == (self.n, 4, self.lh, self.lw) and tuple(x.shape) == tuple(il.shape), (il.shape, x.shape)
        assert guide.numel() == self.n * self.H * self.W and mask.numel() == guide.numel()
        gmm = np.ascontiguousarray(np.asarray(guide_minmax, dtype=np.float32).reshape(self.n, 2))
        dmm = np.ascontiguousarray(np.asarray(depth_minmax, dtype=np.float32).reshape(self.n, 2))
        check(self.lib.mdc_begin(self._h, ptr(il), ptr(x), ptr(guide), ptr(mask), gmm.ctypes.data_as(C.c_void_p),
                                 dmm.ctypes.data_as(C.c_void_p), float(lr_latent), float(lr_scaling)))

    def run(self, n_steps: int):
        self._on_stream()
        check(self.lib.mdc_run(self._h, int(n_steps)))

    def sample(self, n_steps: int):
        """n plain DDIM steps without guidance: the train_latents=False branch (marigold_dc.py:905-909)."""
        self._on_stream()
        check(self.lib.mdc_sample(self._h, int(n_steps)))

    def get_state(self):
        self._on_stream()
        x = torch.empty(self.n, 4, self.lh, self.lw, device=self.device, dtype=torch.bfloat16)
        sc = np.zeros(self.n, np.float32)
        sh = np.zeros(self.n, np.float32)
        ls = np.zeros(self.n, np.float32)
        check(self.lib.mdc_get_state(self._h, ptr(x), sc.ctypes.data_as(C.c_void_p), sh.ctypes.data_as(C.c_void_p),
                                     ls.ctypes.data_as(C.c_void_p)))
        return x, torch.from_numpy(sc), torch.from_numpy(sh), torch.from_numpy(ls)

    def encode(self, imgs: torch.Tensor) -> torch.Tensor:
        """Per-frame prologue in the library: preprocess + VAE encoder -> img_latents [N,4,EH,EW] bf16
        (marigold_dc.py:687-698).  imgs: [N, 1|3, H, W] uint8, or floating point in [0, 1]."""
        self._on_stream()
        imgs, dt = self._image_arg(imgs)
        out = torch.empty(self.n, 4, self.lh, self.lw, device=self.device, dtype=torch.bfloat16)
        check(self.lib.mdc_encode(self._h, ptr(imgs), dt, int(imgs.shape[1]), ptr(out)))
        return out

    def _image_arg(self, imgs: torch.Tensor):
        if imgs.ndim != 4 or imgs.shape[0] != self.n or tuple(imgs.shape[-2:]) != (self.H, self.W):
            raise ValueError(f"imgs shape {tuple(imgs.shape)} does not match the engine ({self.n}, C, {self.H}, {self.W})")
        if imgs.dtype == torch.uint8:
            dt = 0
        elif torch.is_floating_point(imgs):
            imgs, dt = imgs.float(), 1
        else:
            raise ValueError(f"Image dtype={imgs.dtype} is not supported.")
        return imgs.to(self.device).contiguous(), dt

    PROJECTIONS = {"linear": 0, "log": 1, "log10": 2}
    OPTIMIZERS = {"adam": 0, "sgd": 1, "adagrad": 2}
    KLD_MODES = {"simple": 1, "strict": 2}
    NORMS = {"minmax": 0, "const": 1, "percentile": 2}

    def set_options(self, projection="linear", inv=False, opt="adam", loss_funcs=("l1", "l2"), kld=False, kld_weight=0.1,
                    kld_mode="simple", percentile=(0.01, 0.99), closed_form=False, interp_mode="bilinear"):
        """Non-default branches of the reference call (mdc_set_options; marigold_dc.py:467-493): they apply to the next
        begin / begin_frame.  loss_funcs is the reference's list (a term listed twice counts twice, :177-236)."""
        self._on_stream()
        if projection not in self.PROJECTIONS:
            raise ValueError(f"Unknown projection method: {projection}")
        if opt not in self.OPTIMIZERS:
            raise ValueError(f"Unknown optimizer: {opt}")
        if interp_mode not in ("bilinear", "nearest"):
            raise NotImplementedError(f"interp_mode='{interp_mode}' (the reference CLI offers bilinear and nearest)")
        if kld and kld_mode not in self.KLD_MODES:
            raise ValueError(f"Unknown mode: {kld_mode}")
        w = np.array([sum(f == k for f in loss_funcs) for k in ("l1", "l2", "edge", "smooth")], dtype=np.float32)
        check(self.lib.mdc_set_options(self._h, self.PROJECTIONS[projection], int(bool(inv)), self.OPTIMIZERS[opt],
                                       w.ctypes.data_as(C.c_void_p), self.KLD_MODES[kld_mode] if kld else 0,
                                       float(kld_weight), float(percentile[0]), float(percentile[1]), int(bool(closed_form)),
                                       int(interp_mode == "nearest")))

    def begin_frame(self, imgs, sparses, x, max_depth, min_depth=0.0, norm="minmax", lr_latent=0.05, lr_scaling=0.005,
                    img_latents=None):
        """The per-frame prologue in one library call (mdc_begin_frame): image preprocess + VAE encoder, sparse-depth
        normalisation, per-call state (marigold_dc.py:687-789).  Raises ValueError for a sample with an empty mask.
        `img_latents` ([N,4,EH,EW] bf16 from an earlier `encode`, possibly of another engine on another stream) skips the
        encoder (mdc_begin_frame_encoded)."""
        self._on_stream()
        imgs, dt = self._image_arg(imgs)
        sparses = sparses.to(self.device, torch.float32).contiguous()
        x = x.to(self.device, torch.bfloat16).contiguous()
        assert sparses.numel() == self.n * self.H * self.W and tuple(x.shape) == (self.n, 4, self.lh, self.lw)
        try:
            if img_latents is None:
                check(self.lib.mdc_begin_frame(self._h, ptr(imgs), dt, int(imgs.shape[1]), ptr(sparses), ptr(x), float(max_depth),
                                               float(min_depth), self.NORMS[norm], float(lr_latent), float(lr_scaling)))
            else:
                if tuple(img_latents.shape) != (self.n, 4, self.lh, self.lw) or img_latents.dtype != torch.bfloat16:
                    raise ValueError(f"img_latents must be bf16 [{self.n}, 4, {self.lh}, {self.lw}], got {img_latents.dtype} "
                                     f"{tuple(img_latents.shape)}")
                img_latents = img_latents.to(self.device).contiguous()
                check(self.lib.mdc_begin_frame_encoded(self._h, ptr(img_latents), ptr(imgs), dt, int(imgs.shape[1]), ptr(sparses),
                                                       ptr(x), float(max_depth), float(min_depth), self.NORMS[norm],
                                                       float(lr_latent), float(lr_scaling)))
        except MdcError as e:
            if "No valid values found in mask" in str(e) or "min_depth must be" in str(e):
                raise ValueError(str(e)) from None
            raise

    def decode_final(self, closed_form: bool = False) -> torch.Tensor:
        """Final decode + affine + clamp + de-normalisation (marigold_dc.py:970-984); closed_form=True fits scale / shift
        by masked least squares (:53-128) instead of using the learned ones."""
        self._on_stream()
        out = torch.empty(self.n, 1, self.H, self.W, device=self.device, dtype=torch.float32)
        fn = self.lib.mdc_decode_final_closed_form if closed_form else self.lib.mdc_decode_final
        check(fn(self._h, ptr(out)))
        return out

    def launch_count(self) -> int:
        return int(self.lib.mdc_launch_count(self._h))

    def profile_gemm_step(self) -> dict:
        """In-situ CUDA-event timing of every tcgen05 GEMM / conv launch of one guided step's launch sequence."""
        self._on_stream()
        ms, fl, n = C.c_float(0), C.c_double(0), C.c_int(0)
        check(self.lib.mdc_dbg_profile_gemm_step(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return dict(ms=ms.value, flops=fl.value, launches=n.value, tflops=fl.value / max(ms.value, 1e-9) / 1e9)

    def device_bytes(self) -> int:
        return int(self.lib.mdc_device_bytes(self._h))

    # ------------------------------------------------------------------ debug (tests / profiling)
    def _io_shapes(self, which):
        if which == 0:
            return (self.n, 8, self.lh, self.lw), (self.n, 4, self.lh, self.lw)
        return (self.n, 4, self.lh, self.lw), (self.n, 3, self.lh * 8, self.lw * 8)

    def dbg_forward(self, which: int, step: int, x: torch.Tensor) -> torch.Tensor:
        self._on_stream()
        ishape, oshape = self._io_shapes(which)
        x = x.to(self.device, torch.float32).contiguous()
        assert tuple(x.shape) == ishape, (x.shape, ishape)
        out = torch.empty(oshape, device=self.device, dtype=torch.float32)
        check(self.lib.mdc_dbg_forward(self._h, which, step, ptr(x), ptr(out)))
        return out

    def dbg_backward(self, which: int, dout: torch.Tensor) -> torch.Tensor:
        self._on_stream()
        ishape, oshape = self._io_shapes(which)
        dout = dout.to(self.device, torch.float32).contiguous()
        assert tuple(dout.shape) == oshape
        din = torch.empty(ishape, device=self.device, dtype=torch.float32)
        check(self.lib.mdc_dbg_backward(self._h, which, ptr(dout), ptr(din)))
        return din

    def dbg_tensor_names(self):
        n = self.lib.mdc_dbg_num_tensors(self._h)
        return [self.lib.mdc_dbg_tensor_name(self._h, i).decode() for i in range(n)]

    def dbg_read(self, name: str, grad: bool = False) -> torch.Tensor:
        self._on_stream()
        shp = (C.c_int * 4)()
        check(self.lib.mdc_dbg_tensor_shape(self._h, name.encode(), shp))
        out = torch.empty(tuple(shp), device=self.device, dtype=torch.float32)
        check(self.lib.mdc_dbg_read_tensor(self._h, name.encode(), int(grad), ptr(out)))
        return out

    def dbg_x_adam(self) -> torch.Tensor:
        self._on_stream()
        x = torch.empty(self.n, 4, self.lh, self.lw, device=self.device, dtype=torch.bfloat16)
        check(self.lib.mdc_dbg_read_x_adam(self._h, ptr(x)))
        return x

    def dbg_buffer(self, which: str) -> torch.Tensor:
        self._on_stream()
        out = torch.empty(self.n, 4, self.lh, self.lw, device=self.device, dtype=torch.float32)
        check(self.lib.mdc_dbg_read_buffer(self._h, which.encode(), ptr(out)))
        return out

    def dbg_frame_state(self):
        self._on_stream()
        guide = torch.empty(self.n, 1, self.H, self.W, device=self.device, dtype=torch.float32)
        mask = torch.empty(self.n, 1, self.H, self.W, device=self.device, dtype=torch.uint8)
        st = np.zeros((self.n, 5), np.float32)
        check(self.lib.mdc_dbg_frame_state(self._h, ptr(guide), ptr(mask), st.ctypes.data_as(C.c_void_p)))
        return guide, mask.bool(), st

    def dbg_loss(self, dec_nchw: torch.Tensor):
        self._on_stream()
        dec = dec_nchw.to(self.device, torch.float32).contiguous()
        assert tuple(dec.shape) == (self.n, 3, self.lh * 8, self.lw * 8)
        ddec = torch.empty_like(dec)
        ls, gs, gt = (np.zeros(self.n, np.float32) for _ in range(3))
        check(self.lib.mdc_dbg_loss(self._h, ptr(dec), ptr(ddec), ls.ctypes.data_as(C.c_void_p),
                                    gs.ctypes.data_as(C.c_void_p), gt.ctypes.data_as(C.c_void_p)))
        return ddec, torch.from_numpy(ls), torch.from_numpy(gs), torch.from_numpy(gt)

    def dbg_update(self, v, dz, dunet_in):
        self._on_stream()
        v, dz, du = (t.to(self.device, torch.float32).contiguous() for t in (v, dz, dunet_in))
        assert tuple(v.shape) == (self.n, 4, self.lh, self.lw) and tuple(du.shape) == (self.n, 8, self.lh, self.lw)
        check(self.lib.mdc_dbg_update(self._h, ptr(v), ptr(dz), ptr(du)))

    def dbg_set_state(self, step: int, x=None, exp_avg=None, exp_avg_sq=None, affine6=None):
        """Teacher forcing (mdc_dbg_set_state): latent and its Adam moments [N,4,EH,EW], affine6 = [6, N] floats (scale,
        shift, and the fp32 Adam moments s_m, s_v, t_m, t_v); the next run(1) executes guided step `step`."""
        self._on_stream()
        keep = [None if t is None else t.to(self.device, torch.bfloat16).contiguous() for t in (x, exp_avg, exp_avg_sq)]
        a6 = None
        if affine6 is not None:
            a6 = np.ascontiguousarray(np.asarray(affine6, dtype=np.float32).reshape(6, self.n))
        check(self.lib.mdc_dbg_set_state(self._h, int(step), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]),
                                         a6.ctypes.data_as(C.c_void_p) if a6 is not None else C.c_void_p(0)))

    def dbg_profile_ops(self, csv_path: str, iters: int = 5):
        self._on_stream()
        check(self.lib.mdc_dbg_profile_ops(self._h, csv_path.encode(), iters))

    def dbg_time_tapes(self, iters: int = 3):
        self._on_stream()
        ms = (C.c_float * 4)()
        check(self.lib.mdc_dbg_time_tapes(self._h, iters, ms))
        return dict(unet_fwd=ms[0], unet_bwd=ms[1], dec_fwd=ms[2], dec_bwd=ms[3])
