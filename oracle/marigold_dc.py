"""ORACLE (test infrastructure): restatement of the reference's guided denoising path,
MarigoldDepthCompletionPipeline.__call__ (marigold_dc.py:467-985), default branch only
(projection="linear", inv=False, opt="adam", train_latents=True, train_method="per-step",
closed_form=False, loss l1+l2, no kld), plus the helpers it uses (marigold_dc.py:53-128, :131-193,
:284-371; utils.py:89-138, :692-739).

Differences from the reference, none of which change results:
  * UNet/VAE weights are frozen (the reference also accumulates unused weight gradients);
  * the empty-prompt embedding is passed in instead of being produced by a CLIP text encoder;
  * optional `trace` callback to expose per-step tensors for teacher-forced parity tests.

PARITY UNPINNED: see oracle/__init__.py.
"""
from __future__ import annotations

import torch

from . import image_processor as ip
from .scheduler import DDIMScheduler

EPSILON = 1e-7


# ------------------------------------------------------------------ helpers (utils.py / marigold_dc.py)
def masked_minmax(x: torch.Tensor, mask: torch.Tensor, dim=None):
    """utils.py:89-138 -- masked min / max, ValueError when a row has no valid element."""
    if x.shape != mask.shape:
        raise ValueError(f"Shape of x {x.shape} must be equal to shape of mask {mask.shape}")
    lo = torch.where(mask, x, torch.full_like(x, float("inf")))
    hi = torch.where(mask, x, torch.full_like(x, float("-inf")))
    mins, maxs = (lo.min(), hi.max()) if dim is None else (lo.amin(dim=dim), hi.amax(dim=dim))
    if torch.isinf(mins).any() or torch.isinf(maxs).any():
        raise ValueError("No valid values found in mask for some positions.")
    return mins, maxs


def compute_affine_params(affines, guides, masks):
    """marigold_dc.py:53-128 -- closed-form masked least-squares scale / shift per sample."""
    n = affines.shape[0]
    a, g, m = affines.reshape(n, -1), guides.reshape(n, -1), masks.reshape(n, -1)
    cnt = m.sum(dim=1, keepdim=True)
    if torch.any(cnt == 0):
        raise ValueError("At least one mask in the batch has no valid points")
    a_mean = (a * m).sum(1, keepdim=True) / cnt
    g_mean = (g * m).sum(1, keepdim=True) / cnt
    ac, gc = (a - a_mean) * m, (g - g_mean) * m
    scales = (ac * gc).sum(1, keepdim=True) / (ac.pow(2).sum(1, keepdim=True) + EPSILON)
    shifts = g_mean - scales * a_mean
    return scales.squeeze(1), shifts.squeeze(1)


def compute_loss(denses, sparses, masks, loss_funcs=("l1", "l2")):
    """marigold_dc.py:171-193 -- per-sample masked mean L1 + masked mean L2 -> [N]."""
    if len(loss_funcs) == 0:
        raise ValueError("loss_funcs must contain at least one loss function")
    total = torch.zeros(denses.shape[0], device=denses.device)
    cnt = masks.sum(dim=(1, 2, 3))
    for f in loss_funcs:
        if f == "l1":
            total = total + ((denses - sparses).abs() * masks).sum(dim=(1, 2, 3)) / cnt
        elif f == "l2":
            total = total + (((denses - sparses) ** 2) * masks).sum(dim=(1, 2, 3)) / cnt
        else:
            raise ValueError(f"Unknown loss function: {f}")
    return total


def mae(preds, targets, masks=None):
    """utils.py:692-714."""
    if masks is not None:
        preds, targets = preds[masks], targets[masks]
    return (preds - targets).abs().mean()


def rmse(preds, targets, masks=None):
    """utils.py:717-739."""
    if masks is not None:
        preds, targets = preds[masks], targets[masks]
    return ((preds - targets) ** 2).mean().sqrt()


def latent_size(H: int, W: int, resolution: int):
    """marigold_dc.py:596-597 (the reference's own latent-size formula)."""
    return resolution * H // (8 * max(H, W)), resolution * W // (8 * max(H, W))


def make_empty_text_embedding(dim: int = 1024, seed: int = 4321, device="cpu", dtype=torch.float32):
    """Stand-in for text_encoder(tokenizer(""))[0] -> [1, 2, dim] (marigold_dc.py:664-674); seeded randn."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(1, 2, dim, generator=g).to(device=device, dtype=dtype)


# ------------------------------------------------------------------ the pipeline
class OraclePipeline:
    """Call-compatible with the reference pipeline for the default guided path."""

    def __init__(self, unet, vae, empty_text_embedding, scheduler: DDIMScheduler | None = None):
        self.unet = unet.requires_grad_(False)
        self.vae = vae.requires_grad_(False)
        self.scheduler = scheduler or DDIMScheduler()
        self.empty_text_embedding = empty_text_embedding
        p = next(unet.parameters())
        self.device, self.dtype = p.device, p.dtype

    # marigold_dc.py:366 -> MarigoldDepthPipeline.decode_prediction (Appendix A.2)
    def decode_prediction(self, z):
        y = self.vae.decode(z / self.vae.cfg.scaling_factor)
        y = y.mean(dim=1, keepdim=True)
        y = torch.clip(y, -1.0, 1.0)
        return (y + 1.0) / 2.0

    # marigold_dc.py:338-371
    def latent_to_affine(self, z, orig_res, padding, interp_mode="bilinear"):
        a = self.decode_prediction(z)
        a = ip.unpad_image(a, padding)
        return ip.resize_antialias(a, orig_res, interp_mode)

    # marigold_dc.py:320-331 (learned scale/shift branch)
    @staticmethod
    def affine_to_metric(affines, guides, masks, scales, shifts):
        n = affines.shape[0]
        mins, maxs = masked_minmax(guides.view(n, -1), masks.view(n, -1), dim=-1)
        mins, maxs = mins.view(n, 1, 1, 1), maxs.view(n, 1, 1, 1)
        return (scales ** 2) * (maxs - mins) * affines + (shifts ** 2) * mins

    def predict_noise(self, img_latents, x, t):
        n = x.shape[0]
        return self.unet(torch.cat([img_latents, x], dim=1), t, self.empty_text_embedding.repeat(n, 1, 1))

    def preprocess(self, imgs, sparses, max_depth, min_depth, norm, resolution, seed, pred_latents_prev, beta):
        """marigold_dc.py:659-756 (linear projection).  Returns a dict of per-call constants."""
        N, _, H, W = imgs.shape
        EH, EW = latent_size(H, W, resolution)
        with torch.no_grad():
            gen = torch.Generator(device=self.device).manual_seed(seed)
            common = torch.randn((1, 4, EH, EW), device=imgs.device, dtype=self.dtype, generator=gen).repeat(N, 1, 1, 1)
            imgs_resized, padding, orig_res = ip.preprocess(imgs, resolution, self.device, self.dtype)
            img_latents = self.vae.encode_mode(imgs_resized) * self.vae.cfg.scaling_factor
            x = common if pred_latents_prev is None else beta * common + (1 - beta) * pred_latents_prev
            masks = sparses > 0
            if norm == "minmax":
                lo, hi = masked_minmax(sparses.view(N, -1), masks.view(N, -1), dim=-1)
                lo, hi = lo.view(N, 1, 1, 1), hi.view(N, 1, 1, 1)
            elif norm == "const":
                lo = torch.full((N, 1, 1, 1), min_depth, device=sparses.device)
                hi = torch.full((N, 1, 1, 1), max_depth, device=sparses.device)
            else:
                raise ValueError(f"Unknown norm method: {norm}")
            clamped = sparses.clamp(min=lo, max=hi)
            if norm == "minmax":
                lo, hi = lo.clamp(min=min_depth), hi.clamp(max=max_depth)
            normed = (clamped - lo) / (hi - lo)
        return dict(x=x, img_latents=img_latents, masks=masks, sparses_normed=normed, min_depths=lo, max_depths=hi,
                    padding=padding, orig_res=orig_res)

    def guided_step(self, st, t, x, scales, shifts, optimizer, trace=None, idx=0):
        """One iteration of marigold_dc.py:801-904.  x, scales, shifts are Parameters updated in place."""
        N = x.shape[0]
        optimizer.zero_grad()
        v = self.predict_noise(st["img_latents"], x, t)
        with torch.no_grad():
            a_t = self.scheduler.alphas_cumprod[int(t)]
            eps_hat = (a_t ** 0.5) * v + ((1 - a_t) ** 0.5) * x
        x0 = self.scheduler.step(v, t, x).pred_original_sample
        aff = self.latent_to_affine(x0, st["orig_res"], st["padding"])
        dense = self.affine_to_metric(aff, st["sparses_normed"], st["masks"], scales, shifts).clamp(min=0.0, max=1.0)
        losses = compute_loss(dense, st["sparses_normed"], st["masks"])
        losses.backward(torch.ones_like(losses))
        with torch.no_grad():
            raw_grad = x.grad.detach().clone()
            en = torch.linalg.norm(eps_hat.reshape(N, -1), dim=1)
            gn = torch.linalg.norm(x.grad.reshape(N, -1), dim=1)
            x.grad *= (en / gn.clamp(min=EPSILON)).view(N, 1, 1, 1)
        x_before = x.detach().clone() if trace is not None else None
        optimizer.step()
        with torch.no_grad():
            x_adam = x.detach().clone() if trace is not None else None
            x.data = self.scheduler.step(v, t, x).prev_sample
        if trace is not None:
            trace(dict(idx=idx, t=int(t), x_in=x_before, v=v.detach(), x0=x0.detach(), losses=losses.detach(),
                       grad=raw_grad, s_grad=scales.grad.detach().clone(), t_grad=shifts.grad.detach().clone(),
                       x_adam=x_adam, x_out=x.detach().clone(), scales=scales.detach().clone(),
                       shifts=shifts.detach().clone(), eps_norm=en, grad_norm=gn))
        return losses.detach()

    def __call__(self, imgs, sparses, max_depth, min_depth=0.0, norm="minmax", pred_latents_prev=None, beta=0.9,
                 steps=50, resolution=768, lr=None, seed=2024, trace=None, max_steps=None):
        if imgs.ndim != 4 or sparses.ndim != 4 or imgs.shape[0] != sparses.shape[0] or imgs.shape[-2:] != sparses.shape[-2:]:
            raise ValueError("Shape of image must be [N, C, H, W] and shape of sparse must be [N, 1, H, W]")
        N = imgs.shape[0]
        lr_latent, lr_scaling = (0.05, 0.005) if lr is None else lr
        st = self.preprocess(imgs, sparses, max_depth, min_depth, norm, resolution, seed, pred_latents_prev, beta)
        x = torch.nn.Parameter(st["x"])
        scales = torch.nn.Parameter(torch.ones(N, 1, 1, 1, device=self.device))
        shifts = torch.nn.Parameter(torch.zeros(N, 1, 1, 1, device=self.device))
        opt = torch.optim.Adam([{"params": [x], "lr": lr_latent}, {"params": [scales, shifts], "lr": lr_scaling}])
        self.scheduler.set_timesteps(steps, device=self.device)
        for i, t in enumerate(self.scheduler.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            self.guided_step(st, t, x, scales, shifts, opt, trace, i)
        with torch.no_grad():
            xd = x.detach()
            aff = self.latent_to_affine(xd, st["orig_res"], st["padding"])
            dense = self.affine_to_metric(aff, st["sparses_normed"], st["masks"], scales, shifts).clamp(min=0.0, max=1.0)
            denses = dense * (st["max_depths"] - st["min_depths"]) + st["min_depths"]
        return denses, xd
