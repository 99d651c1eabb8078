"""ORACLE (test infrastructure): restatement of the reference's guided denoising path,
MarigoldDepthCompletionPipeline.__call__ (marigold_dc.py:467-985) with train_method="per-step" and either
train_latents=True (the guided loop, learned or closed-form scale / shift) or train_latents=False (plain sampling + closed-form affine):
every projection (linear / log / log10, inv), norm (minmax /
percentile / const), optimiser (adam / sgd / adagrad), loss term (l1, l2, edge, smooth) and the kld penalty,
plus the helpers it uses (marigold_dc.py:23-50, :53-128, :131-243, :284-371; utils.py:28-86, :89-138, :692-739).

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


def get_projection_fn(projection):
    """marigold_dc.py:23-50."""
    if projection == "log":
        return torch.log
    if projection == "log10":
        return torch.log10
    if projection == "linear":
        return lambda x: x
    raise ValueError(f"Unknown projection method: {projection}")


def kld_stdnorm(x, reduction="mean", mode="simple"):
    """utils.py:28-86 -- KL divergence of the latent from N(0, 1), per sample ("none") or reduced."""
    n = x.shape[0]
    x_ = x.reshape(n, -1)
    eps = torch.finfo(x.dtype).eps
    if mode == "simple":
        dist = x_.square().mean(dim=-1)
    elif mode == "strict":
        mu = x_.mean(dim=-1)
        var = x_.var(dim=-1, unbiased=False)
        dist = 0.5 * (mu.square() + var - torch.log(var + eps) - 1)
    else:
        raise ValueError(f"Unknown mode: {mode}")
    if reduction == "mean":
        return dist.mean()
    if reduction == "sum":
        return dist.sum()
    if reduction == "none":
        return dist
    raise ValueError(f"Unknown reduction: {reduction}")


def compute_loss(denses, sparses, masks, loss_funcs=("l1", "l2"), images=None, kld=False, kld_weight=0.1,
                 kld_mode="simple", pred_latents=None):
    """marigold_dc.py:131-243 -- per-sample sum of the listed terms (a term listed twice counts twice) -> [N]."""
    if len(loss_funcs) == 0:
        raise ValueError("loss_funcs must contain at least one loss function")
    if kld and pred_latents is None:
        raise ValueError("pred_latents must be provided when kl-divergence constraint is enabled")
    total = torch.zeros(denses.shape[0], device=denses.device)
    cnt = masks.sum(dim=(1, 2, 3))
    for f in loss_funcs:
        if f == "l1":
            total = total + ((denses - sparses).abs() * masks).sum(dim=(1, 2, 3)) / cnt
        elif f == "l2":
            total = total + (((denses - sparses) ** 2) * masks).sum(dim=(1, 2, 3)) / cnt
        elif f == "edge":
            if images is None:
                raise ValueError("image must be provided for edge loss")
            c = images.shape[1]
            if c == 3:
                gray = 0.299 * images[:, 0:1] + 0.587 * images[:, 1:2] + 0.114 * images[:, 2:3]
            elif c == 1:
                gray = images
            else:
                raise ValueError(f"Image must have 1 or 3 channels, got {c}")
            px = (denses[:, :, :, :-1] - denses[:, :, :, 1:]).abs()
            py = (denses[:, :, :-1, :] - denses[:, :, 1:, :]).abs()
            gx = (gray[:, :, :, :-1] - gray[:, :, :, 1:]).abs()
            gy = (gray[:, :, :-1, :] - gray[:, :, 1:, :]).abs()
            total = total + (px - gx).abs().mean(dim=(1, 2, 3)) + (py - gy).abs().mean(dim=(1, 2, 3))
        elif f == "smooth":
            if images is None:
                raise ValueError("image must be provided for smooth loss")
            total = total + (denses[:, :, :-1, :] - denses[:, :, 1:, :]).abs().mean(dim=(1, 2, 3)) \
                + (denses[:, :, :, :-1] - denses[:, :, :, 1:]).abs().mean(dim=(1, 2, 3))
        else:
            raise ValueError(f"Unknown loss function: {f}")
    if kld:
        total = total + kld_weight * kld_stdnorm(pred_latents, reduction="none", mode=kld_mode)
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

    def preprocess(self, imgs, sparses, max_depth, min_depth, norm, resolution, seed, pred_latents_prev, beta,
                   projection="linear", inv=False, percentile=(0.01, 0.99)):
        """marigold_dc.py:659-756.  Returns a dict of per-call constants."""
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
            elif norm == "percentile":
                p = torch.tensor(percentile, device=sparses.device)
                ranges = torch.stack([torch.quantile(s[m], p) for s, m in zip(sparses, masks, strict=True)])
                lo, hi = ranges[:, 0].view(N, 1, 1, 1), ranges[:, 1].view(N, 1, 1, 1)
            elif norm == "const":
                lo = torch.full((N, 1, 1, 1), min_depth, device=sparses.device)
                hi = torch.full((N, 1, 1, 1), max_depth, device=sparses.device)
            else:
                raise ValueError(f"Unknown norm method: {norm}")
            clamped = sparses.clamp(min=lo, max=hi)
            if norm in ("minmax", "percentile"):
                lo, hi = lo.clamp(min=min_depth), hi.clamp(max=max_depth)
            proj = get_projection_fn(projection)
            lo_p, hi_p, clamped_p = proj(lo), proj(hi), proj(clamped)
            if inv:
                lo_p, hi_p = 1 / hi_p, 1 / lo_p
                clamped_p = 1 / clamped_p
            normed = (clamped_p - lo_p) / (hi_p - lo_p)
        return dict(x=x, img_latents=img_latents, masks=masks, sparses_normed=normed, min_depths=lo, max_depths=hi,
                    min_depths_proj=lo_p, max_depths_proj=hi_p, projection=projection, inv=inv, imgs=imgs,
                    padding=padding, orig_res=orig_res)

    @staticmethod
    def to_guide_space(dense, st):
        """marigold_dc.py:842-862 -- normalised linear depth -> the projected / inverted space of the guide."""
        if st.get("projection", "linear") == "linear" and not st.get("inv", False):
            return dense
        d = dense * (st["max_depths"] - st["min_depths"]) + st["min_depths"]
        d = get_projection_fn(st["projection"])(d)
        if st["inv"]:
            d = 1 / d
        return (d - st["min_depths_proj"]) / (st["max_depths_proj"] - st["min_depths_proj"])

    def guided_step(self, st, t, x, scales, shifts, optimizer, trace=None, idx=0, loss_kw=None, closed_form=False):
        """One iteration of marigold_dc.py:801-904.  x, scales, shifts are Parameters updated in place."""
        N = x.shape[0]
        optimizer.zero_grad()
        v = self.predict_noise(st["img_latents"], x, t)
        with torch.no_grad():
            a_t = self.scheduler.alphas_cumprod[int(t)]
            eps_hat = (a_t ** 0.5) * v + ((1 - a_t) ** 0.5) * x
        x0 = self.scheduler.step(v, t, x).pred_original_sample
        aff = self.latent_to_affine(x0, st["orig_res"], st["padding"], st.get("interp_mode", "bilinear"))
        if closed_form:  # marigold_dc.py:332-336: refit by least squares, differentiably
            cs, ct = compute_affine_params(aff, st["sparses_normed"], st["masks"])
            dense = (cs.view(N, 1, 1, 1) * aff + ct.view(N, 1, 1, 1)).clamp(min=0.0, max=1.0)
        else:
            dense = self.affine_to_metric(aff, st["sparses_normed"], st["masks"], scales, shifts).clamp(min=0.0, max=1.0)
        dense = self.to_guide_space(dense, st)
        losses = compute_loss(dense, st["sparses_normed"], st["masks"], images=st.get("imgs"), pred_latents=x,
                              **(loss_kw or {}))
        losses.backward(torch.ones_like(losses))
        with torch.no_grad():
            raw_grad = x.grad.detach().clone()
            en = torch.linalg.norm(eps_hat.reshape(N, -1), dim=1)
            gn = torch.linalg.norm(x.grad.reshape(N, -1), dim=1)
            x.grad *= (en / gn.clamp(min=EPSILON)).view(N, 1, 1, 1)
        x_before = x.detach().clone() if trace is not None else None
        opt_in = None
        if trace is not None:  # optimiser state BEFORE this step's update (teacher forcing at any step, tests only)
            def _st(p, k):
                s_ = optimizer.state.get(p, {})
                return s_[k].detach().clone() if k in s_ else torch.zeros_like(p)
            opt_in = {n_: dict(exp_avg=_st(p_, "exp_avg"), exp_avg_sq=_st(p_, "exp_avg_sq"))
                      for n_, p_ in (("x", x), ("scales", scales), ("shifts", shifts))}
            opt_in["scales_in"], opt_in["shifts_in"] = scales.detach().clone(), shifts.detach().clone()
        optimizer.step()
        with torch.no_grad():
            x_adam = x.detach().clone() if trace is not None else None
            x.data = self.scheduler.step(v, t, x).prev_sample
        if trace is not None:
            trace(dict(idx=idx, t=int(t), x_in=x_before, v=v.detach(), x0=x0.detach(), losses=losses.detach(),
                       grad=raw_grad, s_grad=None if scales.grad is None else scales.grad.detach().clone(),
                       t_grad=None if shifts.grad is None else shifts.grad.detach().clone(),
                       x_adam=x_adam, x_out=x.detach().clone(), scales=scales.detach().clone(),
                       shifts=shifts.detach().clone(), eps_norm=en, grad_norm=gn, opt_in=opt_in))
        return losses.detach()

    @torch.no_grad()
    def sample_closed_form(self, st, steps, max_steps=None):
        """train_latents=False (hence closed_form=True, marigold_dc.py:605-613): plain DDIM sampling (:805-809, :905-909)
        and the final decode with the closed-form least-squares scale / shift (:332-336, :970-984)."""
        x = st["x"]
        self.scheduler.set_timesteps(steps, device=self.device)
        for i, t in enumerate(self.scheduler.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            x = self.scheduler.step(self.predict_noise(st["img_latents"], x, t), t, x).prev_sample
        aff = self.latent_to_affine(x, st["orig_res"], st["padding"], st.get("interp_mode", "bilinear"))
        n = aff.shape[0]
        scales, shifts = compute_affine_params(aff, st["sparses_normed"], st["masks"])
        dense = (scales.view(n, 1, 1, 1) * aff + shifts.view(n, 1, 1, 1)).clamp(min=0.0, max=1.0)
        return dense * (st["max_depths"] - st["min_depths"]) + st["min_depths"], x

    def __call__(self, imgs, sparses, max_depth, min_depth=0.0, norm="minmax", pred_latents_prev=None, beta=0.9,
                 steps=50, resolution=768, lr=None, seed=2024, trace=None, max_steps=None, projection="linear",
                 inv=False, percentile=(0.01, 0.99), opt="adam", loss_funcs=None, kld=False, kld_weight=0.1,
                 kld_mode="simple", train_latents=True, closed_form=None, interp_mode="bilinear"):
        if imgs.ndim != 4 or sparses.ndim != 4 or imgs.shape[0] != sparses.shape[0] or imgs.shape[-2:] != sparses.shape[-2:]:
            raise ValueError("Shape of image must be [N, C, H, W] and shape of sparse must be [N, 1, H, W]")
        N = imgs.shape[0]
        lr_latent, lr_scaling = (0.05, 0.005) if lr is None else lr
        st = self.preprocess(imgs, sparses, max_depth, min_depth, norm, resolution, seed, pred_latents_prev, beta,
                             projection, inv, percentile)
        loss_kw = dict(loss_funcs=tuple(loss_funcs or ("l1", "l2")), kld=kld, kld_weight=kld_weight, kld_mode=kld_mode)
        st["interp_mode"] = interp_mode
        closed_form = (not train_latents) if closed_form is None else closed_form
        if not train_latents:
            return self.sample_closed_form(st, steps, max_steps)
        x = torch.nn.Parameter(st["x"])
        scales = torch.nn.Parameter(torch.ones(N, 1, 1, 1, device=self.device))
        shifts = torch.nn.Parameter(torch.zeros(N, 1, 1, 1, device=self.device))
        groups = [{"params": [x], "lr": lr_latent}]
        if not closed_form:  # marigold_dc.py:764-783: no affine parameters in closed-form mode
            groups.append({"params": [scales, shifts], "lr": lr_scaling})
        opt = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad}[opt](groups)  # :776-789
        self.scheduler.set_timesteps(steps, device=self.device)
        for i, t in enumerate(self.scheduler.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            self.guided_step(st, t, x, scales, shifts, opt, trace, i, loss_kw, closed_form)
        with torch.no_grad():
            xd = x.detach()
            aff = self.latent_to_affine(xd, st["orig_res"], st["padding"], interp_mode)
            if closed_form:
                cs, ct = compute_affine_params(aff, st["sparses_normed"], st["masks"])
                dense = (cs.view(N, 1, 1, 1) * aff + ct.view(N, 1, 1, 1)).clamp(min=0.0, max=1.0)
            else:
                dense = self.affine_to_metric(aff, st["sparses_normed"], st["masks"], scales, shifts).clamp(min=0.0, max=1.0)
            denses = dense * (st["max_depths"] - st["min_depths"]) + st["min_depths"]
        return denses, xd
