"""DDIM tables the engine needs (host side): alphas_cumprod and the trailing timestep grid.

Mirrors the scheduler the reference installs at predict.py:491-494 (DDIMScheduler, scaled_linear betas,
v_prediction, timestep_spacing="trailing", set_alpha_to_one=False); the per-step algebra itself
(marigold_dc.py:813-826, :901-904) runs in csrc/tail.cuh.
"""
from __future__ import annotations

import numpy as np
import torch


def alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def trailing_timesteps(num_inference_steps: int, num_train_timesteps=1000) -> np.ndarray:
    ratio = num_train_timesteps / num_inference_steps
    return (np.round(np.arange(num_train_timesteps, 0, -ratio)).astype(np.int64) - 1).astype(np.int32)


def tables_from_scheduler(scheduler, steps: int):
    """Use a caller-provided (diffusers-like) scheduler when given, our own tables otherwise.

    The engine implements exactly the scheduler the reference installs for the guided path (predict.py:491-494):
    deterministic DDIM (eta = 0) with v-prediction, `alpha_prev = alphas_cumprod[t - 1000 // steps]` and
    `final_alpha_cumprod = alphas_cumprod[0]` (set_alpha_to_one=False).  Anything else -- the LCMScheduler arm of
    predict.py:495-498, epsilon / sample prediction, set_alpha_to_one=True, clipping / thresholding -- would silently be
    computed as that default case, so it is refused here."""
    if scheduler is not None and hasattr(scheduler, "alphas_cumprod") and hasattr(scheduler, "set_timesteps"):
        name = type(scheduler).__name__
        if "LCM" in name or hasattr(scheduler, "original_inference_steps"):
            raise NotImplementedError(f"{name}: only the DDIM scheduler of the guided path is implemented (predict.py:491-494); "
                                      "the LCM arm (predict.py:495-498) is outside the B200 hot path")
        cfg = getattr(scheduler, "config", None)
        get = (lambda k, d=None: (cfg.get(k, d) if isinstance(cfg, dict) else getattr(cfg, k, d))) if cfg is not None else (lambda k, d=None: d)
        pt = get("prediction_type", getattr(scheduler, "prediction_type", "v_prediction"))
        if pt not in (None, "v_prediction"):
            raise NotImplementedError(f"prediction_type={pt!r}: the engine implements v_prediction (marigold_dc.py:816-818)")
        if get("clip_sample", False) or get("thresholding", False):
            raise NotImplementedError("clip_sample / thresholding schedulers are not implemented")
        ac = torch.as_tensor(scheduler.alphas_cumprod).float().cpu()
        if ac.ndim != 1 or ac.shape[0] != 1000:
            raise ValueError(f"expected 1000 alphas_cumprod (num_train_timesteps), got shape {tuple(ac.shape)}")
        fa = getattr(scheduler, "final_alpha_cumprod", None)
        if fa is not None and abs(float(fa) - float(ac[0])) > 1e-6:
            raise NotImplementedError(f"final_alpha_cumprod={float(fa):.6f} (set_alpha_to_one=True?): the engine uses "
                                      f"alphas_cumprod[0]={float(ac[0]):.6f} for the last step (set_alpha_to_one=False)")
        scheduler.set_timesteps(steps)
        ts = np.asarray(torch.as_tensor(scheduler.timesteps).cpu().numpy(), dtype=np.int32)
        if ts.shape[0] != steps:
            raise ValueError(f"scheduler produced {ts.shape[0]} timesteps for {steps} steps")
        return ac, ts
    return alphas_cumprod(), trailing_timesteps(steps)
