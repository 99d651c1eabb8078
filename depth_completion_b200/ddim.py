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
    """Use a caller-provided (diffusers-like) scheduler when given, our own tables otherwise."""
    if scheduler is not None and hasattr(scheduler, "alphas_cumprod") and hasattr(scheduler, "set_timesteps"):
        scheduler.set_timesteps(steps)
        ts = np.asarray(torch.as_tensor(scheduler.timesteps).cpu().numpy(), dtype=np.int32)
        ac = torch.as_tensor(scheduler.alphas_cumprod).float().cpu()
        return ac, ts
    return alphas_cumprod(), trailing_timesteps(steps)
