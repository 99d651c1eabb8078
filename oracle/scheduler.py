"""ORACLE (test infrastructure): DDIMScheduler of diffusers==0.31.0 as configured by the reference
(prs-eth/marigold-v1-0 scheduler config + predict.py:491-494): scaled_linear betas, v_prediction,
eta = 0, set_alpha_to_one=False, timestep_spacing="trailing".  SURVEY.md Appendix A.3.  PARITY UNPINNED
(diffusers absent); pinned by the known-answer values in tests/test_oracle_formulas.py.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
        self.num_train_timesteps = num_train_timesteps
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)  # CPU fp32, like the reference's
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.num_train_timesteps / num_inference_steps
        ts = np.round(np.arange(self.num_train_timesteps, 0, -step_ratio)).astype(np.int64) - 1  # trailing
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample):
        """eta = 0 DDIM update for v-prediction; scalars are 0-dim fp32 tensors so bf16 samples stay bf16."""
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        pred_original_sample = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        pred_epsilon = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        pred_sample_direction = (1 - a_prev) ** 0.5 * pred_epsilon
        prev_sample = a_prev ** 0.5 * pred_original_sample + pred_sample_direction
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=pred_original_sample)
