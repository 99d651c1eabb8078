"""Sequence driver: the batch loop of the reference CLI over one image sequence (`/root/reference/predict.py:585-700`),
on top of the drop-in pipeline class -- SURVEY.md section 8(f)-3 (temporal prior + batched video driver).

* frames go through `pipe(...)` in batches of `batch_size` (`predict.py:599-603`);
* `use_prev_latent=True` feeds every call's `pred_latents` to the next call as `pred_latents_prev`
  (`predict.py:697-699`, blended with weight `beta` at `marigold_dc.py:699-704`); like the reference this forces
  `batch_size = 1` (`predict.py:423-430`) and makes the sequence a serial chain, so it is never sharded;
* independent frames (`use_prev_latent=False`) shard over ranks with `pipeline.shard_frames`.
"""
from __future__ import annotations

import warnings

import torch

from .pipeline import shard_frames


def sequence_batches(n_frames: int, batch_size: int, use_prev_latent: bool, rank: int = 0, world: int = 1):
    """[(start, stop), ...] frame ranges this rank runs, in order.  Pure host logic (unit-tested on the CPU)."""
    if n_frames < 0 or batch_size < 1:
        raise ValueError(f"n_frames={n_frames}, batch_size={batch_size}")
    if use_prev_latent:
        if world > 1:
            raise ValueError("use_prev_latent chains frames serially (predict.py:697-699): give each rank its own sequence "
                             "instead of sharding one")
        batch_size = 1
        frames = range(n_frames)
    else:
        frames = shard_frames(n_frames, rank, world)
    return [(s, min(s + batch_size, frames.stop)) for s in range(frames.start, frames.stop, batch_size)]


def complete_sequence(pipe, imgs: torch.Tensor, sparses: torch.Tensor, max_depth: float, *, batch_size: int = 1,
                      use_prev_latent: bool = False, beta: float = 0.9, rank: int = 0, world: int = 1, **pipe_kwargs):
    """Runs `pipe` over imgs [F,C,H,W] / sparses [F,1,H,W]; returns (denses of this rank's frames [f,1,H,W] fp32,
    their (start, stop) frame range, last pred_latents)."""
    if imgs.shape[0] != sparses.shape[0]:
        raise ValueError(f"{imgs.shape[0]} images vs {sparses.shape[0]} sparse maps")
    if use_prev_latent and batch_size > 1:
        warnings.warn("batch_size is forced to 1 when use_prev_latent=True (predict.py:423-430)")
    plan = sequence_batches(imgs.shape[0], batch_size, use_prev_latent, rank, world)
    outs, prev = [], None
    for s, e in plan:
        dense, lat = pipe(imgs[s:e], sparses[s:e], max_depth, pred_latents_prev=prev if use_prev_latent else None, beta=beta,
                          **pipe_kwargs)
        outs.append(dense)
        if use_prev_latent:
            prev = lat
    rng = (plan[0][0], plan[-1][1]) if plan else (0, 0)
    dense_all = torch.cat(outs, 0) if outs else torch.empty(0, 1, *imgs.shape[-2:])
    return dense_all, rng, (lat if plan else None)
