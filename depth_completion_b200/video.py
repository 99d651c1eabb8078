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
                      use_prev_latent: bool = False, beta: float = 0.9, rank: int = 0, world: int = 1,
                      frames_in_flight: int = 1, overlap_prologue: bool = False, **pipe_kwargs):
    """Runs `pipe` over imgs [F,C,H,W] / sparses [F,1,H,W]; returns (denses of this rank's frames [f,1,H,W] fp32,
    their (start, stop) frame range, last pred_latents).

    frames_in_flight > 1 (independent frames only) keeps that many pipeline calls in flight on separate CUDA streams,
    each on its own engine (shared weights): call k+1's prologue -- H2D copies, image preprocessing, VAE encoder -- and
    its guided steps overlap call k's loop, whose batch-1 UNet layers leave most SMs idle.  This is a throughput mode:
    a frame's latency grows, frames per second go up.

    overlap_prologue=True (one call in flight; works WITH use_prev_latent, whose chain only constrains the depth latent):
    call k+1's image prologue -- H2D copy, preprocessing, VAE encoder (marigold_dc.py:687-698) -- runs on a side stream on a
    second engine while call k is in its guided loop, and call k+1 starts from the encoded latents
    (`mdc_begin_frame_encoded`): SURVEY.md section 8(f)-3, "pipeline frame k+1's encoder under frame k's loop"."""
    if imgs.shape[0] != sparses.shape[0]:
        raise ValueError(f"{imgs.shape[0]} images vs {sparses.shape[0]} sparse maps")
    if use_prev_latent and batch_size > 1:
        warnings.warn("batch_size is forced to 1 when use_prev_latent=True (predict.py:423-430)")
    if frames_in_flight < 1:
        raise ValueError(f"frames_in_flight={frames_in_flight}")
    if frames_in_flight > 1 and use_prev_latent:
        raise ValueError("use_prev_latent chains every frame on the previous one (predict.py:697-699): nothing can be in flight "
                         "besides the current frame; use complete_sequences for several independent sequences")
    plan = sequence_batches(imgs.shape[0], batch_size, use_prev_latent, rank, world)
    outs, prev, lat = [], None, None
    if overlap_prologue and frames_in_flight > 1:
        raise ValueError("overlap_prologue is the one-call-in-flight mode; frames_in_flight > 1 already overlaps whole calls")
    if frames_in_flight > 1 and plan:
        return _complete_in_flight(pipe, imgs, sparses, max_depth, plan, frames_in_flight, pipe_kwargs)
    if overlap_prologue and plan:
        return _complete_overlapped(pipe, imgs, sparses, max_depth, plan, use_prev_latent, beta, pipe_kwargs)
    for s, e in plan:
        dense, lat = pipe(imgs[s:e], sparses[s:e], max_depth, pred_latents_prev=prev if use_prev_latent else None, beta=beta,
                          **pipe_kwargs)
        outs.append(dense)
        if use_prev_latent:
            prev = lat
    rng = (plan[0][0], plan[-1][1]) if plan else (0, 0)
    dense_all = torch.cat(outs, 0) if outs else torch.empty(0, 1, *imgs.shape[-2:])
    return dense_all, rng, (lat if plan else None)


def _complete_overlapped(pipe, imgs, sparses, max_depth, plan, use_prev_latent, beta, pipe_kwargs):
    """One call in flight, the NEXT call's image prologue under it: while the main engine replays call k's guided steps
    on the caller's stream, a side stream copies call k+1's images to the device and encodes them on a second engine
    (`pipe.encode_ahead`); call k+1 then begins from those latents.  The previous-latent chain (predict.py:697-699) is
    untouched: only the depth latent depends on call k's result, the image latents never do."""
    dev = pipe.device
    cur = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    geo = {k: pipe_kwargs[k] for k in ("steps", "resolution") if k in pipe_kwargs}

    def encode(i):
        s, e = plan[i]
        with torch.cuda.stream(side):
            lat = pipe.encode_ahead(imgs[s:e], **geo)
            done = torch.cuda.Event()
            done.record(side)
        return lat, done

    side.wait_stream(cur)  # inputs produced on the caller's stream
    nxt = encode(0)
    outs, prev, lat = [], None, None
    for i, (s, e) in enumerate(plan):
        img_lat, done = nxt
        cur.wait_event(done)
        ticket = pipe.submit(imgs[s:e], sparses[s:e], max_depth, pred_latents_prev=prev if use_prev_latent else None, beta=beta,
                             _img_latents=img_lat, **pipe_kwargs)
        if i + 1 < len(plan):
            nxt = encode(i + 1)  # enqueued while call i's guided steps are running
        dense, lat = pipe.collect(ticket)
        img_lat.record_stream(cur)
        outs.append(dense)
        if use_prev_latent:
            prev = lat
    return torch.cat(outs, 0), (plan[0][0], plan[-1][1]), lat


def _complete_in_flight(pipe, imgs, sparses, max_depth, plan, k, pipe_kwargs):
    """`k` calls in flight, round robin over `k` engine slots and CUDA streams (submit / collect halves of the call)."""
    dev = pipe.device
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(dev))
    tickets, outs, lat = {}, [None] * len(plan), None

    def collect(i):
        nonlocal lat
        with torch.cuda.stream(streams[i % k]):
            outs[i], lat = pipe.collect(tickets.pop(i))

    for i, (s, e) in enumerate(plan):
        if i - k in tickets:
            collect(i - k)  # frees slot i % k; the other slots keep the GPU busy meanwhile
        st = streams[i % k]
        st.wait_event(start)  # inputs produced on the caller's stream before this call
        with torch.cuda.stream(st):
            tickets[i] = pipe.submit(imgs[s:e], sparses[s:e], max_depth, _slot=i % k, _concurrent=True, **pipe_kwargs)
    for i in sorted(tickets):
        collect(i)
    cur = torch.cuda.current_stream(dev)
    for st in streams:
        cur.wait_stream(st)
    rng = (plan[0][0], plan[-1][1])
    return torch.cat(outs, 0), rng, lat


def complete_sequences(pipe, imgs: torch.Tensor, sparses: torch.Tensor, max_depth: float, *, beta: float = 0.9, rank: int = 0,
                       world: int = 1, seq_batch: int | None = None, **pipe_kwargs):
    """Temporal prior WITHOUT the reference's batch-1 restriction (predict.py:423-430): S independent sequences (cameras,
    clips) of F frames each, imgs [S,F,C,H,W] / sparses [S,F,1,H,W].  Frame k of `seq_batch` sequences goes through ONE
    batched call whose `pred_latents_prev` holds every sequence's own previous latent (marigold_dc.py:598-603, :699-704 blend
    per sample), so each sequence is the same serial chain the reference runs, and the batch dimension carries sequences
    instead of consecutive frames.  Sequences are sharded over the ranks (they are independent); returns (denses
    [s,F,1,H,W] of this rank's sequences, their (start, stop) sequence range)."""
    if imgs.ndim != 5 or sparses.ndim != 5 or imgs.shape[:2] != sparses.shape[:2]:
        raise ValueError(f"expected imgs [S,F,C,H,W] and sparses [S,F,1,H,W], got {tuple(imgs.shape)} and {tuple(sparses.shape)}")
    S, Fr = imgs.shape[:2]
    mine = shard_frames(S, rank, world)
    seq_batch = seq_batch or max(1, len(mine))
    outs = []
    for b0 in range(mine.start, mine.stop, seq_batch):
        b1 = min(b0 + seq_batch, mine.stop)
        prev, per_frame = None, []
        for f in range(Fr):
            dense, prev = pipe(imgs[b0:b1, f], sparses[b0:b1, f], max_depth, pred_latents_prev=prev, beta=beta, **pipe_kwargs)
            per_frame.append(dense)
        outs.append(torch.stack(per_frame, 1))
    dense_all = torch.cat(outs, 0) if outs else torch.empty(0, Fr, 1, *imgs.shape[-2:])
    return dense_all, (mine.start, mine.stop)
