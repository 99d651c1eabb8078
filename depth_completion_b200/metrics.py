"""MAE / RMSE of completed depth maps against sparse measurements: the accuracy half of BASELINE.json's metric.

`mae` / `rmse` follow `/root/reference/utils.py:692-739` (mean over the masked elements), `calc_bins` `utils.py:162-190`,
and `evaluate_dataset` the scoring loop of `analyze.py:225-300`: dense maps written by `dataset_io.complete_dataset`
are compared with the depth-coded sparse PNGs of the same dataset directory, both clamped to [min_depth, max_depth],
batch by batch; a dataset's score is the mean of its per-batch scores (as the reference averages them), overall and per
depth bin.  Host / torch code on purpose: it runs once per dataset, after the hot path.
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import dataset_io as dio


def mae(preds: torch.Tensor, targets: torch.Tensor, masks: torch.Tensor | None = None) -> torch.Tensor:
    if masks is not None:
        preds, targets = preds[masks], targets[masks]
    return (preds - targets).abs().mean()


def rmse(preds: torch.Tensor, targets: torch.Tensor, masks: torch.Tensor | None = None) -> torch.Tensor:
    if masks is not None:
        preds, targets = preds[masks], targets[masks]
    return (preds - targets).square().mean().sqrt()


METRICS = {"mae": mae, "rmse": rmse}


def calc_bins(lower: float, upper: float, size: float) -> list[tuple[float, float]]:
    """[(lower, lower + size), ...] up to `upper`, the last bin cut at `upper`."""
    if lower >= upper:
        raise ValueError(f"Lower bound {lower} must be less than upper bound {upper}")
    bins = []
    while lower < upper:
        bins.append((lower, min(lower + size, upper)))
        lower += size
    return bins


def evaluate_dataset(dataset_dir, result_dir, max_sparse_depth: float = 120.0, min_depth: float = 0.0,
                     max_depth: float = 120.0, metrics=("mae", "rmse"), bin_size: float | None = None, batch_size: int = 1,
                     device="cpu", keep_batches: bool = False, num_threads: int = 1) -> dict:
    """{"overall": {metric: score}, "num_points": n, "bins": [{"range": (lo, hi), metric: score, "num_points": n}, ...]}.

    `result_dir` is the directory holding `dense/` (what `complete_dataset` wrote for this dataset directory).
    keep_batches=True adds "batches": the per-batch score tensors and point counts, overall and per bin, which is what
    the reference pools over ALL datasets for `results_all.json` (analyze.py:163-170, :322-348)."""
    for m in metrics:
        if m not in METRICS:
            raise ValueError(f"Unknown metric: {m}")
    dataset_dir, result_dir = Path(dataset_dir), Path(result_dir)
    sparse_dir, dense_dir = dataset_dir / dio.SPARSE_DIR, result_dir / dio.DENSE_DIR
    pairs = []
    for _, sp in dio.find_pairs(dataset_dir):
        found = dio.find_dense(dense_dir / sp.relative_to(sparse_dir))
        if found is not None:
            pairs.append((sp, found))
    if not pairs:
        raise FileNotFoundError(f"No dense maps found under {dense_dir} for the sparse maps of {dataset_dir}")
    bins = calc_bins(min_depth, max_depth, bin_size) if bin_size else []
    overall = {m: [] for m in metrics}
    binned = [{m: [] for m in metrics} for _ in bins]
    n_pts, n_binned = 0, [0] * len(bins)
    for i in range(0, len(pairs), batch_size):
        chunk = pairs[i:i + batch_size]
        sparses = dio.to_depth(torch.stack(dio.load_many([sp for sp, _ in chunk], num_threads=num_threads)),
                               max_distance=max_sparse_depth).to(device)
        denses = torch.stack([torch.from_numpy(a).reshape(1, *sparses.shape[-2:])
                              for a in dio.load_many([d for _, d in chunk], dio.load_dense, num_threads)]).to(device)
        mask = sparses > 0
        if not mask.any():
            continue
        sparses, denses = sparses.clamp(min_depth, max_depth), denses.float().clamp(min_depth, max_depth)
        for m in metrics:
            overall[m].append(METRICS[m](denses, sparses, mask))
        n_pts += int(mask.sum())
        for b, (lo, hi) in enumerate(bins):
            mb = mask & (sparses >= lo) & (sparses <= hi)
            if mb.any():
                for m in metrics:
                    binned[b][m].append(METRICS[m](denses, sparses, mb))
                n_binned[b] += int(mb.sum())
    out = {"overall": {m: float(torch.stack(v).mean()) for m, v in overall.items() if v}, "num_points": n_pts, "bins": []}
    for b, rng in enumerate(bins):
        if n_binned[b]:
            out["bins"].append({"range": rng, "num_points": n_binned[b],
                                **{m: float(torch.stack(v).mean()) for m, v in binned[b].items()}})
    if keep_batches:
        out["batches"] = {"overall": overall, "binned": binned, "num_points": n_pts, "num_binned": n_binned, "ranges": bins}
    return out
