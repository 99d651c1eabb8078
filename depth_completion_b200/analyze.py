"""`python -m depth_completion_b200.analyze DATASET_ROOT RESULT_ROOT [options]`: the command line of
`/root/reference/analyze.py` (MAE / RMSE of the dense maps `predict` wrote against the sparse measurements, overall and
per depth bin, one `results.json` per dataset and a `results_all.json` pooled over all datasets).

Options, defaults, fall-backs and the JSON layout follow `analyze.py:16-127` and `:300-361`; the scoring itself is
`metrics.evaluate_dataset` (the loop of `analyze.py:225-300`).  A dataset's score is the mean of its per-batch scores and
the pooled score is the mean over the batches of all datasets, exactly as the reference averages them.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import click
import torch

from . import dataset_io as dio
from . import metrics as mt
from .cli_common import LOG_LEVELS, CommaSeparated, get_logger


def _report(name, overall, binned, n_binned, n_pts, ranges, metrics, min_depth, max_depth, calc_binned, log) -> dict:
    """One results dictionary in the reference's layout (analyze.py:300-330), logged the way it logs it."""
    log.info(f"[{name}]:")
    log.info(f"  {min_depth:.1f} <= x <= {max_depth:.1f}:")
    res = {"overall": {}}
    for m in metrics:
        res["overall"][m] = float(torch.stack(overall[m]).mean())
        log.info(f"    {m}: {res['overall'][m]:.2f}")
    if calc_binned:
        res["binned"] = []
        for b, (lo, hi) in enumerate(ranges):
            pct = float(n_binned[b] / n_pts) * 100 if n_pts else float("nan")
            entry = {"range": (lo, hi), "metrics": {}, "percentage": pct}
            log.info(f"  {lo:.1f} <= x <= {hi:.1f} ({pct:.1f}%):")
            for m in metrics:
                # an empty bin is NaN, like torch.stack([]).mean() would be if the reference guarded it
                entry["metrics"][m] = float(torch.stack(binned[b][m]).mean()) if binned[b][m] else float("nan")
                log.info(f"    {m}: {entry['metrics'][m]:.2f}")
            res["binned"].append(entry)
    return res


def run(dataset_root: Path, result_root: Path, metrics, calc_binned_scores=True, bin_size=10.0, max_sparse_depth=120.0,
        max_depth=120.0, min_depth=0.0, batch_size=32, device="cpu", log=None, num_threads: int = 1) -> dict:
    """The body of analyze.py:153-361 without click; returns what goes into `results_all.json`."""
    log = log or get_logger()
    dataset_root, result_root = Path(dataset_root), Path(result_root)
    datasets = dio.find_dataset_dirs(dataset_root)
    if not datasets:
        raise FileNotFoundError("No dataset directories found")
    log.info(f"Found {len(datasets):,} datasets")
    ranges = mt.calc_bins(min_depth, max_depth, bin_size)
    all_overall = {m: [] for m in metrics}
    all_binned = [{m: [] for m in metrics} for _ in ranges]
    all_pts, all_nb = 0, [0] * len(ranges)
    for ds in datasets:
        result_dir = result_root / (ds.relative_to(dataset_root) if ds != dataset_root else Path("."))
        if not result_dir.exists():
            log.warning(f"No result directory found for {ds.name}. Skip this dataset")
            continue
        try:
            r = mt.evaluate_dataset(ds, result_dir, max_sparse_depth, min_depth, max_depth, metrics,
                                    bin_size if calc_binned_scores else None, batch_size, device, keep_batches=True,
                                    num_threads=num_threads)
        except FileNotFoundError:
            log.warning(f"No dense & sparse depth map pairs found for {ds.name}. Skip this dataset")
            continue
        bt = r["batches"]
        binned = bt["binned"] if calc_binned_scores else [{m: [] for m in metrics} for _ in ranges]
        n_binned = bt["num_binned"] if calc_binned_scores else [0] * len(ranges)
        res = _report(ds.name, bt["overall"], binned, n_binned, bt["num_points"], ranges, metrics, min_depth, max_depth,
                      calc_binned_scores, log)
        with (result_dir / "results.json").open("w") as f:
            json.dump(res, f, indent=2)
        log.success(f"Saved results to {result_dir / 'results.json'}")
        for m in metrics:
            all_overall[m] += bt["overall"][m]
            for b in range(len(ranges)):
                all_binned[b][m] += binned[b][m]
        all_pts += bt["num_points"]
        all_nb = [a + b for a, b in zip(all_nb, n_binned)]
    if not any(all_overall.values()):
        raise FileNotFoundError(f"No dataset under {dataset_root} has results under {result_root}")
    res_all = _report("All", all_overall, all_binned, all_nb, all_pts, ranges, metrics, min_depth, max_depth,
                      calc_binned_scores, log)
    res_all.setdefault("binned", [])   # analyze.py:319 always writes the key
    with (result_root / "results_all.json").open("w") as f:
        json.dump(res_all, f, indent=2)
    log.success(f"Saved results for all datasets to {result_root / 'results_all.json'}")
    return res_all


@click.command(help="Analyze results of depth completion.")
@click.argument("dataset_root", type=click.Path(exists=True, path_type=Path, file_okay=False, dir_okay=True))
@click.argument("result_root", type=click.Path(exists=True, path_type=Path, file_okay=False, dir_okay=True))
@click.option("--log", type=click.Path(path_type=Path), default=None, help="Path to save logs.")
@click.option("--log-level", type=click.Choice(LOG_LEVELS), default="INFO", show_default=True)
@click.option("--metrics", type=CommaSeparated(str), default="mae,rmse", show_default=True,
              help="Comma-separated list of metrics to compute. Available options: mae, rmse")
@click.option("--calc-binned-scores", type=bool, default=True, show_default=True, help="Whether to compute binned scores.")
@click.option("--bin-size", type=click.FloatRange(min=0, min_open=True), default=10.0, show_default=True,
              help="Bin size in meters.")
@click.option("--max-sparse-depth", type=click.FloatRange(min=0, min_open=True), default=120.0, show_default=True)
@click.option("--max-depth", type=click.FloatRange(min=0, min_open=True), default=120.0, show_default=True)
@click.option("--min-depth", type=click.FloatRange(min=0), default=0.0, show_default=True)
@click.option("-bs", "--batch-size", type=click.IntRange(min=1), default=32, show_default=True)
@click.option("-nt", "--num-threads", type=click.IntRange(min=1), default=8, show_default=True,
              help="Number of threads for loading sparse & dense depth maps.")
@click.option("--cuda", type=bool, default=True, show_default=True, help="Whether to use CUDA for faster processing.")
def main(dataset_root, result_root, log, log_level, metrics, calc_binned_scores, bin_size, max_sparse_depth, max_depth,
         min_depth, batch_size, num_threads, cuda) -> None:
    logger = get_logger(log_level, log)
    if log is not None:
        logger.info(f"Saving logs to {log}")
    if cuda and not torch.cuda.is_available():
        logger.warning("CUDA is not available. Using CPU instead.")
        cuda = False
    kept = []
    for m in metrics:
        if m not in mt.METRICS:
            logger.error(f"Invalid metric: {m} (skipped)")
        else:
            kept.append(m)
    if not kept:
        logger.critical("No valid metrics provided")
        sys.exit(1)
    try:
        run(dataset_root, result_root, kept, calc_binned_scores, bin_size, max_sparse_depth, max_depth, min_depth, batch_size,
            "cuda" if cuda else "cpu", logger, num_threads)
    except FileNotFoundError as e:
        logger.critical(str(e))
        sys.exit(1)


if __name__ == "__main__":
    main()
