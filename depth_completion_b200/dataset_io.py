"""Dataset front end: the files either side of the hot path.

What `/root/reference/predict.py:512-773` does around the pipeline call, without the CLI: find dataset directories
(`utils.py:193-227`: a directory holding `image/` and `sparse/`), pair every image with `sparse/<same relative path>.png`
(`predict.py:547-575`), decode the sparse PNG's first channel as `max_sparse_depth * v / 255` metres with 0 = missing
(`utils.py:1137-1158`), run the pipeline batch by batch (`video.sequence_batches` cuts the batches, so `use_prev_latent` and
rank sharding behave as in `predict.py:599-700`; every batch is loaded and moved to the GPU on its own), skip maps containing NaN (`predict.py:712-714`) and store each dense map as
`<dst>/<dataset>/dense/<relative path>.npy|npz` (`predict.py:717-728`, `utils.py:592-690`).  `compress="bl2"` needs
blosc2, which this image does not have: it raises instead of silently writing another format.  Visualisation
(`predict.py:731-764`) and segmentation masks are outside the hot path's data formats and are not built.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch


IMAGE_DIR, SPARSE_DIR, DENSE_DIR = "image", "sparse", "dense"  # utils.py:20-23
_SAVE_SUFFIX = {None: ".npy", "npy": ".npy", "npz": ".npz", "bl2": ".bl2"}


def is_dataset_dir(path: Path) -> bool:
    path = Path(path)
    return path.is_dir() and (path / IMAGE_DIR).is_dir() and (path / SPARSE_DIR).is_dir()


def find_dataset_dirs(root: Path) -> list[Path]:
    """The root itself when it is a dataset directory, else every dataset directory below it (utils.py:210-227)."""
    root = Path(root)
    if is_dataset_dir(root):
        return [root]
    return sorted(p for p in root.rglob("*") if is_dataset_dir(p))


def _is_image(path: Path) -> bool:
    if not path.is_file():
        return False
    try:
        from PIL import Image

        with Image.open(path) as im:
            return im.size[0] > 0 and im.size[1] > 0
    except Exception:
        return False


def find_pairs(dataset_dir: Path) -> list[tuple[Path, Path]]:
    """(image, sparse) paths sorted by image file name; images without `sparse/<rel>.png` are dropped
    (predict.py:547-575)."""
    img_dir, sparse_dir = Path(dataset_dir) / IMAGE_DIR, Path(dataset_dir) / SPARSE_DIR
    pairs = []
    for img in sorted((p for p in img_dir.rglob("*") if _is_image(p)), key=lambda p: p.name):
        sp = sparse_dir / img.relative_to(img_dir).with_suffix(".png")
        if sp.exists():
            pairs.append((img, sp))
    return pairs


def load_rgb(path: Path) -> torch.Tensor | None:
    """[3, H, W] uint8 (mode "RGB", as predict.py:613-625 asks for), or None when the file cannot be decoded."""
    try:
        from PIL import Image

        with Image.open(path) as im:
            arr = np.asarray(im.convert("RGB"), dtype=np.uint8)
    except Exception:
        return None
    return torch.from_numpy(arr.copy()).permute(2, 0, 1).contiguous()


def to_depth(imgs: torch.Tensor, max_distance: float = 120.0, dtype=torch.float32) -> torch.Tensor:
    """[N, 3, H, W] depth-coded images -> [N, 1, H, W] metres: max_distance * channel0 / 255 (utils.py:1137-1158)."""
    return max_distance * (imgs.to(dtype)[:, 0] / 255.0).unsqueeze(1)


def encode_depth_png(depth: torch.Tensor, max_distance: float = 120.0) -> np.ndarray:
    """Inverse of `to_depth` for writing fixtures: [H, W] metres -> [H, W, 3] uint8 (256 levels)."""
    q = torch.round(depth.clamp(0, max_distance) / max_distance * 255.0).to(torch.uint8)
    return q.unsqueeze(-1).repeat(1, 1, 3).cpu().numpy()


def save_tensor(x: torch.Tensor, path: Path, compress: str | None = None) -> None:
    """utils.py:592-690: .npy (None / "npy"), .npz ("npz", numpy's compressed container) or .bl2 (blosc2)."""
    path = Path(path)
    if compress not in _SAVE_SUFFIX:
        raise ValueError(f"Unknown compression: {compress}")
    if path.suffix != _SAVE_SUFFIX[compress]:
        raise ValueError(f"Invalid extension: {path.suffix} (must be {_SAVE_SUFFIX[compress]})")
    if compress == "bl2":
        raise RuntimeError("compress='bl2' needs the blosc2 package, which is not installed here; use 'npz' or 'npy'")
    path.parent.mkdir(parents=True, exist_ok=True)
    if torch.is_floating_point(x) and x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    arr = x.detach().cpu().numpy()
    if compress == "npz":
        np.savez_compressed(path, arr)
    else:
        np.save(path, arr)


def load_dense(path: Path) -> np.ndarray:
    path = Path(path)
    if path.suffix == ".npz":
        with np.load(path) as z:
            return z[z.files[0]]
    return np.load(path)


def complete_dataset(pipe, src_root, dst_root, max_depth: float = 120.0, max_sparse_depth: float = 120.0, *,
                     batch_size: int = 1, use_prev_latent: bool = False, beta: float = 0.9, compress: str | None = "npz",
                     device=None, rank: int = 0, world: int = 1, **pipe_kwargs) -> dict:
    """Runs `pipe` over every dataset directory under `src_root`; returns {dataset name: [saved dense paths]}.

    Like `predict.py:599-690`, one BATCH at a time is decoded, moved to the device, completed and written, so host and
    device memory hold `batch_size` frames, never a whole dataset; frames of different resolutions may share a dataset
    (a batch is cut where the resolution changes; every geometry keeps its engine resident on the shared weight bank).
    With `use_prev_latent` the frames chain (batch 1, `predict.py:423-430, :697-699`) and the chain restarts where the
    resolution changes.  With world > 1 the independent frames are sharded across ranks and every rank writes its own
    files (no collective at all).  Unreadable pairs are skipped like `predict.py:636-655`."""
    from .video import sequence_batches

    src_root, dst_root = Path(src_root), Path(dst_root)
    datasets = find_dataset_dirs(src_root)
    if not datasets:
        raise FileNotFoundError(f"No dataset directories found at {src_root}")
    device = device if device is not None else getattr(pipe, "device", "cpu")
    suffix = _SAVE_SUFFIX[compress]
    saved = {}
    for ds in datasets:
        pairs = find_pairs(ds)
        if not pairs:
            raise FileNotFoundError(f"No valid input pairs found in {ds}")
        out, prev, prev_shape = [], None, None
        sparse_dir = ds / SPARSE_DIR
        rel_ds = ds.relative_to(src_root) if ds != src_root else Path(".")
        for b0, b1 in sequence_batches(len(pairs), batch_size, use_prev_latent, rank, world):
            loaded = [(load_rgb(i), load_rgb(s), s) for i, s in pairs[b0:b1]]
            loaded = [t for t in loaded if t[0] is not None and t[1] is not None]
            # cut the batch where the resolution changes (one pipeline call per run of equal shapes)
            runs, start = [], 0
            for k in range(1, len(loaded) + 1):
                if k == len(loaded) or loaded[k][0].shape != loaded[start][0].shape:
                    runs.append(loaded[start:k])
                    start = k
            for run in runs:
                if not run:
                    continue
                imgs = torch.stack([t[0] for t in run]).to(device)
                sparses = to_depth(torch.stack([t[1] for t in run]).to(device), max_distance=max_sparse_depth)
                if use_prev_latent and prev is not None and prev_shape != tuple(imgs.shape[-2:]):
                    prev = None
                denses, lat = pipe(imgs, sparses, max_depth, pred_latents_prev=prev if use_prev_latent else None, beta=beta,
                                   **pipe_kwargs)
                if use_prev_latent:
                    prev, prev_shape = lat, tuple(imgs.shape[-2:])
                for (_, _, sp), dense in zip(run, denses):
                    if torch.isnan(dense).any():
                        continue
                    path = (dst_root / rel_ds / DENSE_DIR / sp.relative_to(sparse_dir)).with_suffix(suffix)
                    save_tensor(dense, path, compress=compress)
                    out.append(path)
        saved[ds.name] = out
    return saved
