"""Dataset front end: the files either side of the hot path.

What `/root/reference/predict.py:512-773` does around the pipeline call, without the CLI: find dataset directories
(`utils.py:193-227`: a directory holding `image/` and `sparse/`), pair every image with `sparse/<same relative path>.png`
(`predict.py:547-575`), decode the sparse PNG's first channel as `max_sparse_depth * v / 255` metres with 0 = missing
(`utils.py:1137-1158`), run the pipeline batch by batch (`video.sequence_batches` cuts the batches, so `use_prev_latent` and
rank sharding behave as in `predict.py:599-700`; every batch is loaded and moved to the GPU on its own), skip maps containing NaN (`predict.py:712-714`) and store each dense map as
`<dst>/<dataset>/dense/<relative path>.npy|npz|bl2` (`predict.py:717-728`, `utils.py:592-690`).  `compress="bl2"` goes
through the `blosc2` package exactly like the reference (`blosc2.save_array` / `load_array`); where that package is not
installed (this build image) it raises instead of silently writing another format.  The visualisation panels
(`predict.py:731-764`) are `vis.py`; the command line around all of this is `predict.py` / `analyze.py` of this package.
Segmentation masks are loaded by the reference CLI but never reach the pipeline call (`predict.py:659-694`), so they have no
counterpart here.
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


def load_many(paths, loader=load_rgb, num_threads: int = 1) -> list:
    """`loader` over `paths`, in order, on up to `num_threads` threads (utils.py:817-970 loads a batch's files in parallel:
    PIL and numpy release the GIL while decoding / decompressing).  Failed loads stay None."""
    paths = list(paths)
    if num_threads <= 1 or len(paths) <= 1:
        return [loader(p) for p in paths]
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(num_threads, len(paths))) as ex:
        return list(ex.map(loader, paths))


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
    if compress == "bl2" and not have_blosc2():
        raise RuntimeError("compress='bl2' needs the blosc2 package, which is not installed here; use 'npz' or 'npy'")
    path.parent.mkdir(parents=True, exist_ok=True)
    if torch.is_floating_point(x) and x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    arr = x.detach().cpu().numpy()
    if compress == "bl2":
        import blosc2

        blosc2.save_array(arr, str(path), mode="w")
    elif compress == "npz":
        np.savez_compressed(path, arr)
    else:
        np.save(path, arr)


def have_blosc2() -> bool:
    try:
        import blosc2  # noqa: F401
    except ImportError:
        return False
    return True


def find_dense(stem: Path) -> Path | None:
    """`stem` with the first of .npy / .npz / .bl2 that exists (utils.py:1189-1218 with NPARRAY_EXTS)."""
    for ext in (".npy", ".npz", ".bl2"):
        if Path(stem).with_suffix(ext).exists():
            return Path(stem).with_suffix(ext)
    return None


def load_dense(path: Path) -> np.ndarray:
    path = Path(path)
    if path.suffix == ".bl2":
        if not have_blosc2():
            raise RuntimeError(f"{path}: reading .bl2 needs the blosc2 package, which is not installed here")
        import blosc2

        return blosc2.load_array(str(path))
    if path.suffix == ".npz":
        with np.load(path) as z:
            return z[z.files[0]]
    return np.load(path)


def complete_dataset(pipe, src_root, dst_root, max_depth: float = 120.0, max_sparse_depth: float = 120.0, *,
                     batch_size: int = 1, use_prev_latent: bool = False, beta: float = 0.9, compress: str | None = "npz",
                     device=None, rank: int = 0, world: int = 1, save_dense: bool = True, on_frame=None, on_batch=None,
                     overlap_prologue: bool = False, **pipe_kwargs) -> dict:
    """Runs `pipe` over every dataset directory under `src_root`; returns {dataset name: [saved dense paths]}.

    Like `predict.py:599-690`, one BATCH at a time is decoded, moved to the device, completed and written, so host and
    device memory hold `batch_size` frames, never a whole dataset; frames of different resolutions may share a dataset
    (a batch is cut where the resolution changes; every geometry keeps its engine resident on the shared weight bank).
    With `use_prev_latent` the frames chain (batch 1, `predict.py:423-430, :697-699`) and the chain restarts where the
    resolution changes.  With world > 1 the independent frames are sharded across ranks and every rank writes its own
    files (no collective at all).  Unreadable pairs are skipped like `predict.py:636-655`.

    Hooks for the command line (`predict.py` of this package): `save_dense=False` skips the files (`--save-dense`),
    `on_frame(dataset_dir, img_path, sparse_path, img, sparse, dense)` is called per completed frame with device tensors
    (visualisation, `predict.py:731-764`), `on_batch(dataset_dir, n_frames, seconds)` per pipeline call (progress bar).
    `overlap_prologue=True` decodes, uploads and VAE-encodes the NEXT batch on a side stream while the current one is in
    its guided loop (`video._complete_overlapped` is the in-memory form of the same loop)."""
    import time

    from .video import sequence_batches

    src_root, dst_root = Path(src_root), Path(dst_root)
    datasets = find_dataset_dirs(src_root)
    if not datasets:
        raise FileNotFoundError(f"No dataset directories found at {src_root}")
    device = device if device is not None else getattr(pipe, "device", "cpu")
    suffix = _SAVE_SUFFIX[compress]
    geo = {k: pipe_kwargs[k] for k in ("steps", "resolution") if k in pipe_kwargs}
    overlap = bool(overlap_prologue) and torch.device(device).type == "cuda"
    side = torch.cuda.Stream(device=device) if overlap else None

    def runs_of(pairs):
        """Batches of decoded files, cut where the resolution changes (one pipeline call per run of equal shapes)."""
        for b0, b1 in sequence_batches(len(pairs), batch_size, use_prev_latent, rank, world):
            chunk = pairs[b0:b1]  # like predict.py:612-625: one thread per file of the batch
            files = load_many([i for i, _ in chunk] + [sp for _, sp in chunk], num_threads=2 * len(chunk))
            loaded = [(files[k], files[len(chunk) + k], i, sp) for k, (i, sp) in enumerate(chunk)]
            loaded = [t for t in loaded if t[0] is not None and t[1] is not None]
            start = 0
            for k in range(1, len(loaded) + 1):
                if k == len(loaded) or loaded[k][0].shape != loaded[start][0].shape:
                    yield loaded[start:k]
                    start = k

    def stage(run):
        """Host -> device (and, when overlapping, the VAE encoder on the side stream) for one run."""
        if run is None:
            return None
        if not overlap:
            imgs = torch.stack([t[0] for t in run]).to(device)
            sparses = to_depth(torch.stack([t[1] for t in run]).to(device), max_distance=max_sparse_depth)
            return run, imgs, sparses, None, None
        with torch.cuda.stream(side):
            imgs = torch.stack([t[0] for t in run]).pin_memory().to(device, non_blocking=True)
            sparses = to_depth(torch.stack([t[1] for t in run]).pin_memory().to(device, non_blocking=True),
                               max_distance=max_sparse_depth)
            lat = pipe.encode_ahead(imgs, **geo)
            done = torch.cuda.Event()
            done.record(side)
        return run, imgs, sparses, lat, done

    saved = {}
    for ds in datasets:
        pairs = find_pairs(ds)
        if not pairs:
            raise FileNotFoundError(f"No valid input pairs found in {ds}")
        out, prev, prev_shape = [], None, None
        img_dir, sparse_dir = ds / IMAGE_DIR, ds / SPARSE_DIR
        rel_ds = ds.relative_to(src_root) if ds != src_root else Path(".")
        it = runs_of(pairs)
        nxt = stage(next(it, None))
        while nxt is not None:
            run, imgs, sparses, img_lat, done = nxt
            t0 = time.time()
            if use_prev_latent and prev is not None and prev_shape != tuple(imgs.shape[-2:]):
                prev = None
            kw = dict(pred_latents_prev=prev if use_prev_latent else None, beta=beta, **pipe_kwargs)
            if overlap:
                cur = torch.cuda.current_stream(device)
                cur.wait_event(done)
                ticket = pipe.submit(imgs, sparses, max_depth, _img_latents=img_lat, **kw)
                nxt = stage(next(it, None))  # decoded, uploaded and encoded while this call's guided steps run
                denses, lat = pipe.collect(ticket)
                for t in (imgs, sparses, img_lat):
                    t.record_stream(cur)
            else:
                denses, lat = pipe(imgs, sparses, max_depth, **kw)
                nxt = stage(next(it, None))
            if use_prev_latent:
                prev, prev_shape = lat, tuple(imgs.shape[-2:])
            if on_batch is not None:
                on_batch(ds, len(run), time.time() - t0)
            for (_, _, ip, sp), img, sparse, dense in zip(run, imgs, sparses, denses):
                if torch.isnan(dense).any():
                    continue
                if save_dense:
                    path = (dst_root / rel_ds / DENSE_DIR / sp.relative_to(sparse_dir)).with_suffix(suffix)
                    save_tensor(dense, path, compress=compress)
                    out.append(path)
                if on_frame is not None:
                    on_frame(ds, ip.relative_to(img_dir), sp.relative_to(sparse_dir), img, sparse, dense)
        saved[ds.name] = out
    return saved
