"""Visualisation of the inputs and outputs of a completed frame: what `/root/reference/predict.py:731-764` writes next to the
dense maps (`<dst>/<dataset>/vis/<image stem>_vis.jpg`, a row of image | sparse | dense panels).

* `visualize_depth` follows `/root/reference/utils.py:370-432`: clamp to [min_depth, max_depth], normalise to [0, 1], colour
  with the "Spectral" map.  The reference colours through diffusers' `MarigoldImageProcessor.colormap`; without matplotlib
  that routine interpolates linearly between the eleven ColorBrewer "Spectral" anchors, which is what is done here (on
  whatever device the depth map lives on, so nothing leaves the GPU before the grid is assembled).
* `make_grid` follows `utils.py:973-1066` (one row by default, optional resize where -1 keeps the aspect ratio) and
  `save_img_tensor` `utils.py:533-589`, both on torchvision like the reference.

Host / torch code on purpose: it runs once per frame, after the hot path.
"""
from __future__ import annotations

from pathlib import Path

import torch

# ColorBrewer 11-class "Spectral" (the anchors matplotlib's `Spectral` colormap is built from), 0.0 -> first, 1.0 -> last
_SPECTRAL_255 = ((158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
                 (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189), (94, 79, 162))
_INTERP = ("nearest", "bilinear", "bicubic", "lanczos")


def colormap_spectral(x01: torch.Tensor, as_bytes: bool = True) -> torch.Tensor:
    """[...] values in [0, 1] -> [..., 3] colours: piecewise-linear interpolation between the anchors."""
    anchors = torch.tensor(_SPECTRAL_255, dtype=torch.float32, device=x01.device) / 255.0
    k = anchors.shape[0]
    pos = x01.float().clamp(0.0, 1.0) * (k - 1)
    left = pos.long().clamp(max=k - 1)
    right = (left + 1).clamp(max=k - 1)
    frac = (pos - left.float()).unsqueeze(-1)
    out = (1.0 - frac) * anchors[left] + frac * anchors[right]
    return (out * 255.0).to(torch.uint8) if as_bytes else out


def visualize_depth(depth_maps: torch.Tensor, max_depth: float, min_depth: float = 0.0) -> torch.Tensor:
    """[N,1,H,W] metres -> [N,3,H,W] uint8 (utils.py:370-432, same ValueErrors)."""
    if min_depth >= max_depth:
        raise ValueError(f"Invalid values range: [{min_depth}, {max_depth}].")
    if depth_maps.ndim != 4 or depth_maps.shape[1] != 1:
        raise ValueError(f"Input depth maps must have shape [N,1,H,W], got {depth_maps.shape}")
    d = depth_maps.float().clamp(min=min_depth, max=max_depth)
    d = ((d - min_depth) / (max_depth - min_depth)).clamp(0.0, 1.0)
    return colormap_spectral(d[:, 0]).permute(0, 3, 1, 2).contiguous()


def make_grid(imgs, nrow: int | None = None, resize: tuple[int, int] | None = None, interpolation: str = "bilinear",
              antialias: bool = False) -> torch.Tensor:
    """[N,C,H,W] (or a list of [C,H,W]) -> [C, grid_h, grid_w]; one row unless `nrow` is given (utils.py:973-1066)."""
    import torchvision
    import torchvision.transforms.functional as TF

    if isinstance(imgs, list):
        if not imgs:
            raise ValueError("Empty list of images provided")
        if any(not isinstance(i, torch.Tensor) or i.dim() != 3 for i in imgs):
            raise ValueError("Each image in the list must be a 3D tensor (C,H,W)")
        imgs = torch.stack(imgs)
    if imgs.dim() != 4:
        raise ValueError("Images must be 4D tensor (N,C,H,W)")
    grid = torchvision.utils.make_grid(imgs, nrow=len(imgs) if nrow is None else nrow)
    if resize is not None and tuple(resize) != (-1, -1):
        th, tw = resize
        _, h, w = grid.shape
        target = [th if th != -1 else int(tw * h / w), tw if tw != -1 else int(th * w / h)]
        if interpolation.lower() not in _INTERP:
            raise ValueError(f"Unsupported interpolation mode: {interpolation}. Supported modes are: 'nearest', 'bilinear', "
                             "'bicubic', 'lanczos'.")
        mode = getattr(torchvision.transforms.InterpolationMode, interpolation.upper())
        grid = TF.resize(grid.unsqueeze(0), target, interpolation=mode, antialias=antialias).squeeze(0)
    return grid


def save_img_tensor(img: torch.Tensor, path: Path) -> None:
    """[C,H,W] uint8 (0-255) or float32 (0-1) -> image file, format from the extension (utils.py:533-589)."""
    import torchvision

    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    img = img.detach().cpu()
    if img.dtype == torch.uint8:
        img = img.float() / 255.0
    elif img.dtype == torch.float32:
        if img.max() > 1.0 or img.min() < 0.0:
            raise ValueError("Image tensor must be in the range [0, 1] if dtype is float32")
    else:
        raise ValueError(f"Unsupported image type: {img.dtype}")
    torchvision.utils.save_image(img, path)


def frame_panels(img: torch.Tensor, sparse: torch.Tensor, dense: torch.Tensor, order, max_depth: float, min_depth: float = 0.0):
    """The panels of one frame in `order` (predict.py:734-757): the image, the sparse map coloured with missing pixels
    black, the dense map coloured.  img [C,H,W] uint8, sparse / dense [1,H,W] metres."""
    panels = []
    for view in order:
        if view == "image":
            panels.append(img)
        elif view == "sparse":
            v = visualize_depth(sparse[None], max_depth=max_depth, min_depth=min_depth)[0]
            v[(sparse <= 0.0).repeat(img.shape[0], 1, 1)] = 0
            panels.append(v)
        elif view == "dense":
            panels.append(visualize_depth(dense[None], max_depth=max_depth, min_depth=min_depth)[0])
        else:
            raise ValueError(f"Invalid order: {view}")
    return panels
