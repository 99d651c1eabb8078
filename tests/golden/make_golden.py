"""Generates tests/golden/tiny_96x128.npz (and tiny_96x128_taesd.npz, the same with the AutoencoderTiny VAE): a fixed run of the fp32 oracle ON THE CPU (tiny UNet / VAE, weights seed
1234 rounded to bf16-representable values, synthetic 96x128 frame seed 0, resolution 128, the first 3 of 50 guided
steps).  The reference itself cannot be imported in this container (diffusers is absent), so these
vectors pin the ORACLE: `tests/test_oracle_formulas.py` re-runs it on the CPU against them, and the GPU tests drive
the CUDA engine from the stored state (teacher forcing) and compare with the stored results.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run():
    from helpers import build_models
    from depth_completion_b200.synthetic import make_frame
    from oracle.marigold_dc import OraclePipeline

    torch.set_num_threads(1)  # fixed reduction order
    unet, vae, ctx, _, _ = build_models("cpu", tiny=True, seed=1234)
    fr = make_frame(H=96, W=128, n_points=100, seed=0)
    pipe = OraclePipeline(unet, vae, ctx)
    st = pipe.preprocess(fr["img"], fr["sparse"], fr["max_depth"], 0.0, "minmax", 128, 2024, None, 0.9)
    tr = []
    dense, lat = pipe(fr["img"], fr["sparse"], fr["max_depth"], steps=50, resolution=128, trace=tr.append, max_steps=3)
    out = dict(img=fr["img"].numpy(), sparse=fr["sparse"].numpy(), gt=fr["gt"].numpy(), holdout=fr["holdout"].numpy(),
               max_depth=np.float32(fr["max_depth"]), x_init=st["x"].numpy(), img_latents=st["img_latents"].numpy(),
               guide=st["sparses_normed"].numpy(), mask=st["masks"].numpy(), depth_min=st["min_depths"].numpy(),
               depth_max=st["max_depths"].numpy(), dense=dense.numpy(), latents=lat.numpy())
    for k in ("v", "x0", "losses", "grad", "scales", "shifts", "x_out", "x_adam"):
        out["step_" + k] = np.stack([t[k].float().numpy() for t in tr])
    return out


def run_taesd():
    """Same frame and UNet, AutoencoderTiny (oracle/taesd.py, seed 77 scaled by 1.5 as in the GPU tests) as the VAE: the
    reference CLI's default configuration (predict.py:484-488)."""
    from helpers import build_models
    from depth_completion_b200.synthetic import make_frame
    from oracle.marigold_dc import OraclePipeline
    from oracle.taesd import AutoencoderTiny

    torch.set_num_threads(1)
    unet, _, ctx, _, _ = build_models("cpu", tiny=True, seed=1234)
    torch.manual_seed(77)
    vae = AutoencoderTiny()
    with torch.no_grad():
        for p in vae.parameters():
            p.copy_((p * 1.5).bfloat16().float())
    vae.requires_grad_(False)
    fr = make_frame(H=96, W=128, n_points=100, seed=0)
    pipe = OraclePipeline(unet, vae, ctx)
    st = pipe.preprocess(fr["img"], fr["sparse"], fr["max_depth"], 0.0, "minmax", 128, 2024, None, 0.9)
    tr = []
    dense, lat = pipe(fr["img"], fr["sparse"], fr["max_depth"], steps=50, resolution=128, trace=tr.append, max_steps=2)
    out = dict(x_init=st["x"].numpy(), img_latents=st["img_latents"].numpy(), dense=dense.numpy())
    for k in ("v", "x0", "losses", "grad", "scales", "x_adam"):
        out["step_" + k] = np.stack([t[k].float().numpy() for t in tr])
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("tiny_96x128.npz", run), ("tiny_96x128_taesd.npz", run_taesd)):
        out = fn()
        path = os.path.join(here, name)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
