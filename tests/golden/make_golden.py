"""Generates tests/golden/tiny_96x128.npz: a fixed run of the fp32 oracle ON THE CPU (tiny UNet / VAE, weights seed
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


if __name__ == "__main__":
    out = run()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_96x128.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
