"""Run by tests/test_gpu_switches.py::test_sparse_output_head_matches_dense_head in two subprocesses (with and without
MDC_NO_SPARSEHEAD, which is read once per process): ONE guided step of the full-width engine on a BASELINE-shaped frame
from the same state; the loss, the total latent gradient and the decoder-input gradient are saved for comparison."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]


def main(out_path, kind, n):
    from depth_completion_b200.config import UNetConfig, VAEConfig
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_batch, random_init_modules

    dev = torch.device("cuda:0")
    unet, vae, ctx = random_init_modules(UNetConfig(), VAEConfig(), dev)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    geo = dict(nyu=dict(H=480, W=640, res=768, max_depth=10.0, min_field=0.5), kitti=dict(H=352, W=1216, res=1216, max_depth=80.0, min_field=1.0))[kind]
    fr = make_batch(n, H=geo["H"], W=geo["W"], kind=kind, n_points=500, max_depth=geo["max_depth"], min_field=geo["min_field"], seed=3)
    pipe(fr["img"].to(dev), fr["sparse"].to(dev), geo["max_depth"], steps=50, resolution=geo["res"], _begin_only=True)
    eng = list(pipe._engines.values())[-1]
    rec = {}
    for step in range(2):
        eng.run(1)
        x, sc, sh, ls = eng.get_state()
        rec[f"loss{step}"] = ls.clone()
        rec[f"grad{step}"] = eng.dbg_buffer("grad").cpu()
        rec[f"dz{step}"] = eng.dbg_read("vae.in", grad=True).cpu()
        rec[f"x{step}"] = x.float().cpu()
    torch.save(rec, out_path)
    print("HEAD_CASE_OK")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
