"""Run by tests/test_gpu_switches.py in a subprocess whose environment carries MDC_NO_* switches (they are read once per
process): tape parity (UNet + decoder, forward and input gradient) and a short pipeline run on the tiny models."""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), HERE]

from helpers import build_engine, build_models, rel_l2  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=True)
    eng = build_engine(unet, vae, ctx, ucfg, vcfg, 1, 96, 128, 128, 50, dev)
    g = torch.Generator(device=dev).manual_seed(3)
    z = torch.randn(1, 4, eng.lh, eng.lw, device=dev, generator=g).bfloat16().float()
    x = z.clone().requires_grad_(True)
    y = vae.decode(x)
    dout = torch.randn(y.shape, device=dev, generator=g).bfloat16().float()
    y.backward(dout)
    e_f, e_b = rel_l2(eng.dbg_forward(1, 0, z), y), rel_l2(eng.dbg_backward(1, dout), x.grad)
    assert e_f < 4e-2 and e_b < 6e-2, ("decoder", e_f, e_b)
    xin = torch.randn(1, 8, eng.lh, eng.lw, device=dev, generator=g).bfloat16().float()
    x = xin.clone().requires_grad_(True)
    y = unet(x, torch.tensor(999, device=dev), ctx)
    dout = torch.randn(y.shape, device=dev, generator=g).bfloat16().float()
    y.backward(dout)
    u_f, u_b = rel_l2(eng.dbg_forward(0, 0, xin), y), rel_l2(eng.dbg_backward(0, dout), x.grad)
    assert u_f < 4e-2 and u_b < 6e-2, ("unet", u_f, u_b)
    eng.close()
    if os.environ.get("MDC_GNEPI") or os.environ.get("MDC_NO_ROWSHARE"):
        # the switches that only act on large images: decoder tape on a 512x768 frame (two-pass GroupNorm with the
        # statistics from the conv epilogues / tap-by-tap instead of row-shared convolutions)
        big = build_engine(unet, vae, ctx, ucfg, vcfg, 2, 512, 768, 768, 50, dev)
        z = torch.randn(2, 4, big.lh, big.lw, device=dev, generator=g).bfloat16().float()
        x = z.clone().requires_grad_(True)
        y = vae.decode(x)
        dout = torch.randn(y.shape, device=dev, generator=g).bfloat16().float()
        y.backward(dout)
        b_f, b_b = rel_l2(big.dbg_forward(1, 0, z), y), rel_l2(big.dbg_backward(1, dout), x.grad)
        assert b_f < 4e-2 and b_b < 6e-2, ("large decoder", b_f, b_b)
        big.close()
    from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
    from depth_completion_b200.synthetic import make_frame
    from oracle.marigold_dc import OraclePipeline

    fr = make_frame(H=96, W=128, n_points=100)
    img, sp = fr["img"].to(dev), fr["sparse"].to(dev)
    pipe = MarigoldDepthCompletionPipeline(unet, vae)
    pipe.empty_text_embedding = ctx
    dense, _ = pipe(img, sp, fr["max_depth"], steps=5, resolution=128)
    ref, _ = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())(
        img, sp, fr["max_depth"], steps=5, resolution=128)
    d = ((dense - ref).abs().mean() / fr["max_depth"]).item()
    assert torch.isfinite(dense).all() and d < 3e-2, d
    import hashlib
    print("DENSE_SHA", hashlib.sha256(dense.detach().float().cpu().numpy().tobytes()).hexdigest())
    print(f"SWITCH_CASE_OK decoder {e_f:.3e}/{e_b:.3e} unet {u_f:.3e}/{u_b:.3e} pipeline {d:.3e}")


if __name__ == "__main__":
    main()
