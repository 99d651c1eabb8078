"""Manual probe: error of the engine vs the fp32 oracle next to the error of the oracle run in bf16 (torch kernels)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_models, build_engine, rel_err, rel_l2
from depth_completion_b200 import ddim

dev = torch.device("cuda:0")
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=True)
eng = build_engine(unet, vae, ctx, ucfg, vcfg, 1, 96, 128, 128, 50, dev)
u16, v16 = copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16()
ts = ddim.trailing_timesteps(50)
for scale in (1.0, 5.0):
    z = (torch.randn(1, 4, eng.lh, eng.lw, device=dev) * scale).bfloat16().float()
    x = z.clone().requires_grad_(True); y = vae.decode(x)
    dout = torch.randn_like(y).bfloat16().float(); y.backward(dout)
    x16 = z.bfloat16().requires_grad_(True); y16 = v16.decode(x16); y16.backward(dout.bfloat16())
    got = eng.dbg_forward(1, 0, z); din = eng.dbg_backward(1, dout)
    print(f"decoder z*{scale}: fwd ours {rel_l2(got, y):.4e} torch-bf16 {rel_l2(y16, y):.4e} | bwd ours {rel_l2(din, x.grad):.4e} torch-bf16 {rel_l2(x16.grad, x.grad):.4e}")
for step in (0, 30):
    t = torch.tensor(int(ts[step]), device=dev)
    xin = torch.randn(1, 8, eng.lh, eng.lw, device=dev).bfloat16().float()
    x = xin.clone().requires_grad_(True); y = unet(x, t, ctx)
    dout = torch.randn_like(y).bfloat16().float(); y.backward(dout)
    x16 = xin.bfloat16().requires_grad_(True); y16 = u16(x16, t, ctx.bfloat16()); y16.backward(dout.bfloat16())
    got = eng.dbg_forward(0, step, xin); din = eng.dbg_backward(0, dout)
    print(f"unet step {step}: fwd ours {rel_l2(got, y):.4e} torch-bf16 {rel_l2(y16, y):.4e} | bwd ours {rel_l2(din, x.grad):.4e} torch-bf16 {rel_l2(x16.grad, x.grad):.4e}")
