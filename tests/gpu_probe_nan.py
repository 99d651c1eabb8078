import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_models, build_engine
dev = torch.device("cuda:0")
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=True)
eng = build_engine(unet, vae, ctx, ucfg, vcfg, 1, 96, 128, 128, 50, dev)
g = torch.Generator(device=dev).manual_seed(0)
xin = torch.randn(1, 8, eng.lh, eng.lw, device=dev, generator=g).bfloat16().float()
out = eng.dbg_forward(0, 0, xin)
print("fwd nan:", torch.isnan(out).any().item())
din = eng.dbg_backward(0, torch.randn(out.shape, device=dev, generator=g).bfloat16().float())
print("bwd nan:", torch.isnan(din).any().item())
for name in eng.dbg_tensor_names():
    if not name.startswith("unet"): continue
    try:
        gr = eng.dbg_read(name, grad=True)
    except Exception as e:
        continue
    a = eng.dbg_read(name)
    print(f"{name:50s} act nan {torch.isnan(a).any().item()} grad nan {torch.isnan(gr).float().mean().item():.4f} shape {tuple(gr.shape)}")
