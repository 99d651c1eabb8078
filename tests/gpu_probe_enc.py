import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_engine, build_models
dev = torch.device("cuda:0")
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=True)
for (H, W, res) in [(96, 128, 128), (60, 80, 128)]:
    eng = build_engine(unet, vae, ctx, ucfg, vcfg, 2, H, W, res, 50, dev)
    imgs = torch.randint(0, 256, (2, 3, H, W), device=dev, dtype=torch.uint8)
    got = eng.encode(imgs)
    print(H, W, "latent norm", got.float().norm().item(), "geometry", eng.lh, eng.lw)
    for name in eng.dbg_tensor_names():
        if name.startswith("vae.enc") :
            t = eng.dbg_read(name)
            print("   ", name, tuple(t.shape), f"{t.norm().item():.4f}", "nan" if torch.isnan(t).any() else "")
