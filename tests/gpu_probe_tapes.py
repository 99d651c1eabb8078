"""Manual probe (not collected by pytest): per-module forward/backward errors of the UNet / decoder tapes vs the oracle."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_models, build_engine, rel_err, rel_l2

dev = torch.device("cuda:0")
tiny = os.environ.get("FULL", "0") != "1"
H, W, RES = (96, 128, 128) if tiny else (480, 640, 768)
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=tiny)
eng = build_engine(unet, vae, ctx, ucfg, vcfg, 1, H, W, RES, 50, dev)
print("engine device MB", eng.device_bytes() / 2**20, "latent", eng.lh, eng.lw, flush=True)
names = set(eng.dbg_tensor_names())

def probe(which, model, prefix, inp, fwd_fn, step=0):
    acts, grads = {}, {}
    hooks = []
    for name, mod in model.named_modules():
        key = prefix + name
        if key in names:
            def fh(m, i, o, key=key):
                acts[key] = o.detach()
                if o.requires_grad:
                    o.register_hook(lambda g, key=key: grads.__setitem__(key, g.detach()))
            hooks.append(mod.register_forward_hook(fh))
    x = inp.clone().requires_grad_(True)
    y = fwd_fn(x)
    got = eng.dbg_forward(which, step, inp)
    print(f"[{prefix}] forward out: rel_max {rel_err(got, y):.3e} rel_l2 {rel_l2(got, y):.3e}", flush=True)
    for k in acts:
        g = eng.dbg_read(k)
        print(f"   fwd {k:60s} max {rel_err(g, acts[k]):.3e} l2 {rel_l2(g, acts[k]):.3e}")
    dout = torch.randn_like(y).bfloat16().float()
    y.backward(dout)
    din = eng.dbg_backward(which, dout)
    print(f"[{prefix}] backward din: rel_max {rel_err(din, x.grad):.3e} rel_l2 {rel_l2(din, x.grad):.3e}", flush=True)
    for k in reversed(list(acts)):
        if k in grads:
            try:
                g = eng.dbg_read(k, grad=True)
                print(f"   bwd {k:60s} max {rel_err(g, grads[k]):.3e} l2 {rel_l2(g, grads[k]):.3e}")
            except Exception as e:
                print("   bwd", k, "n/a", e)
    for h in hooks: h.remove()

try:
    z = torch.randn(1, 4, eng.lh, eng.lw, device=dev).bfloat16().float()
    probe(1, vae, "vae.", z, lambda x: vae.decode(x))
except Exception:
    traceback.print_exc()
try:
    xin = torch.randn(1, 8, eng.lh, eng.lw, device=dev).bfloat16().float()
    from depth_completion_b200 import ddim
    ts = ddim.trailing_timesteps(50)
    for step in (0, 30):
        probe(0, unet, "unet.", xin, lambda x: unet(x, torch.tensor(int(ts[step]), device=dev), ctx), step=step)
except Exception:
    traceback.print_exc()
print(eng.dbg_time_tapes(3))
