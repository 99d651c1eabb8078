import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_gpu_tapes import _tiny_vae_setup
from helpers import rel_l2
dev = torch.device("cuda:0")
unet, vae, ctx, eng = _tiny_vae_setup(dev)
g = torch.Generator(device=dev).manual_seed(5)
z = torch.randn(1, 4, eng.lh, eng.lw, device=dev, generator=g).bfloat16().float()
x = z.clone().requires_grad_(True)
acts = {}
layers = vae.decoder.layers
h = torch.tanh(x / 3) * 3
for i, l in enumerate(layers):
    h = l(h)
    h.retain_grad()
    acts[i] = h
y = h.mul(2).sub(1)
dout = torch.randn(y.shape, device=dev, generator=g).bfloat16().float()
y.backward(dout)
got = eng.dbg_forward(1, 0, z)
din = eng.dbg_backward(1, dout)
print("fwd", rel_l2(got, y), "bwd", rel_l2(din, x.grad))
names = eng.dbg_tensor_names()
for i, l in enumerate(layers):
    key = f"vae.decoder.layers.{i}"
    cand = [key + ".conv.4", key]
    for k in cand:
        if k in names:
            a = eng.dbg_read(k)
            ga = eng.dbg_read(k, grad=True)
            ref_a = acts[i]
            ref_g = acts[i].grad
            relu_out = k.endswith("conv.4") or i == 0
            if i == 0:
                ref_a = acts[1]; ref_g = acts[1].grad
            if relu_out:
                ref_g = ref_g * (ref_a > 0)
            print(f"{k:32s} act {rel_l2(a, ref_a):.4f} grad {rel_l2(ga, ref_g):.4f}")
            break
