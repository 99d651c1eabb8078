"""Manual probe: localise gradient differences of one guided step between the engine and the fp32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_models, build_engine, rel_err, rel_l2
from depth_completion_b200 import prologue
from depth_completion_b200.synthetic import make_frame
from oracle.marigold_dc import OraclePipeline, compute_loss
from oracle import image_processor as ip

dev = torch.device("cuda:0")
H, W, RES, NPTS, STEPS = 96, 128, 128, 100, 50
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=True)
fr = make_frame(H=H, W=W, n_points=NPTS)
img, sparse = fr["img"].to(dev), fr["sparse"].to(dev)
eng = build_engine(unet, vae, ctx, ucfg, vcfg, 1, H, W, RES, STEPS, dev)
op = OraclePipeline(unet, vae, ctx)   # fp32 oracle (weights are bf16-representable)
st = op.preprocess(img, sparse, fr["max_depth"], 0.0, "minmax", RES, 2024, None, 0.9)
x_in = st["x"].bfloat16().float()
img_lat = st["img_latents"].bfloat16().float()
N = 1
masks = st["masks"]; guide = st["sparses_normed"]
gmin, gmax = prologue.masked_minmax(guide.view(N, -1), masks.view(N, -1))
lo, hi = st["min_depths"].view(N), st["max_depths"].view(N)
eng.begin(img_lat, x_in, guide, masks, torch.stack([gmin, gmax], 1).cpu().numpy(), torch.stack([lo, hi], 1).cpu().numpy())
eng.run(1)
xo, sc, sh, ls = eng.get_state()

# oracle, with every interface tensor retained
op.scheduler.set_timesteps(STEPS, device=dev)
t = op.scheduler.timesteps[0]
x = x_in.clone().requires_grad_(True)
cat = torch.cat([img_lat, x], 1); cat.retain_grad()
v = unet(cat, t, ctx); v.retain_grad()
x0 = op.scheduler.step(v, t, x).pred_original_sample
z = x0 / vae.cfg.scaling_factor; z.retain_grad()
dec = vae.decode(z); dec.retain_grad()
y = dec.mean(1, keepdim=True).clip(-1, 1); y = (y + 1) / 2
aff = ip.resize_antialias(ip.unpad_image(y, st["padding"]), st["orig_res"], "bilinear")
scales = torch.ones(1, 1, 1, 1, device=dev, requires_grad=True); shifts = torch.zeros(1, 1, 1, 1, device=dev, requires_grad=True)
dense = op.affine_to_metric(aff, guide, masks, scales, shifts).clamp(0, 1)
loss = compute_loss(dense, guide, masks)
loss.backward(torch.ones_like(loss))
print("loss ours", ls.tolist(), "oracle", loss.tolist(), "s_grad", scales.grad.item(), "t_grad", shifts.grad.item())
def cmp(name, got, ref):
    cos = torch.nn.functional.cosine_similarity(got.flatten().float(), ref.flatten().float(), dim=0).item()
    print(f"{name:28s} rel_l2 {rel_l2(got, ref):.4e} cos {cos:.6f} |got| {got.norm().item():.4e} |ref| {ref.norm().item():.4e}")
cmp("v (unet.out)", eng.dbg_read("unet.out"), v)
cmp("z (vae.in)", eng.dbg_read("vae.in"), z)
cmp("dec (vae.out)", eng.dbg_read("vae.out"), dec)
cmp("d dec", eng.dbg_read("vae.out", True), dec.grad)
cmp("d z", eng.dbg_read("vae.in", True), z.grad)
cmp("d v", eng.dbg_read("unet.out", True), v.grad)
cmp("d unet_in[4:8]", eng.dbg_read("unet.in", True)[:, 4:8], cat.grad[:, 4:8])
cmp("dx_direct", eng.dbg_buffer("dx_direct"), (x.grad - cat.grad[:, 4:8]))
cmp("grad total", eng.dbg_buffer("grad"), x.grad)
# feed the oracle's d dec into the engine's decoder backward alone
dz2 = eng.dbg_backward(1, dec.grad)
cmp("dec bwd from oracle d dec", dz2, z.grad)
dz3 = eng.dbg_backward(1, torch.randn_like(dec.grad) * dec.grad.abs().max())
# nonzero pattern
nz_o = (dec.grad.abs() > 0).float().mean().item(); nz_e = (eng.dbg_read("vae.out", True).abs() > 0).float().mean().item()
print("nonzero fraction d dec: oracle", nz_o, "engine", nz_e)
m = dec.mean(1); print("fraction |mean|>1:", (m.abs() > 1).float().mean().item())
# module-level backward errors inside the decoder for this sparse gradient
acts, grads = {}, {}
names = set(eng.dbg_tensor_names())
hooks = []
for name, mod in vae.named_modules():
    key = "vae." + name
    if key in names:
        def fh(mm, i, o, key=key):
            if o.requires_grad:
                o.register_hook(lambda g, key=key: grads.__setitem__(key, g.detach()))
        hooks.append(mod.register_forward_hook(fh))
z2 = z.detach().clone().requires_grad_(True)
d2 = vae.decode(z2); d2.backward(dec.grad)
eng.dbg_forward(1, 0, z.detach()); eng.dbg_backward(1, dec.grad)
for k in reversed(list(grads)):
    cmp("  bwd " + k[4:], eng.dbg_read(k, True), grads[k])
