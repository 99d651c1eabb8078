"""Manual probe: guided loop of the product vs the oracle (fp32 and bf16) on the tiny or full config."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from helpers import build_models, rel_err, rel_l2
from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
from depth_completion_b200.synthetic import make_frame
from oracle.marigold_dc import OraclePipeline, mae, rmse
import copy

dev = torch.device("cuda:0")
tiny = os.environ.get("FULL", "0") != "1"
STEPS = int(os.environ.get("STEPS", "50"))
H, W, RES, NPTS = (96, 128, 128, 100) if tiny else (480, 640, 768, 500)
unet, vae, ctx, ucfg, vcfg = build_models(dev, tiny=tiny)
fr = make_frame(H=H, W=W, n_points=NPTS)
img, sparse, gt, hold = fr["img"].to(dev), fr["sparse"].to(dev), fr["gt"].to(dev), fr["holdout"].to(dev)

pipe = MarigoldDepthCompletionPipeline(unet, vae)
pipe.empty_text_embedding = ctx
t0 = time.time()
dense, lat = pipe(img, sparse, fr["max_depth"], steps=STEPS, resolution=RES)
torch.cuda.synchronize(); t1 = time.time()
print(f"product first call {t1 - t0:.2f}s (includes engine build + weight packing)", flush=True)
t0 = time.time()
dense, lat = pipe(img, sparse, fr["max_depth"], steps=STEPS, resolution=RES)
torch.cuda.synchronize(); t1 = time.time()
print(f"product second call {t1 - t0:.3f}s for {STEPS} steps; loss {pipe.last_losses.tolist()} scale {pipe.last_scales.tolist()} shift {pipe.last_shifts.tolist()}", flush=True)
eng = next(iter(pipe._engines.values()))
print("device MB", eng.device_bytes() / 2**20, eng.dbg_time_tapes(2), flush=True)

def run_oracle(dtype, trace=None, max_steps=None):
    u, v = copy.deepcopy(unet).to(dtype), copy.deepcopy(vae).to(dtype)
    op = OraclePipeline(u, v, ctx.to(dtype))
    t0 = time.time()
    d, x = op(img, sparse, fr["max_depth"], steps=STEPS, resolution=RES, trace=trace, max_steps=max_steps)
    torch.cuda.synchronize()
    print(f"oracle {dtype} {time.time() - t0:.2f}s", flush=True)
    return d, x

tr16 = []
d16, x16 = run_oracle(torch.bfloat16, trace=tr16.append)
print("oracle bf16 losses first/last", tr16[0]["losses"].tolist(), tr16[-1]["losses"].tolist())
rng = float(fr["max_depth"])
def report(name, a, b):
    print(f"{name}: max|diff|/range {((a - b).abs().max() / rng).item():.4e}  mean|diff|/range {((a - b).abs().mean() / rng).item():.4e}")
report("product vs oracle-bf16 dense", dense, d16)
print("MAE/RMSE holdout product", mae(dense, gt, hold).item(), rmse(dense, gt, hold).item(), " oracle-bf16", mae(d16, gt, hold).item(), rmse(d16, gt, hold).item())
print("MAE/RMSE guide   product", mae(dense, gt, sparse > 0).item(), rmse(dense, gt, sparse > 0).item(), " oracle-bf16", mae(d16, gt, sparse > 0).item(), rmse(d16, gt, sparse > 0).item())
if tiny or os.environ.get("FP32", "0") == "1":
    tr32 = []
    d32, x32 = run_oracle(torch.float32, trace=tr32.append)
    report("product vs oracle-fp32 dense", dense, d32)
    report("oracle-bf16 vs oracle-fp32 dense", d16, d32)
    print("MAE/RMSE holdout oracle-fp32", mae(d32, gt, hold).item(), rmse(d32, gt, hold).item())

# teacher-forced single step: feed the oracle-bf16 state of step k, compare one step
for k in (0, min(STEPS - 1, 25)):
    st = tr16[k]
    # rebuild engine state at step k: x_in, and (for k>0) Adam moments are not restorable -> only compare raw gradient & v
    eng = next(iter(pipe._engines.values()))
    # use the pipeline internals: begin() with x_in then fast-forward the device step counter via k no-op? not available,
    # so only step 0 is exact; for k > 0 compare against a fresh oracle step at timestep index 0 is meaningless -> skip
    if k != 0:
        continue
    pipe2 = MarigoldDepthCompletionPipeline(unet, vae); pipe2.empty_text_embedding = ctx
    d1, x1 = pipe2(img, sparse, fr["max_depth"], steps=STEPS, resolution=RES) if False else (None, None)
    # run exactly one step through the same cached engine
    import numpy as np
    from depth_completion_b200 import prologue
    # re-begin by calling the pipeline with steps but intercept: simplest is a 1-step rerun using engine directly
    N = 1
    masks = sparse > 0
    lo, hi = prologue.masked_minmax(sparse.view(N, -1), masks.view(N, -1))
    guide = (sparse.clamp(min=lo.view(N,1,1,1), max=hi.view(N,1,1,1)) - lo.view(N,1,1,1)) / (hi - lo).view(N,1,1,1)
    gmin, gmax = prologue.masked_minmax(guide.view(N, -1), masks.view(N, -1))
    op = OraclePipeline(copy.deepcopy(unet).bfloat16(), copy.deepcopy(vae).bfloat16(), ctx.bfloat16())
    stt = op.preprocess(img, sparse, fr["max_depth"], 0.0, "minmax", RES, 2024, None, 0.9)
    eng.begin(stt["img_latents"], st["x_in"], guide, masks, torch.stack([gmin, gmax], 1).cpu().numpy(), torch.stack([lo, hi], 1).cpu().numpy())
    eng.run(1)
    xo, sc, sh, ls = eng.get_state()
    g = eng.dbg_buffer("grad")
    print(f"step0 loss ours {ls.tolist()} oracle {st['losses'].tolist()}")
    og = st["grad"].float()
    cos = torch.nn.functional.cosine_similarity(g.flatten(), og.flatten(), dim=0).item()
    print(f"step0 raw grad: rel_l2 {rel_l2(g, og):.4e} cosine {cos:.6f} norm ours {g.norm().item():.4e} oracle {og.norm().item():.4e} (oracle grad_norm {st['grad_norm'].tolist()})")
    print(f"step0 v: rel_l2 {rel_l2(eng.dbg_read('unet.out'), st['v']):.4e}")
    xa = eng.dbg_x_adam().float()
    same = ((xa - st["x_adam"].float()).abs() < 1e-2).float().mean().item()
    print(f"step0 x_adam agreement (|d|<1e-2): {same:.4f}; x_out rel_l2 {rel_l2(xo, st['x_out']):.4e}; scale ours {sc.tolist()} oracle {st['scales'].flatten().tolist()} shift {sh.tolist()} {st['shifts'].flatten().tolist()}")
    if tiny or os.environ.get("FP32", "0") == "1":
        og32 = tr32[0]["grad"].float()
        print(f"step0 raw grad vs fp32 oracle: ours rel_l2 {rel_l2(g, og32):.4e}; oracle-bf16 rel_l2 {rel_l2(og, og32):.4e}")
