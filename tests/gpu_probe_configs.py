"""Manual probe: BASELINE.json configs (c) KITTI-shaped and (e) high-res batch-4 at full model size: builds, runs, timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
from depth_completion_b200.synthetic import make_batch

dev = torch.device("cuda:0")
unet, vae, ctx = bench.make_models(dev, False)
pipe = MarigoldDepthCompletionPipeline(unet, vae)
pipe.empty_text_embedding = ctx
for name, n, H, W, res, kind, md in [("c_kitti", 1, 352, 1216, 1216, "kitti", 80.0), ("e_highres_b4", 4, 768, 1024, 1024, "nyu", 10.0),
                                     ("b_res640", 1, 480, 640, 640, "nyu", 10.0), ("b2_two_frames", 2, 480, 640, 768, "nyu", 10.0)]:
    b = make_batch(n, H=H, W=W, kind=kind, n_points=500, max_depth=md, min_field=1.0 if kind == "kitti" else 0.5)
    img, sp = b["img"].to(dev), b["sparse"].to(dev)
    t0 = time.time()
    d, lat = pipe(img, sp, md, steps=50, resolution=res)
    torch.cuda.synchronize(); t1 = time.time()
    d, lat = pipe(img, sp, md, steps=50, resolution=res)
    torch.cuda.synchronize(); t2 = time.time()
    eng = next(iter(pipe._engines.values()))
    valid = (sp > 0)
    mae = (d - b["gt"].to(dev)).abs()[valid].mean().item()
    print(f"{name}: points/frame {int(valid.sum().item()) // n} latent {eng.lh}x{eng.lw} mem {eng.device_bytes() / 2**30:.1f} GB first {t1 - t0:.2f}s "
          f"second {t2 - t1:.3f}s ({(t2 - t1) / 50 * 1e3:.1f} ms/step, {n / (t2 - t1):.3f} frames/s) finite {bool(torch.isfinite(d).all())} "
          f"loss {pipe.last_losses.tolist()} mae@guides {mae:.4f}", flush=True)
