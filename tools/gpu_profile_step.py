"""Profiling driver (run under ncu): builds the full-size engine and runs a few guided steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
from depth_completion_b200.synthetic import make_frame

tiny = os.environ.get("TINY", "0") == "1"
steps = int(os.environ.get("STEPS", "2"))
dev = torch.device("cuda:0")
w = bench.workload(tiny)
unet, vae, ctx = bench.make_models(dev, tiny, vae_kind=os.environ.get("VAE", "original"))
pipe = MarigoldDepthCompletionPipeline(unet, vae)
pipe.empty_text_embedding = ctx
fr = make_frame(H=w["H"], W=w["W"], n_points=w["n_points"], max_depth=w["max_depth"])
pipe(fr["img"].to(dev), fr["sparse"].to(dev), w["max_depth"], steps=50, resolution=w["resolution"], _begin_only=True)
eng = next(iter(pipe._engines.values()))
torch.cuda.synchronize()
print("launches per step:", end=" ")
l0 = eng.launch_count()
eng.run(steps)
torch.cuda.synchronize()
print((eng.launch_count() - l0) // steps)
if os.environ.get("OPS_CSV"):
    eng.dbg_profile_ops(os.environ["OPS_CSV"], 5)
    print("wrote", os.environ["OPS_CSV"])
print(eng.dbg_time_tapes(3))
