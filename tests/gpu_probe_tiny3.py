import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_gpu_tapes import _tiny_vae_setup
from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
from depth_completion_b200.synthetic import make_frame
dev = torch.device("cuda:0")
unet, vae, ctx, eng = _tiny_vae_setup(dev)
if os.environ.get("KEEP") != "1":
    eng.close()
fr = make_frame(H=96, W=128, n_points=100, seed=3)
img, sp = fr["img"].to(dev), fr["sparse"].to(dev)
pipe = MarigoldDepthCompletionPipeline(unet, vae)
pipe.empty_text_embedding = ctx
steps = int(os.environ.get("STEPS", "20"))
for rep in range(3):
    try:
        dense, lat = pipe(img, sp, fr["max_depth"], steps=steps, resolution=128)
        torch.cuda.synchronize()
        print("ok rep", rep, float(dense.mean()), flush=True)
    except Exception as e:
        print("FAIL rep", rep, str(e)[:300], flush=True); sys.exit(1)
