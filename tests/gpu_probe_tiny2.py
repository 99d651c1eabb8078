import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from test_gpu_tapes import _tiny_vae_setup
from depth_completion_b200.synthetic import make_frame
dev = torch.device("cuda:0")
unet, vae, ctx, eng = _tiny_vae_setup(dev)
fr = make_frame(H=96, W=128, n_points=100, seed=3)
img, sp = fr["img"].to(dev), fr["sparse"].to(dev)
x = torch.randn(1, 4, eng.lh, eng.lw, device=dev).bfloat16()
def stage(name, fn):
    try:
        r = fn(); torch.cuda.synchronize(); print("ok  ", name, flush=True); return r
    except Exception as e:
        print("FAIL", name, str(e)[:200], flush=True); sys.exit(1)
stage("begin_frame", lambda: eng.begin_frame(img, sp, x, 10.0))
for i in range(4):
    stage(f"run step {i}", lambda: eng.run(1))
stage("get_state", lambda: eng.get_state())
stage("decode_final", lambda: eng.decode_final())
stage("begin_frame 2", lambda: eng.begin_frame(img, sp, x, 10.0))
stage("run 20", lambda: eng.run(20))
stage("decode_final 2", lambda: eng.decode_final())
