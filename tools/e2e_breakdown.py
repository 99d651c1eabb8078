"""Where the end-to-end time of one call goes besides the 50 guided steps (config b, pinned host inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from depth_completion_b200.pipeline import MarigoldDepthCompletionPipeline
from depth_completion_b200.synthetic import make_frame

dev = torch.device("cuda:0")
w = bench.workload(False)
unet, vae, ctx = bench.make_models(dev, False)
pipe = MarigoldDepthCompletionPipeline(unet, vae)
pipe.empty_text_embedding = ctx
fr = make_frame(H=w["H"], W=w["W"], n_points=w["n_points"], max_depth=w["max_depth"])
img, sp = fr["img"].pin_memory(), fr["sparse"].pin_memory()
for _ in range(2):
    pipe(img, sp, w["max_depth"], steps=50, resolution=w["resolution"])
torch.cuda.synchronize()


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(3):
    t0 = t()
    tk = pipe.submit(img, sp, w["max_depth"], steps=50, resolution=w["resolution"], _begin_only=True)
    t1 = t()
    eng = tk["eng"] if isinstance(tk, dict) and "eng" in tk else next(iter(pipe._engines.values()))
    eng.run(50)
    t2 = t()
    dn = eng.decode_final(closed_form=False)
    t3 = t()
    st = eng.get_state()
    host = dn.cpu()
    t4 = t()
    t5 = time.perf_counter()
    d2, _ = pipe(img, sp, w["max_depth"], steps=50, resolution=w["resolution"])
    h2 = d2.cpu()
    t6 = t()
    print(f"rep {rep}: submit(begin only: H2D + prologue + encoder) {1e3 * (t1 - t0):.2f} ms | 50 steps {1e3 * (t2 - t1):.2f} | "
          f"decode_final {1e3 * (t3 - t2):.2f} | get_state + D2H {1e3 * (t4 - t3):.2f} | whole call {1e3 * (t6 - t5):.2f}", flush=True)
# finer: the pieces of submit
import torch.cuda.nvtx  # noqa
x = None
for rep in range(2):
    t0 = t(); a, b = img.to(dev), sp.to(dev); t1 = t()
    gen = torch.Generator(device=dev).manual_seed(2024); c = torch.randn((1, 4, 72, 96), device=dev, dtype=torch.bfloat16, generator=gen); t2 = t()
    from depth_completion_b200.pipeline import check_image
    check_image(a); t3 = t()
    print(f"H2D {1e3 * (t1 - t0):.2f} ms, generator + randn {1e3 * (t2 - t1):.2f}, check_image {1e3 * (t3 - t2):.2f}")
