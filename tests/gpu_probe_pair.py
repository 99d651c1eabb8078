"""CTA-pair (cta_group::2) vs single-CTA vs multicast on the large decoder convolutions and a square GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
names = {1: "single", 2: "pair", 3: "mcast2"}
for (H, W, Ci, Co) in [(576, 768, 128, 128), (576, 768, 256, 256), (288, 384, 512, 512), (144, 192, 512, 512), (576, 768, 256, 128),
                       (72, 96, 320, 320), (72, 96, 960, 320), (36, 48, 640, 640), (36, 48, 1920, 640)]:
    x = torch.randn(1, H, W, Ci, device=dev).bfloat16()
    w = torch.randn(Co, Ci, 3, 3, device=dev) / (3 * Ci ** 0.5)
    debug.tune(cs=1)
    ref, _ = debug.conv3x3(x, w)
    out = []
    for cs in (1, 2, 3):
        debug.tune(cs=cs, wcopies=4)
        y, ms = debug.conv3x3(x, w, iters=20)
        err = (y.float() - ref.float()).abs().max().item()
        out.append(f"{names[cs]} {ms*1e3:7.1f}us {2*H*W*9*Ci*Co/ms/1e9:6.0f}TF/s" + (f" ERR {err:.3f}" if err > 0.05 else ""))
    print(f"conv {H}x{W} {Ci}->{Co}: " + " | ".join(out), flush=True)
for (M, N, K) in [(8192, 8192, 8192), (6912, 2560, 320), (6912, 320, 1280), (6912, 960, 320), (1728, 5120, 640), (1728, 640, 2560)]:
    A = torch.randn(1, 1, M, K, device=dev).bfloat16()
    B = torch.randn(1, 1, N, K, device=dev).bfloat16()
    out = []
    for cs in (1, 2, 3):
        debug.tune(cs=cs)
        y, ms = debug.gemm(A, B, iters=20)
        out.append(f"{names[cs]} {ms*1e3:7.1f}us {2*M*N*K/ms/1e9:6.0f}TF/s")
    print(f"gemm {M}x{N}x{K}: " + " | ".join(out), flush=True)
debug.tune()
