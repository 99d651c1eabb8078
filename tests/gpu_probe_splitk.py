"""Split-K / tile-shape sweep for the low-resolution UNet convolutions (weights rotated so they stream from HBM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
shapes = [  # (H, W, Cin, Cout)
    (9, 12, 1280, 1280), (9, 12, 2560, 1280), (18, 24, 1280, 1280), (18, 24, 2560, 1280), (18, 24, 640, 1280),
    (36, 48, 640, 640), (36, 48, 1280, 640), (36, 48, 320, 640), (72, 96, 320, 320), (72, 96, 640, 320),
]
only = os.environ.get("ONLY")
for (H, W, Ci, Co) in shapes:
    if only and only != f"{H}x{W}x{Ci}x{Co}":
        continue
    x = torch.randn(1, H, W, Ci, device=dev).bfloat16()
    w = torch.randn(Co, Ci, 3, 3, device=dev) / (3 * Ci ** 0.5)
    ncopy = max(2, int(200e6 / (Co * Ci * 18)))
    debug.tune(cs=1)
    ref, _ = debug.conv3x3(x, w)
    line = [f"{H}x{W} {Ci}->{Co} gflop {2*H*W*9*Ci*Co/1e9:6.2f} |"]
    for bn in (0, 128, 64):
        for cs in [int(c) for c in os.environ.get("CS", "1,2,4").split(",")]:
            for ks in (1, -1, 2, 3, 4, 5, 6, 7, 9, 12, 14, 18, 24, 29):
                if bn and Co % bn:
                    continue
                try:
                    debug.tune(bn=bn, cs=cs, ksplit=ks, wcopies=ncopy)
                    out, ms = debug.conv3x3(x, w, iters=40)
                    err = (out.float() - ref.float()).abs().max().item()
                    line.append(f"bn{bn} cs{cs} ks{ks}: {ms*1e3:6.1f}us" + (f" ERR{err:.3f}" if err > 0.06 else ""))
                except Exception as e:
                    line.append(f"bn{bn} cs{cs} ks{ks}: FAIL {str(e)[:60]}")
    print("\n  ".join(line), flush=True)
debug.tune()
