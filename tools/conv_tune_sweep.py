"""Output-tile width x split-K sweep of the low-resolution 3x3 convolutions against the planner's own choice
(mdc_dbg_tune: bn, ksplit; weights rotate through 8 copies so they stream from HBM as in the real step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
SHAPES = [(1920, 640, 36, 48), (960, 320, 72, 96), (2560, 1280, 18, 24), (1280, 640, 36, 48), (640, 320, 72, 96), (512, 512, 72, 96),
          (1280, 1280, 18, 24), (1280, 1280, 9, 12), (320, 320, 72, 96), (640, 640, 36, 48)]
for (C, Cout, H, W) in SHAPES:
    for dgrad in (False, True):
        cin = Cout if dgrad else C
        x = torch.randn(1, H, W, cin, device=dev, generator=g).bfloat16()
        w = torch.randn(Cout, C, 3, 3, device=dev, generator=g) * 0.02
        gf = 2.0 * H * W * 9 * C * Cout / 1e9
        debug.tune(ksplit=-1, wcopies=8)
        _, auto = debug.conv3x3(x, w, dgrad=dgrad, iters=30)
        best = (auto, "auto")
        res = []
        for bn in (0, 64, 96, 128, 160, 192, 256):
            for ks in (1, 2, 3, 4, 6, 8, 12):
                try:
                    debug.tune(bn=bn, ksplit=ks, wcopies=8)
                    _, ms = debug.conv3x3(x, w, dgrad=dgrad, iters=30)
                except Exception as e:  # illegal combination for this shape
                    continue
                res.append((ms, bn, ks))
                if ms < best[0]:
                    best = (ms, f"bn {bn} ksplit {ks}")
        debug.tune()
        res.sort()
        top = ", ".join(f"bn {b} ks {k}: {m * 1e3:.1f}" for m, b, k in res[:4])
        print(f"{'dgrad' if dgrad else 'fwd  '} {C}->{Cout} @{H}x{W} ({gf:.1f} GF): planner {auto * 1e3:.1f} us ({gf / auto / 1e3:.2f} TF/s) | "
              f"best {best[0] * 1e3:.1f} us ({best[1]}) | {top}", flush=True)
