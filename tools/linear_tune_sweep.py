"""Output-tile width x split-K sweep of the UNet's linears (y = x W^T, forward and input gradient) against the planner's
own choice (mdc_dbg_tune: ksplit; bn through mdc_dbg_gemm; weights rotate through 8 copies so they stream from HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
LEVELS = [(6912, 320), (1728, 640), (432, 1280), (108, 1280)]
if os.environ.get("SMALL_ONLY"):
    LEVELS = LEVELS[1:]
for (M, d) in LEVELS:
    for name, N, K in (("qkv", 3 * d, d), ("to_out/proj", d, d), ("ff_in", 8 * d, d), ("ff_in dgrad", d, 8 * d), ("ff_out", d, 4 * d),
                       ("ff_out dgrad", 4 * d, d), ("qkv dgrad", d, 3 * d)):
        A = torch.randn(1, 1, M, K, device=dev, generator=g).bfloat16()
        B = (torch.randn(1, 1, N, K, device=dev, generator=g) * 0.02).bfloat16()
        gf = 2.0 * M * N * K / 1e9
        debug.tune(ksplit=-1, wcopies=8)
        _, auto = debug.gemm(A, B, iters=40)
        res = []
        for bn in (0, 64, 96, 128, 160, 192, 256, 320):
            if bn and N % bn:
                continue
            for ks in (1, 2, 3, 4, 5, 6, 8, 10, 12):
                if ks > 1 and (K // 64 < 4 * ks or M > 432):  # the planner only splits launches with <= 100 output tiles
                    continue
                if os.environ.get("TRACE"):
                    print("  trying", M, N, K, "bn", bn, "ks", ks, flush=True)
                try:
                    debug.tune(ksplit=ks, wcopies=8)
                    _, ms = debug.gemm(A, B, bn=bn, iters=40)
                    torch.cuda.synchronize()
                except Exception:
                    continue
                res.append((ms, bn, ks))
        debug.tune()
        if os.environ.get("SWEEP_CSV"):
            with open(os.environ["SWEEP_CSV"], "a") as f:
                f.write(f"{M},{N},{K},auto,auto,{auto * 1e3:.2f}\n")
                for ms, bn, ks in res:
                    f.write(f"{M},{N},{K},{bn},{ks},{ms * 1e3:.2f}\n")
        res.sort()
        top = ", ".join(f"bn {b} ks {k}: {m * 1e3:.1f}" for m, b, k in res[:4])
        print(f"M {M} {name:13s} N {N} K {K} ({gf:.1f} GF): planner {auto * 1e3:.1f} us | best {res[0][0] * 1e3:.1f} | {top}", flush=True)
