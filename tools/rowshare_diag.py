"""Diagnostic: which row-offset view of the shared A box is wrong in the row-shared-taps conv mode?
mode 1 = tap by tap, 2 = row-shared (descriptor start address shifted, matrix base offset 0), 3 = shifted AND base offset s.
Measured on a B200 (profiles/r02_rowshare_diag.log, taken when modes 2 / 3 were still swapped): only base offset 0 is exact."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (H, W, C, Cout) in ((4, 128, 64, 64), (4, 256, 128, 128)):
    x = torch.randn(1, C, H, W, device=dev, generator=g).bfloat16()
    for mode in (1, 2, 3):
        for tap in list(range(9)) + [-1]:
            w = torch.randn(Cout, C, 3, 3, device=dev, generator=g) * 0.05
            if tap >= 0:
                m = torch.zeros(3, 3, device=dev)
                m.view(-1)[tap] = 1
                w = w * m
            ref = torch.nn.functional.conv2d(x.float(), w.bfloat16().float(), None, padding=1).permute(0, 2, 3, 1)
            debug.tune_rowshare(mode)
            out, _ = debug.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w)
            torch.cuda.synchronize()
            debug.tune_rowshare(0)
            err = ((out.float() - ref).norm() / ref.norm()).item()
            # per-column error profile for the all-taps case
            extra = ""
            if err > 1e-2:
                e = (out.float() - ref).abs().amax(dim=(0, 1, 3))  # per w
                bad = (e > 0.05 * ref.abs().max()).nonzero().flatten().tolist()
                extra = f" bad w: {bad[:12]}{'...' if len(bad) > 12 else ''} ({len(bad)} of {W})"
            print(f"H{H} W{W} C{C}->{Cout} mode {mode} tap {tap} (r={tap // 3 if tap >= 0 else '*'}, s={tap % 3 if tap >= 0 else '*'}): rel_l2 {err:.3e}{extra}", flush=True)
