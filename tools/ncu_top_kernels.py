"""The kernels that dominate a guided step, one launch each at the sizes of config b, for `ncu --set full` captures:
conv3x3 128->128 and 256->256 at 576x768 (forward, row-shared taps / CTA pairs), the 128-channel dgrad, flash attention
forward / backward at T = 6912 x 5 heads, GroupNorm+SiLU forward / backward on the 113 MB tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for cin, cout in ((128, 128), (256, 256)):
    x = torch.randn(1, 576, 768, cin, device=dev, generator=g).bfloat16()
    w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05
    y, ms = debug.conv3x3(x, w, iters=0)
    print(f"conv {cin}->{cout} fwd {ms * 1e3:.1f} us")
    y, ms = debug.conv3x3(x, w, dgrad=True, iters=0)
    print(f"conv {cin}->{cout} dgrad {ms * 1e3:.1f} us")
heads, T = 5, 6912
qkv = torch.randn(1, T, 3 * heads * 64, device=dev, generator=g).bfloat16()
dout = torch.randn(1, T, heads * 64, device=dev, generator=g).bfloat16()
o, dq, ms = debug.attention(qkv, heads, dout, iters=0)
print("attention", ms)
x = torch.randn(1, 576 * 768, 128, device=dev, generator=g).bfloat16()
dy = torch.randn(1, 576 * 768, 128, device=dev, generator=g).bfloat16()
gamma = torch.randn(128, device=dev, generator=g)
beta = torch.randn(128, device=dev, generator=g)
y, dx, st, ms = debug.groupnorm(x, gamma, beta, 32, 1e-6, True, dy=dy, iters=0)
print("groupnorm", ms)
