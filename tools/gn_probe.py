"""Runs the two-pass GroupNorm kernels on a decoder-sized tensor (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
C = int(os.environ.get("C", "128"))
H, W = (576, 768)
x = torch.randn(1, H * W, C, device=dev, generator=g).bfloat16()
dy = torch.randn(1, H * W, C, device=dev, generator=g).bfloat16()
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for it in range(3):
    flush.zero_()
    y, dx, st, ms = debug.groupnorm(x, gamma, beta, 32, 1e-6, 1, dy, mode=1, iters=0)
torch.cuda.synchronize()
y, dx, st, ms = debug.groupnorm(x, gamma, beta, 32, 1e-6, 1, dy, mode=1, iters=10)
print(f"C={C}: fwd {ms[0]*1e3:.1f} us bwd {ms[1]*1e3:.1f} us")
