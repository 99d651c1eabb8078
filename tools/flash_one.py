"""One forward + backward of the level-0 self-attention (T = 6912, 5 heads) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
heads, T = 5, 6912
qkv = torch.randn(1, T, 3 * heads * 64, device=dev, generator=g).bfloat16()
dout = torch.randn(1, T, heads * 64, device=dev, generator=g).bfloat16()
o, dq, ms = debug.attention(qkv, heads, dout, iters=int(os.environ.get("ITERS", "3")))
print(ms)
