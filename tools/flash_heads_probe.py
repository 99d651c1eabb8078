"""Flash kernel time against the number of CTAs per SM (heads x 54 query tiles at T = 6912) and at the other UNet levels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (n, T, heads) in ((1, 6912, 1), (1, 6912, 2), (1, 6912, 3), (1, 6912, 5), (1, 6912, 8), (1, 1728, 10), (1, 432, 20), (1, 108, 20)):
    d = heads * 64
    qkv = torch.randn(n, T, 3 * d, device=dev, generator=g).bfloat16()
    dout = torch.randn(n, T, d, device=dev, generator=g).bfloat16()
    o, dq, ms = debug.attention(qkv, heads, dout, iters=20)
    print(f"T{T} h{heads} CTAs {heads * ((T + 127) // 128)}: fwd {ms[0]*1e3:7.1f} us  bwd {ms[1]*1e3:7.1f} us", flush=True)
