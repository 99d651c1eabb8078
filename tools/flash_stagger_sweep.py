"""Sweep of the start-up stagger between the two co-resident CTAs of the flash kernels (FlashParams::stagger_ns)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (n, T, heads) in ((1, 6912, 5), (1, 1728, 10), (1, 432, 20), (2, 6912, 5)):
    d = heads * 64
    qkv = torch.randn(n, T, 3 * d, device=dev, generator=g).bfloat16()
    dout = torch.randn(n, T, d, device=dev, generator=g).bfloat16()
    base = None
    for ns in (0, 100, 200, 300, 400, 500, 700, 1000, 1500):
        os.environ["MDC_FLASH_STAGGER_NS"] = str(ns)
        o, dq, ms = debug.attention(qkv, heads, dout, iters=20)
        if base is None:
            base = o.float().clone()
        same = torch.equal(base, o.float())
        print(f"n{n} T{T} h{heads} stagger {ns:5d} ns: fwd {ms[0] * 1e3:7.1f} us  bwd {ms[1] * 1e3:7.1f} us  identical {same}", flush=True)
