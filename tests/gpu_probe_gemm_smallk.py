"""ncu target: the short-K linear-layer GEMMs of the UNet transformer blocks (epilogue-dominated)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
debug.tune(cs=int(os.environ.get("CS", "1")))
for (M, N, K) in [(6912, 960, 320), (6912, 2560, 320), (6912, 320, 1280)]:
    A = torch.randn(1, 1, M, K, device=dev).bfloat16()
    B = torch.randn(1, 1, N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    y, ms = debug.gemm(A, B, bias=bias, iters=3)
    print(M, N, K, ms * 1e3, "us")
torch.cuda.synchronize()
