"""ncu target: three representative launches of umma_gemm_kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
w = torch.randn(128, 128, 3, 3, device=dev) / 34
x = torch.randn(1, 576, 768, 128, device=dev).bfloat16()
out, ms = debug.conv3x3(x, w, iters=2); print("conv128", ms)
A = torch.randn(1, 1, 6912, 320, device=dev).bfloat16(); B = torch.randn(1, 1, 2560, 320, device=dev).bfloat16()
out, ms = debug.gemm(A, B, iters=2); print("gemm k320", ms)
w = torch.randn(256, 256, 3, 3, device=dev) / 48
x = torch.randn(1, 576, 768, 256, device=dev).bfloat16()
out, ms = debug.conv3x3(x, w, iters=2); print("conv256", ms)
torch.cuda.synchronize()
