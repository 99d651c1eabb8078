"""Epilogue / fixed-cost probe: GEMMs with tiny K so the kernel time is launch + fill + epilogue."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug
dev = torch.device("cuda:0")
debug.tune(cs=int(os.environ.get("CS", "1")))
for (M, N, K) in [(128 * 148, 256, 64), (128 * 148, 512, 64), (128 * 148, 1024, 64), (128 * 148, 2048, 64), (128 * 148, 256, 1024), (128 * 148, 512, 1024),
                  (128 * 148, 128, 64), (128 * 148, 64, 64), (128 * 148, 32, 64), (6912, 960, 64), (6912, 960, 320), (6912, 960, 1280)]:
    A = torch.randn(1, 1, M, K, device=dev).bfloat16()
    B = torch.randn(1, 1, N, K, device=dev).bfloat16()
    for f32 in (False, True):
        y, ms = debug.gemm(A, B, out_f32=f32, iters=50)
        print(f"M{M} N{N} K{K} f32={int(f32)}: {ms*1e3:7.1f} us", flush=True)
