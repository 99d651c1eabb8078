"""Manual probe (not collected by pytest): prints errors and TFLOP/s for a few GEMM/conv shapes."""
import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_completion_b200 import debug

dev = torch.device("cuda:0")
def rel(got, ref): return ((got.float() - ref).abs().max() / ref.abs().max()).item()

def try_gemm(M, N, K, a_mn, b_mn, nb0=1, iters=0):
    try:
        A = torch.randn(1, nb0, M, K, device=dev).bfloat16(); B = torch.randn(1, nb0, N, K, device=dev).bfloat16()
        ref = A.float() @ B.float().transpose(-1, -2)
        Ain = A.transpose(-1, -2).contiguous() if a_mn else A
        Bin = B.transpose(-1, -2).contiguous() if b_mn else B
        out, ms = debug.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_f32=True, iters=iters)
        torch.cuda.synchronize()
        tf = 2.0 * M * N * K * nb0 / (ms * 1e-3) / 1e12 if ms > 0 else 0
        print(f"gemm M{M} N{N} K{K} a_mn{int(a_mn)} b_mn{int(b_mn)} nb{nb0}: rel_err {rel(out, ref):.3e}  {ms:.4f} ms  {tf:.1f} TF/s", flush=True)
    except Exception as e:
        print("gemm FAILED", M, N, K, a_mn, b_mn, repr(e), flush=True); traceback.print_exc()

def try_conv(NB, H, W, C, Cout, dgrad=False, iters=0):
    try:
        w = torch.randn(Cout, C, 3, 3, device=dev) / (3 * C ** 0.5)
        Cx = Cout if dgrad else C
        x = torch.randn(NB, Cx, H, W, device=dev).bfloat16()
        if dgrad:
            ref = torch.nn.grad.conv2d_input((NB, C, H, W), w.bfloat16().float(), x.float(), padding=1).permute(0, 2, 3, 1)
        else:
            ref = torch.nn.functional.conv2d(x.float(), w.bfloat16().float(), padding=1).permute(0, 2, 3, 1)
        out, ms = debug.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w, dgrad=dgrad, iters=iters)
        torch.cuda.synchronize()
        tf = 2.0 * NB * H * W * 9 * C * Cout / (ms * 1e-3) / 1e12 if ms > 0 else 0
        print(f"conv NB{NB} {H}x{W} C{C}->{Cout} dgrad{int(dgrad)}: rel_err {rel(out, ref):.3e}  {ms:.4f} ms  {tf:.1f} TF/s", flush=True)
    except Exception as e:
        print("conv FAILED", NB, H, W, C, Cout, repr(e), flush=True); traceback.print_exc()

try_gemm(128, 64, 64, False, False)
try_gemm(256, 256, 512, False, False)
try_gemm(256, 256, 512, False, True)
try_gemm(256, 256, 512, True, False)
try_gemm(256, 256, 512, True, True)
try_gemm(8192, 8192, 8192, False, False, iters=5)
try_gemm(6912, 2560, 320, False, False, iters=10)
try_gemm(6912, 6912, 64, False, False, nb0=5, iters=10)
try_gemm(6912, 64, 6912, False, True, nb0=5, iters=10)
try_conv(1, 8, 16, 64, 64)
try_conv(1, 72, 96, 320, 320, iters=10)
try_conv(1, 72, 96, 320, 320, dgrad=True, iters=10)
try_conv(1, 576, 768, 128, 128, iters=5)
try_conv(1, 576, 768, 256, 256, iters=5)
try_conv(1, 288, 384, 512, 512, iters=5)
try_conv(1, 9, 12, 1280, 1280, iters=10)
