"""Per-kernel-class comparison on one B200: this library's kernels vs the kernels PyTorch runs for the same op in bf16
(cuDNN benchmark mode, cuBLAS, SDPA, ATen GroupNorm) -- the kernel-class bar of SURVEY.md section 2.1.
Writes a markdown table (argv[1], default profiles/r02_kernel_classes.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from depth_completion_b200 import debug  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)


def time_torch(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3  # us


rows = []


def conv_case(C, Cout, H, W):
    x = torch.randn(1, C, H, W, device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, C, 3, 3, device=dev, generator=g) * 0.03)
    w16 = w.bfloat16().contiguous(memory_format=torch.channels_last)
    dy = torch.randn(1, Cout, H, W, device=dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    t_f = time_torch(lambda: F.conv2d(x, w16, None, padding=1))
    t_b = time_torch(lambda: torch.nn.grad.conv2d_input(x.shape, w16, dy, padding=1))
    xn = x.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    small = H * W < 4096  # low-resolution UNet layers: the engine runs these split-K, weights stream from HBM (not L2)
    if small:
        debug.tune(ksplit=-1, wcopies=8)
    _, o_f = debug.conv3x3(xn, w, iters=16 if small else 10)
    _, o_b = debug.conv3x3(dyn, w, dgrad=True, iters=16 if small else 10)
    if small:
        debug.tune()
    gf = 2.0 * H * W * 9 * C * Cout / 1e9
    rows.append((f"conv3x3 {C}->{Cout} @{H}x{W} fwd", o_f * 1e3, t_f, gf))
    rows.append((f"conv3x3 {C}->{Cout} @{H}x{W} dgrad", o_b * 1e3, t_b, gf))


def attn_case(T, heads, dh):
    d = heads * dh
    qkv = (torch.randn(1, T, 3 * d, device=dev, generator=g) * 1.0).bfloat16()
    dout = torch.randn(1, T, d, device=dev, generator=g).bfloat16()
    _, _, ms = debug.attention(qkv, heads, dout, iters=10)
    x = qkv.clone().requires_grad_(True)

    def fwd():
        q, k, v = (x[..., i * d:(i + 1) * d].view(1, T, heads, dh).transpose(1, 2) for i in range(3))
        return F.scaled_dot_product_attention(q, k, v)

    t_f = time_torch(lambda: fwd())
    o = fwd()
    do = dout.view(1, T, heads, dh).transpose(1, 2)
    t_b = time_torch(lambda: torch.autograd.grad(o, x, do, retain_graph=True))
    gf = 4.0 * T * T * dh * heads / 1e9
    rows.append((f"attention T={T} heads={heads} d={dh} fwd", ms[0] * 1e3, t_f, gf))
    rows.append((f"attention T={T} heads={heads} d={dh} bwd", ms[1] * 1e3, t_b, 2.5 * gf))


def gn_case(C, H, W):
    x = torch.randn(1, H * W, C, device=dev, generator=g).bfloat16()
    dy = torch.randn(1, H * W, C, device=dev, generator=g).bfloat16()
    ga, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    _, _, _, ms = debug.groupnorm(x, ga, be, 32, 1e-6, 1, dy, mode=0, iters=10)
    xt = x.view(1, H, W, C).permute(0, 3, 1, 2).requires_grad_(True)  # channels-last NCHW view
    ga16, be16 = ga.bfloat16(), be.bfloat16()
    fwd = lambda: F.silu(F.group_norm(xt, 32, ga16, be16, 1e-6))
    t_f = time_torch(lambda: fwd())
    y = fwd()
    dyt = dy.view(1, H, W, C).permute(0, 3, 1, 2)
    t_b = time_torch(lambda: torch.autograd.grad(y, xt, dyt, retain_graph=True))
    mb = x.numel() * 2 / 1e6
    rows.append((f"GroupNorm+SiLU {C}ch @{H}x{W} ({mb:.0f} MB) fwd", ms[0] * 1e3, t_f, 0.0))
    rows.append((f"GroupNorm+SiLU {C}ch @{H}x{W} ({mb:.0f} MB) bwd", ms[1] * 1e3, t_b, 0.0))


conv_case(256, 256, 576, 768)
conv_case(128, 128, 576, 768)
conv_case(512, 512, 288, 384)
conv_case(1280, 1280, 18, 24)
attn_case(6912, 5, 64)
attn_case(1728, 10, 64)
attn_case(6912, 1, 512)
gn_case(256, 576, 768)
gn_case(128, 576, 768)
gn_case(320, 72, 96)
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_kernel_classes.md")
with open(out_path, "w") as f:
    f.write("# Kernel classes: this library vs PyTorch's kernels for the same op, bf16, one B200\n\n"
            "`python tools/kernel_class_table.py`.  Ours: back-to-back launches timed by the library (CUDA events, warm L2); torch: mean of 10\n"
            "calls with an L2 flush before each (cudnn.benchmark=True, channels_last for the convolutions; SDPA picks its own backend).\n"
            "The torch timings therefore include cold-cache effects ours do not: read the table for the class, not the last 10 %.\n\n"
            "| op | ours us | torch us | ours TF/s | torch TF/s |\n|---|---|---|---|---|\n")
    for name, o, t, gf in rows:
        f.write(f"| {name} | {o:.1f} | {t:.1f} | {gf / o * 1e3 if gf else 0:.0f} | {gf / t * 1e3 if gf else 0:.0f} |\n")
print(open(out_path).read())
