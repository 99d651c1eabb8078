"""Kernel-level GPU parity at the sizes the benchmark actually runs (BASELINE.json config b: latent 72x96, image 576x768;
res-640: T = 4800; KITTI-shaped config c: T = 6688), through the C ABI's debug entry points.

The model-level tests use a narrow UNet on 96x128 frames, where the engine's selectors pick the single-launch GroupNorm,
flash attention sees at most 3 key tiles and the fused upsample-conv only runs at toy sizes.  These tests put the kernels
that dominate a full-size guided step under the gated suite:
  * fused flash attention head_dim 64, forward + backward (flash.cuh) at T in {6912, 4800, 6688, 1728}, 5 / 10 heads,
    batch 2, including a case whose row maxima keep growing along the keys (the lazy-rescale path);
  * the unfused head_dim-512 attention of the VAE mid block at T = 6912;
  * the two-pass GroupNorm(+SiLU) kernels at 128 / 256 ch x 576x768 and 512 ch x 288x384 (56 - 226 MB tensors), and the
    single-launch variant at a UNet size, with and without gradient accumulation;
  * the fused nearest-2x-upsample + conv3x3 and its input gradient at 512 ch @ 144x192.
Reference: torch fp32 on the same bf16-rounded inputs (F.scaled_dot_product_attention / F.group_norm / F.conv2d +
autograd); yardstick: the same op run by torch in bf16 (what the reference's bf16 mode executes).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(got, ref):
    return ((got.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-20)).item()


def _sdpa_ref(qkv, heads, dout, dtype):
    n, T, c3 = qkv.shape
    d = c3 // 3
    dh = d // heads
    x = qkv.to(dtype).detach().clone().requires_grad_(True)
    q, k, v = (x[..., i * d:(i + 1) * d].view(n, T, heads, dh).transpose(1, 2) for i in range(3))
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, T, d)
    o.backward(dout.to(dtype))
    return o.detach(), x.grad.detach()


def _check(name, ours, t16, ref, cap, slack=2e-3, factor=1.35):
    e_ours, e_16 = rel_l2(ours, ref), rel_l2(t16, ref)
    assert e_ours < cap, f"{name}: kernel rel_l2 {e_ours:.3e} (torch-bf16 {e_16:.3e})"
    assert e_ours <= factor * e_16 + slack, f"{name}: kernel {e_ours:.3e} vs torch-bf16 {e_16:.3e}"


@pytest.mark.parametrize("n,T,heads,ramp", [(1, 6912, 5, 0.0),    # config (b) level 0: 72x96 tokens, 5 heads
                                            (1, 6912, 5, 6.0),    # ... with key norms growing along the sequence (lazy rescale)
                                            (2, 4800, 5, 0.0),    # res-640 level 0 (60x80), batch 2; T % 128 = 64
                                            (1, 6688, 5, 0.0),    # config (c) level 0 (44x152); T % 128 = 32
                                            (2, 1728, 10, 2.0),   # level 1 (36x48), 10 heads, batch 2; T % 128 = 64
                                            (1, 432, 20, 0.0),    # level 2 (18x24), 20 heads
                                            (2, 108, 20, 0.0)])   # mid block (9x12): a single partial query tile
def test_flash_attention_fullsize(cuda, n, T, heads, ramp):
    from depth_completion_b200 import debug

    g = torch.Generator(device=cuda).manual_seed(T + heads)
    d = heads * 64
    qkv = torch.randn(n, T, 3 * d, device=cuda, generator=g) * 1.2
    if ramp:  # keys later in the sequence score higher: every query's running maximum grows by many powers of two
        t = torch.linspace(0.0, 1.0, T, device=cuda).view(1, T, 1)
        qkv[..., d:2 * d] *= 1.0 + ramp * t
    qkv = qkv.bfloat16()
    dout = torch.randn(n, T, d, device=cuda, generator=g).bfloat16()
    o, dqkv, _ = debug.attention(qkv, heads, dout)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all() and torch.isfinite(dqkv.float()).all()
    o32, g32 = _sdpa_ref(qkv, heads, dout, torch.float32)
    o16, g16 = _sdpa_ref(qkv, heads, dout, torch.bfloat16)
    _check("flash fwd", o, o16, o32, cap=2e-2)
    for i, nm in enumerate(("dq", "dk", "dv")):
        sl = slice(i * d, (i + 1) * d)
        _check(f"flash {nm}", dqkv[..., sl], g16[..., sl], g32[..., sl], cap=5e-2)


def test_vae_mid_attention_fullsize(cuda):
    """head_dim 512, one head, T = 6912 (the VAE decoder's mid block at config b)."""
    from depth_completion_b200 import debug

    g = torch.Generator(device=cuda).manual_seed(512)
    n, T, d = 1, 6912, 512
    qkv = (torch.randn(n, T, 3 * d, device=cuda, generator=g) * 0.7).bfloat16()
    dout = torch.randn(n, T, d, device=cuda, generator=g).bfloat16()
    o, dqkv, _ = debug.attention(qkv, 1, dout)
    torch.cuda.synchronize()
    o32, g32 = _sdpa_ref(qkv, 1, dout, torch.float32)
    o16, g16 = _sdpa_ref(qkv, 1, dout, torch.bfloat16)
    _check("d512 fwd", o, o16, o32, cap=2e-2)
    for i, nm in enumerate(("dq", "dk", "dv")):
        sl = slice(i * d, (i + 1) * d)
        _check(f"d512 {nm}", dqkv[..., sl], g16[..., sl], g32[..., sl], cap=5e-2)


def _gn_ref(x_nhwc, gamma, beta, groups, eps, silu, dy, dtype):
    n, HW, C = x_nhwc.shape
    x = x_nhwc.to(dtype).transpose(1, 2).reshape(n, C, HW, 1).detach().clone().requires_grad_(True)
    y = F.group_norm(x, groups, gamma.to(dtype), beta.to(dtype), eps)
    if silu:
        y = F.silu(y)
    y.backward(dy.to(dtype).transpose(1, 2).reshape(n, C, HW, 1))
    back = lambda t: t.detach().reshape(n, C, HW).transpose(1, 2)
    return back(y), back(x.grad)


@pytest.mark.parametrize("n,H,W,C,mode,silu,eps", [
    (1, 576, 768, 128, 1, 1, 1e-6),   # decoder up_blocks.3: 113 MB, two-pass kernels
    (1, 576, 768, 256, 1, 1, 1e-6),   # up_blocks.3.resnets.0.norm1: 226 MB
    (1, 288, 384, 512, 1, 1, 1e-6),   # up_blocks.2.resnets.0.norm1: 113 MB
    (2, 288, 384, 256, 0, 1, 1e-6),   # batch 2, automatic selection (two-pass at this size)
    (1, 72, 96, 512, 0, 0, 1e-6),     # mid-block attention norm (no SiLU), automatic = single launch
    (1, 72, 96, 320, 2, 1, 1e-5),     # UNet level 0, single-launch (cluster) kernels required
    (1, 72, 96, 320, 3, 1, 1e-5),     # ... the grid-barrier variant of the same (kept for comparison)
    (1, 72, 96, 960, 2, 1, 1e-5),     # UNet up path concat at level 0 (13 MB): cluster of 8 per group
    (1, 72, 96, 960, 3, 1, 1e-5),
    (1, 36, 48, 640, 2, 1, 1e-5),     # level 1
    (1, 18, 24, 1280, 2, 1, 1e-5),    # level 2
    (1, 18, 24, 1280, 3, 1, 1e-5),
    (1, 9, 12, 2560, 2, 1, 1e-5),     # level 3 concat: 80 channels per group (lanes stride the pairs)
    (1, 9, 12, 2560, 3, 1, 1e-5),
    (2, 36, 48, 1920, 2, 1, 1e-5),    # UNet up path concat width, batch 2
    (1, 144, 192, 512, 0, 1, 1e-6),   # decoder up_blocks.1 (28 MB): cluster forward, two-pass backward
    (1, 44, 152, 960, 0, 1, 1e-5)])   # KITTI-shaped level 0 concat
def test_groupnorm_fullsize(cuda, n, H, W, C, mode, silu, eps):
    from depth_completion_b200 import debug

    g = torch.Generator(device=cuda).manual_seed(C + H)
    HW = H * W
    # per-channel offsets and scales so that the statistics are not trivially (0, 1)
    x = torch.randn(n, HW, C, device=cuda, generator=g) * (0.5 + torch.rand(C, device=cuda, generator=g)) + torch.randn(C, device=cuda, generator=g)
    x = x.bfloat16()
    gamma = 1.0 + 0.3 * torch.randn(C, device=cuda, generator=g)
    beta = 0.2 * torch.randn(C, device=cuda, generator=g)
    dy = torch.randn(n, HW, C, device=cuda, generator=g).bfloat16()
    y, dx, stats, ms = debug.groupnorm(x, gamma, beta, 32, eps, silu, dy, mode=mode, iters=5)
    torch.cuda.synchronize()
    mb = x.numel() * 2 / 1e6
    print(f"[measured] GroupNorm {n}x{H}x{W}x{C} mode {mode} ({mb:.0f} MB): fwd {ms[0] * 1e3:.1f} us, bwd {ms[1] * 1e3:.1f} us")
    y32, dx32 = _gn_ref(x, gamma, beta, 32, eps, silu, dy, torch.float32)
    y16, dx16 = _gn_ref(x, gamma, beta, 32, eps, silu, dy, torch.bfloat16)
    _check("groupnorm fwd", y, y16, y32, cap=6e-3, slack=1.5e-3)
    _check("groupnorm bwd", dx, dx16, dx32, cap=1.2e-2, slack=3e-3)
    # statistics against torch in fp64
    xf = x.double().view(n, HW, 32, C // 32)
    mean, var = xf.mean(dim=(1, 3)), xf.var(dim=(1, 3), unbiased=False)
    assert (stats[..., 0].double() - mean).abs().max().item() < 1e-4 * (1 + mean.abs().max().item())
    assert ((stats[..., 1].double() - (var + eps).rsqrt()).abs() / (var + eps).rsqrt()).max().item() < 1e-4
    if n == 1 and C <= 256:  # gradient accumulation at fan-out points (acc flag)
        base = torch.randn(n, HW, C, device=cuda, generator=g).bfloat16()
        _, dx_acc, _, _ = debug.groupnorm(x, gamma, beta, 32, eps, silu, dy, dx_init=base, mode=mode)
        assert rel_l2(dx_acc, base.float() + dx32) < 8e-3


@pytest.mark.parametrize("NB,H,W,C", [(1, 144, 192, 512),   # decoder up_blocks.1.upsamplers.0 -> 288x384 (521.8 GF)
                                      (2, 36, 48, 640),     # UNet up_blocks.2.upsamplers.0, batch 2
                                      (1, 22, 76, 256)])    # KITTI-shaped aspect ratio, partial tiles
def test_fused_upsample_conv_fullsize(cuda, NB, H, W, C):
    from depth_completion_b200 import debug

    g = torch.Generator(device=cuda).manual_seed(H * 7 + C)
    w = torch.randn(C, C, 3, 3, device=cuda, generator=g) * (1.0 / (3 * C ** 0.5))
    wq = w.bfloat16().float()
    bias = torch.randn(C, device=cuda, generator=g)
    x = torch.randn(NB, C, H, W, device=cuda, generator=g).bfloat16()
    ref = F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), wq, bias, padding=1).permute(0, 2, 3, 1)
    out, _ = debug.upconv(x.permute(0, 2, 3, 1).contiguous(), w, bias=bias)
    torch.cuda.synchronize()
    # the kernel sums the 3x3 taps that fall on the same low-resolution pixel BEFORE rounding the weights to bf16
    assert rel_l2(out, ref) < 4e-3
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 2 ** -6
    dy = torch.randn(NB, C, 2 * H, 2 * W, device=cuda, generator=g).bfloat16()
    xr = x.float().detach().clone().requires_grad_(True)
    F.conv2d(F.interpolate(xr, scale_factor=2.0, mode="nearest"), wq, None, padding=1).backward(dy.float())
    dref = xr.grad.permute(0, 2, 3, 1)
    din, _ = debug.upconv(dy.permute(0, 2, 3, 1).contiguous(), w, dgrad=True)
    torch.cuda.synchronize()
    assert rel_l2(din, dref) < 4e-3
    assert ((din.float() - dref).abs().max() / dref.abs().max()).item() < 2 ** -6


@pytest.mark.parametrize("H,W,C,ksplit", [(9, 12, 1280, -1), (18, 24, 1280, -1), (9, 12, 1280, 7)])
def test_fused_upsample_conv_dgrad_splitk(cuda, H, W, C, ksplit):
    """The UNet's upsamplers at 9x12 / 18x24: one or four output tiles and 320 k-chunks, so the input gradient runs
    split-K (cost model: ksplit = -1) -- round 1 ran these two launches unsplit at 100-118 us each."""
    from depth_completion_b200 import debug

    g = torch.Generator(device=cuda).manual_seed(H + C)
    w = torch.randn(C, C, 3, 3, device=cuda, generator=g) * (1.0 / (3 * C ** 0.5))
    dy = torch.randn(1, C, 2 * H, 2 * W, device=cuda, generator=g).bfloat16()
    xr = torch.zeros(1, C, H, W, device=cuda, requires_grad=True)
    F.conv2d(F.interpolate(xr, scale_factor=2.0, mode="nearest"), w.bfloat16().float(), None, padding=1).backward(dy.float())
    dref = xr.grad.permute(0, 2, 3, 1)
    try:
        debug.tune(ksplit=ksplit)
        din, ms = debug.upconv(dy.permute(0, 2, 3, 1).contiguous(), w, dgrad=True, iters=5)
        torch.cuda.synchronize()
    finally:
        debug.tune()
    print(f"[measured] upconv dgrad {C}@{H}x{W} ksplit {ksplit}: {ms * 1e3:.1f} us")
    assert rel_l2(din, dref) < 4e-3
