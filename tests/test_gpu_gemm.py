"""GPU parity of the tcgen05 GEMM / implicit-GEMM conv kernel against torch fp32 references.

bf16 inputs, fp32 accumulation: the reference is computed in fp32 from the same bf16-rounded inputs,
so the only differences are accumulation order and the final bf16 rounding of the output
(tolerance: 2^-7 relative to the output scale, stated per test).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel_err(got, ref):
    return ((got.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K,nb0,nb1", [(128, 64, 64, 1, 1), (300, 320, 200, 1, 1), (6912, 64, 1000, 5, 1),
                                           (257, 512, 512, 1, 2), (640, 256, 2560, 2, 2)])
def test_gemm_majors(cuda, a_mn, b_mn, M, N, K, nb0, nb1):
    from depth_completion_b200 import debug

    if (a_mn and M % 8) or (b_mn and N % 8) or K % 8:
        pytest.skip("TMA needs 16-byte row strides")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(nb1, nb0, M, K, device=cuda, generator=g).bfloat16()
    B = torch.randn(nb1, nb0, N, K, device=cuda, generator=g).bfloat16()
    ref = torch.matmul(A.float(), B.float().transpose(-1, -2)) * 0.125
    Ain = A.transpose(-1, -2).contiguous() if a_mn else A
    Bin = B.transpose(-1, -2).contiguous() if b_mn else B
    out, _ = debug.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, alpha=0.125, out_f32=True)
    torch.cuda.synchronize()
    err = _rel_err(out, ref)
    assert err < 1e-4, f"fp32-out rel err {err}"
    out16, _ = debug.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, alpha=0.125)
    assert _rel_err(out16, ref) < 2 ** -7


@pytest.mark.parametrize("M,N,K", [(1000, 320, 320), (433, 1280, 640), (128, 16, 64), (77, 24, 72)])
def test_gemm_epilogue(cuda, M, N, K):
    from depth_completion_b200 import debug

    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(1, 1, M, K, device=cuda, generator=g).bfloat16()
    B = torch.randn(1, 1, N, K, device=cuda, generator=g).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    res = torch.randn(1, 1, M, N, device=cuda, generator=g).bfloat16()
    ref = torch.matmul(A.float(), B.float().transpose(-1, -2)) + bias + res.float()
    out, _ = debug.gemm(A, B, bias=bias, res=res)
    torch.cuda.synchronize()
    assert _rel_err(out, ref) < 2 ** -7


@pytest.mark.parametrize("NB,H,W,C,Cout", [(1, 8, 16, 64, 64), (2, 9, 12, 128, 320), (1, 36, 48, 320, 320),
                                           (1, 15, 20, 192, 64), (1, 72, 96, 8, 320), (1, 24, 40, 128, 3),
                                           (1, 72, 96, 4, 128), (1, 11, 19, 64, 128),
                                           (1, 192, 256, 64, 128), (2, 160, 224, 128, 64)])  # single n-tile, many tiles: B-resident mode
@pytest.mark.parametrize("dgrad", [False, True])
def test_conv3x3(cuda, NB, H, W, C, Cout, dgrad):
    from depth_completion_b200 import debug

    g = torch.Generator(device="cuda").manual_seed(H * 31 + W)
    w = torch.randn(Cout, C, 3, 3, device=cuda, generator=g) * (1.0 / (3 * C ** 0.5))
    wq = w.bfloat16().float()
    if not dgrad:
        x = torch.randn(NB, C, H, W, device=cuda, generator=g).bfloat16()
        bias = torch.randn(Cout, device=cuda, generator=g)
        bias_img = torch.randn(NB, Cout, device=cuda, generator=g)
        res = torch.randn(NB, H, W, Cout, device=cuda, generator=g).bfloat16()
        ref = torch.nn.functional.conv2d(x.float(), wq, bias, padding=1) + bias_img[:, :, None, None]
        ref = ref.permute(0, 2, 3, 1) + res.float()
        out, _ = debug.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w, bias=bias, bias_img=bias_img, res=res)
    else:
        dy = torch.randn(NB, Cout, H, W, device=cuda, generator=g).bfloat16()
        ref = torch.nn.grad.conv2d_input((NB, C, H, W), wq, dy.float(), padding=1).permute(0, 2, 3, 1)
        out, _ = debug.conv3x3(dy.permute(0, 2, 3, 1).contiguous(), w, dgrad=True)
    torch.cuda.synchronize()
    err = _rel_err(out, ref)
    assert err < 2 ** -7, f"conv rel err {err}"


@pytest.mark.parametrize("cs,ksplit", [(1, 0), (2, 0), (3, 0), (4, 0), (1, 3), (2, 5), (1, -1), (2, -1)])
@pytest.mark.parametrize("H,W,C,Cout", [(18, 24, 640, 320), (9, 12, 1280, 256), (23, 17, 192, 160)])
def test_conv3x3_cluster_and_splitk_modes(cuda, H, W, C, Cout, cs, ksplit):
    """Every launch mode of the GEMM kernel (single CTA, CTA pair = cta_group::2 MMA, B multicast over 2 / 4 CTAs,
    forced and cost-model split-K) computes the same convolution; checked against torch fp32."""
    from depth_completion_b200 import debug

    g = torch.Generator(device="cuda").manual_seed(H * 131 + C)
    w = torch.randn(Cout, C, 3, 3, device=cuda, generator=g) * (1.0 / (3 * C ** 0.5))
    x = torch.randn(1, C, H, W, device=cuda, generator=g).bfloat16()
    bias = torch.randn(Cout, device=cuda, generator=g)
    res = torch.randn(1, H, W, Cout, device=cuda, generator=g).bfloat16()
    ref = torch.nn.functional.conv2d(x.float(), w.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    try:
        debug.tune(cs=cs, ksplit=ksplit)
        out, _ = debug.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w, bias=bias, res=res)
        torch.cuda.synchronize()
    finally:
        debug.tune()
    err = _rel_err(out, ref)
    assert err < 2 ** -7, f"conv rel err {err} (cs={cs}, ksplit={ksplit})"


@pytest.mark.parametrize("NB,H,W,C,Cout", [(1, 6, 256, 64, 128), (2, 5, 300, 128, 128), (1, 9, 128, 256, 64), (1, 4, 517, 192, 96),
                                           (1, 3, 384, 128, 320)])  # the last one: BN > 128, forced anyway
@pytest.mark.parametrize("dgrad", [False, True])
def test_conv3x3_rowshare(cuda, NB, H, W, C, Cout, dgrad):
    """Row-shared-taps mode (GemmParams::rowshare): one 130-pixel activation box per kernel row, the three taps as MMAs
    over row-offset views of it (matrix base offset 0 / 1 / 2).  Same answer as the tap-by-tap mode and as torch fp32,
    including the zero padding at both image borders and partial tiles (W not a multiple of 128)."""
    from depth_completion_b200 import debug

    g = torch.Generator(device="cuda").manual_seed(W * 13 + C)
    w = torch.randn(Cout, C, 3, 3, device=cuda, generator=g) * (1.0 / (3 * C ** 0.5))
    wq = w.bfloat16().float()
    try:
        debug.tune_rowshare(2)
        if not dgrad:
            x = torch.randn(NB, C, H, W, device=cuda, generator=g).bfloat16()
            bias = torch.randn(Cout, device=cuda, generator=g)
            res = torch.randn(NB, H, W, Cout, device=cuda, generator=g).bfloat16()
            ref = torch.nn.functional.conv2d(x.float(), wq, bias, padding=1).permute(0, 2, 3, 1) + res.float()
            out, _ = debug.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w, bias=bias, res=res)
        else:
            dy = torch.randn(NB, Cout, H, W, device=cuda, generator=g).bfloat16()
            ref = torch.nn.grad.conv2d_input((NB, C, H, W), wq, dy.float(), padding=1).permute(0, 2, 3, 1)
            out, _ = debug.conv3x3(dy.permute(0, 2, 3, 1).contiguous(), w, dgrad=True)
        torch.cuda.synchronize()
    finally:
        debug.tune_rowshare(0)
    err = _rel_err(out, ref)
    assert err < 2 ** -7, f"rowshare conv rel err {err}"


def test_conv3x3_rowshare_speed(cuda):
    """The decoder's 128-channel convolutions at 576x768 (14 launches per guided step): time with and without the mode."""
    from depth_completion_b200 import debug

    g = torch.Generator(device="cuda").manual_seed(1)
    w = torch.randn(128, 128, 3, 3, device=cuda, generator=g) * 0.03
    x = torch.randn(1, 576, 768, 128, device=cuda, generator=g).bfloat16()
    out = {}
    try:
        for mode in (1, 2):
            debug.tune_rowshare(mode)
            o, ms = debug.conv3x3(x, w, iters=20)
            out[mode] = (o.float().clone(), ms)
    finally:
        debug.tune_rowshare(0)
    gf = 2.0 * 576 * 768 * 9 * 128 * 128 / 1e9
    print(f"[measured] conv 128->128 @576x768: tap-by-tap {out[1][1] * 1e3:.1f} us ({gf / out[1][1]:.0f} TF/s), "
          f"row-shared {out[2][1] * 1e3:.1f} us ({gf / out[2][1]:.0f} TF/s)")
    # same products, another summation order: at most a bf16 rounding step apart
    assert ((out[1][0] - out[2][0]).abs().max() / out[1][0].abs().max()).item() < 2 ** -7
