"""Python handles on the kernel-level debug entry points (include/mdc_debug.h).

Used by tests/ and profiling scripts to run ONE kernel of the hot path on torch-owned device
buffers.  These wrappers only marshal pointers; all arithmetic happens in libmdc_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib, ptr


def _ll(v):
    return C.c_longlong(int(v))


def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, res=None, alpha=1.0, out_f32=False, bn=0, iters=0):
    """Batched D = alpha * A @ B^T (+bias) (+res).

    A: [nb1, nb0, M, K] bf16 (or [nb1, nb0, K, M] when a_mn), B: [nb1, nb0, N, K] (or [.., K, N] when b_mn).
    Returns (D [nb1, nb0, M, N], ms_per_launch).
    """
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and A.is_cuda and B.is_cuda
    A = A.contiguous()
    B = B.contiguous()
    nb1, nb0 = A.shape[0], A.shape[1]
    M, K = (A.shape[3], A.shape[2]) if a_mn else (A.shape[2], A.shape[3])
    N = B.shape[3] if b_mn else B.shape[2]
    out = torch.empty(nb1, nb0, M, N, device=A.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    ms = C.c_float(0)
    lda = A.shape[3]
    ldb = B.shape[3]
    if res is not None:
        res = res.contiguous()
        assert res.shape == out.shape and res.dtype == torch.bfloat16
    check(lib().mdc_dbg_gemm(
        C.c_int(M), C.c_int(N), C.c_int(K),
        ptr(A), C.c_int(int(a_mn)), _ll(lda), _ll(A.stride(1)), _ll(A.stride(0)),
        ptr(B), C.c_int(int(b_mn)), _ll(ldb), _ll(B.stride(1)), _ll(B.stride(0)),
        ptr(out), C.c_int(int(out_f32)), _ll(N), _ll(M * N), _ll(nb0 * M * N),
        ptr(bias), ptr(res), _ll(N), _ll(M * N), _ll(nb0 * M * N),
        C.c_float(alpha), C.c_int(nb0), C.c_int(nb1), C.c_int(bn), C.c_int(iters), C.byref(ms)))
    return out, ms.value


def conv3x3(x_nhwc, w_oihw, *, dgrad=False, bias=None, bias_img=None, res=None, iters=0):
    """3x3 / stride 1 / pad 1 convolution (or its input gradient) on an NHWC bf16 tensor.

    x_nhwc: [NB, H, W, Cx] bf16; w_oihw: [Cout, C, 3, 3] fp32.  Forward: Cx == C, result has Cout
    channels.  dgrad: Cx == Cout, result has C channels.
    """
    assert x_nhwc.dtype == torch.bfloat16 and w_oihw.dtype == torch.float32
    NB, H, W, Cx = x_nhwc.shape
    ldx = x_nhwc.stride(2)
    if x_nhwc.stride(3) != 1 or x_nhwc.stride(1) != W * ldx or x_nhwc.stride(0) != H * W * ldx or ldx % 8:
        pad = torch.zeros(NB, H, W, (Cx + 7) // 8 * 8, device=x_nhwc.device, dtype=torch.bfloat16)
        pad[..., :Cx] = x_nhwc
        x_nhwc = pad[..., :Cx]
    w_oihw = w_oihw.contiguous()
    Cout, Cc = w_oihw.shape[0], w_oihw.shape[1]
    Cres = Cc if dgrad else Cout
    assert Cx == (Cout if dgrad else Cc)
    ldo = (Cres + 7) // 8 * 8
    out = torch.zeros(NB, H, W, ldo, device=x_nhwc.device, dtype=torch.bfloat16)
    ms = C.c_float(0)
    if res is not None:
        res = res.contiguous()
    check(lib().mdc_dbg_conv3x3(
        C.c_int(NB), C.c_int(H), C.c_int(W), C.c_int(Cc), C.c_int(Cout), ptr(x_nhwc), _ll(x_nhwc.stride(2)),
        ptr(w_oihw), C.c_int(int(dgrad)), ptr(bias), ptr(bias_img), ptr(res),
        _ll(res.stride(2) if res is not None else 0), ptr(out), _ll(ldo), C.c_int(iters), C.byref(ms)))
    return out[..., :Cres], ms.value


def tune(bn=0, cs=0, ksplit=0, wcopies=1):
    """Planner overrides for tests / sweeps (include/mdc_debug.h: mdc_dbg_tune); call tune() to reset."""
    check(lib().mdc_dbg_tune(C.c_int(bn), C.c_int(cs), C.c_int(ksplit), C.c_int(wcopies)))


def tune_rowshare(mode=0):
    """0 automatic, 1 off, 2 force the row-shared-taps convolution mode wherever legal (mdc_dbg_tune_rowshare)."""
    check(lib().mdc_dbg_tune_rowshare(C.c_int(mode)))


def attention(qkv, heads, dout=None, iters=0):
    """Self-attention as the engine plans it (mdc_dbg_attention): head_dim 64 runs the fused tcgen05 flash kernels, other
    head dims GEMM + softmax + GEMM.  qkv: [n, T, 3 * heads * dh] bf16 (q | k | v).  Returns (o, dqkv or None, (ms_fwd,
    ms_bwd))."""
    assert qkv.dtype == torch.bfloat16 and qkv.is_cuda and qkv.is_contiguous() and qkv.ndim == 3
    n, T, c3 = qkv.shape
    d = c3 // 3
    dh = d // heads
    o = torch.empty(n, T, d, device=qkv.device, dtype=torch.bfloat16)
    dqkv = None
    if dout is not None:
        dout = dout.contiguous()
        assert dout.shape == o.shape and dout.dtype == torch.bfloat16
        dqkv = torch.zeros_like(qkv)
    ms = (C.c_float * 2)()
    check(lib().mdc_dbg_attention(C.c_int(n), C.c_int(T), C.c_int(heads), C.c_int(dh), ptr(qkv), ptr(o), ptr(dout), ptr(dqkv),
                                  C.c_int(iters), ms))
    return o, dqkv, (ms[0], ms[1])


def groupnorm(x_nhwc, gamma, beta, groups, eps, silu, dy=None, dx_init=None, mode=0, iters=0):
    """GroupNorm (+SiLU) forward / input-gradient backward with the engine's kernel selection (mdc_dbg_groupnorm).
    x_nhwc: [n, HW, C] bf16; mode 0 auto, 1 two-pass kernels, 2 single-launch kernels; dx_init: accumulate into it.
    Returns (y, dx or None, stats [n, groups, 2], (ms_fwd, ms_bwd))."""
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and x_nhwc.ndim == 3
    n, HW, Cc = x_nhwc.shape
    gamma, beta = gamma.float().contiguous(), beta.float().contiguous()
    y = torch.empty_like(x_nhwc)
    stats = torch.empty(n, groups, 2, device=x_nhwc.device, dtype=torch.float32)
    dx = None
    if dy is not None:
        dy = dy.contiguous()
        dx = dx_init.clone().contiguous() if dx_init is not None else torch.empty_like(x_nhwc)
    ms = (C.c_float * 2)()
    check(lib().mdc_dbg_groupnorm(C.c_int(n), C.c_int(HW), C.c_int(Cc), C.c_int(groups), C.c_float(eps), C.c_int(int(silu)),
                                  ptr(x_nhwc), ptr(gamma), ptr(beta), ptr(y), ptr(dy), ptr(dx), C.c_int(int(dx_init is not None)),
                                  C.c_int(mode), ptr(stats), C.c_int(iters), ms))
    return y, dx, stats, (ms[0], ms[1])


def upconv(x_nhwc, w_oihw, *, dgrad=False, bias=None, iters=0):
    """Fused nearest-2x upsample + conv3x3 (mdc_dbg_upconv).  Forward: x [NB, H, W, C] -> [NB, 2H, 2W, Cout]; dgrad: x is
    dy [NB, 2H, 2W, Cout] -> [NB, H, W, C].  Channel counts must be multiples of 8."""
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous() and w_oihw.dtype == torch.float32
    w_oihw = w_oihw.contiguous()
    Cout, Cc = w_oihw.shape[0], w_oihw.shape[1]
    NB, Hx, Wx, Cx = x_nhwc.shape
    H, W = (Hx // 2, Wx // 2) if dgrad else (Hx, Wx)
    Cres = Cc if dgrad else Cout
    assert Cx == (Cout if dgrad else Cc) and Cx % 8 == 0 and Cres % 8 == 0
    out = torch.zeros(NB, H if dgrad else 2 * H, W if dgrad else 2 * W, Cres, device=x_nhwc.device, dtype=torch.bfloat16)
    ms = C.c_float(0)
    check(lib().mdc_dbg_upconv(C.c_int(NB), C.c_int(H), C.c_int(W), C.c_int(Cc), C.c_int(Cout), ptr(x_nhwc), _ll(Cx), ptr(w_oihw),
                               C.c_int(int(dgrad)), ptr(bias), ptr(out), _ll(Cres), C.c_int(iters), C.byref(ms)))
    return out, ms.value
