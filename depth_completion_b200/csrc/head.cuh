// Sparse output head of the KL decoder inside a guided step.
//
// The loss of a guided step (marigold_dc.py:829-840, default loss_funcs l1 + l2) reads the decoded map only at the four
// bilinear taps of every valid sparse point (marigold_dc.py:366-370) -- a few thousand of the 442 368 decoder pixels of
// BASELINE config b -- and its gradient is non-zero only there.  The last two decoder layers (conv_norm_out + SiLU,
// conv_out: vae.decoder) therefore do not have to be evaluated densely inside the loop:
//
//   forward   GroupNorm statistics of the last feature map (dense, one read), then conv_out(silu(norm(x))) ONLY at the
//             tap pixels (head_points_fwd_kernel: one warp per tap, 9 x C products) -- the 113 MB normalised map is
//             neither written nor read, the dense 128 -> 3 convolution disappears;
//   backward  d(conv_out input) is non-zero only in the 3x3 neighbourhoods of the tap pixels, so the input gradient of
//             conv_out is evaluated on the fly where some neighbour carries a gradient (never materialised), the
//             GroupNorm backward sums run over those pixels only (gn_bwd_stats_sp_kernel) and the dense pass
//             dx = rstd (dxhat - m1 - xhat m2) reads x and writes dx without reading a dy tensor (gn_bwd_apply_sp_kernel).
//
// Arithmetic and rounding points are those of the dense kernels (gn_apply_s_kernel, the conv epilogue, the dgrad GEMM's
// bf16 output, gn_bwd_*): values agree to fp32 summation order.  The dense path stays for the final decode, the dense
// "edge" / "smooth" losses and every debug entry point (Engine::sparse_head_allowed).
#pragma once
#include "kernels.cuh"
#include "tail.cuh"

namespace mdc {

// conv_out(silu(groupnorm(x))) at the bilinear tap pixels of the valid points.  x [N, PPH*PPW, ldx] bf16 (the GroupNorm
// input), stats (mean, rstd) per (image, group), wf the forward pack [3][9 * Cp] (k = tap * Cp + c, tap = r * 3 + s),
// dec [N, PPH*PPW, ld_dec] receives channels 0..2 of the tap pixels (duplicated taps write identical values).
__global__ void __launch_bounds__(256) head_points_fwd_kernel(const bf16* __restrict__ x, long long ldx, int C, int G, int Cp,
                                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int silu, const bf16* __restrict__ wf,
                                                              const float* __restrict__ bias, TailGeom g, const int* __restrict__ pt_idx,
                                                              const int* __restrict__ pt_off, bf16* __restrict__ dec) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = pt_off[g.N];
  const int cpg = C / G;
  const float sy = static_cast<float>(g.ph) / g.H, sx = static_cast<float>(g.pw) / g.W;
  for (int slot = warp; slot < 4 * total; slot += nwarps) {
    const int i = slot >> 2, k = slot & 3;
    int n = 0;
    while (n + 1 < g.N && i >= pt_off[n + 1]) ++n;
    const int pix = pt_idx[i];
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(pix / g.W, sy, g.ph, y0, y1, ly, g.nearest);
    bilinear_src(pix % g.W, sx, g.pw, x0, x1, lx, g.nearest);
    const int yy = (k & 2) ? y1 : y0, xx = (k & 1) ? x1 : x0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int c = lane * 2; c < C; c += 64) {
      const int gi = c / cpg;
      const float mean = stats[2 * (n * G + gi)], rstd = stats[2 * (n * G + gi) + 1];
      const float2 sc = make_float2(rstd * gamma[c], rstd * gamma[c + 1]);
      const float2 sf = make_float2(beta[c] - mean * sc.x, beta[c + 1] - mean * sc.y);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int y2 = yy + t / 3 - 1, x2 = xx + t % 3 - 1;
        if (y2 < 0 || y2 >= g.PPH || x2 < 0 || x2 >= g.PPW) continue;  // zero padding of the convolution input
        const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(x + ((1LL * n * g.PPH + y2) * g.PPW + x2) * ldx + c);
        float2 h = __ffma2_rn(__bfloat1622float2(xv), sc, sf);
        if (silu) h = silu2(bf16r2(h));
        h = bf16r2(h);  // the dense path stores the normalised map in bf16
        const float2 w0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wf + (0 * 9 + t) * Cp + c));
        const float2 w1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wf + (1 * 9 + t) * Cp + c));
        const float2 w2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(wf + (2 * 9 + t) * Cp + c));
        a0 = fmaf(h.x, w0.x, fmaf(h.y, w0.y, a0));
        a1 = fmaf(h.x, w1.x, fmaf(h.y, w1.y, a1));
        a2 = fmaf(h.x, w2.x, fmaf(h.y, w2.y, a2));
      }
    }
    a0 = warp_sum(a0), a1 = warp_sum(a1), a2 = warp_sum(a2);
    if (lane == 0) {
      bf16* o = dec + ((1LL * n * g.PPH + yy) * g.PPW + xx) * g.ld_dec;
      o[0] = __float2bfloat16(a0 + bias[0]), o[1] = __float2bfloat16(a1 + bias[1]), o[2] = __float2bfloat16(a2 + bias[2]);
    }
  }
}

// 16-byte loads / stores of a channel vector (a plain BF8 struct copy is scalarised into four 32-bit accesses)
__device__ __forceinline__ BF8 head_ld8(const bf16* __restrict__ p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  BF8 b;
  *reinterpret_cast<uint4*>(&b) = u;
  return b;
}
__device__ __forceinline__ void head_st8(bf16* __restrict__ p, const BF8& b) { *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&b); }

// The output gradient of conv_out as the tail leaves it (dec_grad_kernel): ddec [N, H*W, ld] bf16 whose channels 0..2
// all hold dmean / 3 and which is zero away from the tap pixels.
struct SparseDy {
  const bf16* ddec;
  long long ld;
  int H, W;
};
// dv[t]: output gradient at the pixel that read input pixel (y, x) through tap t = r * 3 + s, i.e. (y - r + 1, x - s + 1).
__device__ __forceinline__ bool head_load_dv9(const SparseDy& sp, int n, int y, int x, float (&dv)[9]) {
  bool any = false;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int qy = y - (t / 3 - 1), qx = x - (t % 3 - 1);
    float v = 0.f;
    if (qy >= 0 && qy < sp.H && qx >= 0 && qx < sp.W) v = __bfloat162float(sp.ddec[((1LL * n * sp.H + qy) * sp.W + qx) * sp.ld]);
    dv[t] = v;
    any = any || (v != 0.f);
  }
  return any;
}
// wsum[t][c] = sum over the three output channels of w[o][c][tap t] (all three carry the same gradient), fp32.
__device__ __forceinline__ void head_fill_wsum(float* __restrict__ wsum, const bf16* __restrict__ wf, int C, int Cp) {
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const int t = i / C, c = i % C;
    wsum[i] = __bfloat162float(wf[(0 * 9 + t) * Cp + c]) + __bfloat162float(wf[(1 * 9 + t) * Cp + c]) +
              __bfloat162float(wf[(2 * 9 + t) * Cp + c]);
  }
}
// d(conv_out input)[pixel][cv * 8 .. + 8), rounded to bf16 like the dense input-gradient GEMM's output
__device__ __forceinline__ BF8 head_dy8(const float (&dv)[9], const float* __restrict__ wsum, int C, int cv) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (dv[t] == 0.f) continue;
    const float4 w0 = *reinterpret_cast<const float4*>(wsum + t * C + cv * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(wsum + t * C + cv * 8 + 4);
    a[0] = fmaf(dv[t], w0.x, a[0]), a[1] = fmaf(dv[t], w0.y, a[1]), a[2] = fmaf(dv[t], w0.z, a[2]), a[3] = fmaf(dv[t], w0.w, a[3]);
    a[4] = fmaf(dv[t], w1.x, a[4]), a[5] = fmaf(dv[t], w1.y, a[5]), a[6] = fmaf(dv[t], w1.z, a[6]), a[7] = fmaf(dv[t], w1.w, a[7]);
  }
  return f_to_bf8(a);
}

// GroupNorm backward pass 1 over the pixels that receive a gradient from conv_out: same sums, layout and finalisation
// as gn_bwd_stats_kernel, with dy evaluated on the fly (head_dy8) and x read only where dy is non-zero.
// Phase A marks the block's pixels that have a gradient in their 3x3 neighbourhood (one pixel per thread and iteration,
// nine independent 2-byte loads each -- a (channel vector, row) thread layout would walk the pixels serially at one L2
// round trip per step) into shared memory and into act[N, H*W] for pass 2; phase B does the arithmetic of the marked ones.
// dynamic shared memory: [warps][2 G] floats, wsum [9][C] floats, then pix_per_block flag bytes.
__global__ void __launch_bounds__(384, 2) gn_bwd_stats_sp_kernel(const bf16* __restrict__ x, SparseDy sp, const bf16* __restrict__ wf, int Cp,
                                                                 GNShape s, const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, int silu, float* __restrict__ partial,
                                                                 float* __restrict__ gstats_out, unsigned int* __restrict__ ticket,
                                                                 unsigned char* __restrict__ act) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) float head_sh[];
  float* sh = head_sh;
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5;
  float* wsum = sh + ((nwarps_ * 2 * s.G + 3) & ~3);
  unsigned char* flag = reinterpret_cast<unsigned char*>(wsum + 9 * s.C);
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  head_fill_wsum(wsum, wf, s.C, Cp);
  constexpr int UA = 4;  // pixels per thread and iteration: 36 independent loads in flight
  for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x * UA) {
    bool on[UA];
#pragma unroll
    for (int u = 0; u < UA; ++u) {
      const int q = p + u * blockDim.x;
      float dv[9];
      on[u] = q < p1 && head_load_dv9(sp, n, q / sp.W, q % sp.W, dv);
    }
#pragma unroll
    for (int u = 0; u < UA; ++u) {
      const int q = p + u * blockDim.x;
      if (q < p1) flag[q - p0] = on[u] ? 1 : 0, act[1LL * n * s.HW + q] = on[u] ? 1 : 0;
    }
  }
  __syncthreads();
  GNBwdPairs k;
  gn_load_pairs(k, s, n, cv, stats, gamma, beta);
  float2 sa2[4], sb2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sa2[i] = sb2[i] = make_float2(0.f, 0.f);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  for (int p = p0 + r; p < p1; p += R) {
    if (!flag[p - p0]) continue;
    float dv[9];
    head_load_dv9(sp, n, p / sp.W, p % sp.W, dv);
    const BF8 vd = head_dy8(dv, wsum, s.C, cv);
    const BF8 vx = head_ld8(xb + 1LL * p * s.ld);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 xv = __bfloat1622float2(vx.v[i]);
      const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd.v[i]), k, i, silu);
      const float2 xh = __ffma2_rn(xv, make_float2(k.r[i], k.r[i]), make_float2(k.M[i], k.M[i]));
      sa2[i] = __fadd2_rn(sa2[i], d);
      sb2[i] = __ffma2_rn(d, xh, sb2[i]);
    }
  }
  float sa[4], sb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sa[i] = sa2[i].x + sa2[i].y, sb[i] = sb2[i].x + sb2[i].y;
  gn_block_partial(sa, sb, cv, cpg, s.G, sh, partial);
  gn_finalize_last_block(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, 0.f, 1, gstats_out, ticket);
}

// GroupNorm backward pass 2 without a dy tensor: dx (+)= dxhat * r + (x * C1 + C2) (constants as in
// gn_bwd_apply_kernel), where dxhat = 0 away from the 3x3 neighbourhoods of the tap pixels.  Two phases per block:
// a streaming loop over the pixels WITHOUT a gradient (reads x, writes dx; two constants per channel pair live in
// registers; the block's activity bytes sit in shared memory), then the few pixels WITH one (full GroupNorm backward
// arithmetic).  dynamic shared memory: wsum [9][C] floats, then pix_per_block flag bytes.
template <bool ACC>
__global__ void __launch_bounds__(384, 2) gn_bwd_apply_sp_kernel(const bf16* __restrict__ x, SparseDy sp, const bf16* __restrict__ wf, int Cp,
                                                                 GNShape s, const float* __restrict__ stats, const float* __restrict__ gstats,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                                                 const unsigned char* __restrict__ act, bf16* __restrict__ dx, long long lddx) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) float head_sh[];
  float* wsum = head_sh;
  unsigned char* flag = reinterpret_cast<unsigned char*>(wsum + 9 * s.C);
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  bf16* ob = dx + (1LL * n * s.HW) * lddx + cv * 8;
  for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) flag[p - p0] = act[1LL * n * s.HW + p];
  head_fill_wsum(wsum, wf, s.C, Cp);
  __syncthreads();
  {
    float2 C1[4], C2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = (cv * 8 + 2 * i) / cpg;
      const float mean = stats[2 * (n * s.G + g)], rstd = stats[2 * (n * s.G + g) + 1];
      const float m1 = gstats[2 * (n * s.G + g)], m2 = gstats[2 * (n * s.G + g) + 1];
      const float M = -mean * rstd;
      const float c1 = -rstd * rstd * m2, c2 = -rstd * m1 - M * rstd * m2;
      C1[i] = make_float2(c1, c1), C2[i] = make_float2(c2, c2);
    }
    constexpr int U = 4;
    for (int p = p0 + r; p < p1; p += R * U) {
      BF8 vx[U], vo[U];
      bool off[U];  // in range and without a gradient
#pragma unroll
      for (int u = 0; u < U; ++u) {
        off[u] = (p + u * R < p1) && flag[p + u * R - p0] == 0;
        if (off[u]) {
          vx[u] = head_ld8(xb + 1LL * (p + u * R) * s.ld);
          if constexpr (ACC) vo[u] = head_ld8(ob + 1LL * (p + u * R) * lddx);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (off[u]) {
          BF8 o;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // d = 0: the dense kernel computes fma(0, r, fma(x, C1, C2)), which is the inner fma exactly
            float2 g = __ffma2_rn(__bfloat1622float2(vx[u].v[i]), C1[i], C2[i]);
            if constexpr (ACC) g = __fadd2_rn(g, __bfloat1622float2(vo[u].v[i]));
            o.v[i] = __floats2bfloat162_rn(g.x, g.y);
          }
          head_st8(ob + 1LL * (p + u * R) * lddx, o);
        }
    }
  }
  // phase 2: the 3x3 neighbourhoods of the tap pixels (a few per cent of the map)
  GNBwdPairs k;
  gn_load_pairs(k, s, n, cv, stats, gamma, beta);
  float C1[4], C2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int g = (cv * 8 + 2 * i) / cpg;
    const float m1 = gstats[2 * (n * s.G + g)], m2 = gstats[2 * (n * s.G + g) + 1];
    C1[i] = -k.r[i] * k.r[i] * m2;
    C2[i] = -k.r[i] * m1 - k.M[i] * k.r[i] * m2;
  }
  for (int p = p0 + r; p < p1; p += R) {
    if (!flag[p - p0]) continue;
    float dv[9];
    head_load_dv9(sp, n, p / sp.W, p % sp.W, dv);
    const BF8 vd = head_dy8(dv, wsum, s.C, cv);
    const BF8 vx = head_ld8(xb + 1LL * p * s.ld);
    BF8 vo, o;
    if constexpr (ACC) vo = head_ld8(ob + 1LL * p * lddx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 xv = __bfloat1622float2(vx.v[i]);
      const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd.v[i]), k, i, silu);
      float2 g = __ffma2_rn(d, make_float2(k.r[i], k.r[i]), __ffma2_rn(xv, make_float2(C1[i], C1[i]), make_float2(C2[i], C2[i])));
      if constexpr (ACC) g = __fadd2_rn(g, __bfloat1622float2(vo.v[i]));
      o.v[i] = __floats2bfloat162_rn(g.x, g.y);
    }
    head_st8(ob + 1LL * p * lddx, o);
  }
}

}  // namespace mdc
