// Bandwidth-bound kernels of the hot path (everything that is not a dense contraction):
// GroupNorm(+SiLU) fwd/bwd, LayerNorm fwd/bwd, GEGLU fwd/bwd, softmax fwd/bwd, the 2-token
// cross-attention, nearest resampling and its adjoint, channel-slice add/copy.
// Activations are NHWC bf16 ("rows" = n*h*w pixels/tokens, "ld" = elements between rows).
// All kernels use 16-byte vector loads/stores, fp32 math, warp-shuffle reductions, and an
// `acc` flag (accumulate into the destination) so the backward tape needs no separate add kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace mdc {

typedef __nv_bfloat16 bf16;

// Eight bf16 channels = one 16-byte access.  The copy operations move the whole vector as a uint4: a member-wise copy of
// four __nv_bfloat162 is compiled into four 32-bit loads / stores (cuobjdump: LDG.E / STG.E instead of LDG.E.128 /
// STG.E.128), i.e. four partially filled sector requests per warp instruction on every HBM stream of this file.
struct alignas(16) BF8 {
  __nv_bfloat162 v[4];
  __host__ __device__ __forceinline__ BF8() {}
  __host__ __device__ __forceinline__ BF8(const BF8& o) { *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(o.v); }
  __host__ __device__ __forceinline__ BF8& operator=(const BF8& o) {
    *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(o.v);
    return *this;
  }
};
// by value on purpose: the argument is usually a global / shared memory reference, and the copy is ONE 16-byte load
__device__ __forceinline__ void bf8_to_f(const BF8 b, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(b.v[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ BF8 f_to_bf8(const float (&f)[8]) {
  BF8 b;
#pragma unroll
  for (int i = 0; i < 4; ++i) b.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return b;
}
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum for blockDim.x <= 1024; `sh` must hold 32 floats. Result valid in all threads.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}
// sigmoid(x) = 0.5 tanh(x/2) + 0.5: one MUFU op (tanh.approx, rel. error ~2^-11, far below bf16 resolution)
__device__ __forceinline__ float sigmoidf_(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(t, 0.5f, 0.5f);
}
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float silu_grad(float x) {
  float s = sigmoidf_(x);
  return s * (1.f + x * (1.f - s));
}
// ---- packed fp32x2 forms (Blackwell issues two fp32 FMAs per instruction): the two-pass GroupNorm kernels on the
// 56-226 MB decoder tensors are instruction-bound as much as HBM-bound, so their per-element math is written on channel
// pairs with every per-channel affine step folded into one FMA.
__device__ __forceinline__ float2 bf16r2(float2 v) { return __bfloat1622float2(__floats2bfloat162_rn(v.x, v.y)); }
__device__ __forceinline__ float2 sigmoid2(float2 x) {
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h.y));
  return __ffma2_rn(make_float2(t0, t1), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ float2 silu2(float2 x) { return __fmul2_rn(x, sigmoid2(x)); }
__device__ __forceinline__ float2 silu_grad2(float2 x) {  // s (1 + x (1 - s))
  const float2 s = sigmoid2(x);
  const float2 om = __ffma2_rn(s, make_float2(-1.f, -1.f), make_float2(1.f, 1.f));
  return __fmul2_rn(s, __ffma2_rn(x, om, make_float2(1.f, 1.f)));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// =========================================================================== GroupNorm
// Thread layout shared by all GroupNorm kernels: blockDim = CV * R (CV = C/8 channel vectors per pixel),
// thread -> (cv = tid % CV, r = tid / CV); each block walks `pix_per_block` pixels of one image.  A thread's
// channel vector, hence the groups of its 4 channel pairs, is fixed, so statistics accumulate in registers and the
// per-group constants are loaded once.  Loops are unrolled by GN_UNROLL pixels so every thread keeps several
// independent 16-byte loads in flight (these kernels are pure HBM streams).
constexpr int GN_UNROLL = 4;
struct GNShape {
  int N, HW, C, G;
  long long ld;        // pixel stride of x / y (elements)
  int pix_per_block;   // pixels handled by one block
  int blocks_per_img;  // gridDim.x = N * blocks_per_img
  int stages1 = 0;     // ring depth of the streamed ONE-tensor kernels (gn_stats_s / gn_apply_s); 0 = GNS_STAGES
};

// The last block of an image to finish (atomic ticket) reduces the per-block partials of that image in a FIXED order,
// in double, and writes the per-group results: mode 0 -> (mean, rstd), mode 1 -> (sum / m, sumxy / m).  This folds the
// former finalize launch into the statistics kernels (182 fewer launches per guided step).
__device__ __forceinline__ void gn_finalize_last_block(const float* __restrict__ partial, int n, int G, int bpi, double m,
                                                      float eps, int mode, float* __restrict__ out,
                                                      unsigned int* __restrict__ ticket) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&ticket[n], 1u);
    s_last = (t == static_cast<unsigned int>(bpi - 1)) ? 1u : 0u;
    if (s_last) ticket[n] = 0u;  // self-cleaning for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = warp; g < G; g += nwarps) {
    double a = 0, b = 0;
    // the whole GPU waits for this one block: issue all of a lane's loads before the first add (same order of adds)
    constexpr int RK = 10;  // 32 x 10 = 320 partials per pass (two blocks per SM)
    for (int base = 0; base < bpi; base += 32 * RK) {
      float2 pv[RK];
#pragma unroll
      for (int i = 0; i < RK; ++i) {
        const int k = base + lane + 32 * i;
        pv[i] = k < bpi ? __ldcg(reinterpret_cast<const float2*>(partial + (1LL * (n * bpi + k) * G + g) * 2)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < RK; ++i)
        if (base + lane + 32 * i < bpi) a += pv[i].x, b += pv[i].y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      const int i = n * G + g;
      if (mode == 0) {
        double mean = a / m, var = b / m - mean * mean;
        if (var < 0) var = 0;
        out[2 * i] = static_cast<float>(mean);
        out[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + eps));
      } else {
        out[2 * i] = static_cast<float>(a / m);
        out[2 * i + 1] = static_cast<float>(b / m);
      }
    }
  }
}

// pass 1 of forward: per-block partial (sum, sumsq) per group -> partial[(n*bpi + b)*G*2 + g*2 + {0,1}]
__global__ void __launch_bounds__(512, 2) gn_stats_kernel(const bf16* __restrict__ x, GNShape s,
                                                          float* __restrict__ partial, float eps,
                                                          float* __restrict__ stats_out,
                                                          unsigned int* __restrict__ ticket) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ float sh[];  // 2*G
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  // sh: [warps][2G] per-warp accumulators (low-contention shared atomics), reduced across warps at the end
  const int nwarps_ = (blockDim.x + 31) >> 5, wid_ = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* base = x + (1LL * n * s.HW) * s.ld + cv * 8;
  for (int p = p0 + r; p < p1; p += R * GN_UNROLL) {
    BF8 v[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) v[u] = *reinterpret_cast<const BF8*>(base + 1LL * (p + u * R) * s.ld);
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 t = __bfloat1622float2(v[u].v[i]);
          sum[i] += t.x + t.y;
          sq[i] += t.x * t.x + t.y * t.y;
        }
      }
  }
  {
    float* mine = sh + wid_ * 2 * s.G;
    int gprev = -1;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // merge this thread's pairs that fall into the same group before touching smem
      int g = (cv * 8 + 2 * i) / cpg;
      if (g != gprev && gprev >= 0) {
        atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
        a = b = 0.f;
      }
      gprev = g, a += sum[i], b += sq[i];
    }
    atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nwarps_; ++w) t += sh[w * 2 * s.G + i];
    partial[1LL * blockIdx.x * 2 * s.G + i] = t;
  }
  gn_finalize_last_block(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, eps, 0, stats_out, ticket);
}

// Reduce the per-block partials in a fixed order (deterministic), in double: one warp per (n, group).
// mode 0: (sum, sumsq) -> (mean, rstd);  mode 1: plain sums scaled by 1/m (used by the backward pass).
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int N, int G, int bpi, double m, float eps,
                                   int mode, float* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N * G) return;
  const int n = i / G, g = i % G;
  double a = 0, b = 0;
  for (int k = lane; k < bpi; k += 32) {
    const float2 p = *reinterpret_cast<const float2*>(partial + (1LL * (n * bpi + k) * G + g) * 2);
    a += p.x, b += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane) return;
  if (mode == 0) {
    double mean = a / m, var = b / m - mean * mean;
    if (var < 0) var = 0;
    out[2 * i] = static_cast<float>(mean);
    out[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + eps));
  } else {
    out[2 * i] = static_cast<float>(a / m);
    out[2 * i + 1] = static_cast<float>(b / m);
  }
}

// pass 2 of forward: y = act((x - mean) * rstd * gamma + beta).
// With `partial` != nullptr the statistics come from the epilogue of the GEMM that produced x (GemmParams::gn_partial):
// `nparts` rows of (sum, sum of squares) per group and image, reduced here in a fixed order in double (every block
// does it for its image: a few KB from L2), and block 0 of the image publishes (mean, rstd) to `stats` for the backward.
__global__ void __launch_bounds__(512, 2) gn_apply_kernel(const bf16* __restrict__ x, GNShape s,
                                                          float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int silu,
                                                          bf16* __restrict__ y, long long ldy,
                                                          const float* __restrict__ partial, int nparts, float eps) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float sstat[128];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  if (partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double m = 1.0 * s.HW * cpg;
    for (int g = warp; g < s.G; g += nwarps) {
      double a = 0, q = 0;
      for (int k = lane; k < nparts; k += 32) {
        const float2 pv = __ldcg(reinterpret_cast<const float2*>(partial + (1LL * n * nparts + k) * 2 * s.G + 2 * g));
        a += pv.x, q += pv.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        const double mean = a / m;
        double var = q / m - mean * mean;
        if (var < 0) var = 0;
        sstat[2 * g] = static_cast<float>(mean);
        sstat[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + eps));
      }
    }
    __syncthreads();
    if (b == 0)
      for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) stats[2 * n * s.G + i] = sstat[i];
  }
  float2 sc[4], sf[4];  // per channel pair: y = act(x * sc + sf)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cv * 8 + 2 * i, g = c / cpg;  // a pair never straddles a group (channels per group is even)
    const float mean = partial ? sstat[2 * g] : stats[2 * (n * s.G + g)], rstd = partial ? sstat[2 * g + 1] : stats[2 * (n * s.G + g) + 1];
    sc[i] = make_float2(rstd * gamma[c], rstd * gamma[c + 1]);
    sf[i] = make_float2(beta[c] - mean * sc[i].x, beta[c + 1] - mean * sc[i].y);
  }
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  bf16* yb = y + (1LL * n * s.HW) * ldy + cv * 8;
  for (int p = p0 + r; p < p1; p += R * GN_UNROLL) {
    BF8 v[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) v[u] = *reinterpret_cast<const BF8*>(xb + 1LL * (p + u * R) * s.ld);
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) {
        BF8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 h = __ffma2_rn(__bfloat1622float2(v[u].v[i]), sc[i], sf[i]);
          if (silu) h = silu2(bf16r2(h));
          o.v[i] = __floats2bfloat162_rn(h.x, h.y);
        }
        *reinterpret_cast<BF8*>(yb + 1LL * (p + u * R) * ldy) = o;
      }
  }
}

// Per-thread constants of the backward kernels: group statistics per channel PAIR (a pair never straddles a group
// because channels-per-group is even), affine parameters per channel.
struct GNBwdConst {
  float mean[4], rstd[4], ga[8], be[8];
};
__device__ __forceinline__ void gn_load_const(GNBwdConst& k, const GNShape& s, int n, int cv,
                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                              const float* __restrict__ beta) {
  const int cpg = s.C / s.G;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int g = (cv * 8 + 2 * i) / cpg;
    k.mean[i] = stats[2 * (n * s.G + g)], k.rstd[i] = stats[2 * (n * s.G + g) + 1];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) k.ga[i] = gamma[cv * 8 + i], k.be[i] = beta[cv * 8 + i];
}

// Per-thread constants of the two-pass backward kernels, per channel pair, with every affine step folded:
//   u  = x * A + B          (pre-activation: A = rstd gamma, B = beta - mean rstd gamma)
//   xh = x * r + M          (normalised value: r = rstd, M = -mean rstd)
//   d  = dy * act'(u) * gamma
struct GNBwdPairs {
  float2 A[4], B[4], G[4];
  float r[4], M[4];
};
__device__ __forceinline__ void gn_load_pairs(GNBwdPairs& k, const GNShape& s, int n, int cv, const float* __restrict__ stats,
                                              const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int cpg = s.C / s.G;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cv * 8 + 2 * i, g = c / cpg;
    const float mean = stats[2 * (n * s.G + g)], rstd = stats[2 * (n * s.G + g) + 1];
    k.G[i] = make_float2(gamma[c], gamma[c + 1]);
    k.A[i] = make_float2(rstd * k.G[i].x, rstd * k.G[i].y);
    k.B[i] = make_float2(beta[c] - mean * k.A[i].x, beta[c + 1] - mean * k.A[i].y);
    k.r[i] = rstd, k.M[i] = -mean * rstd;
  }
}
// d = dy * act'(bf16(u)) * gamma for one channel pair
__device__ __forceinline__ float2 gn_dxhat2(float2 x, float2 dy, const GNBwdPairs& k, int i, int silu) {
  float2 d = __fmul2_rn(dy, k.G[i]);
  if (silu) d = __fmul2_rn(d, silu_grad2(bf16r2(__ffma2_rn(x, k.A[i], k.B[i]))));
  return d;
}

// backward pass 1: per-block partial (sum dxhat, sum dxhat*xhat) per group, dxhat = dy * act'(h) * gamma
__global__ void __launch_bounds__(384, 2) gn_bwd_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              long long lddy, GNShape s,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int silu,
                                                              float* __restrict__ partial,
                                                              float* __restrict__ gstats_out,
                                                              unsigned int* __restrict__ ticket) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ float sh[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5, wid_ = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  GNBwdPairs k;
  gn_load_pairs(k, s, n, cv, stats, gamma, beta);
  float2 sa2[4], sb2[4];  // lanes of a pair are summed at the end
#pragma unroll
  for (int i = 0; i < 4; ++i) sa2[i] = sb2[i] = make_float2(0.f, 0.f);
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  const bf16* db = dy + (1LL * n * s.HW) * lddy + cv * 8;
  constexpr int U = 2;
  for (int p = p0 + r; p < p1; p += R * U) {
    BF8 vx[U], vd[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
        vx[u] = *reinterpret_cast<const BF8*>(xb + 1LL * (p + u * R) * s.ld);
        vd[u] = *reinterpret_cast<const BF8*>(db + 1LL * (p + u * R) * lddy);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xv = __bfloat1622float2(vx[u].v[i]);
          const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd[u].v[i]), k, i, silu);
          const float2 xh = __ffma2_rn(xv, make_float2(k.r[i], k.r[i]), make_float2(k.M[i], k.M[i]));
          sa2[i] = __fadd2_rn(sa2[i], d);
          sb2[i] = __ffma2_rn(d, xh, sb2[i]);
        }
      }
  }
  {
    float* mine = sh + wid_ * 2 * s.G;
    int gprev = -1;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int g = (cv * 8 + 2 * i) / cpg;
      if (g != gprev && gprev >= 0) {
        atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
        a = b = 0.f;
      }
      gprev = g, a += sa2[i].x + sa2[i].y, b += sb2[i].x + sb2[i].y;
    }
    atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nwarps_; ++w) t += sh[w * 2 * s.G + i];
    partial[1LL * blockIdx.x * 2 * s.G + i] = t;
  }
  gn_finalize_last_block(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, 0.f, 1, gstats_out, ticket);
}

// backward pass 2: dx (+)= rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat*xhat)) = dxhat * r + (x * C1 + C2) with
// C1 = -r^2 m2, C2 = -r m1 - M r m2 per group
__global__ void __launch_bounds__(384, 2) gn_bwd_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              long long lddy, GNShape s,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gstats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int silu,
                                                              bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
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
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  const bf16* db = dy + (1LL * n * s.HW) * lddy + cv * 8;
  bf16* ob = dx + (1LL * n * s.HW) * lddx + cv * 8;
  constexpr int U = 2;
  for (int p = p0 + r; p < p1; p += R * U) {
    BF8 vx[U], vd[U], vo[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
        vx[u] = *reinterpret_cast<const BF8*>(xb + 1LL * (p + u * R) * s.ld);
        vd[u] = *reinterpret_cast<const BF8*>(db + 1LL * (p + u * R) * lddy);
        if (acc) vo[u] = *reinterpret_cast<const BF8*>(ob + 1LL * (p + u * R) * lddx);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
        BF8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xv = __bfloat1622float2(vx[u].v[i]);
          const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd[u].v[i]), k, i, silu);
          float2 g = __ffma2_rn(d, make_float2(k.r[i], k.r[i]), __ffma2_rn(xv, make_float2(C1[i], C1[i]), make_float2(C2[i], C2[i])));
          if (acc) g = __fadd2_rn(g, __bfloat1622float2(vo[u].v[i]));
          o.v[i] = __floats2bfloat162_rn(g.x, g.y);
        }
        *reinterpret_cast<BF8*>(ob + 1LL * (p + u * R) * lddx) = o;
      }
  }
}

// --------------------------------------------------------------------------- single-launch GroupNorm
// For activations whose per-CTA slab fits in shared memory (every UNet tensor, the low-resolution decoder tensors) the
// two passes run in ONE kernel: each CTA stages its pixel slab in shared memory while accumulating the group
// partials, all CTAs meet at a grid barrier, every CTA reduces the partials of its image in a fixed order and applies
// the normalisation from shared memory.  The tensor is read from global memory once and a launch disappears.
// The grid never exceeds the SM count (one CTA per SM), so all CTAs are co-resident; dependents are released
// (griddepcontrol.launch_dependents) only after the barrier, so no later grid can take an SM away before that.

// Sense-reversing grid barrier: bar[0] arrival counter, bar[1] generation, bar[2] sticky timeout flag.  Self-cleaning.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* vgen = bar + 1;
    const unsigned int gen = *vgen;  // read before arriving: the generation cannot advance without this CTA
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      const long long t0 = clock64();
      while (*vgen == gen) {
        if (clock64() - t0 > (4LL << 30)) {  // ~2 s: never hang the device; results are then garbage and flagged
          atomicExch(bar + 2, 1u);
          break;
        }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// Fixed-order reduction of the per-block partials of image n: 8 lanes per group, each summing every 8th block in
// double, combined by an xor tree.  Writes out2[2g], out2[2g+1] (shared memory) for all groups; needs >= 8*G threads.
__device__ __forceinline__ void gn_reduce_partials(const float* __restrict__ partial, int n, int G, int bpi, double m,
                                                   float eps, int mode, float* __restrict__ out2) {
  const int g = threadIdx.x >> 3, j = threadIdx.x & 7;
  if (g < G) {
    double a = 0, b = 0;
    // every CTA of the grid sits in this reduction right after the barrier: all of a lane's loads are issued before the
    // first add (one L2 round trip instead of one per four partials); the adds keep their order, so the sums are the same
    constexpr int RK = 20;  // 8 x 20 = 160 partials per pass (the grid never exceeds the SM count)
    for (int base = 0; base < bpi; base += 8 * RK) {
      float2 pv[RK];
#pragma unroll
      for (int i = 0; i < RK; ++i) {
        const int k = base + j + 8 * i;
        pv[i] = k < bpi ? __ldcg(reinterpret_cast<const float2*>(partial + (1LL * (n * bpi + k) * G + g) * 2)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < RK; ++i)
        if (base + j + 8 * i < bpi) a += pv[i].x, b += pv[i].y;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (j == 0) {
      if (mode == 0) {
        double mean = a / m, var = b / m - mean * mean;
        if (var < 0) var = 0;
        out2[2 * g] = static_cast<float>(mean);
        out2[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + eps));
      } else {
        out2[2 * g] = static_cast<float>(a / m);
        out2[2 * g + 1] = static_cast<float>(b / m);
      }
    }
  }
}

// Merge a thread's 4 channel-pair sums into the per-warp shared accumulators, then the warps into partial[blockIdx].
__device__ __forceinline__ void gn_block_partial(const float (&sa)[4], const float (&sb)[4], int cv, int cpg, int G,
                                                 float* __restrict__ sh, float* __restrict__ partial) {
  const int nwarps_ = (blockDim.x + 31) >> 5, wid_ = threadIdx.x >> 5;
  float* mine = sh + wid_ * 2 * G;
  int gprev = -1;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int g = (cv * 8 + 2 * i) / cpg;
    if (g != gprev && gprev >= 0) {
      atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
      a = b = 0.f;
    }
    gprev = g, a += sa[i], b += sb[i];
  }
  atomicAdd(&mine[2 * gprev], a), atomicAdd(&mine[2 * gprev + 1], b);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nwarps_; ++w) t += sh[w * 2 * G + i];
    partial[1LL * blockIdx.x * 2 * G + i] = t;
  }
}

// shared memory: [pix_per_block * C bf16 slab][nwarps * 2G floats][2G floats]
__global__ void __launch_bounds__(512, 1) gn_fused_fwd_kernel(const bf16* __restrict__ x, GNShape s,
                                                              float* __restrict__ partial, float eps,
                                                              float* __restrict__ stats_out, unsigned int* __restrict__ bar,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int silu,
                                                              bf16* __restrict__ y, long long ldy) {
  ptx::pdl_wait();
  extern __shared__ __align__(16) uint8_t gn_smem[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5;
  BF8* slab = reinterpret_cast<BF8*>(gn_smem);
  float* sh = reinterpret_cast<float*>(gn_smem + static_cast<size_t>(s.pix_per_block) * s.C * 2);
  float* sstat = sh + nwarps_ * 2 * s.G;
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  float sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  for (int p = p0 + r; p < p1; p += R * GN_UNROLL) {
    BF8 v[GN_UNROLL];
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) v[u] = *reinterpret_cast<const BF8*>(xb + 1LL * (p + u * R) * s.ld);
#pragma unroll
    for (int u = 0; u < GN_UNROLL; ++u)
      if (p + u * R < p1) {
        slab[(p + u * R - p0) * CV + cv] = v[u];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 t = __bfloat1622float2(v[u].v[i]);
          sum[i] += t.x + t.y;
          sq[i] += t.x * t.x + t.y * t.y;
        }
      }
  }
  gn_block_partial(sum, sq, cv, cpg, s.G, sh, partial);
  grid_barrier(bar, gridDim.x);
  ptx::pdl_launch();
  gn_reduce_partials(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, eps, 0, sstat);
  __syncthreads();
  if (b == 0)
    for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) stats_out[2 * n * s.G + i] = sstat[i];
  float sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int c = cv * 8 + i, g = c / cpg;
    sc[i] = sstat[2 * g + 1] * gamma[c];
    sf[i] = beta[c] - sstat[2 * g] * sc[i];
  }
  bf16* yb = y + (1LL * n * s.HW) * ldy + cv * 8;
  for (int p = p0 + r; p < p1; p += R) {  // each thread re-reads exactly the slab entries it wrote
    float f[8];
    bf8_to_f(slab[(p - p0) * CV + cv], f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float h = f[i] * sc[i] + sf[i];
      f[i] = silu ? siluf_(bf16r(h)) : h;
    }
    *reinterpret_cast<BF8*>(yb + 1LL * p * ldy) = f_to_bf8(f);
  }
}

// shared memory: [x slab][dy slab][nwarps * 2G floats][2G floats]
__global__ void __launch_bounds__(512, 1) gn_fused_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                              long long lddy, GNShape s,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int silu,
                                                              float* __restrict__ partial, unsigned int* __restrict__ bar,
                                                              bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  extern __shared__ __align__(16) uint8_t gn_smem[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5;
  const size_t slab_bytes = static_cast<size_t>(s.pix_per_block) * s.C * 2;
  BF8* slab_x = reinterpret_cast<BF8*>(gn_smem);
  BF8* slab_d = reinterpret_cast<BF8*>(gn_smem + slab_bytes);
  float* sh = reinterpret_cast<float*>(gn_smem + 2 * slab_bytes);
  float* sstat = sh + nwarps_ * 2 * s.G;
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  GNBwdConst k;
  gn_load_const(k, s, n, cv, stats, gamma, beta);
  float sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
  const int p0 = b * s.pix_per_block, p1 = min(s.HW, p0 + s.pix_per_block);
  const bf16* xb = x + (1LL * n * s.HW) * s.ld + cv * 8;
  const bf16* db = dy + (1LL * n * s.HW) * lddy + cv * 8;
  constexpr int U = 2;
  for (int p = p0 + r; p < p1; p += R * U) {
    BF8 vx[U], vd[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
        vx[u] = *reinterpret_cast<const BF8*>(xb + 1LL * (p + u * R) * s.ld);
        vd[u] = *reinterpret_cast<const BF8*>(db + 1LL * (p + u * R) * lddy);
      }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (p + u * R < p1) {
        slab_x[(p + u * R - p0) * CV + cv] = vx[u];
        slab_d[(p + u * R - p0) * CV + cv] = vd[u];
        float fx[8], fd[8];
        bf8_to_f(vx[u], fx);
        bf8_to_f(vd[u], fd);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float xh = (fx[i] - k.mean[i >> 1]) * k.rstd[i >> 1];
          float d = fd[i];
          if (silu) d *= silu_grad(bf16r(xh * k.ga[i] + k.be[i]));
          d *= k.ga[i];
          sa[i >> 1] += d;
          sb[i >> 1] += d * xh;
        }
      }
  }
  gn_block_partial(sa, sb, cv, cpg, s.G, sh, partial);
  grid_barrier(bar, gridDim.x);
  ptx::pdl_launch();
  gn_reduce_partials(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, 0.f, 1, sstat);
  __syncthreads();
  float m1[4], m2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int g = (cv * 8 + 2 * i) / cpg;
    m1[i] = sstat[2 * g], m2[i] = sstat[2 * g + 1];
  }
  bf16* ob = dx + (1LL * n * s.HW) * lddx + cv * 8;
  for (int p = p0 + r; p < p1; p += R) {
    float fx[8], fd[8], o[8];
    bf8_to_f(slab_x[(p - p0) * CV + cv], fx);
    bf8_to_f(slab_d[(p - p0) * CV + cv], fd);
    if (acc) bf8_to_f(*reinterpret_cast<const BF8*>(ob + 1LL * p * lddx), o);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float xh = (fx[i] - k.mean[i >> 1]) * k.rstd[i >> 1];
      float d = fd[i];
      if (silu) d *= silu_grad(bf16r(xh * k.ga[i] + k.be[i]));
      d *= k.ga[i];
      float g = k.rstd[i >> 1] * (d - m1[i >> 1] - xh * m2[i >> 1]);
      o[i] = acc ? o[i] + g : g;
    }
    *reinterpret_cast<BF8*>(ob + 1LL * p * lddx) = f_to_bf8(o);
  }
}

// --------------------------------------------------------------------------- streamed two-pass GroupNorm
// ncu on the register-fed kernels above (113 MB tensor, profiles/r02_ncu_groupnorm.md): DRAM 39-54 % busy, issue slots
// 53-59 % busy, 35-49 % of the warp slots occupied -- neither memory nor instruction bound but LATENCY bound: a thread
// keeps only 4 x 16 B in flight and cannot request the next batch before it has consumed the current one.  The streamed
// variants decouple the two: a block's pixel slab is contiguous memory (pixel stride == C), so one thread feeds a ring
// of shared-memory stages with 1-D bulk copies (cp.async.bulk, completion on an mbarrier) that runs GNS_STAGES - 1
// chunks ahead of the arithmetic, and the bytes in flight per SM no longer depend on registers or occupancy.
constexpr int GNS_STAGES = 4;
constexpr int GNS_UNROLL = 2;  // pixel rows (of R pixels) per chunk and thread
struct GnsPipe {
  uint64_t* full;      // [GNS_STAGES]
  uint8_t* buf;        // [GNS_STAGES][ntens][chunk_bytes]
  uint32_t chunk_bytes, ntens;
  int stages;
  const uint8_t* src[2];
  long long total_bytes;  // bytes of this block's slab (per tensor)
  int nchunks;
};
__device__ __forceinline__ void gns_issue(const GnsPipe& q, int k) {  // one thread
  const int st = k % q.stages;
  const long long off = 1LL * k * q.chunk_bytes;
  const uint32_t bytes = static_cast<uint32_t>(min(static_cast<long long>(q.chunk_bytes), q.total_bytes - off));
  ptx::mbar_expect_tx(&q.full[st], bytes * q.ntens);
  for (uint32_t t = 0; t < q.ntens; ++t) ptx::bulk_load_1d(q.buf + (st * q.ntens + t) * q.chunk_bytes, q.src[t] + off, bytes, &q.full[st]);
}
// Sets the ring up and requests the first GNS_STAGES - 1 chunks.  smem: [8 x u64 barriers][stages x ntens x chunk].
__device__ __forceinline__ GnsPipe gns_begin(uint8_t* smem, const bf16* t0, const bf16* t1, long long elem_off, long long elems, int C,
                                             int R, int stages = GNS_STAGES) {
  GnsPipe q;
  q.stages = stages;
  q.full = reinterpret_cast<uint64_t*>(smem);
  q.buf = smem + 128;
  q.ntens = t1 ? 2u : 1u;
  q.chunk_bytes = static_cast<uint32_t>(R) * GNS_UNROLL * C * 2u;
  q.src[0] = reinterpret_cast<const uint8_t*>(t0 + elem_off);
  q.src[1] = t1 ? reinterpret_cast<const uint8_t*>(t1 + elem_off) : nullptr;
  q.total_bytes = elems * 2;
  q.nchunks = static_cast<int>((q.total_bytes + q.chunk_bytes - 1) / q.chunk_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < q.stages; ++i) ptx::mbar_init(&q.full[i], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 0; k < q.stages - 1 && k < q.nchunks; ++k) gns_issue(q, k);
  return q;
}
// Per chunk: the refill of the stage consumed in the PREVIOUS iteration is requested first (every thread passed the
// __syncthreads that ended that iteration), then the block waits for this chunk's bytes.
__device__ __forceinline__ const uint8_t* gns_acquire(const GnsPipe& q, int k) {
  if (threadIdx.x == 0 && k + q.stages - 1 < q.nchunks) gns_issue(q, k + q.stages - 1);
  ptx::mbar_wait(&q.full[k % q.stages], (k / q.stages) & 1);
  return q.buf + (k % q.stages) * q.ntens * q.chunk_bytes;
}
__host__ __device__ inline size_t gns_smem_bytes(int threads, int ntens, int stages = GNS_STAGES) {
  return 128 + static_cast<size_t>(stages) * ntens * threads * GNS_UNROLL * 16;
}
__host__ __device__ inline int gns_stages1(const GNShape& s) { return s.stages1 > 0 ? s.stages1 : GNS_STAGES; }

__global__ void __launch_bounds__(512, 2) gn_stats_s_kernel(const bf16* __restrict__ x, GNShape s, float* __restrict__ partial, float eps,
                                                            float* __restrict__ stats_out, unsigned int* __restrict__ ticket) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(128) uint8_t gns_smem[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5;
  float* sh = reinterpret_cast<float*>(gns_smem + gns_smem_bytes(blockDim.x, 1, gns_stages1(s)));  // [warps][2G]
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  const int p0 = b * s.pix_per_block, npix = min(s.HW, p0 + s.pix_per_block) - p0;
  GnsPipe q = gns_begin(gns_smem, x, nullptr, (1LL * n * s.HW + p0) * s.C, 1LL * npix * s.C, s.C, R, gns_stages1(s));
  float2 sum2[4], sq2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sum2[i] = sq2[i] = make_float2(0.f, 0.f);
  for (int k = 0; k < q.nchunks; ++k) {
    const BF8* st = reinterpret_cast<const BF8*>(gns_acquire(q, k));
#pragma unroll
    for (int u = 0; u < GNS_UNROLL; ++u) {
      const int pl = (k * GNS_UNROLL + u) * R + r;
      if (pl < npix) {
        const BF8 v = st[(u * R + r) * CV + cv];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __bfloat1622float2(v.v[i]);
          sum2[i] = __fadd2_rn(sum2[i], t);
          sq2[i] = __ffma2_rn(t, t, sq2[i]);
        }
      }
    }
    __syncthreads();
  }
  float sa[4], sb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sa[i] = sum2[i].x + sum2[i].y, sb[i] = sq2[i].x + sq2[i].y;
  gn_block_partial(sa, sb, cv, cpg, s.G, sh, partial);
  gn_finalize_last_block(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, eps, 0, stats_out, ticket);
}

__global__ void __launch_bounds__(512, 2) gn_apply_s_kernel(const bf16* __restrict__ x, GNShape s, float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                                            bf16* __restrict__ y, long long ldy, const float* __restrict__ partial,
                                                            int nparts, float eps) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(128) uint8_t gns_smem[];
  __shared__ float sstat[128];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int p0 = b * s.pix_per_block, npix = min(s.HW, p0 + s.pix_per_block) - p0;
  GnsPipe q = gns_begin(gns_smem, x, nullptr, (1LL * n * s.HW + p0) * s.C, 1LL * npix * s.C, s.C, R, gns_stages1(s));  // loads fly during the prologue
  if (partial) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const double m = 1.0 * s.HW * cpg;
    for (int g = warp; g < s.G; g += nwarps) {
      double a = 0, qq = 0;
      for (int k = lane; k < nparts; k += 32) {
        const float2 pv = __ldcg(reinterpret_cast<const float2*>(partial + (1LL * n * nparts + k) * 2 * s.G + 2 * g));
        a += pv.x, qq += pv.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        qq += __shfl_xor_sync(0xffffffffu, qq, o);
      }
      if (lane == 0) {
        const double mean = a / m;
        double var = qq / m - mean * mean;
        if (var < 0) var = 0;
        sstat[2 * g] = static_cast<float>(mean);
        sstat[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + eps));
      }
    }
    __syncthreads();
    if (b == 0)
      for (int i = threadIdx.x; i < 2 * s.G; i += blockDim.x) stats[2 * n * s.G + i] = sstat[i];
  }
  float2 sc[4], sf[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cv * 8 + 2 * i, g = c / cpg;
    const float mean = partial ? sstat[2 * g] : stats[2 * (n * s.G + g)], rstd = partial ? sstat[2 * g + 1] : stats[2 * (n * s.G + g) + 1];
    sc[i] = make_float2(rstd * gamma[c], rstd * gamma[c + 1]);
    sf[i] = make_float2(beta[c] - mean * sc[i].x, beta[c + 1] - mean * sc[i].y);
  }
  bf16* yb = y + (1LL * n * s.HW + p0) * ldy + cv * 8;
  for (int k = 0; k < q.nchunks; ++k) {
    const BF8* st = reinterpret_cast<const BF8*>(gns_acquire(q, k));
#pragma unroll
    for (int u = 0; u < GNS_UNROLL; ++u) {
      const int pl = (k * GNS_UNROLL + u) * R + r;
      if (pl < npix) {
        const BF8 v = st[(u * R + r) * CV + cv];
        BF8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 h = __ffma2_rn(__bfloat1622float2(v.v[i]), sc[i], sf[i]);
          if (silu) h = silu2(bf16r2(h));
          o.v[i] = __floats2bfloat162_rn(h.x, h.y);
        }
        *reinterpret_cast<BF8*>(yb + 1LL * pl * ldy) = o;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(384, 2) gn_bwd_stats_s_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, GNShape s,
                                                                const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, int silu, float* __restrict__ partial,
                                                                float* __restrict__ gstats_out, unsigned int* __restrict__ ticket) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(128) uint8_t gns_smem[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int nwarps_ = (blockDim.x + 31) >> 5;
  float* sh = reinterpret_cast<float*>(gns_smem + gns_smem_bytes(blockDim.x, 2));
  for (int i = threadIdx.x; i < nwarps_ * 2 * s.G; i += blockDim.x) sh[i] = 0.f;
  const int p0 = b * s.pix_per_block, npix = min(s.HW, p0 + s.pix_per_block) - p0;
  GnsPipe q = gns_begin(gns_smem, x, dy, (1LL * n * s.HW + p0) * s.C, 1LL * npix * s.C, s.C, R);
  GNBwdPairs kk;
  gn_load_pairs(kk, s, n, cv, stats, gamma, beta);
  float2 sa2[4], sb2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sa2[i] = sb2[i] = make_float2(0.f, 0.f);
  for (int k = 0; k < q.nchunks; ++k) {
    const BF8* st = reinterpret_cast<const BF8*>(gns_acquire(q, k));
    const BF8* sd = reinterpret_cast<const BF8*>(reinterpret_cast<const uint8_t*>(st) + q.chunk_bytes);
#pragma unroll
    for (int u = 0; u < GNS_UNROLL; ++u) {
      const int pl = (k * GNS_UNROLL + u) * R + r;
      if (pl < npix) {
        const BF8 vx = st[(u * R + r) * CV + cv], vd = sd[(u * R + r) * CV + cv];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xv = __bfloat1622float2(vx.v[i]);
          const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd.v[i]), kk, i, silu);
          const float2 xh = __ffma2_rn(xv, make_float2(kk.r[i], kk.r[i]), make_float2(kk.M[i], kk.M[i]));
          sa2[i] = __fadd2_rn(sa2[i], d);
          sb2[i] = __ffma2_rn(d, xh, sb2[i]);
        }
      }
    }
    __syncthreads();
  }
  float sa[4], sb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sa[i] = sa2[i].x + sa2[i].y, sb[i] = sb2[i].x + sb2[i].y;
  gn_block_partial(sa, sb, cv, cpg, s.G, sh, partial);
  gn_finalize_last_block(partial, n, s.G, s.blocks_per_img, 1.0 * s.HW * cpg, 0.f, 1, gstats_out, ticket);
}

__global__ void __launch_bounds__(384, 2) gn_bwd_apply_s_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, GNShape s,
                                                                const float* __restrict__ stats, const float* __restrict__ gstats,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                                                bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(128) uint8_t gns_smem[];
  const int CV = s.C >> 3, cpg = s.C / s.G;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV, R = blockDim.x / CV;
  const int n = blockIdx.x / s.blocks_per_img, b = blockIdx.x % s.blocks_per_img;
  const int p0 = b * s.pix_per_block, npix = min(s.HW, p0 + s.pix_per_block) - p0;
  GnsPipe q = gns_begin(gns_smem, x, dy, (1LL * n * s.HW + p0) * s.C, 1LL * npix * s.C, s.C, R);
  GNBwdPairs kk;
  gn_load_pairs(kk, s, n, cv, stats, gamma, beta);
  float C1[4], C2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int g = (cv * 8 + 2 * i) / cpg;
    const float m1 = gstats[2 * (n * s.G + g)], m2 = gstats[2 * (n * s.G + g) + 1];
    C1[i] = -kk.r[i] * kk.r[i] * m2;
    C2[i] = -kk.r[i] * m1 - kk.M[i] * kk.r[i] * m2;
  }
  bf16* ob = dx + (1LL * n * s.HW + p0) * lddx + cv * 8;
  for (int k = 0; k < q.nchunks; ++k) {
    const BF8* st = reinterpret_cast<const BF8*>(gns_acquire(q, k));
    const BF8* sd = reinterpret_cast<const BF8*>(reinterpret_cast<const uint8_t*>(st) + q.chunk_bytes);
#pragma unroll
    for (int u = 0; u < GNS_UNROLL; ++u) {
      const int pl = (k * GNS_UNROLL + u) * R + r;
      if (pl < npix) {
        const BF8 vx = st[(u * R + r) * CV + cv], vd = sd[(u * R + r) * CV + cv];
        BF8 vo, o;
        if (acc) vo = *reinterpret_cast<const BF8*>(ob + 1LL * pl * lddx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xv = __bfloat1622float2(vx.v[i]);
          const float2 d = gn_dxhat2(xv, __bfloat1622float2(vd.v[i]), kk, i, silu);
          float2 g = __ffma2_rn(d, make_float2(kk.r[i], kk.r[i]), __ffma2_rn(xv, make_float2(C1[i], C1[i]), make_float2(C2[i], C2[i])));
          if (acc) g = __fadd2_rn(g, __bfloat1622float2(vo.v[i]));
          o.v[i] = __floats2bfloat162_rn(g.x, g.y);
        }
        *reinterpret_cast<BF8*>(ob + 1LL * pl * lddx) = o;
      }
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------- cluster GroupNorm (no grid barrier)
// GroupNorm statistics are independent per (image, group), so the single-launch form needs no grid-wide rendezvous: a
// CLUSTER of K CTAs owns one (image, group), CTA k stages its slice of the pixels -- only the group's C/G channels of
// each, a 20-160 byte segment per pixel -- in shared memory, the K partial sums meet through distributed shared memory
// behind one hardware cluster barrier, and every CTA normalises its slice from shared memory.  Unlike the grid-barrier
// kernels above this cannot starve when several kernels share the GPU, and its critical path is one cluster barrier
// instead of ~150 CTAs polling one L2 word.  Thread mapping: a lane owns one bf16 PAIR position j of the group
// (PP = C/G/2 pairs per pixel) and, with PP <= 16, one of 32/PP pixels of the warp's round.
struct GNClusterShape {
  int N, HW, C, G, K;     // K = cluster size (CTAs per (image, group))
  long long ld;
  int pix_per_cta;
};
__device__ __forceinline__ void gnc_map(int PP, int lane, int& ppw, int& sp, int& j0, int& jstep) {
  // PP <= 32: 32 / PP pixels per warp round, lane -> (sub-pixel sp, pair j0); PP > 32: one pixel per round, lanes stride the pairs
  if (PP <= 32) {
    ppw = 32 / PP, sp = lane / PP, j0 = lane - sp * PP, jstep = PP;  // jstep = PP: a single pair per lane
    if (sp >= ppw) sp = -1;                                            // idle lane
  } else {
    ppw = 1, sp = 0, j0 = lane, jstep = 32;
  }
}
// block-wide sum of two values -> all threads; `red` holds 64 floats
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a), b = warp_sum(b);
  __syncthreads();
  if (lane == 0) red[w] = a, red[32 + w] = b;
  __syncthreads();
  float ra = lane < nw ? red[lane] : 0.f, rb = lane < nw ? red[32 + lane] : 0.f;
  a = warp_sum(ra), b = warp_sum(rb);
}
// sum of the K CTAs' (a, b) through distributed shared memory, fixed order; `slot` = 2 floats of THIS CTA's shared memory
__device__ __forceinline__ void cluster_sum2(float& a, float& b, float* slot, int K) {
  if (threadIdx.x == 0) slot[0] = a, slot[1] = b;
  ptx::cluster_sync_all();
  double sa = 0, sb = 0;
  for (int k = 0; k < K; ++k) {
    const uint32_t addr = ptx::mapa_u32(slot, static_cast<uint32_t>(k));
    sa += ptx::ld_shared_cluster_f32(addr), sb += ptx::ld_shared_cluster_f32(addr + 4);
  }
  a = static_cast<float>(sa), b = static_cast<float>(sb);
  ptx::cluster_sync_all();  // nobody leaves (or reuses the slot) while a peer may still read it
}

// shared memory: [slab: pix_per_cta x C/G bf16][red 64 floats][slot 2 floats (+2 pad)]
__global__ void __launch_bounds__(256) gn_cluster_fwd_kernel(const bf16* __restrict__ x, GNClusterShape s, float eps,
                                                             float* __restrict__ stats_out, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int silu, bf16* __restrict__ y,
                                                             long long ldy) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) uint8_t gnc_smem[];
  const int cpg = s.C / s.G, PP = cpg >> 1;
  const int cl = blockIdx.x / s.K, rank = static_cast<int>(ptx::cluster_ctarank());
  const int n = cl / s.G, g = cl % s.G;
  const int p0 = rank * s.pix_per_cta, npix = max(0, min(s.HW, p0 + s.pix_per_cta) - p0);
  __nv_bfloat162* slab = reinterpret_cast<__nv_bfloat162*>(gnc_smem);
  float* red = reinterpret_cast<float*>(gnc_smem + ((static_cast<size_t>(s.pix_per_cta) * cpg * 2 + 15) & ~size_t(15)));
  float* slot = red + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int ppw, sp, j0, jstep;
  gnc_map(PP, lane, ppw, sp, j0, jstep);
  const bf16* xb = x + (1LL * n * s.HW + p0) * s.ld + g * cpg;
  float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
  if (sp >= 0)
    for (int pl = warp * ppw + sp; pl < npix; pl += nw * ppw)
      for (int j = j0; j < PP; j += jstep) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(xb + 1LL * pl * s.ld + 2 * j);
        slab[pl * PP + j] = v;
        const float2 t = __bfloat1622float2(v);
        sum2 = __fadd2_rn(sum2, t);
        sq2 = __ffma2_rn(t, t, sq2);
      }
  float a = sum2.x + sum2.y, b = sq2.x + sq2.y;
  block_sum2(a, b, red);
  cluster_sum2(a, b, slot, s.K);
  const double m = 1.0 * s.HW * cpg, mean_d = a / m;
  double var = b / m - mean_d * mean_d;
  if (var < 0) var = 0;
  const float mean = static_cast<float>(mean_d), rstd = static_cast<float>(1.0 / sqrt(var + eps));
  if (rank == 0 && threadIdx.x == 0) stats_out[2 * (n * s.G + g)] = mean, stats_out[2 * (n * s.G + g) + 1] = rstd;
  bf16* yb = y + (1LL * n * s.HW + p0) * ldy + g * cpg;
  if (sp >= 0)
    for (int j = j0; j < PP; j += jstep) {  // (a single j per lane when PP <= 32)
      const int c = g * cpg + 2 * j;
      const float2 sc = make_float2(rstd * gamma[c], rstd * gamma[c + 1]);
      const float2 sf = make_float2(beta[c] - mean * sc.x, beta[c + 1] - mean * sc.y);
      for (int pl = warp * ppw + sp; pl < npix; pl += nw * ppw) {
        float2 h = __ffma2_rn(__bfloat1622float2(slab[pl * PP + j]), sc, sf);
        if (silu) h = silu2(bf16r2(h));
        *reinterpret_cast<__nv_bfloat162*>(yb + 1LL * pl * ldy + 2 * j) = __floats2bfloat162_rn(h.x, h.y);
      }
    }
}

// shared memory: [x slab][dy slab][red 64][slot 2]
__global__ void __launch_bounds__(256) gn_cluster_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, long long lddy,
                                                             GNClusterShape s, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int silu, bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) uint8_t gnc_smem[];
  const int cpg = s.C / s.G, PP = cpg >> 1;
  const int cl = blockIdx.x / s.K, rank = static_cast<int>(ptx::cluster_ctarank());
  const int n = cl / s.G, g = cl % s.G;
  const int p0 = rank * s.pix_per_cta, npix = max(0, min(s.HW, p0 + s.pix_per_cta) - p0);
  const size_t slab_bytes = (static_cast<size_t>(s.pix_per_cta) * cpg * 2 + 15) & ~size_t(15);
  __nv_bfloat162* slab_x = reinterpret_cast<__nv_bfloat162*>(gnc_smem);
  float2* slab_d = reinterpret_cast<float2*>(gnc_smem + slab_bytes);  // dxhat kept in fp32: no second activation-derivative pass
  float* red = reinterpret_cast<float*>(gnc_smem + slab_bytes + static_cast<size_t>(s.pix_per_cta) * cpg * 4);
  float* slot = red + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int ppw, sp, j0, jstep;
  gnc_map(PP, lane, ppw, sp, j0, jstep);
  const float mean = stats[2 * (n * s.G + g)], rstd = stats[2 * (n * s.G + g) + 1], Mr = -mean * rstd;
  const bf16* xb = x + (1LL * n * s.HW + p0) * s.ld + g * cpg;
  const bf16* db = dy + (1LL * n * s.HW + p0) * lddy + g * cpg;
  float2 sa2 = make_float2(0.f, 0.f), sb2 = make_float2(0.f, 0.f);
  if (sp >= 0)
    for (int j = j0; j < PP; j += jstep) {
      const int c = g * cpg + 2 * j;
      const float2 G2 = make_float2(gamma[c], gamma[c + 1]);
      const float2 A2 = make_float2(rstd * G2.x, rstd * G2.y), B2 = make_float2(beta[c] - mean * A2.x, beta[c + 1] - mean * A2.y);
      for (int pl = warp * ppw + sp; pl < npix; pl += nw * ppw) {
        const __nv_bfloat162 vx = *reinterpret_cast<const __nv_bfloat162*>(xb + 1LL * pl * s.ld + 2 * j);
        const float2 xv = __bfloat1622float2(vx);
        float2 d = __fmul2_rn(__bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(db + 1LL * pl * lddy + 2 * j)), G2);
        if (silu) d = __fmul2_rn(d, silu_grad2(bf16r2(__ffma2_rn(xv, A2, B2))));
        const float2 xh = __ffma2_rn(xv, make_float2(rstd, rstd), make_float2(Mr, Mr));
        slab_x[pl * PP + j] = vx;
        slab_d[pl * PP + j] = d;
        sa2 = __fadd2_rn(sa2, d);
        sb2 = __ffma2_rn(d, xh, sb2);
      }
    }
  float a = sa2.x + sa2.y, b = sb2.x + sb2.y;
  block_sum2(a, b, red);
  cluster_sum2(a, b, slot, s.K);
  const float m = 1.f * s.HW * cpg, m1 = a / m, m2 = b / m;
  const float C1 = -rstd * rstd * m2, C2 = -rstd * m1 - Mr * rstd * m2;
  bf16* ob = dx + (1LL * n * s.HW + p0) * lddx + g * cpg;
  if (sp >= 0)
    for (int j = j0; j < PP; j += jstep)
      for (int pl = warp * ppw + sp; pl < npix; pl += nw * ppw) {
        const float2 xv = __bfloat1622float2(slab_x[pl * PP + j]);
        float2 gx = __ffma2_rn(slab_d[pl * PP + j], make_float2(rstd, rstd), __ffma2_rn(xv, make_float2(C1, C1), make_float2(C2, C2)));
        __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(ob + 1LL * pl * lddx + 2 * j);
        if (acc) gx = __fadd2_rn(gx, __bfloat1622float2(*dst));
        *dst = __floats2bfloat162_rn(gx.x, gx.y);
      }
}

// =========================================================================== AutoencoderTiny element-wise pieces
// g = (y > 0) ? g : 0 in place: backward of the ReLU that produced y (applied once y's gradient is complete)
__global__ void relu_mask_kernel(bf16* __restrict__ g, long long ldg, const bf16* __restrict__ y, long long ldy,
                                 long long rows, int C) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int v = static_cast<int>(i % cv) * 8;
    float fy[8], fg[8];
    bf8_to_f(*reinterpret_cast<const BF8*>(y + r * ldy + v), fy);
    bf8_to_f(*reinterpret_cast<const BF8*>(g + r * ldg + v), fg);
#pragma unroll
    for (int k = 0; k < 8; ++k) fg[k] = fy[k] > 0.f ? fg[k] : 0.f;
    *reinterpret_cast<BF8*>(g + r * ldg + v) = f_to_bf8(fg);
  }
}
// DecoderTiny input clamp: y = tanh(x / m) * m (each step rounded to bf16 like the reference's bf16 tensors);
// backward: dx (+)= dy * (1 - tanh(x / m)^2).  C <= 8 channels per row (latents), one thread per element.
__global__ void tanh_clamp_fwd_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                      long long rows, int C, float m) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  const float t = bf16r(tanhf(bf16r(__bfloat162float(x[r * ldx + c]) / m)));
  y[r * ldy + c] = __float2bfloat16(t * m);
}
__global__ void tanh_clamp_bwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ dy, long long lddy,
                                      bf16* __restrict__ dx, long long lddx, long long rows, int C, float m, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  const float t = tanhf(__bfloat162float(x[r * ldx + c]) / m);
  float g = __bfloat162float(dy[r * lddy + c]) * (1.f - t * t);
  if (acc) g += __bfloat162float(dx[r * lddx + c]);
  dx[r * lddx + c] = __float2bfloat16(g);
}
// EncoderTiny input map: y = (x + 1) / 2, rounded to bf16 after each step
__global__ void unit_range_kernel(const bf16* __restrict__ x, long long ldx, bf16* __restrict__ y, long long ldy,
                                  long long rows, int C) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = static_cast<int>(i % C);
  y[r * ldy + c] = __float2bfloat16(bf16r(__bfloat162float(x[r * ldx + c]) + 1.f) * 0.5f);
}
// v = scale * v + shift on a small fp32 vector (biases folded into a scaled / shifted epilogue)
__global__ void vec_affine_kernel(float* __restrict__ v, long long n, float scale, float shift) {
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i < n) v[i] = scale * v[i] + shift;
}

// =========================================================================== LayerNorm (one warp per row)
constexpr int LN_MAXV = 5;  // supports d <= 32*8*5 = 1280
// eight consecutive parameters as two 16-byte loads (parameter vectors are 256-byte aligned arena blocks)
__device__ __forceinline__ void ld_f8(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

__global__ void ln_fwd_kernel(const bf16* __restrict__ x, long long ldx, int rows, int d, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, bf16* __restrict__ y, long long ldy,
                              float* __restrict__ stats) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = d >> 3;
  float f[LN_MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      bf8_to_f(*reinterpret_cast<const BF8*>(x + row * ldx + v * 8), f[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[k][i];
    }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = f[k][i] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  if (lane == 0) stats[2 * row] = mean, stats[2 * row + 1] = rstd;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      float o[8], ga[8], be[8];
      ld_f8(gamma + v * 8, ga);
      ld_f8(beta + v * 8, be);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (f[k][i] - mean) * rstd * ga[i] + be[i];
      *reinterpret_cast<BF8*>(y + row * ldy + v * 8) = f_to_bf8(o);
    }
  }
}

__global__ void ln_bwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ dy, long long lddy,
                              int rows, int d, const float* __restrict__ gamma, const float* __restrict__ stats,
                              bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = d >> 3;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float xh[LN_MAXV][8], dh[LN_MAXV][8];
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      float fx[8], fd[8], ga[8];
      bf8_to_f(*reinterpret_cast<const BF8*>(x + row * ldx + v * 8), fx);
      bf8_to_f(*reinterpret_cast<const BF8*>(dy + row * lddy + v * 8), fd);
      ld_f8(gamma + v * 8, ga);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[k][i] = (fx[i] - mean) * rstd;
        dh[k][i] = fd[i] * ga[i];
        a += dh[k][i];
        b += dh[k][i] * xh[k][i];
      }
    }
  }
  a = warp_sum(a) / d;
  b = warp_sum(b) / d;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    int v = lane + 32 * k;
    if (v < nv) {
      float o[8];
      if (acc) bf8_to_f(*reinterpret_cast<const BF8*>(dx + row * lddx + v * 8), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float g = rstd * (dh[k][i] - a - xh[k][i] * b);
        o[i] = acc ? o[i] + g : g;
      }
      *reinterpret_cast<BF8*>(dx + row * lddx + v * 8) = f_to_bf8(o);
    }
  }
}

// =========================================================================== GEGLU
// x [rows, 2F] -> y [rows, F] = x[:, :F] * gelu(x[:, F:])
__global__ void geglu_fwd_kernel(const bf16* __restrict__ x, long long ldx, long long rows, int F, bf16* __restrict__ y,
                                 long long ldy) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int fv = F >> 3;
  const long long total = rows * fv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    long long r = i / fv;
    int v = i % fv;
    float a[8], g[8], o[8];
    bf8_to_f(*reinterpret_cast<const BF8*>(x + r * ldx + v * 8), a);
    bf8_to_f(*reinterpret_cast<const BF8*>(x + r * ldx + F + v * 8), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = a[k] * bf16r(gelu_erf(g[k]));
    *reinterpret_cast<BF8*>(y + r * ldy + v * 8) = f_to_bf8(o);
  }
}
__global__ void geglu_bwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ dy, long long lddy,
                                 long long rows, int F, bf16* __restrict__ dx, long long lddx, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int fv = F >> 3;
  const long long total = rows * fv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    long long r = i / fv;
    int v = i % fv;
    float a[8], g[8], d[8], oa[8], og[8];
    bf8_to_f(*reinterpret_cast<const BF8*>(x + r * ldx + v * 8), a);
    bf8_to_f(*reinterpret_cast<const BF8*>(x + r * ldx + F + v * 8), g);
    bf8_to_f(*reinterpret_cast<const BF8*>(dy + r * lddy + v * 8), d);
    if (acc) {
      bf8_to_f(*reinterpret_cast<const BF8*>(dx + r * lddx + v * 8), oa);
      bf8_to_f(*reinterpret_cast<const BF8*>(dx + r * lddx + F + v * 8), og);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float ga = d[k] * bf16r(gelu_erf(g[k]));
      float gg = d[k] * a[k] * gelu_erf_grad(g[k]);
      oa[k] = acc ? oa[k] + ga : ga;
      og[k] = acc ? og[k] + gg : gg;
    }
    *reinterpret_cast<BF8*>(dx + r * lddx + v * 8) = f_to_bf8(oa);
    *reinterpret_cast<BF8*>(dx + r * lddx + F + v * 8) = f_to_bf8(og);
  }
}

// =========================================================================== softmax over rows of fp32 scores
// S [rows, ld] fp32 (already scaled), T valid columns -> P [rows, ld] bf16.  One block per row, the row is
// cached in shared memory.  ld is a multiple of 8 so rows are 16-byte aligned; T need not be.
__device__ __forceinline__ void st_bf16x4(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&lo), o.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = o;
}
__device__ __forceinline__ float4 ld_bf16x4(const bf16* p) {
  uint2 pr = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pr.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&pr.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__global__ void softmax_fwd_kernel(const float* __restrict__ S, bf16* __restrict__ P, int T, long long ld) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ float row[];  // T4 floats + 32
  const int T4 = T & ~3;
  float* red = row + ((T + 3) & ~3);
  const float* s = S + 1LL * blockIdx.x * ld;
  bf16* p = P + 1LL * blockIdx.x * ld;
  float m = -INFINITY;
  for (int i = threadIdx.x * 4; i < T4; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(s + i);
    *reinterpret_cast<float4*>(row + i) = v;
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int i = T4 + threadIdx.x; i < T; i += blockDim.x) {
    row[i] = s[i];
    m = fmaxf(m, row[i]);
  }
  m = block_max(m, red);
  float sum = 0.f;
  for (int i = threadIdx.x * 4; i < T4; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(row + i);
    v.x = __expf(v.x - m), v.y = __expf(v.y - m), v.z = __expf(v.z - m), v.w = __expf(v.w - m);
    *reinterpret_cast<float4*>(row + i) = v;
    sum += v.x + v.y + v.z + v.w;
  }
  for (int i = T4 + threadIdx.x; i < T; i += blockDim.x) {
    row[i] = __expf(row[i] - m);
    sum += row[i];
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int i = threadIdx.x * 4; i < T4; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(row + i);
    st_bf16x4(p + i, v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
  for (int i = T4 + threadIdx.x; i < T; i += blockDim.x) p[i] = __float2bfloat16(row[i] * inv);
}
// dS = P * (dP - sum_j P dP) * scale, written in place over P (bf16).
__global__ void softmax_bwd_kernel(const float* __restrict__ dP, bf16* __restrict__ P, int T, long long ld,
                                   float scale) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ float row[];
  const int T4 = T & ~3;
  float* red = row + ((T + 3) & ~3);
  const float* d = dP + 1LL * blockIdx.x * ld;
  bf16* p = P + 1LL * blockIdx.x * ld;
  float dot = 0.f;
  for (int i = threadIdx.x * 4; i < T4; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<const float4*>(d + i);
    float4 a = ld_bf16x4(p + i);
    *reinterpret_cast<float4*>(row + i) = v;
    dot += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
  }
  for (int i = T4 + threadIdx.x; i < T; i += blockDim.x) {
    row[i] = d[i];
    dot += __bfloat162float(p[i]) * row[i];
  }
  dot = block_sum(dot, red);
  for (int i = threadIdx.x * 4; i < T4; i += blockDim.x * 4) {
    float4 v = *reinterpret_cast<float4*>(row + i);
    float4 a = ld_bf16x4(p + i);
    st_bf16x4(p + i, a.x * (v.x - dot) * scale, a.y * (v.y - dot) * scale, a.z * (v.z - dot) * scale,
              a.w * (v.w - dot) * scale);
  }
  for (int i = T4 + threadIdx.x; i < T; i += blockDim.x)
    p[i] = __float2bfloat16(__bfloat162float(p[i]) * (row[i] - dot) * scale);
}

// =========================================================================== cross-attention with 2 key tokens
// q [rows, d] (heads x 64), kc/vc [2, d] fp32 (step-invariant).  One warp per (row, head); lane owns 2 channels.
__global__ void xattn2_fwd_kernel(const bf16* __restrict__ q, long long ldq, long long rows, int heads,
                                  const float* __restrict__ kc, const float* __restrict__ vc, float scale,
                                  bf16* __restrict__ o, long long ldo) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long w = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const long long r = w / heads;
  const int h = w % heads, d = heads * 64, c = h * 64 + lane * 2;
  float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q + r * ldq + c));
  float s0 = qv.x * kc[c] + qv.y * kc[c + 1], s1 = qv.x * kc[d + c] + qv.y * kc[d + c + 1];
  s0 = warp_sum(s0) * scale, s1 = warp_sum(s1) * scale;
  float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
  float p0 = bf16r(e0 * inv), p1 = bf16r(e1 * inv);
  *reinterpret_cast<__nv_bfloat162*>(o + r * ldo + c) =
      __floats2bfloat162_rn(p0 * vc[c] + p1 * vc[d + c], p0 * vc[c + 1] + p1 * vc[d + c + 1]);
}
__global__ void xattn2_bwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ dout,
                                  long long lddo, long long rows, int heads, const float* __restrict__ kc,
                                  const float* __restrict__ vc, float scale, bf16* __restrict__ dq, long long lddq,
                                  int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long w = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const long long r = w / heads;
  const int h = w % heads, d = heads * 64, c = h * 64 + lane * 2;
  float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q + r * ldq + c));
  float2 gv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dout + r * lddo + c));
  float s0 = qv.x * kc[c] + qv.y * kc[c + 1], s1 = qv.x * kc[d + c] + qv.y * kc[d + c + 1];
  float dp0 = gv.x * vc[c] + gv.y * vc[c + 1], dp1 = gv.x * vc[d + c] + gv.y * vc[d + c + 1];
  s0 = warp_sum(s0) * scale, s1 = warp_sum(s1) * scale;
  dp0 = warp_sum(dp0), dp1 = warp_sum(dp1);
  float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
  float p0 = e0 * inv, p1 = e1 * inv, dot = p0 * dp0 + p1 * dp1;
  float ds0 = p0 * (dp0 - dot) * scale, ds1 = p1 * (dp1 - dot) * scale;
  float gx = ds0 * kc[c] + ds1 * kc[d + c], gy = ds0 * kc[c + 1] + ds1 * kc[d + c + 1];
  __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(dq + r * lddq + c);
  if (acc) {
    float2 old = __bfloat1622float2(*dst);
    gx += old.x, gy += old.y;
  }
  *dst = __floats2bfloat162_rn(gx, gy);
}

// =========================================================================== collapsed cross-attention block
// attn2 of a BasicTransformerBlock attends to the 2 tokens of the (constant) empty-prompt embedding, so its four ops
//   n2 = LN2(h);  q = n2 Wq^T;  o = softmax_2(scale q k^T) v  (per head);  h' = h + o Wo^T + bo
// collapse algebraically into two skinny products with step-invariant matrices (built once in prepare()):
//   S = n2 At^T   with At[h*2+j][:] = scale * Wq[head h rows]^T k_{h,j}      [2H, d]
//   h' = h + bo + P U  with U[h*2+j][:] = Wo[:, head h cols] v_{h,j}          [2H, d],  P = pairwise softmax(S)
// i.e. 2*2H*d MACs per token instead of 2*d*d, and ONE kernel (LayerNorm, both products, softmax, bias, residual)
// instead of four.  One warp per token; the row lives in registers.
constexpr int XA_MAXC = 40;  // 2 * heads <= 40

template <int NV>  // NV = ceil(d / 256): 8-element vectors per lane
__global__ void xattn_fused_fwd_kernel(const bf16* __restrict__ h, long long ldh, long long rows, int d, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ At, const float* __restrict__ U,
                                       const float* __restrict__ bo, bf16* __restrict__ out, long long ldo,
                                       float* __restrict__ stats) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long row = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = d >> 3;
  float x[NV][8], n2[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      bf8_to_f(*reinterpret_cast<const BF8*>(h + row * ldh + v * 8), x[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += x[k][i];
    }
  }
  const float mean = warp_sum(sum) / d;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = x[k][i] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
  if (lane == 0) stats[2 * row] = mean, stats[2 * row + 1] = rstd;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) n2[k][i] = bf16r((x[k][i] - mean) * rstd * gamma[v * 8 + i] + beta[v * 8 + i]);
    }
  }
  float o[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[k][i] = x[k][i] + bo[v * 8 + i];
    }
  }
  for (int c = 0; c < C; c += 2) {  // one head (2 key tokens) at a time
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float4* a0 = reinterpret_cast<const float4*>(At + 1LL * c * d + v * 8);
        const float4* a1 = reinterpret_cast<const float4*>(At + 1LL * (c + 1) * d + v * 8);
        const float4 a00 = __ldg(a0), a01 = __ldg(a0 + 1), a10 = __ldg(a1), a11 = __ldg(a1 + 1);
        s0 += n2[k][0] * a00.x + n2[k][1] * a00.y + n2[k][2] * a00.z + n2[k][3] * a00.w + n2[k][4] * a01.x +
              n2[k][5] * a01.y + n2[k][6] * a01.z + n2[k][7] * a01.w;
        s1 += n2[k][0] * a10.x + n2[k][1] * a10.y + n2[k][2] * a10.z + n2[k][3] * a10.w + n2[k][4] * a11.x +
              n2[k][5] * a11.y + n2[k][6] * a11.z + n2[k][7] * a11.w;
      }
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1);
    const float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float4* u0 = reinterpret_cast<const float4*>(U + 1LL * c * d + v * 8);
        const float4* u1 = reinterpret_cast<const float4*>(U + 1LL * (c + 1) * d + v * 8);
        const float4 u00 = __ldg(u0), u01 = __ldg(u0 + 1), u10 = __ldg(u1), u11 = __ldg(u1 + 1);
        o[k][0] += p0 * u00.x + p1 * u10.x, o[k][1] += p0 * u00.y + p1 * u10.y;
        o[k][2] += p0 * u00.z + p1 * u10.z, o[k][3] += p0 * u00.w + p1 * u10.w;
        o[k][4] += p0 * u01.x + p1 * u11.x, o[k][5] += p0 * u01.y + p1 * u11.y;
        o[k][6] += p0 * u01.z + p1 * u11.z, o[k][7] += p0 * u01.w + p1 * u11.w;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) *reinterpret_cast<BF8*>(out + row * ldo + v * 8) = f_to_bf8(o[k]);
  }
}

// Backward: dh (+)= dy + LN2_bwd( dS At ),  dS = pairwise-softmax-bwd(P, dy U^T).  Recomputes n2, S, P from h.
template <int NV>
__global__ void xattn_fused_bwd_kernel(const bf16* __restrict__ h, long long ldh, const bf16* __restrict__ dy,
                                       long long lddy, long long rows, int d, int C, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ At,
                                       const float* __restrict__ U, const float* __restrict__ stats,
                                       bf16* __restrict__ dh, long long lddh, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long row = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = d >> 3;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  float xh[NV][8], n2[NV][8], g[NV][8], dn[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      float x[8];
      bf8_to_f(*reinterpret_cast<const BF8*>(h + row * ldh + v * 8), x);
      bf8_to_f(*reinterpret_cast<const BF8*>(dy + row * lddy + v * 8), g[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[k][i] = (x[i] - mean) * rstd;
        n2[k][i] = bf16r(xh[k][i] * gamma[v * 8 + i] + beta[v * 8 + i]);
        dn[k][i] = 0.f;
      }
    }
  }
  for (int c = 0; c < C; c += 2) {
    float s0 = 0.f, s1 = 0.f, dp0 = 0.f, dp1 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float* a0 = At + 1LL * c * d + v * 8;
        const float* a1 = a0 + d;
        const float* u0 = U + 1LL * c * d + v * 8;
        const float* u1 = u0 + d;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0 += n2[k][i] * __ldg(a0 + i), s1 += n2[k][i] * __ldg(a1 + i);
          dp0 += g[k][i] * __ldg(u0 + i), dp1 += g[k][i] * __ldg(u1 + i);
        }
      }
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1), dp0 = warp_sum(dp0), dp1 = warp_sum(dp1);
    const float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv, dot = p0 * dp0 + p1 * dp1;
    const float ds0 = p0 * (dp0 - dot), ds1 = p1 * (dp1 - dot);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float* a0 = At + 1LL * c * d + v * 8;
        const float* a1 = a0 + d;
#pragma unroll
        for (int i = 0; i < 8; ++i) dn[k][i] += ds0 * __ldg(a0 + i) + ds1 * __ldg(a1 + i);
      }
    }
  }
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dn[k][i] *= gamma[v * 8 + i];
        a += dn[k][i];
        b += dn[k][i] * xh[k][i];
      }
    }
  }
  a = warp_sum(a) / d, b = warp_sum(b) / d;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int v = lane + 32 * k;
    if (v < nv) {
      float o[8];
      if (acc) bf8_to_f(*reinterpret_cast<const BF8*>(dh + row * lddh + v * 8), o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float gx = g[k][i] + rstd * (dn[k][i] - a - xh[k][i] * b);
        o[i] = acc ? o[i] + gx : gx;
      }
      *reinterpret_cast<BF8*>(dh + row * lddh + v * 8) = f_to_bf8(o);
    }
  }
}

// ---- block-parallel form (the one the engine uses).  The one-warp-per-token kernels above serialise 2H dot products of
// length d behind L2-latency loads, which is slow exactly where tokens are few (9x12 ... 18x24 maps, d = 1280, 2H = 40).
// Here a 256-thread CTA owns RB tokens and every phase is spread over all of its warps:
//   A  LayerNorm of the RB rows (warp per row)                              -> shared memory
//   B  scores S[r][c] = n2[r] . At[c]: a warp takes two columns per round, their At rows held in registers across the
//      RB rows; the 2 RB partial sums of a lane are reduced with ONE butterfly (2 RB + 4 shuffles instead of 10 RB)
//   C  pairwise softmax over the 2 key tokens of each head (thread per (row, head))
//   D  out[r] = h[r] + bo + sum_c P[r][c] U[c]: thread per 8-channel vector, all RB rows (U is read once per CTA)
// Backward recomputes n2 / S / P from h and the saved LayerNorm statistics, then
//   dP = dy U^T (as B),  dS = softmax-bwd(P, dP),  dn = dS At (as D),  dh = dy + LayerNorm-bwd(dn).
// K4 = float4 chunks per lane and row (d <= 128 K4): the loops are sized for the width class instead of predicated --
// ncu on the first version (every loop sized for d = 1280): 13.5 M warp instructions for 6912 x 320 tokens, issue bound.
constexpr int XB_THREADS = 256;
constexpr int XB_LDS = XA_MAXC + 1;  // row stride of the small score arrays in shared memory

// Sum each of the NVAL per-lane values over the 32 lanes; afterwards lane l holds the total of value (l >> (5 - log2 NVAL)).
template <int NVAL>
__device__ __forceinline__ float warp_multi_sum(float (&v)[NVAL], int lane) {
  int o = 16;
#pragma unroll
  for (int half = NVAL / 2; half >= 1; half /= 2, o /= 2) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half], keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  float t = v[0];
  for (; o >= 1; o /= 2) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}
// S[r][c], S[r][c1] for r < RB from rows_s [RB][d] (shared) and the matrix rows Mx[c], Mx[c1] (global, L2 / L1)
template <int RB, int K4>
__device__ __forceinline__ void xb_two_columns(const float* __restrict__ rows_s, const float* __restrict__ Mx, int d, int nq, int c, int c1,
                                               int C, int lane, float* __restrict__ dst) {
  float4 a0[K4], a1[K4];
#pragma unroll
  for (int k = 0; k < K4; ++k) {
    const int qd = lane + 32 * k;
    if (qd < nq) {
      a0[k] = __ldg(reinterpret_cast<const float4*>(Mx + 1LL * c * d) + qd);
      a1[k] = __ldg(reinterpret_cast<const float4*>(Mx + 1LL * min(c1, C - 1) * d) + qd);
    } else {
      a0[k] = a1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  float v[2 * RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < K4; ++k) {
      const int qd = min(lane + 32 * k, nq - 1);  // (lanes past the row read a valid address; their a0 / a1 are zero)
      const float4 x = reinterpret_cast<const float4*>(rows_s + r * d)[qd];
      s0 += x.x * a0[k].x + x.y * a0[k].y + x.z * a0[k].z + x.w * a0[k].w;
      s1 += x.x * a1[k].x + x.y * a1[k].y + x.z * a1[k].z + x.w * a1[k].w;
    }
    v[2 * r] = s0, v[2 * r + 1] = s1;
  }
  const float tot = warp_multi_sum<2 * RB>(v, lane);
  constexpr int SH = (RB == 8 ? 1 : (RB == 4 ? 2 : 3));  // lane l holds value l >> SH
  if ((lane & ((1 << SH) - 1)) == 0) {
    const int idx = lane >> SH, r = idx >> 1, cc = (idx & 1) ? c1 : c;
    if (cc < C) dst[r * XB_LDS + cc] = tot;
  }
}

template <int RB, int K4>
__global__ void __launch_bounds__(XB_THREADS) xattn_block_fwd_kernel(const bf16* __restrict__ h, long long ldh, int rows, int d, int C,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ At, const float* __restrict__ U,
                                                                     const float* __restrict__ bo, bf16* __restrict__ out,
                                                                     long long ldo, float* __restrict__ stats) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) float xb_smem[];
  float* n2s = xb_smem;            // [RB][d]
  float* Ss = xb_smem + RB * d;    // [RB][XB_LDS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = XB_THREADS / 32, NV = (K4 + 1) / 2;
  const int row0 = blockIdx.x * RB, nv = d >> 3, nq = d >> 2;
  // ---- A: LayerNorm (eps 1e-5), rounded to bf16 like the reference's bf16 LayerNorm output
  for (int r = warp; r < RB; r += NW) {
    const int row = row0 + r;
    float f[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv && row < rows) {
        bf8_to_f(*reinterpret_cast<const BF8*>(h + 1LL * row * ldh + v * 8), f[k]);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += f[k][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[k][i] = 0.f;
      }
    }
    const float mean = warp_sum(sum) / d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = f[k][i] - mean;
          q += t * t;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
    if (lane == 0 && row < rows) stats[2LL * row] = mean, stats[2LL * row + 1] = rstd;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int v = lane + 32 * k;
      if (v < nv) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
        const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = row < rows ? bf16r((f[k][i] - mean) * rstd * ga[i] + be[i]) : 0.f;
        float4* dst = reinterpret_cast<float4*>(n2s + r * d + v * 8);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]), dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
  __syncthreads();
  // ---- B: scores
  for (int c = warp; c < C; c += 2 * NW) xb_two_columns<RB, K4>(n2s, At, d, nq, c, c + NW, C, lane, Ss);
  __syncthreads();
  // ---- C: softmax over the two key tokens of each head
  for (int t = threadIdx.x; t < RB * (C >> 1); t += XB_THREADS) {
    const int r = t / (C >> 1), hd = t % (C >> 1);
    const float s0 = Ss[r * XB_LDS + 2 * hd], s1 = Ss[r * XB_LDS + 2 * hd + 1];
    const float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
    Ss[r * XB_LDS + 2 * hd] = e0 * inv, Ss[r * XB_LDS + 2 * hd + 1] = e1 * inv;
  }
  __syncthreads();
  // ---- D: output rows; a thread owns one 8-channel vector of ALL RB rows, so U crosses L2 -> SM once per CTA
  for (int v = threadIdx.x; v < nv; v += XB_THREADS) {
    float o[RB][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bo + v * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bo + v * 8 + 4));
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (row0 + r < rows) {
          bf8_to_f(*reinterpret_cast<const BF8*>(h + 1LL * (row0 + r) * ldh + v * 8), o[r]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) o[r][i] = 0.f;
        }
        o[r][0] += b0.x, o[r][1] += b0.y, o[r][2] += b0.z, o[r][3] += b0.w, o[r][4] += b1.x, o[r][5] += b1.y, o[r][6] += b1.z, o[r][7] += b1.w;
      }
    }
#pragma unroll 5
    for (int c = 0; c < C; ++c) {
      const float4 u0 = __ldg(reinterpret_cast<const float4*>(U + 1LL * c * d + v * 8)), u1 = __ldg(reinterpret_cast<const float4*>(U + 1LL * c * d + v * 8 + 4));
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float pc = Ss[r * XB_LDS + c];
        o[r][0] += pc * u0.x, o[r][1] += pc * u0.y, o[r][2] += pc * u0.z, o[r][3] += pc * u0.w;
        o[r][4] += pc * u1.x, o[r][5] += pc * u1.y, o[r][6] += pc * u1.z, o[r][7] += pc * u1.w;
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (row0 + r < rows) *reinterpret_cast<BF8*>(out + 1LL * (row0 + r) * ldo + v * 8) = f_to_bf8(o[r]);
  }
}

template <int RB, int K4>
__global__ void __launch_bounds__(XB_THREADS) xattn_block_bwd_kernel(const bf16* __restrict__ h, long long ldh, const bf16* __restrict__ dy,
                                                                     long long lddy, int rows, int d, int C,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const float* __restrict__ At, const float* __restrict__ U,
                                                                     const float* __restrict__ stats, bf16* __restrict__ dh,
                                                                     long long lddh, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  extern __shared__ __align__(16) float xb_smem[];
  float* xhs = xb_smem;                 // [RB][d] normalised rows
  float* gs = xhs + RB * d;             // [RB][d] dy
  float* n2s = gs + RB * d;             // [RB][d] bf16(xh gamma + beta); later reused for dn * gamma
  float* Ss = n2s + RB * d;             // [RB][XB_LDS] scores -> dS
  float* dPs = Ss + RB * XB_LDS;        // [RB][XB_LDS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = XB_THREADS / 32;
  const int row0 = blockIdx.x * RB, nv = d >> 3, nq = d >> 2;
  // ---- A: stage xh, n2 and dy
  for (int item = threadIdx.x; item < RB * nv; item += XB_THREADS) {
    const int r = item / nv, v = item % nv, row = row0 + r;
    float x[8], g[8], n2[8];
    if (row < rows) {
      const float mean = stats[2LL * row], rstd = stats[2LL * row + 1];
      bf8_to_f(*reinterpret_cast<const BF8*>(h + 1LL * row * ldh + v * 8), x);
      bf8_to_f(*reinterpret_cast<const BF8*>(dy + 1LL * row * lddy + v * 8), g);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = (x[i] - mean) * rstd;
        n2[i] = bf16r(x[i] * __ldg(gamma + v * 8 + i) + __ldg(beta + v * 8 + i));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = g[i] = n2[i] = 0.f;
    }
    float4* dx = reinterpret_cast<float4*>(xhs + r * d + v * 8);
    float4* dg = reinterpret_cast<float4*>(gs + r * d + v * 8);
    float4* dn = reinterpret_cast<float4*>(n2s + r * d + v * 8);
    dx[0] = make_float4(x[0], x[1], x[2], x[3]), dx[1] = make_float4(x[4], x[5], x[6], x[7]);
    dg[0] = make_float4(g[0], g[1], g[2], g[3]), dg[1] = make_float4(g[4], g[5], g[6], g[7]);
    dn[0] = make_float4(n2[0], n2[1], n2[2], n2[3]), dn[1] = make_float4(n2[4], n2[5], n2[6], n2[7]);
  }
  __syncthreads();
  // ---- B: S = n2 At^T and dP = dy U^T
  for (int c = warp; c < C; c += 2 * NW) xb_two_columns<RB, K4>(n2s, At, d, nq, c, c + NW, C, lane, Ss);
  for (int c = warp; c < C; c += 2 * NW) xb_two_columns<RB, K4>(gs, U, d, nq, c, c + NW, C, lane, dPs);
  __syncthreads();
  // ---- C: dS of the pairwise softmax (in place over S)
  for (int t = threadIdx.x; t < RB * (C >> 1); t += XB_THREADS) {
    const int r = t / (C >> 1), hd = t % (C >> 1);
    const float s0 = Ss[r * XB_LDS + 2 * hd], s1 = Ss[r * XB_LDS + 2 * hd + 1];
    const float dp0 = dPs[r * XB_LDS + 2 * hd], dp1 = dPs[r * XB_LDS + 2 * hd + 1];
    const float m = fmaxf(s0, s1), e0 = __expf(s0 - m), e1 = __expf(s1 - m), inv = 1.f / (e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv, dot = p0 * dp0 + p1 * dp1;
    Ss[r * XB_LDS + 2 * hd] = p0 * (dp0 - dot), Ss[r * XB_LDS + 2 * hd + 1] = p1 * (dp1 - dot);
  }
  __syncthreads();
  // ---- D: dn gamma = (dS At) gamma over n2's storage (dead by now); a thread owns one 8-channel vector of all rows
  for (int v = threadIdx.x; v < nv; v += XB_THREADS) {
    float o[RB][8];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) o[r][i] = 0.f;
#pragma unroll 5
    for (int c = 0; c < C; ++c) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(At + 1LL * c * d + v * 8)), a1 = __ldg(reinterpret_cast<const float4*>(At + 1LL * c * d + v * 8 + 4));
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float ds = Ss[r * XB_LDS + c];
        o[r][0] += ds * a0.x, o[r][1] += ds * a0.y, o[r][2] += ds * a0.z, o[r][3] += ds * a0.w;
        o[r][4] += ds * a1.x, o[r][5] += ds * a1.y, o[r][6] += ds * a1.z, o[r][7] += ds * a1.w;
      }
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float4* dst = reinterpret_cast<float4*>(n2s + r * d + v * 8);
      dst[0] = make_float4(o[r][0] * g0.x, o[r][1] * g0.y, o[r][2] * g0.z, o[r][3] * g0.w);
      dst[1] = make_float4(o[r][4] * g1.x, o[r][5] * g1.y, o[r][6] * g1.z, o[r][7] * g1.w);
    }
  }
  __syncthreads();
  // ---- E: LayerNorm backward per row, plus the residual path (dh = dy + ...)
  for (int r = warp; r < RB; r += NW) {
    const int row = row0 + r;
    if (row >= rows) continue;
    const float rstd = stats[2LL * row + 1];
    float a = 0.f, b = 0.f;
    for (int qd = lane; qd < nq; qd += 32) {
      const float4 t = reinterpret_cast<const float4*>(n2s + r * d)[qd], xh = reinterpret_cast<const float4*>(xhs + r * d)[qd];
      a += t.x + t.y + t.z + t.w;
      b += t.x * xh.x + t.y * xh.y + t.z * xh.z + t.w * xh.w;
    }
    a = warp_sum(a) / d, b = warp_sum(b) / d;
    for (int qd = lane; qd < nq; qd += 32) {  // 4 channels per lane and pass: 8-byte bf16 stores
      const float4 g = reinterpret_cast<const float4*>(gs + r * d)[qd], dn = reinterpret_cast<const float4*>(n2s + r * d)[qd];
      const float4 xh = reinterpret_cast<const float4*>(xhs + r * d)[qd];
      float4 o = make_float4(g.x + rstd * (dn.x - a - xh.x * b), g.y + rstd * (dn.y - a - xh.y * b), g.z + rstd * (dn.z - a - xh.z * b),
                             g.w + rstd * (dn.w - a - xh.w * b));
      bf16* dst = dh + 1LL * row * lddh + qd * 4;
      if (acc) {
        const float4 old = ld_bf16x4(dst);
        o.x += old.x, o.y += old.y, o.z += old.z, o.w += old.w;
      }
      st_bf16x4(dst, o.x, o.y, o.z, o.w);
    }
  }
}

// prepare-time: At[h*2+j][i] = scale * sum_r Wq[h*64+r][i] * kc[j][h*64+r];  U[h*2+j][o] = sum_r Wo[o][h*64+r] * vc[j][h*64+r]
__global__ void xattn_collapse_kernel(const bf16* __restrict__ Wq, long long ldq, const bf16* __restrict__ Wo, long long ldwo,
                                      const float* __restrict__ kc, const float* __restrict__ vc, int d, int heads,
                                      float scale, float* __restrict__ At, float* __restrict__ U) {
  ptx::pdl_wait();  // every kernel launched through launch_k must order itself after its predecessor
  ptx::pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over (c, e)
  const int C = 2 * heads;
  if (i >= C * d) return;
  const int c = i / d, e = i % d, hd = c >> 1, j = c & 1;
  float a = 0.f, u = 0.f;
  for (int r = 0; r < 64; ++r) {
    const int col = hd * 64 + r;
    a += __bfloat162float(Wq[1LL * col * ldq + e]) * kc[j * d + col];
    u += __bfloat162float(Wo[1LL * e * ldwo + col]) * vc[j * d + col];
  }
  At[i] = a * scale;
  U[i] = u;
}

// =========================================================================== resampling
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  return min(static_cast<int>(floorf(dst * scale)), in_size - 1);
}
// nearest upsample [N, h, w, C] -> [N, H, W, C] (torch F.interpolate(mode="nearest") index rule)
__global__ void upsample_nearest_fwd_kernel(const bf16* __restrict__ x, long long ldx, int N, int h, int w, int C,
                                            bf16* __restrict__ y, long long ldy, int H, int W) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = 1LL * N * H * W * cv;
  const float sh = static_cast<float>(h) / H, sw = static_cast<float>(w) / W;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int v = i % cv;
    long long p = i / cv;
    int X = p % W, Y = (p / W) % H, n = p / (1LL * W * H);
    int sy = nearest_src(Y, sh, h), sx = nearest_src(X, sw, w);
    *reinterpret_cast<BF8*>(y + p * ldy + v * 8) =
        *reinterpret_cast<const BF8*>(x + ((1LL * n * h + sy) * w + sx) * ldx + v * 8);
  }
}
// adjoint: dx[n, sy, sx] (+)= sum of dy over the destination pixels that read (sy, sx)
__global__ void upsample_nearest_bwd_kernel(const bf16* __restrict__ dy, long long lddy, int N, int h, int w, int C,
                                            bf16* __restrict__ dx, long long lddx, int H, int W, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = 1LL * N * h * w * cv;
  const float sh = static_cast<float>(h) / H, sw = static_cast<float>(w) / W;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int v = i % cv;
    long long p = i / cv;
    int sx = p % w, sy = (p / w) % h, n = p / (1LL * w * h);
    float o[8];
    if (acc)
      bf8_to_f(*reinterpret_cast<const BF8*>(dx + p * lddx + v * 8), o);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
    int y0 = max(0, static_cast<int>(sy / sh) - 2), y1 = min(H - 1, static_cast<int>((sy + 1) / sh) + 2);
    int x0 = max(0, static_cast<int>(sx / sw) - 2), x1 = min(W - 1, static_cast<int>((sx + 1) / sw) + 2);
    for (int Y = y0; Y <= y1; ++Y) {
      if (nearest_src(Y, sh, h) != sy) continue;
      for (int X = x0; X <= x1; ++X) {
        if (nearest_src(X, sw, w) != sx) continue;
        float f[8];
        bf8_to_f(*reinterpret_cast<const BF8*>(dy + ((1LL * n * H + Y) * W + X) * lddy + v * 8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += f[k];
      }
    }
    *reinterpret_cast<BF8*>(dx + p * lddx + v * 8) = f_to_bf8(o);
  }
}
// y[n, oy, ox] = x[n, 2*oy + off, 2*ox + off]   (stride-2 conv = stride-1 conv + this subsample)
__global__ void subsample2_fwd_kernel(const bf16* __restrict__ x, long long ldx, int N, int H, int W, int C, int off,
                                      bf16* __restrict__ y, long long ldy, int Ho, int Wo) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = 1LL * N * Ho * Wo * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int v = i % cv;
    long long p = i / cv;
    int ox = p % Wo, oy = (p / Wo) % Ho, n = p / (1LL * Wo * Ho);
    *reinterpret_cast<BF8*>(y + p * ldy + v * 8) =
        *reinterpret_cast<const BF8*>(x + ((1LL * n * H + 2 * oy + off) * W + 2 * ox + off) * ldx + v * 8);
  }
}
// adjoint: zero-insert
__global__ void subsample2_bwd_kernel(const bf16* __restrict__ dy, long long lddy, int N, int H, int W, int C, int off,
                                      bf16* __restrict__ dx, long long lddx, int Ho, int Wo, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = 1LL * N * H * W * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int v = i % cv;
    long long p = i / cv;
    int X = p % W, Y = (p / W) % H, n = p / (1LL * W * H);
    float o[8];
    if (acc)
      bf8_to_f(*reinterpret_cast<const BF8*>(dx + p * lddx + v * 8), o);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = 0.f;
    }
    int yy = Y - off, xx = X - off;
    if (yy >= 0 && xx >= 0 && !(yy & 1) && !(xx & 1) && (yy >> 1) < Ho && (xx >> 1) < Wo) {
      float f[8];
      bf8_to_f(*reinterpret_cast<const BF8*>(dy + ((1LL * n * Ho + (yy >> 1)) * Wo + (xx >> 1)) * lddy + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += f[k];
    }
    *reinterpret_cast<BF8*>(dx + p * lddx + v * 8) = f_to_bf8(o);
  }
}

// dst[rows, C] (+)= src[rows, C]  (channel-slice copy/accumulate with independent row strides)
__global__ void add_rows_kernel(const bf16* __restrict__ src, long long lds, bf16* __restrict__ dst, long long ldd,
                                long long rows, int C, int acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int cv = C >> 3;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    long long r = i / cv;
    int v = i % cv;
    BF8 s = *reinterpret_cast<const BF8*>(src + r * lds + v * 8);
    if (acc) {
      float a[8], b[8];
      bf8_to_f(s, a);
      bf8_to_f(*reinterpret_cast<const BF8*>(dst + r * ldd + v * 8), b);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += b[k];
      s = f_to_bf8(a);
    }
    *reinterpret_cast<BF8*>(dst + r * ldd + v * 8) = s;
  }
}

inline int ew_grid(long long work_items, int block = 256) {
  long long b = (work_items + block - 1) / block;
  long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mdc
