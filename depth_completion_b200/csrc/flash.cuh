// Fused self-attention for head_dim 64 on tcgen05 / TMEM, forward and backward (the UNet's attn1 blocks).
//
// Replaces the unfused S = QK^T -> softmax -> PV chain (which streams a T x T matrix through HBM) for the
// reference's F.scaled_dot_product_attention call (diffusers AttnProcessor2_0).
//
// All three kernels share one structure: a CTA owns a 128-row tile ("M side"), streams 64-row tiles of the other
// side through a TMA pipeline, and alternates   MMA group 1 (scores)  ->  softmax / elementwise threads  ->  MMA group
// 2 (consumes the bf16 probabilities the softmax threads wrote to swizzled shared memory).  Two CTAs share an SM
// (<= 96 KB smem, 256 TMEM columns each) so one CTA's tensor work overlaps the other's exponentials.
// Forward: four softmax warps, one per TMEM lane quadrant (a thread owns a whole row: the row maximum and sum need no
// exchange).  Backward: eight elementwise warps, two per quadrant with 32 columns each -- nothing there reduces over a
// row, a lone warp's per-tile chain (TMEM read -> exponentials -> two shared-memory tiles) is latency-bound, and twice
// the warps at half the registers shortens it (dK/dV + dQ at T = 6912, 2 heads: 288 -> 229 us).  The same split of the
// forward needs a per-tile exchange of row maxima and measured slower with two CTAs per SM (107 -> 121 us).
//
//   flash_fwd_kernel : CTA = 128 queries.  S = Q K^T (N = 64 keys), online softmax in registers, O += P V in TMEM
//                      (rescaled lazily); writes O (bf16) and LSE2 = m2 + log2(l) (base-2, scale folded).
//   flash_dkv_kernel : CTA = 128 keys.  S^T = K Q^T, dP^T = V dO^T, P^T = exp2(S^T c - LSE2[q]),
//                      dS^T = P^T (dP^T - delta[q]) scale;  dV += P^T dO,  dK += dS^T Q  (TMEM accumulators).
//   flash_dq_kernel  : CTA = 128 queries.  S = Q K^T, dP = dO V^T, dS likewise;  dQ += dS K.
//
// Shared-memory operand tiles are [rows x 64] bf16 = 128-byte rows in the 128B-swizzled UMMA layout: written either
// by TMA (Q, K, V, dO) or by the softmax threads (P, dS; chunk c of row r lives at chunk c ^ (r & 7)).  A [keys x d]
// tile serves both as a K-major operand (contraction over d) and as an MN-major operand (contraction over keys).
#pragma once
#include "gemm.cuh"
#include "kernels.cuh"

namespace mdc {

constexpr int FA_THREADS_FWD = 192;  // forward: producer, MMA issuer, 4 softmax warps (one per TMEM lane quadrant)
constexpr int FA_THREADS = 320;      // backward: producer, MMA issuer, 8 elementwise warps (two per quadrant, 32 columns each)
constexpr int FA_STAGES = 2;
constexpr float FA_LOG2E = 1.4426950408889634f;

struct FlashParams {
  CUtensorMap tmM1, tmM2;  // the two 128-row resident tiles (box 128 rows)   fwd: Q, -    dkv: K, V    dq: Q, dO
  CUtensorMap tmS1, tmS2;  // the two streamed 64-row tiles (box 64 rows)     fwd: K, V    dkv: Q, dO   dq: K, V
  int T, heads;
  float scale;             // 1/sqrt(head_dim)
  float lazy;              // forward: O is rescaled only when a row maximum grows by more than 2^lazy (0 = always)
  int Tp;                  // row stride of lse2 / delta: T rounded up to 64 (padding: lse2 = +inf, delta = 0)
  // outputs / side inputs
  bf16* out1;              // fwd: O      dkv: dK     dq: dQ
  bf16* out2;              //             dkv: dV
  long long ld_out, img_stride_out;  // row stride and per-image stride (elements) of out1/out2 (head h at column h*64)
  float* lse2;             // [n, heads, T]
  const float* delta;      // [n, heads, T]
};

__device__ __forceinline__ void fa_store_row_chunk(uint8_t* tile, int r, int c, const float (&v)[8]) {
  // 8 bf16 = one 16-byte chunk c of row r in a 128B-swizzled [rows x 64] tile
  BF8 b = f_to_bf8(v);
  *reinterpret_cast<BF8*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = b;
}
// The MMA helpers take PRECOMPUTED 64-bit shared-memory descriptors (128B swizzle, SBO 1024); only the 14-bit start
// address field advances: +2 (32 B) per 16-wide K step of a K-major tile, +128 (2048 B = 16 rows) for an MN-major tile.
__device__ __forceinline__ uint64_t fa_desc(const void* smem_ptr) {
  return ptx::make_smem_desc_sw128(ptx::smem_u32(smem_ptr), 16, 1024);
}
constexpr uint64_t FA_MN_LBO_DELTA = static_cast<uint64_t>((8192u - 16u) >> 4) << 16;  // LBO field 16 B -> 8192 B
__device__ __forceinline__ void fa_mma_kmajor(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              bool accumulate_first) {
  // D[128 x 64] (+)= A[128 x 64k] . B[64 x 64k]^T, both K-major
  ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate_first ? 1u : 0u);
  ptx::umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1u);
  ptx::umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1u);
  ptx::umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1u);
}
__device__ __forceinline__ void fa_mma_bmn(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           bool accumulate_first) {
  // D[128 x 64n] (+)= A[128 x 64k] (K-major) . B[64k rows x 64n] (MN-major: rows are the contraction index)
  const uint64_t b = bdesc + FA_MN_LBO_DELTA;
  ptx::umma_bf16(d_tmem, adesc, b, idesc, accumulate_first ? 1u : 0u);
  ptx::umma_bf16(d_tmem, adesc + 2, b + 128, idesc, 1u);
  ptx::umma_bf16(d_tmem, adesc + 4, b + 256, idesc, 1u);
  ptx::umma_bf16(d_tmem, adesc + 6, b + 384, idesc, 1u);
}

__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fa_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

struct FaSmem {
  uint64_t m_full, s_full[FA_STAGES], s_empty[FA_STAGES], acc_full, acc_empty, p_full, p_empty, o_full, o_empty;
  uint32_t tmem_slot;
  alignas(16) float lse_s[2][64];
  alignas(16) float del_s[2][64];  // per-iteration staging of LSE2 / delta of the streamed queries (dkv kernel)
};

// Common prologue: carve shared memory, init barriers, allocate TMEM.  Layout after the 2 KiB header:
// [M1 16K][M2 16K][S stages: FA_STAGES x (8K + 8K)][P1 16K][P2 16K]
struct FaCtx {
  FaSmem* b;
  uint8_t *m1, *m2, *st, *p1, *p2;
  uint32_t tmem;
  int warp, lane;
};
__device__ __forceinline__ FaCtx fa_setup(uint8_t* smem_raw, uint32_t tmem_cols, uint32_t ew = 8) {
  FaCtx c;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  c.b = reinterpret_cast<FaSmem*>(smem);
  static_assert(sizeof(FaSmem) <= 2048, "FaSmem must fit the 2 KiB header");
  c.m1 = smem + 2048, c.m2 = c.m1 + 16384, c.st = c.m2 + 16384;
  c.p1 = c.st + FA_STAGES * 16384, c.p2 = c.p1 + 16384;
  c.warp = threadIdx.x >> 5, c.lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&c.b->m_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) ptx::mbar_init(&c.b->s_full[i], 1), ptx::mbar_init(&c.b->s_empty[i], 1);
    ptx::mbar_init(&c.b->acc_full, 1), ptx::mbar_init(&c.b->acc_empty, ew);
    ptx::mbar_init(&c.b->p_full, ew), ptx::mbar_init(&c.b->p_empty, 1);
    ptx::mbar_init(&c.b->o_full, 1), ptx::mbar_init(&c.b->o_empty, ew);
    ptx::fence_mbar_init();
  }
  if (c.warp == 1) ptx::tmem_alloc(&c.b->tmem_slot, tmem_cols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  c.tmem = c.b->tmem_slot;
  ptx::pdl_wait();  // prologue above overlaps the previous kernel's tail; every CTA holds its TMEM before dependents start
  ptx::pdl_launch();
  return c;
}
__device__ __forceinline__ void fa_teardown(const FaCtx& c, uint32_t tmem_cols) {
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (c.warp == 1) ptx::tmem_dealloc(c.tmem, tmem_cols);
}
// TMA producer shared by the three kernels: the resident tile(s) once, then the streamed pairs.
// `stats` (dK/dV kernel): LSE2 and delta of the streamed query tile ride on the same stage barrier as two 256-byte bulk
// copies (rows are padded to a multiple of 64 entries, see FlashParams::Tp).
__device__ __forceinline__ void fa_producer(const FaCtx& c, const FlashParams& p, int m_row0, int h, int n, int n_iter,
                                            bool two_m, bool stats = false) {
  const long long srow = (static_cast<long long>(n) * p.heads + h) * p.Tp;
  ptx::mbar_expect_tx(&c.b->m_full, two_m ? 32768u : 16384u);
  ptx::tma_load_4d(&p.tmM1, &c.b->m_full, c.m1, 0, m_row0, h, n);
  if (two_m) ptx::tma_load_4d(&p.tmM2, &c.b->m_full, c.m2, 0, m_row0, h, n);
  for (int i = 0; i < n_iter; ++i) {
    const int s = i % FA_STAGES;
    ptx::mbar_wait(&c.b->s_empty[s], ((i / FA_STAGES) & 1) ^ 1);
    ptx::mbar_expect_tx(&c.b->s_full[s], stats ? 16384u + 512u : 16384u);
    ptx::tma_load_4d(&p.tmS1, &c.b->s_full[s], c.st + s * 16384, 0, i * 64, h, n);
    ptx::tma_load_4d(&p.tmS2, &c.b->s_full[s], c.st + s * 16384 + 8192, 0, i * 64, h, n);
    if (stats) {
      ptx::bulk_load_1d(&c.b->lse_s[s][0], p.lse2 + srow + i * 64, 256u, &c.b->s_full[s]);
      ptx::bulk_load_1d(&c.b->del_s[s][0], p.delta + srow + i * 64, 256u, &c.b->s_full[s]);
    }
  }
}
__device__ __forceinline__ void fa_warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(bar);
}

// ---------------------------------------------------------------------------------------------------- forward
// Software-pipelined: S is double-buffered in TMEM (S(i+1) = Q K(i+1)^T is issued before P(i) V(i)), P is
// double-buffered in shared memory, O accumulates in TMEM (lazy rescaling, see the softmax loop).
__global__ void __launch_bounds__(FA_THREADS_FWD, 2) flash_fwd_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TCOLS = 256;  // S0: [0,64)  S1: [64,128)  O tile: [128,192)
  FaCtx c = fa_setup(smem_raw, TCOLS, 4);
  // extra barriers for the double buffers live in the spare m2 tile slot's first bytes? no: reuse FaSmem fields:
  //   acc_full/acc_empty -> S buffer 0, o_full/o_empty -> O tile, p_full/p_empty -> P buffer 0; buffer-1 barriers below.
  // [0] s1_full [1] s1_empty [2] p1_full [3] p1_empty [4] o_final (m2 tile unused here).  o_full completes once per key
  // tile and may only be waited on by a thread that is at most one phase behind (a parity wait cannot tell phase k from
  // phase k - 2): the softmax threads are that close inside the loop (S(i) complete implies P(i-2) V(i-2) complete), but
  // NOT after it -- the final accumulator is therefore published on its own single-use barrier.
  uint64_t* x_bar = reinterpret_cast<uint64_t*>(c.m2);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&x_bar[0], 1), ptx::mbar_init(&x_bar[1], 4), ptx::mbar_init(&x_bar[2], 4), ptx::mbar_init(&x_bar[3], 1);
    ptx::mbar_init(&x_bar[4], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  auto sacc_full = [&](int b) { return b ? &x_bar[0] : &c.b->acc_full; };
  auto sacc_empty = [&](int b) { return b ? &x_bar[1] : &c.b->acc_empty; };
  auto pf = [&](int b) { return b ? &x_bar[2] : &c.b->p_full; };
  auto pe = [&](int b) { return b ? &x_bar[3] : &c.b->p_empty; };
  auto pbuf = [&](int b) { return b ? c.p2 : c.p1; };
  const int q0 = blockIdx.x * 128, h = blockIdx.y, n = blockIdx.z;
  const int n_iter = (p.T + 63) / 64;
  const uint32_t idesc_s = ptx::make_idesc_bf16(128, 64, 0, 0), idesc_o = ptx::make_idesc_bf16(128, 64, 0, 1);
  if (c.warp == 0) {
    if (c.lane == 0) fa_producer(c, p, q0, h, n, n_iter, false);
    __syncwarp();
  } else if (c.warp == 1) {
    const bool leader = ptx::elect_one();
    const uint64_t d_q = fa_desc(c.m1), d_st = fa_desc(c.st);
    const uint64_t d_p0 = fa_desc(c.p1), d_p1 = fa_desc(c.p2);
    ptx::mbar_wait(&c.b->m_full, 0);
    ptx::mbar_wait(&c.b->s_full[0], 0);
    ptx::tc_fence_after();
    if (leader) {
      fa_mma_kmajor(c.tmem, d_q, d_st, idesc_s, false);  // S(0)
      ptx::umma_commit(sacc_full(0));
    }
    __syncwarp();
    for (int i = 0; i < n_iter; ++i) {
      const int s = i % FA_STAGES, b = i & 1;
      if (i + 1 < n_iter) {  // S(i+1) = Q K(i+1)^T into the other S buffer, ahead of P(i) V(i)
        const int s1 = (i + 1) % FA_STAGES, b1 = (i + 1) & 1;
        ptx::mbar_wait(&c.b->s_full[s1], ((i + 1) / FA_STAGES) & 1);
        ptx::mbar_wait(sacc_empty(b1), (((i + 1) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (leader) {
          fa_mma_kmajor(c.tmem + b1 * 64, d_q, d_st + s1 * 1024, idesc_s, false);
          ptx::umma_commit(sacc_full(b1));
        }
        __syncwarp();
      }
      ptx::mbar_wait(pf(b), (i >> 1) & 1);
      ptx::tc_fence_after();
      if (leader) {
        fa_mma_bmn(c.tmem + 128, b ? d_p1 : d_p0, d_st + s * 1024 + 512, idesc_o, i > 0);  // O += P(i) V(i)
        ptx::umma_commit(&c.b->o_full);
        ptx::umma_commit(pe(b));
        ptx::umma_commit(&c.b->s_empty[s]);
        if (i + 1 == n_iter) ptx::umma_commit(&x_bar[4]);  // every P V has completed: O is final
      }
      __syncwarp();
    }
  } else {
    const int q = c.warp & 3, row = q * 32 + c.lane;
    const uint32_t t_row = c.tmem + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = p.scale * FA_LOG2E;
    // m is the reference maximum the stored exponentials are relative to.  It only moves when a row's maximum grows
    // by more than 2^8 (the probabilities then stay below 256, exact enough in bf16 / fp32): O accumulates in TMEM
    // across key tiles and is touched by the softmax threads only on those rare rescales and once at the end, so the
    // per-tile TMEM read traffic (64 B/clk per SM, the bound of this kernel) is the 32 KB score tile alone.
    float m = -INFINITY, l = 0.f;
    for (int i = 0; i < n_iter; ++i) {
      const int b = i & 1;
      ptx::mbar_wait(sacc_full(b), (i >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t raw[64];
      ptx::tmem_ld32(t_row + b * 64, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
      ptx::tmem_ld32(t_row + b * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      fa_warp_arrive(sacc_empty(b), c.lane);
      const int valid = p.T - i * 64;  // keys of this tile that exist
      if (valid < 64) {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j >= valid) raw[j] = 0xff800000u;  // -inf
      }
      // row maximum on the raw scores (the scale is positive), three-input max: 32 instructions for 64 columns
      float mx = -INFINITY, mx1 = -INFINITY;  // two independent chains
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        mx = fa_max3(mx, __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
        mx1 = fa_max3(mx1, __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3]));
      }
      mx = fmaxf(mx, mx1) * c2;
      const bool grow = mx > m + p.lazy;  // always true for the first tile (m = -inf)
      const float m_new = grow ? mx : m;
      const float alpha = grow ? fa_exp2(m - m_new) : 1.f;
      if (i > 0 && __any_sync(0xffffffffu, grow)) {  // rescale the TMEM accumulator rows of this warp
        uint32_t acc[64];
        ptx::mbar_wait(&c.b->o_full, (i - 1) & 1);  // P(i-1) V(i-1) and everything before it has completed
        ptx::tc_fence_after();
        ptx::tmem_ld32(t_row + 128, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
        ptx::tmem_ld32(t_row + 160, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) * alpha);
        ptx::tmem_st32(t_row + 128, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
        ptx::tmem_st32(t_row + 160, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
      }
      // p = exp2(s * c2 - m_new): one packed FMA per two columns, MUFU.EX2, packed row-sum
      const float2 cc = make_float2(c2, c2), neg = make_float2(-m_new, -m_new);
      float2 sum2 = make_float2(0.f, 0.f);
      ptx::mbar_wait(pe(b), ((i >> 1) & 1) ^ 1);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(raw[ch * 8 + j]), __uint_as_float(raw[ch * 8 + j + 1])), cc, neg);
          const float2 e = make_float2(fa_exp2(a.x), fa_exp2(a.y));
          v[j] = e.x, v[j + 1] = e.y;
          sum2 = __fadd2_rn(sum2, e);
        }
        fa_store_row_chunk(pbuf(b), row, ch, v);
      }
      ptx::fence_proxy_async_smem();
      fa_warp_arrive(pf(b), c.lane);  // also orders the rescale above before P(i) V(i) is issued
      l = l * alpha + (sum2.x + sum2.y);
      m = m_new;
    }
    {
      uint32_t raw[64];
      ptx::mbar_wait(&x_bar[4], 0);
      ptx::tc_fence_after();
      ptx::tmem_ld32(t_row + 128, *reinterpret_cast<uint32_t(*)[32]>(&raw[0]));
      ptx::tmem_ld32(t_row + 160, *reinterpret_cast<uint32_t(*)[32]>(&raw[32]));
      ptx::tmem_ld_wait();
      if (q0 + row < p.T) {
        const float inv = 1.f / l;
        bf16* dst = p.out1 + n * p.img_stride_out + static_cast<long long>(q0 + row) * p.ld_out + h * 64;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[ch * 8 + j]) * inv;
          *reinterpret_cast<BF8*>(dst + ch * 8) = f_to_bf8(v);
        }
        p.lse2[(static_cast<long long>(n) * p.heads + h) * p.Tp + q0 + row] = m + log2f(l);
      }
    }
  }
  fa_teardown(c, TCOLS);
}

// delta[n, h, t] = sum_c dO[n, t, h*64 + c] * O[n, t, h*64 + c]     (one warp per (token, head))
__global__ void flash_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, long long ld, int N, int T,
                                   int Tp, int heads, float* __restrict__ delta) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long w = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= 1LL * N * T * heads) return;
  const int h = w % heads;
  const long long tok = w / heads;  // n * T + t
  const long long off = tok * ld + h * 64 + lane * 2;
  float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o + off));
  float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dout + off));
  float s = warp_sum(a.x * b.x + a.y * b.y);
  if (lane == 0) {
    const long long n = tok / T, t = tok % T;
    delta[(n * heads + h) * Tp + t] = s;
  }
}

// ---------------------------------------------------------------------------------------------------- dK, dV
// Eight elementwise warps: two per TMEM lane quadrant, each owning 32 of the tile's 64 columns.  The per-warp chain
// (TMEM read -> exponentials -> two shared-memory tiles) is latency-bound, not throughput-bound, so twice the warps at
// half the registers is what shortens it; nothing in the backward needs a row reduction, the split costs no exchange.
__global__ void __launch_bounds__(FA_THREADS, 2) flash_dkv_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TCOLS = 256;  // S^T [0,64)  dP^T [64,128)  dV [128,192)  dK [192,256)
  FaCtx c = fa_setup(smem_raw, TCOLS, 8);
  const int k0 = blockIdx.x * 128, h = blockIdx.y, n = blockIdx.z;
  const int n_iter = (p.T + 63) / 64;  // query tiles
  const uint32_t idesc_k = ptx::make_idesc_bf16(128, 64, 0, 0), idesc_mn = ptx::make_idesc_bf16(128, 64, 0, 1);
  if (c.warp == 0) {
    if (c.lane == 0) fa_producer(c, p, k0, h, n, n_iter, true, true);
    __syncwarp();
  } else if (c.warp == 1) {
    const bool leader = ptx::elect_one();
    const uint64_t d_k = fa_desc(c.m1), d_v = fa_desc(c.m2), d_p = fa_desc(c.p1), d_ds = fa_desc(c.p2), d_st = fa_desc(c.st);
    ptx::mbar_wait(&c.b->m_full, 0);
    ptx::mbar_wait(&c.b->s_full[0], 0);
    ptx::tc_fence_after();
    if (leader) {
      fa_mma_kmajor(c.tmem, d_k, d_st, idesc_k, false);             // S^T(0)  = K Q(0)^T
      fa_mma_kmajor(c.tmem + 64, d_v, d_st + 512, idesc_k, false);  // dP^T(0) = V dO(0)^T
      ptx::umma_commit(&c.b->acc_full);
    }
    __syncwarp();
    for (int i = 0; i < n_iter; ++i) {
      const int s = i % FA_STAGES;
      const uint64_t d_q = d_st + s * 1024, d_do = d_q + 512;
      if (i + 1 < n_iter) {  // scores of the next query tile as soon as the softmax threads hold tile i in registers
        const int s1 = (i + 1) % FA_STAGES;
        ptx::mbar_wait(&c.b->s_full[s1], ((i + 1) / FA_STAGES) & 1);
        ptx::mbar_wait(&c.b->acc_empty, i & 1);
        ptx::tc_fence_after();
        if (leader) {
          fa_mma_kmajor(c.tmem, d_k, d_st + s1 * 1024, idesc_k, false);
          fa_mma_kmajor(c.tmem + 64, d_v, d_st + s1 * 1024 + 512, idesc_k, false);
          ptx::umma_commit(&c.b->acc_full);
        }
        __syncwarp();
      }
      ptx::mbar_wait(&c.b->p_full, i & 1);
      ptx::tc_fence_after();
      if (leader) {
        fa_mma_bmn(c.tmem + 128, d_p, d_do, idesc_mn, i > 0);  // dV += P^T  dO
        fa_mma_bmn(c.tmem + 192, d_ds, d_q, idesc_mn, i > 0);  // dK += dS^T Q
        ptx::umma_commit(&c.b->p_empty);
        ptx::umma_commit(&c.b->s_empty[s]);
      }
      __syncwarp();
    }
    if (leader) ptx::umma_commit(&c.b->o_full);  // accumulators final
    __syncwarp();
  } else {
    const int q = c.warp & 3, row = q * 32 + c.lane, hf = (c.warp - 2) >> 2;  // column half of this warp
    const uint32_t t_row = c.tmem + (static_cast<uint32_t>(q * 32) << 16) + hf * 32;
    const float c2 = p.scale * FA_LOG2E;
    const float2 cc2 = make_float2(c2, c2), sc2 = make_float2(p.scale, p.scale);
    for (int i = 0; i < n_iter; ++i) {
      const int s = i % FA_STAGES;
      ptx::mbar_wait(&c.b->acc_full, i & 1);
      ptx::tc_fence_after();
      uint32_t rs[32], rd[32];  // this warp's half of the S^T / dP^T rows; frees the TMEM buffers for the next tile's MMAs
      ptx::tmem_ld32(t_row, rs);
      ptx::tmem_ld32(t_row + 64, rd);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      fa_warp_arrive(&c.b->acc_empty, c.lane);
      // LSE2 / delta of the tile's queries arrived with Q / dO (same stage barrier; S^T complete implies they have landed,
      // the wait only makes the bulk-copied bytes visible to this thread).  Padding queries carry LSE2 = +inf: probability 0.
      ptx::mbar_wait(&c.b->s_full[s], (i / FA_STAGES) & 1);
      ptx::mbar_wait(&c.b->p_empty, (i & 1) ^ 1);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float pv[8], dv[8];
        const int c0 = hf * 32 + ch * 8;
        const float4 l0 = *reinterpret_cast<const float4*>(&c.b->lse_s[s][c0]);
        const float4 l1 = *reinterpret_cast<const float4*>(&c.b->lse_s[s][c0 + 4]);
        const float4 d0 = *reinterpret_cast<const float4*>(&c.b->del_s[s][c0]);
        const float4 d1 = *reinterpret_cast<const float4*>(&c.b->del_s[s][c0 + 4]);
        const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float dl[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int j = 0; j < 8; j += 2) {  // packed fp32x2: p = exp2(s c - lse), ds = p (dp - delta) scale
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(rs[ch * 8 + j]), __uint_as_float(rs[ch * 8 + j + 1])), cc2,
                                      make_float2(-ls[j], -ls[j + 1]));
          const float2 pr = make_float2(fa_exp2(a.x), fa_exp2(a.y));
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(rd[ch * 8 + j]), __uint_as_float(rd[ch * 8 + j + 1])), sc2,
                                      make_float2(-dl[j] * p.scale, -dl[j + 1] * p.scale));
          const float2 d = __fmul2_rn(pr, t);
          pv[j] = pr.x, pv[j + 1] = pr.y;
          dv[j] = d.x, dv[j + 1] = d.y;
        }
        fa_store_row_chunk(c.p1, row, hf * 4 + ch, pv);
        fa_store_row_chunk(c.p2, row, hf * 4 + ch, dv);
      }
      ptx::fence_proxy_async_smem();
      fa_warp_arrive(&c.b->p_full, c.lane);
    }
    ptx::mbar_wait(&c.b->o_full, 0);
    ptx::tc_fence_after();
    const bool ok = k0 + row < p.T;
    const long long off = n * p.img_stride_out + static_cast<long long>(k0 + row) * p.ld_out + h * 64 + hf * 32;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* dst = (which == 0 ? p.out2 : p.out1) + off;  // TMEM [128,192) = dV -> out2 ; [192,256) = dK -> out1
      uint32_t r[32];
      ptx::tmem_ld32(t_row + 128 + which * 64, r);
      ptx::tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[ch * 8 + j]);
          *reinterpret_cast<BF8*>(dst + ch * 8) = f_to_bf8(v);
        }
      }
    }
  }
  fa_teardown(c, TCOLS);
}

// ---------------------------------------------------------------------------------------------------- dQ
__global__ void __launch_bounds__(FA_THREADS, 2) flash_dq_kernel(const __grid_constant__ FlashParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr uint32_t TCOLS = 256;  // S [0,64)  dP [64,128)  dQ [128,192)
  FaCtx c = fa_setup(smem_raw, TCOLS, 8);
  const int q0 = blockIdx.x * 128, h = blockIdx.y, n = blockIdx.z;
  const int n_iter = (p.T + 63) / 64;  // key tiles
  const uint32_t idesc_k = ptx::make_idesc_bf16(128, 64, 0, 0), idesc_mn = ptx::make_idesc_bf16(128, 64, 0, 1);
  if (c.warp == 0) {
    if (c.lane == 0) fa_producer(c, p, q0, h, n, n_iter, true);
    __syncwarp();
  } else if (c.warp == 1) {
    const bool leader = ptx::elect_one();
    const uint64_t d_q = fa_desc(c.m1), d_do = fa_desc(c.m2), d_ds = fa_desc(c.p1), d_st = fa_desc(c.st);
    ptx::mbar_wait(&c.b->m_full, 0);
    ptx::mbar_wait(&c.b->s_full[0], 0);
    ptx::tc_fence_after();
    if (leader) {
      fa_mma_kmajor(c.tmem, d_q, d_st, idesc_k, false);              // S(0)  = Q  K(0)^T
      fa_mma_kmajor(c.tmem + 64, d_do, d_st + 512, idesc_k, false);  // dP(0) = dO V(0)^T
      ptx::umma_commit(&c.b->acc_full);
    }
    __syncwarp();
    for (int i = 0; i < n_iter; ++i) {
      const int s = i % FA_STAGES;
      const uint64_t d_k = d_st + s * 1024;
      if (i + 1 < n_iter) {
        const int s1 = (i + 1) % FA_STAGES;
        ptx::mbar_wait(&c.b->s_full[s1], ((i + 1) / FA_STAGES) & 1);
        ptx::mbar_wait(&c.b->acc_empty, i & 1);
        ptx::tc_fence_after();
        if (leader) {
          fa_mma_kmajor(c.tmem, d_q, d_st + s1 * 1024, idesc_k, false);
          fa_mma_kmajor(c.tmem + 64, d_do, d_st + s1 * 1024 + 512, idesc_k, false);
          ptx::umma_commit(&c.b->acc_full);
        }
        __syncwarp();
      }
      ptx::mbar_wait(&c.b->p_full, i & 1);
      ptx::tc_fence_after();
      if (leader) {
        fa_mma_bmn(c.tmem + 128, d_ds, d_k, idesc_mn, i > 0);  // dQ += dS K
        ptx::umma_commit(&c.b->p_empty);
        ptx::umma_commit(&c.b->s_empty[s]);
      }
      __syncwarp();
    }
    if (leader) ptx::umma_commit(&c.b->o_full);
    __syncwarp();
  } else {
    const int q = c.warp & 3, row = q * 32 + c.lane, hf = (c.warp - 2) >> 2;  // column half of this warp
    const uint32_t t_row = c.tmem + (static_cast<uint32_t>(q * 32) << 16) + hf * 32;
    const float c2 = p.scale * FA_LOG2E;
    const bool ok = q0 + row < p.T;
    const long long sidx = (static_cast<long long>(n) * p.heads + h) * p.Tp + q0 + row;
    const float lse = ok ? p.lse2[sidx] : 0.f, del = ok ? p.delta[sidx] : 0.f;
    const float nlse = ok ? -lse : -INFINITY;  // rows past T produce probability 0
    const float ndel_s = -del * p.scale;
    for (int i = 0; i < n_iter; ++i) {
      ptx::mbar_wait(&c.b->acc_full, i & 1);
      ptx::tc_fence_after();
      uint32_t rs[32], rd[32];
      ptx::tmem_ld32(t_row, rs);
      ptx::tmem_ld32(t_row + 64, rd);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      fa_warp_arrive(&c.b->acc_empty, c.lane);
      ptx::mbar_wait(&c.b->p_empty, (i & 1) ^ 1);
      const int valid = p.T - i * 64 - hf * 32;  // keys of this warp's half that exist
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float dv[8];
        if (valid >= 32) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {  // packed fp32x2
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(rs[ch * 8 + j]), __uint_as_float(rs[ch * 8 + j + 1])),
                                        make_float2(c2, c2), make_float2(nlse, nlse));
            const float2 pr = make_float2(fa_exp2(a.x), fa_exp2(a.y));
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(rd[ch * 8 + j]), __uint_as_float(rd[ch * 8 + j + 1])),
                                        make_float2(p.scale, p.scale), make_float2(ndel_s, ndel_s));
            const float2 d = __fmul2_rn(pr, t);
            dv[j] = d.x, dv[j + 1] = d.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kk = ch * 8 + j;
            const float pr = kk < valid ? fa_exp2(fmaf(__uint_as_float(rs[ch * 8 + j]), c2, nlse)) : 0.f;
            dv[j] = pr * (__uint_as_float(rd[ch * 8 + j]) - del) * p.scale;
          }
        }
        fa_store_row_chunk(c.p1, row, hf * 4 + ch, dv);
      }
      ptx::fence_proxy_async_smem();
      fa_warp_arrive(&c.b->p_full, c.lane);
    }
    ptx::mbar_wait(&c.b->o_full, 0);
    ptx::tc_fence_after();
    bf16* dst = p.out1 + n * p.img_stride_out + static_cast<long long>(q0 + row) * p.ld_out + h * 64 + hf * 32;
    uint32_t r[32];
    ptx::tmem_ld32(t_row + 128, r);
    ptx::tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[ch * 8 + j]);
        *reinterpret_cast<BF8*>(dst + ch * 8) = f_to_bf8(v);
      }
    }
  }
  fa_teardown(c, TCOLS);
}

// ======================================================================================= host side
constexpr int FA_SMEM = 2048 + 2 * 16384 + FA_STAGES * 16384 + 2 * 16384 + 1024;  // + alignment slack

struct FlashPlan {
  FlashParams fwd, dkv, dq;
  dim3 grid;
  // delta kernel
  const bf16 *o, *dout;
  long long ld_o;
  int N, T, heads;
  float* delta;
  double flops_fwd = 0, flops_bwd = 0;
};

// lse2 / delta are [n, heads, Tp] with Tp = T rounded up to the 64-query tile, + 64 floats of slack.
inline int flash_stat_stride(int T) { return (T + 63) & ~63; }
inline size_t flash_stat_floats(int n, int heads, int T) { return static_cast<size_t>(n) * heads * flash_stat_stride(T) + 64; }
__global__ void flash_stat_init_kernel(float* lse2, float* delta, long long tot) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i < tot) lse2[i] = INFINITY, delta[i] = 0.f;
}

inline CUtensorMap fa_map(const bf16* base, long long ld, int T, int heads, int n, int box_rows) {
  uint64_t dims[4] = {64, (uint64_t)T, (uint64_t)heads, (uint64_t)n};
  uint64_t str[3] = {(uint64_t)ld, 64, (uint64_t)T * ld};
  uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return make_tmap_bf16(base, dims, str, box);
}

// q, k, v (and their gradients dq, dk, dv) are column slices of [n, T, ld_qkv] tensors; o / dout of [n, T, ld_o].
inline FlashPlan plan_flash(int n, int T, int heads, const bf16* q, const bf16* k, const bf16* v, long long ld_qkv, bf16* o,
                            const bf16* dout, long long ld_o, bf16* dq, bf16* dk, bf16* dv, long long ld_dqkv, float* lse2,
                            float* delta) {
  FlashPlan f;
  memset(&f.fwd, 0, sizeof(FlashParams));
  f.grid = dim3((T + 127) / 128, heads, n);
  FlashParams base;
  memset(&base, 0, sizeof(base));
  base.T = T, base.Tp = flash_stat_stride(T), base.heads = heads, base.scale = 0.125f, base.lse2 = lse2, base.delta = delta;
  {  // padding entries of the statistics rows: LSE2 = +inf (probability 0), delta = 0; the kernels only write t < T
    const long long tot = 1LL * n * heads * base.Tp;
    flash_stat_init_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256>>>(lse2, delta, tot);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaStreamSynchronize(0));
  }
  base.lazy = getenv("MDC_FLASH_LAZY") ? static_cast<float>(atof(getenv("MDC_FLASH_LAZY"))) : 8.f;
  f.fwd = base;
  f.fwd.tmM1 = fa_map(q, ld_qkv, T, heads, n, 128);
  f.fwd.tmM2 = f.fwd.tmM1;
  f.fwd.tmS1 = fa_map(k, ld_qkv, T, heads, n, 64);
  f.fwd.tmS2 = fa_map(v, ld_qkv, T, heads, n, 64);
  f.fwd.out1 = o, f.fwd.ld_out = ld_o, f.fwd.img_stride_out = 1LL * T * ld_o;
  f.dkv = base;
  f.dkv.tmM1 = fa_map(k, ld_qkv, T, heads, n, 128);
  f.dkv.tmM2 = fa_map(v, ld_qkv, T, heads, n, 128);
  f.dkv.tmS1 = fa_map(q, ld_qkv, T, heads, n, 64);
  f.dkv.tmS2 = fa_map(dout, ld_o, T, heads, n, 64);
  f.dkv.out1 = dk, f.dkv.out2 = dv, f.dkv.ld_out = ld_dqkv, f.dkv.img_stride_out = 1LL * T * ld_dqkv;
  f.dq = base;
  f.dq.tmM1 = fa_map(q, ld_qkv, T, heads, n, 128);
  f.dq.tmM2 = fa_map(dout, ld_o, T, heads, n, 128);
  f.dq.tmS1 = fa_map(k, ld_qkv, T, heads, n, 64);
  f.dq.tmS2 = fa_map(v, ld_qkv, T, heads, n, 64);
  f.dq.out1 = dq, f.dq.ld_out = ld_dqkv, f.dq.img_stride_out = 1LL * T * ld_dqkv;
  f.o = o, f.dout = dout, f.ld_o = ld_o, f.N = n, f.T = T, f.heads = heads, f.delta = delta;
  f.flops_fwd = 4.0 * T * T * 64.0 * heads * n;
  f.flops_bwd = f.flops_fwd;  // counted 1 x forward like every other backward
  return f;
}

inline void flash_set_attrs() {
  static bool done[64] = {false};
  if (!first_use_on_device(done)) return;
  MDC_CUDA(cudaFuncSetAttribute(flash_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  MDC_CUDA(cudaFuncSetAttribute(flash_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  MDC_CUDA(cudaFuncSetAttribute(flash_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
}
inline void run_flash_fwd(const FlashPlan& f, cudaStream_t st) {
  flash_set_attrs();
  launch_k(flash_fwd_kernel, f.grid, dim3(FA_THREADS_FWD), FA_SMEM, st, f.fwd);
}
// dK/dV and dQ are independent once delta exists: with a side stream the two run beside each other (fork / join events
// owned by the caller).  Below one wave of CTA pairs (T <= 1728 here) neither kernel fills the GPU alone; above it the
// second kernel's CTAs fill the SMs the first one's tail leaves idle (20.39 -> 20.19 -> 20.12 ms per step).
// MDC_FLASH_BESIDE_MAX=<ctas> limits this to grids below that size (0 = never).
struct SideBranch {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
inline void run_flash_bwd(const FlashPlan& f, cudaStream_t st, const SideBranch* sb = nullptr) {
  flash_set_attrs();
  const long long warps = 1LL * f.N * f.T * f.heads;
  launch_k(flash_delta_kernel, dim3(static_cast<unsigned>((warps * 32 + 255) / 256)), dim3(256), 0, st, f.o, f.dout, f.ld_o,
           f.N, f.T, f.dq.Tp, f.heads, f.delta);
  static const int beside_max = getenv("MDC_FLASH_BESIDE_MAX") ? atoi(getenv("MDC_FLASH_BESIDE_MAX")) : 0x7fffffff;
  const bool beside = sb && sb->stream && static_cast<int>(f.grid.x * f.grid.y * f.grid.z) < beside_max;
  if (beside) {
    MDC_CUDA(cudaEventRecord(sb->fork, st));
    MDC_CUDA(cudaStreamWaitEvent(sb->stream, sb->fork, 0));
    launch_k(flash_dq_kernel, f.grid, dim3(FA_THREADS), FA_SMEM, sb->stream, f.dq);
    MDC_CUDA(cudaEventRecord(sb->join, sb->stream));
  }
  launch_k(flash_dkv_kernel, f.grid, dim3(FA_THREADS), FA_SMEM, st, f.dkv);
  if (beside)
    MDC_CUDA(cudaStreamWaitEvent(st, sb->join, 0));
  else
    launch_k(flash_dq_kernel, f.grid, dim3(FA_THREADS), FA_SMEM, st, f.dq);
}

}  // namespace mdc
