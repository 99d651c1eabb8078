// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M,N] = alpha * A[M,K] . B[N,K]^T  (+ bias[N]) (+ bias_img[img][N]) (+ residual)      bf16 in, fp32 acc
//
// One CTA per SM loops over 128 x BN output tiles.  Warp 0 (one lane) is the TMA producer, warp 1
// (one lane) issues tcgen05.mma into a double-buffered TMEM accumulator (2 x 256 columns), warps 2..5
// drain TMEM with tcgen05.ld and run the epilogue.  Operands are staged in shared memory in the
// 128-byte-swizzled canonical UMMA layout, written directly by TMA:
//   * K-major operand  : box [64 k, rows]      -> rows x 128 B, 8-row groups 1024 B apart (SBO)
//   * MN-major operand : boxes [64 mn, 64 k]   -> 64 k-rows x 128 B each, chunks 8192 B apart (LBO)
//   * conv3x3 A operand: box [64 c, BW, BH, 1] of the NHWC input at (w0+s-1, h0+r-1); TMA zero-fills
//     out-of-bounds coordinates, which is exactly the conv's zero padding, so no im2col buffer exists.
// This plays the role cuDNN implicit-GEMM / cuBLAS play under the reference.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ptx.cuh"

namespace mdc {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_A_STAGE = GEMM_BM * GEMM_BK * 2;  // 16 KiB
constexpr int GEMM_MAX_STAGES = 8;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BUDGET = 225 * 1024;
constexpr int GEMM_HEADER0 = 4096;   // barriers + staged bias; followed by `npanel` 128 x 32 bf16 store panels (8 KiB each)
constexpr int GEMM_PANEL = 8192;
constexpr int GEMM_TMEM_COLS = 512;

struct GemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  // tiling
  int m_tiles, n_tiles, nb0, nb1;
  int num_k_chunks;
  int BN;
  int stages;
  int a_mn, b_mn;
  uint32_t idesc;
  uint32_t bytesA, bytesB;
  int M, N;  // valid rows per batch (GEMM mode) / valid columns
  // conv mode: 0 = GEMM; 1 = implicit-GEMM conv over `ntaps` taps (3x3, or the four 2x2 phase convs of a fused
  // nearest-2x-upsample + conv3x3, phase = batch slot b0); 2 = input gradient of that fused op (16 taps, 4 phase views)
  int conv;
  int H, W, tiles_h, tiles_w, BW, BH, chunks_per_tap;
  int taps_w, ntaps, off_h0, off_w0, nphase;
  long long out_sh, out_sw, out_pa, out_pb;  // output element strides: pixel row, pixel column, phase row, phase column
  CUtensorMap tmA1, tmA2, tmA3;              // mode 2: phase views 1..3 of dy (tmA is phase 0)
  // epilogue
  void* out;
  int out_f32;
  int vec_ok;
  long long ldc, sc0, sc1;  // element strides of the output: row, batch0, batch1 (conv: sc1 = per image)
  const float* bias;
  const float* bias_img;  // [images][N]; image = conv image or batch1
  const __nv_bfloat16* res;
  long long ldr, sr0, sr1;
  float alpha;
  int relu;  // clamp the finished value at zero (AutoencoderTiny's conv + ReLU, applied after bias / residual)
  // TMA-store epilogue (bf16 outputs): each 128-row x 32-column chunk is staged in 64B-swizzled shared memory and
  // written with one bulk tensor store (coalesced, hardware-clipped) instead of 16-byte scattered stores per thread.
  int tma_store;
  int npanel;  // store panels in flight: 4 (up to three bulk stores outstanding) or 2 when shared memory is short
  CUtensorMap tmC, tmC1, tmC2, tmC3;  // tmC1..3: phases 1..3 of the fused upsample conv
  // cluster B-multicast: `cs` CTAs (consecutive m-tiles of one n-tile) form a cluster; each loads BN/cs rows of the
  // B tile and multicasts them to all, so the weight tile crosses L2 -> SM once per cluster instead of once per CTA.
  int cs;
  // CTA-pair MMA (cs == 2 only): the two CTAs of a cluster run one M = 256 tcgen05.mma.cta_group::2 -- each stages
  // its own 128 A rows and HALF of the B tile, so a k-chunk costs each SM 16 KiB + BN*64 B of L2 -> SM traffic instead
  // of 16 KiB + BN*128 B; only the leader (rank 0) issues MMAs, its commits release stages / publish accumulators in
  // both CTAs, and both CTAs' epilogues hand accumulator stages back to the leader.
  int pair;
  // split-K (nb0 = nb1 = 1 only): CTA (tile, s) accumulates k-chunks [s*kc_per_split, ...) and writes its raw fp32
  // partial to ws[s][row][col]; splitk_reduce_kernel sums the partials and applies the epilogue.
  int ksplit, kc_per_split;
  float* ws;
  long long ws_ld, ws_split_stride;
  // Row-shared taps (conv 3x3, tile = 128 consecutive pixels of ONE image row): the three taps of a kernel row read the
  // same pixels shifted by -1 / 0 / +1, so ONE A box of 130 pixels [64 c, 130 w] is staged per (kernel row, channel
  // chunk) and the three taps are issued as MMAs over row-offset views of it (descriptor start address + 128 B x s),
  // each with its own B tile.  The 128-byte swizzle is a function of the absolute shared-memory address for both the
  // TMA write and the MMA read, so a view that starts s rows into the box needs NO matrix base offset in the descriptor
  // (measured on a B200: setting it to s gives wrong results, leaving it 0 is exact -- tools/rowshare_diag.py).  A k-step then moves 16.6 KB of activations through L2 -> SM for three taps
  // instead of 48 KB, which is what bounds the 128-output-channel convolutions (B is only BN x 128 B per tap).
  int rowshare;
  int a_stage;  // bytes between A stages (GEMM_A_STAGE, or 17 KiB for a 130-row box)
  // GroupNorm statistics of the OUTPUT tensor as an epilogue side product (conv mode, 32 groups, no split-K): every
  // epilogue thread accumulates (sum, sum of squares) of the values it stores, per group, in a private shared-memory
  // column across all tiles of the persistent loop; per image the CTA reduces them once and writes one partial row
  // gn_partial[(img * gridDim.x + cta) * 64 + 2 g + {0, 1}].  The consuming gn_apply kernel reduces the rows in a fixed
  // order (deterministic) -- the separate statistics pass over the 56-226 MB decoder tensors disappears.
  // Residual prefetch (conv mode, large images): the producer warp asks the TMA unit to pull the residual tile of the
  // output tile it is starting into L2; the epilogue reaches that tile one or two accumulator stages later and its
  // per-thread residual loads then hit L2 instead of waiting for HBM (conv2 of a decoder resnet: 150 us vs 100 us for the
  // same convolution without a residual before this).
  CUtensorMap tmR;
  int res_prefetch;
  float* gn_partial;
  int gn_cpg;     // channels per group (2, 4, 8, 16 or 32); 0 = off
  int gn_nimg;
  int gn_smem_off;  // byte offset of the accumulators inside the aligned dynamic shared memory
};
constexpr int GEMM_GN_SMEM = (64 * 128 + 4 * 64) * 4;  // accumulators [64][128] + per-warp sums [4][64]
constexpr int GEMM_A_STAGE_RS = 17 * 1024;  // 130 rows x 128 B, rounded up to the 1024-byte swizzle atom

// work item -> (n_tile, m_tile, b0, b1); with split-K the batch slot b0 carries the split index instead.  With
// clusters a work item is a group of `cs` consecutive m-tiles and CTA rank r takes m_tile = group * cs + r (which may
// be >= m_tiles in the last group: such a CTA still feeds the multicast but stores nothing).
__device__ __forceinline__ void gemm_decode_tile(const GemmParams& p, int tile, int& n_tile, int& m_tile, int& b0,
                                                 int& b1, int rank = 0) {
  n_tile = tile % p.n_tiles;
  int rest = tile / p.n_tiles;
  const int m_groups = (p.m_tiles + p.cs - 1) / p.cs;
  m_tile = (rest % m_groups) * p.cs + rank;
  int b = rest / m_groups;
  const int nb0 = p.ksplit > 1 ? p.ksplit : p.nb0;
  b0 = b % nb0;
  b1 = b / nb0;
}

// (sum, sum of squares) of a 32-column chunk, per GroupNorm group of `cpg` channels, added to this thread's private
// accumulator column acc[(2 g) * 128], acc[(2 g + 1) * 128].  Indexing of v stays static (registers).
template <int CPG>
__device__ __forceinline__ void gn_accumulate_t(const float (&v)[32], float* __restrict__ acc, int g0) {
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += CPG) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int t = 0; t < CPG; ++t) s += v[j0 + t], q = fmaf(v[j0 + t], v[j0 + t], q);
    float* a = acc + (2 * (g0 + j0 / CPG)) * 128;
    a[0] += s;
    a[128] += q;
  }
}
__device__ __forceinline__ void gn_accumulate(const float (&v)[32], float* __restrict__ acc, int cpg, int col0) {
  const int g0 = col0 / cpg;
  switch (cpg) {
    case 2: gn_accumulate_t<2>(v, acc, g0); break;
    case 4: gn_accumulate_t<4>(v, acc, g0); break;
    case 8: gn_accumulate_t<8>(v, acc, g0); break;
    case 16: gn_accumulate_t<16>(v, acc, g0); break;
    default: gn_accumulate_t<32>(v, acc, g0); break;
  }
}

template <bool kPair>
__device__ __forceinline__ void umma_gemm_body(const GemmParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // barriers live in the first 1 KiB of the aligned region, the staged bias in the next 2 KiB; stage buffers at 4 KiB.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias_base = reinterpret_cast<float*>(smem + 1024);  // 2 x 256 floats
  uint8_t* spanel = smem + GEMM_HEADER0;  // npanel x 8 KiB output panels
  uint8_t* sA = spanel + p.npanel * GEMM_PANEL;
  constexpr bool pair = kPair;
  const uint32_t b_tile = static_cast<uint32_t>(pair ? p.BN / 2 : p.BN) * 128u;  // one B tile (one tap)
  const uint32_t b_stage = p.rowshare ? 3u * b_tile : b_tile;
  uint8_t* sB = sA + p.stages * p.a_stage;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = ((p.m_tiles + p.cs - 1) / p.cs) * p.n_tiles * (p.ksplit > 1 ? p.ksplit : p.nb0) * p.nb1;
  const int rank = p.cs > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / p.cs, n_clusters = gridDim.x / p.cs;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full_bar[i], pair ? 2 : 1);  // pair: one arrival per CTA's producer, on the leader's barrier
      // multicast: every CTA of the cluster must release the stage before it is refilled; pair: one multicast commit
      ptx::mbar_init(&empty_bar[i], pair ? 1 : p.cs);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], pair ? 8 : 4);  // pair: the epilogue warps of both CTAs, on the leader's barrier
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB);
  }
  if (pair) ptx::cluster_sync_all();  // both CTAs are resident before the paired TMEM allocation
  if (warp == 1) {
    if (pair)
      ptx::tmem_alloc_pair(tmem_slot, GEMM_TMEM_COLS);
    else
      ptx::tmem_alloc(tmem_slot, GEMM_TMEM_COLS);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.cs > 1) ptx::cluster_sync_all();  // peers' barriers are initialised before any multicast / remote arrive
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail.
  // All CTAs of this grid hold their TMEM before the next grid may start (no allocation deadlock).
  ptx::pdl_wait();
  ptx::pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the (warp-uniform) loop; one elected lane issues the copies.  No divisions inside the K loop:
    // (tap, channel chunk) are carried as counters.
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t phase = 0;
    const int bn_slice = p.BN / p.cs;                       // rows of the B tile this CTA loads (and multicasts)
    const uint32_t b_slice = static_cast<uint32_t>(bn_slice) * 128u;
    const uint32_t ntap_b = p.rowshare ? 3u : 1u;
    const uint32_t tx_bytes = pair ? 2u * (p.bytesA + ntap_b * b_slice) : p.bytesA + ntap_b * p.bytesB;
    const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cs) - 1);
    const uint32_t full_leader = pair ? ptx::mapa_u32(&full_bar[0], 0) : 0u;  // leader's full barriers (cluster address)
    // operand loads: plain / pair (bytes complete on the leader's barrier) / multicast B
    auto load_a = [&](const CUtensorMap* tm, int stage, uint8_t* dst, int c0, int c1, int c2, int c3) {
      if (pair)
        ptx::tma_load_4d_pair(tm, full_leader + stage * 8, dst, c0, c1, c2, c3);
      else
        ptx::tma_load_4d(tm, &full_bar[stage], dst, c0, c1, c2, c3);
    };
    auto load_b = [&](int stage, uint8_t* b_dst, int k0, int n0, int b0, int b1) {
      if (pair)
        ptx::tma_load_4d_pair(&p.tmB, full_leader + stage * 8, b_dst, k0, n0 + rank * bn_slice, b0, b1);
      else if (p.cs > 1)
        ptx::tma_load_4d_mcast(&p.tmB, &full_bar[stage], b_dst + rank * b_slice, k0, n0 + rank * bn_slice, b0, b1, mc_mask);
      else
        ptx::tma_load_4d(&p.tmB, &full_bar[stage], b_dst, k0, n0, b0, b1);
    };
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      int n_tile, m_tile, b0, b1;
      gemm_decode_tile(p, tile, n_tile, m_tile, b0, b1, rank);
      const int n0 = n_tile * p.BN;
      int m0 = m_tile * GEMM_BM, img = 0, h0 = 0, w0 = 0;
      if (p.conv) {
        int tw = m_tile % p.tiles_w;
        int r = m_tile / p.tiles_w;
        int th = r % p.tiles_h;
        img = r / p.tiles_h;
        h0 = th * p.BH;
        w0 = tw * p.BW;
      }
      if (p.res_prefetch && leader && m_tile < p.m_tiles) ptx::tma_prefetch_l2_4d(&p.tmR, n0, w0, h0, img);
      int kc_begin = 0, kc_end = p.num_k_chunks;
      if (p.ksplit > 1) {
        kc_begin = b0 * p.kc_per_split, kc_end = min(p.num_k_chunks, kc_begin + p.kc_per_split);
        b0 = 0;
      }
      // conv modes: tap / channel-chunk counters and the B k-offset of this phase
      int tap = 0, cc = 0;
      if (p.conv) tap = kc_begin / p.chunks_per_tap, cc = kc_begin - tap * p.chunks_per_tap;
      const int pa = (p.conv == 1 && p.nphase > 1) ? (b0 >> 1) : 0, pb = (p.conv == 1 && p.nphase > 1) ? (b0 & 1) : 0;
      const int kb_off = (p.conv == 1 && p.nphase > 1) ? b0 * p.num_k_chunks * GEMM_BK : 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          if (!pair || rank == 0) ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
          uint8_t* a_dst = sA + stage * p.a_stage;
          uint8_t* b_dst = sB + stage * b_stage;
          const int k0 = kc * GEMM_BK;
          if (p.rowshare) {  // `tap` counts kernel rows here: one 130-pixel A box, three B tiles
            load_a(&p.tmA, stage, a_dst, cc * GEMM_BK, w0 - 1, h0 + tap - 1, img);
#pragma unroll
            for (int sx = 0; sx < 3; ++sx)
              load_b(stage, b_dst + sx * b_tile, ((tap * 3 + sx) * p.chunks_per_tap + cc) * GEMM_BK, n0, 0, 0);
          } else if (p.conv == 1) {
            const int r = (p.taps_w == 3) ? (tap >= 6 ? 2 : (tap >= 3 ? 1 : 0)) : (tap >> 1);
            const int sx = tap - r * p.taps_w;
            load_a(&p.tmA, stage, a_dst, cc * GEMM_BK, w0 + sx + p.off_w0 + pb, h0 + r + p.off_h0 + pa, img);
            load_b(stage, b_dst, k0 + kb_off, n0, 0, 0);
          } else if (p.conv == 2) {  // tap = phase * 4 + tr * 2 + ts
            const int ph = tap >> 2, tr = (tap >> 1) & 1, ts = tap & 1;
            const CUtensorMap* tm = ph == 0 ? &p.tmA : (ph == 1 ? &p.tmA1 : (ph == 2 ? &p.tmA2 : &p.tmA3));
            load_a(tm, stage, a_dst, cc * GEMM_BK, w0 - (ts - 1 + (ph & 1)), h0 - (tr - 1 + (ph >> 1)), img);
            load_b(stage, b_dst, k0, n0, 0, 0);
          } else {
            if (!p.a_mn) {
              load_a(&p.tmA, stage, a_dst, k0, m0, b0, b1);
            } else {
              load_a(&p.tmA, stage, a_dst, m0, k0, b0, b1);
              load_a(&p.tmA, stage, a_dst + 8192, m0 + 64, k0, b0, b1);
            }
            if (!p.b_mn) {
              load_b(stage, b_dst, k0, n0, b0, b1);
            } else {
              for (int i = 0; i < p.BN / 64; ++i)
                ptx::tma_load_4d(&p.tmB, &full_bar[stage], b_dst + i * 8192, n0 + i * 64, k0, b0, b1);
            }
          }
          if (pair && rank != 0) ptx::mbar_arrive_cluster(full_leader + stage * 8);  // this CTA's copies are issued
        }
        if (++cc == p.chunks_per_tap) cc = 0, ++tap;
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 && pair && rank != 0) {
    // the pair's MMAs are issued by the leader CTA only
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    // Whole warp runs the loop with warp-uniform values, one elected lane issues tcgen05.mma / commit.  The smem
    // descriptors are linear in (stage, k): only the low word changes, by precomputed increments.
    const bool leader = ptx::elect_one();
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const uint64_t a_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sA), p.a_mn ? 8192u : 16u, 1024);
    const uint64_t b_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(sB), p.b_mn ? 8192u : 16u, 1024);
    const uint32_t a_kinc = (p.a_mn ? 2048u : 32u) >> 4, b_kinc = (p.b_mn ? 2048u : 32u) >> 4;
    const uint32_t a_sinc = static_cast<uint32_t>(p.a_stage) >> 4, b_sinc = b_stage >> 4;
    const uint32_t idesc = p.idesc;
    const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cs) - 1);
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as) * 256u;
      int nkc = p.num_k_chunks;
      if (p.ksplit > 1) {
        int n_tile, m_tile, b0, b1;
        gemm_decode_tile(p, tile, n_tile, m_tile, b0, b1, rank);
        nkc = min(p.num_k_chunks, (b0 + 1) * p.kc_per_split) - b0 * p.kc_per_split;
      }
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (leader) {
          const uint64_t ad = a_desc0 + static_cast<uint64_t>(stage * a_sinc);
          const uint64_t bd = b_desc0 + static_cast<uint64_t>(stage * b_sinc);
          if (p.rowshare) {
#pragma unroll
            for (uint32_t sx = 0; sx < 3; ++sx) {
              // view of the A box shifted by sx pixel rows (128 B each): start address field + 8 (x 16 B)
              const uint64_t as = ad + (sx << 3) + (p.rowshare == 2 ? (static_cast<uint64_t>(sx) << 49) : 0ull);
              const uint64_t bs = bd + sx * (b_tile >> 4);
#pragma unroll
              for (uint32_t kk = 0; kk < 4; ++kk) {
                const uint32_t accu = (kc != 0 || sx != 0 || kk != 0) ? 1u : 0u;
                if (pair)
                  ptx::umma_bf16_pair(d_tmem, as + kk * a_kinc, bs + kk * b_kinc, idesc, accu);
                else
                  ptx::umma_bf16(d_tmem, as + kk * a_kinc, bs + kk * b_kinc, idesc, accu);
              }
            }
            if (pair) ptx::umma_commit_pair(&empty_bar[stage], mc_mask);
          } else if (pair) {
            ptx::umma_bf16_pair(d_tmem, ad, bd, idesc, kc != 0 ? 1u : 0u);
            ptx::umma_bf16_pair(d_tmem, ad + a_kinc, bd + b_kinc, idesc, 1u);
            ptx::umma_bf16_pair(d_tmem, ad + 2 * a_kinc, bd + 2 * b_kinc, idesc, 1u);
            ptx::umma_bf16_pair(d_tmem, ad + 3 * a_kinc, bd + 3 * b_kinc, idesc, 1u);
            ptx::umma_commit_pair(&empty_bar[stage], mc_mask);  // release the stage in both CTAs
          } else {
            ptx::umma_bf16(d_tmem, ad, bd, idesc, kc != 0 ? 1u : 0u);
            ptx::umma_bf16(d_tmem, ad + a_kinc, bd + b_kinc, idesc, 1u);
            ptx::umma_bf16(d_tmem, ad + 2 * a_kinc, bd + 2 * b_kinc, idesc, 1u);
            ptx::umma_bf16(d_tmem, ad + 3 * a_kinc, bd + 3 * b_kinc, idesc, 1u);
          }
          if (pair) {
          } else if (p.cs > 1)
            ptx::umma_commit_mcast(&empty_bar[stage], mc_mask);  // release the stage in every CTA of the cluster
          else
            ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (leader) {  // accumulator complete -> epilogue (of both CTAs in pair mode)
        if (pair)
          ptx::umma_commit_pair(&tfull_bar[as], mc_mask);
        else
          ptx::umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (TMEM -> global)
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t ring = 0;  // chunk counter selecting the store panel
    const uint32_t ring_mask = static_cast<uint32_t>(p.npanel - 1);
    const bool store_leader = (warp == 2) && ptx::elect_one();
    const uint32_t tempty_leader = pair ? ptx::mapa_u32(&tempty_bar[0], 0) : 0u;
    // GroupNorm statistics of the output (see GemmParams::gn_partial)
    const int etid = threadIdx.x - 64;  // 0..127
    float* gacc = reinterpret_cast<float*>(smem + p.gn_smem_off);
    int gn_img = -1;
    if (p.gn_cpg) {
      for (int i = 0; i < 64; ++i) gacc[i * 128 + etid] = 0.f;  // each thread only ever touches its own column
      if (etid < 64)  // this CTA's rows of the partial table start at zero; flushes accumulate into them (a CTA may come
                      // back to an image: the phases of the fused upsample conv sweep the images four times)
        for (int im = 0; im < p.gn_nimg; ++im) p.gn_partial[(static_cast<long long>(im) * gridDim.x + blockIdx.x) * 64 + etid] = 0.f;
    }
    auto gn_flush = [&](int im) {
      float* wsum = gacc + 64 * 128;
      for (int g2 = 0; g2 < 64; ++g2) {
        float v = gacc[g2 * 128 + etid];
        gacc[g2 * 128 + etid] = 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wsum[(etid >> 5) * 64 + g2] = v;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (etid < 64)
        p.gn_partial[(static_cast<long long>(im) * gridDim.x + blockIdx.x) * 64 + etid] +=
            wsum[etid] + wsum[64 + etid] + wsum[128 + etid] + wsum[192 + etid];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      int n_tile, m_tile, b0, b1;
      gemm_decode_tile(p, tile, n_tile, m_tile, b0, b1, rank);
      const int n0 = n_tile * p.BN;
      bool valid;
      long long off_c, off_r;
      int img;
      int sc1 = 0, sc2 = 0, sc3 = 0;  // TMA-store coordinates (dims 1..3) of this tile
      if (p.conv) {
        int tw = m_tile % p.tiles_w;
        int r = m_tile / p.tiles_w;
        int th = r % p.tiles_h;
        img = r / p.tiles_h;
        int hh = th * p.BH + row / p.BW;
        int ww = tw * p.BW + row % p.BW;
        valid = (row < p.BW * p.BH) && hh < p.H && ww < p.W && m_tile < p.m_tiles;
        sc1 = tw * p.BW, sc2 = th * p.BH, sc3 = img;
        long long pix = (static_cast<long long>(hh) * p.W + ww);
        off_c = img * p.sc1 + hh * p.out_sh + ww * p.out_sw;
        if (p.nphase > 1) off_c += (b0 >> 1) * p.out_pa + (b0 & 1) * p.out_pb;
        off_r = img * p.sr1 + pix * p.ldr;
      } else {
        int m = m_tile * GEMM_BM + row;
        valid = m < p.M;
        img = b1;
        sc1 = m_tile * GEMM_BM, sc2 = b0, sc3 = b1;
        // split-K: the batch slot b0 carries the split index (unbatched launches only), it must not move the row
        const int bb0 = p.ksplit > 1 ? 0 : b0;
        off_c = bb0 * p.sc0 + b1 * p.sc1 + static_cast<long long>(m) * p.ldc;
        off_r = bb0 * p.sr0 + b1 * p.sr1 + static_cast<long long>(m) * p.ldr;
      }
      if (p.gn_cpg && m_tile < p.m_tiles && img != gn_img) {  // tiles are visited image by image
        if (gn_img >= 0) gn_flush(gn_img);
        gn_img = img;
      }
      // epilogue operands (split-K partials go to the fp32 workspace, raw)
      void* e_out = p.out;
      int e_f32 = p.out_f32, e_vec = p.vec_ok;
      const float *e_bias = p.bias, *e_bias_img = p.bias_img;
      const __nv_bfloat16* e_res = p.res;
      float e_alpha = p.alpha;
      if (p.ksplit > 1) {
        e_out = p.ws + b0 * p.ws_split_stride;
        off_c = (off_c / p.ldc) * p.ws_ld;  // same row index, workspace row stride
        e_f32 = 1, e_vec = 1, e_bias = nullptr, e_bias_img = nullptr, e_res = nullptr, e_alpha = 1.f;
      }
      // Stage bias (+ per-image bias) of this tile's columns in shared memory once (double-buffered by accumulator
      // stage), so the column loop below has no dependent global loads besides the residual.
      float* sbias = sbias_base + as * 256;
      const bool has_bias = (e_bias != nullptr) || (e_bias_img != nullptr);
      if (has_bias) {
        for (int c = threadIdx.x - 64; c < p.BN; c += 128) {
          const int col = n0 + c;
          float b = 0.f;
          if (col < p.N) {
            if (e_bias) b += __ldg(e_bias + col);
            if (e_bias_img) b += __ldg(e_bias_img + static_cast<long long>(img) * p.N + col);
          }
          sbias[c] = b;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      // residual rows are fetched one 32-column chunk ahead (the first one while the tile's MMAs are still running)
      const int n_chunks = (p.BN + 31) >> 5;
      auto load_res = [&](uint4(&dst)[4], const int ci) {
        const int col0 = n0 + (ci << 5);
        const int ncol = min(32, min(p.BN - (ci << 5), p.N - col0));
        if (e_res && valid && e_vec && ncol == 32) {
          const uint4* r4 = reinterpret_cast<const uint4*>(e_res + off_r + col0);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = r4[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(0, 0, 0, 0);
        }
      };
      uint4 resA[4], resB[4];
      load_res(resA, 0);
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as) * 256u;
      // 32-column chunks, TMEM loads double-buffered in registers: the load of chunk c+1 is in flight while chunk c is
      // converted and stored.  A chunk may extend past BN (BN is a multiple of 16); those columns are masked.
      uint32_t bufA[32], bufB[32];
      ptx::tmem_ld32(t_row, bufA);
      const bool use_ts = p.tma_store && !e_f32;
      const bool tile_ok = m_tile < p.m_tiles || !p.conv;
      auto process = [&](const uint32_t(&raw)[32], uint32_t(&nxt)[32], const uint4(&rres)[4], uint4(&rnxt)[4], const int ci) {
        const int c0 = ci << 5;
        const int col0 = n0 + c0;
        const int ncol = min(32, min(p.BN - c0, p.N - col0));  // valid columns of this chunk (may be <= 0)
        const bool full = ((ncol == 32) || use_ts) && e_vec;
        uint8_t* panel = spanel + (ring & ring_mask) * GEMM_PANEL;
        if (ci + 1 < n_chunks) load_res(rnxt, ci + 1);
        ptx::tmem_ld_wait();
        if (ci + 1 < n_chunks) ptx::tmem_ld32(t_row + c0 + 32, nxt);
        if (use_ts) {
          // stage this 128 x 32 chunk (64-byte rows, 64B swizzle) and hand it to the TMA store engine
          if (valid && ncol > 0) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * e_alpha;
            if (has_bias) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
                v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
              }
            }
            if (e_res) {
              if (ncol == 32) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rres[j]);
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    float2 f = __bfloat1622float2(h[t]);
                    v[8 * j + 2 * t] += f.x;
                    v[8 * j + 2 * t + 1] += f.y;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < ncol) v[j] += __bfloat162float(e_res[off_r + col0 + j]);
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.gn_cpg) gn_accumulate(v, gacc + etid, p.gn_cpg, col0);
            uint4* o4 = reinterpret_cast<uint4*>(panel + row * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
              o4[j ^ ((row >> 1) & 3)] = o;
            }
          }
          ptx::fence_proxy_async_smem();
          // the panel written next (ring + 1) must have been read out by its previous store: with 4 panels that is the
          // store issued three chunks ago, so two may stay in flight; with 2 panels every earlier store must be done
          if (store_leader) {
            if (ring_mask == 3)
              ptx::bulk_wait_read<2>();
            else
              ptx::bulk_wait_read<0>();
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (store_leader) {
            if (tile_ok && ncol > 0) {
              const CUtensorMap* tm = (p.nphase > 1) ? (b0 == 0 ? &p.tmC : (b0 == 1 ? &p.tmC1 : (b0 == 2 ? &p.tmC2 : &p.tmC3))) : &p.tmC;
              ptx::tma_store_4d(tm, panel, col0, sc1, sc2, sc3);
            }
            ptx::bulk_commit();  // one (possibly empty) group per chunk keeps the wait depth above exact
          }
          ++ring;
          return;
        }
        if (!valid || ncol <= 0) return;
        if (full) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * e_alpha;
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(sbias + c0 + 4 * j);
              v[4 * j] += b.x, v[4 * j + 1] += b.y, v[4 * j + 2] += b.z, v[4 * j + 3] += b.w;
            }
          }
          if (e_res) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rres[j]);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float2 f = __bfloat1622float2(h[t]);
                v[8 * j + 2 * t] += f.x;
                v[8 * j + 2 * t + 1] += f.y;
              }
            }
          }
          if (p.relu && p.ksplit <= 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (e_f32) {
            float4* o4 = reinterpret_cast<float4*>(static_cast<float*>(e_out) + off_c + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(e_out) + off_c + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
              o4[j] = o;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j >= ncol) break;
            const int col = col0 + j;
            float x = __uint_as_float(raw[j]) * e_alpha;
            if (has_bias) x += sbias[c0 + j];
            if (e_res) x += __bfloat162float(e_res[off_r + col]);
            if (p.relu && p.ksplit <= 1) x = fmaxf(x, 0.f);
            if (e_f32)
              static_cast<float*>(e_out)[off_c + col] = x;
            else
              static_cast<__nv_bfloat16*>(e_out)[off_c + col] = __float2bfloat16(x);
          }
        }
      };
#pragma unroll 1
      for (int ci = 0; ci < n_chunks; ci += 2) {
        process(bufA, bufB, resA, resB, ci);
        if (ci + 1 < n_chunks) process(bufB, bufA, resB, resA, ci + 1);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair)
          ptx::mbar_arrive_cluster(tempty_leader + as * 8);
        else
          ptx::mbar_arrive(&tempty_bar[as]);
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (p.gn_cpg && gn_img >= 0) gn_flush(gn_img);  // last image of this CTA
  }

  if (p.tma_store && warp == 2) ptx::bulk_wait_all();  // (only the issuing lane has groups pending; a no-op elsewhere)
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (p.cs > 1) ptx::cluster_sync_all();  // no CTA may exit while a peer can still multicast into it / arrive on its barriers
  if (warp == 1) {
    if (pair)
      ptx::tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
    else
      ptx::tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}
// Two entry points: a kernel containing cta_group::2 instructions can only be launched as a cluster of CTA pairs.
__global__ void __launch_bounds__(GEMM_THREADS, 1) umma_gemm_kernel(const __grid_constant__ GemmParams p) {
  umma_gemm_body<false>(p);
}
__global__ void __launch_bounds__(GEMM_THREADS, 1) umma_gemm_pair_kernel(const __grid_constant__ GemmParams p) {
  umma_gemm_body<true>(p);
}

// ======================================================================= host side
struct HostError {
  std::string msg;
};
#define MDC_CHECK(cond, ...)                              \
  do {                                                    \
    if (!(cond)) {                                        \
      char _b[512];                                       \
      snprintf(_b, sizeof(_b), __VA_ARGS__);              \
      throw ::mdc::HostError{std::string(_b)};            \
    }                                                     \
  } while (0)
#define MDC_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw ::mdc::HostError{std::string(#expr) + ": " + cudaGetErrorString(_e)};                  \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MDC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    MDC_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 4-D bf16 tensor map, 128B swizzle, zero fill out of bounds.  dims/box in elements (dim 0 innermost),
// strides in elements for dims 1..3.
inline CUtensorMap make_tmap_bf16(const void* base, const uint64_t dims[4], const uint64_t strides_el[3],
                                  const uint32_t box[4], CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gdims[4], gstr[3];
  cuuint32_t gbox[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) gdims[i] = dims[i], gbox[i] = box[i];
  for (int i = 0; i < 3; ++i) gstr[i] = strides_el[i] * 2;
  MDC_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  for (int i = 0; i < 3; ++i)
    MDC_CHECK(gstr[i] % 16 == 0 && gstr[i] > 0, "TMA stride %d = %llu bytes invalid (dims %llu %llu %llu %llu box %u %u %u %u)", i,
              (unsigned long long)gstr[i], (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
  MDC_CHECK(swz == CU_TENSOR_MAP_SWIZZLE_NONE || box[0] * 2 <= (swz == CU_TENSOR_MAP_SWIZZLE_64B ? 64u : 128u),
            "inner box exceeds the swizzle span");
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdims, gstr, gbox, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MDC_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu box %u %u %u %u", (int)r,
            (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
  return m;
}

// A bf16 matrix operand.  mn_major = 0: element (mn, k) at ptr[mn*ld + k];  1: at ptr[k*ld + mn].
struct Operand {
  const void* ptr = nullptr;
  int mn_major = 0;
  long long ld = 0;
  long long sb0 = 0, sb1 = 0;  // batch strides in elements
};
struct Epilogue {
  void* out = nullptr;
  int out_f32 = 0;
  long long ldc = 0, sc0 = 0, sc1 = 0;
  const float* bias = nullptr;
  const float* bias_img = nullptr;
  const __nv_bfloat16* res = nullptr;
  long long ldr = 0, sr0 = 0, sr1 = 0;
  float alpha = 1.f;
  int relu = 0;
};

struct GemmPlan {
  GemmParams p;
  int grid = 0;
  int smem = 0;
  double flops = 0;
  long long rows = 0;  // output rows (pixels / tokens) -- used by the split-K reduction
};

// out[row, col] = alpha * sum_s ws[s][row][col] (+bias) (+bias_img) (+res) -> bf16 / fp32.  Rows are final output rows.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int ksplit, long long split_stride, long long ws_ld,
                                     long long rows, int N, void* __restrict__ out, int out_f32, long long ldc,
                                     const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res,
                                     long long ldr, float alpha) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int nv = (N + 3) >> 2;
  const long long total = rows * nv;
  // vector epilogue: every operand row starts on an 8-byte (bf16) / 16-byte (fp32) boundary
  const bool vec = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   (!res || ((ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(res) & 7) == 0));
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = static_cast<int>(i % nv) * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4  // four partial loads in flight per thread; the order of the additions is unchanged
    for (int s = 0; s < ksplit; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + s * split_stride + r * ws_ld + c);
      a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
    }
    float v[4] = {a.x * alpha, a.y * alpha, a.z * alpha, a.w * alpha};
    if (vec && c + 3 < N) {  // same arithmetic as the scalar tail below, one 8- / 16-byte access per operand
      if (bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += bias[c + j];
      }
      if (res) {
        const uint2 rr = *reinterpret_cast<const uint2*>(res + r * ldr + c);
        const float2 r0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.x));
        const float2 r1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr.y));
        v[0] += r0.x, v[1] += r0.y, v[2] += r1.x, v[3] += r1.y;
      }
      if (out_f32) {
        *reinterpret_cast<float4*>(static_cast<float*>(out) + r * ldc + c) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * ldc + c) = o;
      }
      continue;
    }
    for (int j = 0; j < 4 && c + j < N; ++j) {
      float x = v[j];
      if (bias) x += bias[c + j];
      if (res) x += __bfloat162float(res[r * ldr + c + j]);
      if (out_f32)
        static_cast<float*>(out)[r * ldc + c + j] = x;
      else
        static_cast<__nv_bfloat16*>(out)[r * ldc + c + j] = __float2bfloat16(x);
    }
  }
}

// Test / tuning overrides (set through mdc_dbg_tune only; 0 = automatic).
struct GemmTune {
  int bn = 0, cs = 0, ksplit = 0, wcopies = 1, rowshare = 0;
};
inline GemmTune& g_tune() {
  static GemmTune t;
  return t;
}

inline int pick_bn(int N) {
  if (g_tune().bn) return g_tune().bn;
  // largest divisor of N that is a multiple of 32 (whole 32-column chunks: the TMA-store epilogue applies), else of 16,
  // and <= 256; otherwise min(256, roundup16(N)).
  for (int bn = 256; bn >= 64; bn -= 32)
    if (N % bn == 0) return bn;
  for (int bn = 240; bn >= 64; bn -= 16)
    if (N % bn == 0) return bn;
  int r = ((N + 15) / 16) * 16;
  return r > 256 ? 256 : r;
}

// SM count of the CURRENT device (cached per device: a process may hold engines on several GPUs).
inline int g_num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}
// true the first time it is called with `done` on the current device (function attributes are per device)
inline bool first_use_on_device(bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

// Cost model (microseconds) shared by the split-K and the output-tile-width choices.  A k-chunk is bound by the L2 -> SM
// fill of its operand tiles (~100 GB/s per SM when every SM pulls), an item (tile x split) pays a pipeline fill +
// epilogue, a split launch pays the reduction kernel and the fp32 partials' round trip through L2.  Work items beyond
// one wave of CTAs serialise.  Returns the cost of the best split count (1 = unsplit) and that count.
inline double gemm_cost_us(int tiles, int nk, double bytesA, int BN, int cs, bool may_split, int* best_split) {
  const int sms = (g_num_sms() / cs) * cs;
  static const int max_tiles = getenv("MDC_SPLITK_MAXTILES") ? atoi(getenv("MDC_SPLITK_MAXTILES")) : 100;
  static const bool no_split = getenv("MDC_NO_SPLITK") != nullptr;
  const double t_chunk = (bytesA + BN * 128.0 / cs) / 100e3, t_item = 2.5, t_red = 4.0, l2_bytes_per_us = 5e6;
  auto waves = [&](int items) { return (items + sms - 1) / sms; };
  int best = 1;
  double best_cost = waves(tiles) * (nk * t_chunk + t_item);
  if (may_split && !no_split && tiles <= max_tiles && nk >= 8) {
    for (int s = 2; s <= std::min(nk / 4, 64); ++s) {
      const int kc = (nk + s - 1) / s;
      if ((nk + kc - 1) / kc != s) continue;  // enable_splitk would round this split count down
      const int items = tiles * s;
      const double cost = waves(items) * (kc * t_chunk + t_item) + t_red + 2.0 * items * GEMM_BM * BN * 4.0 / l2_bytes_per_us;
      if (cost < (best == 1 ? 0.9 * best_cost : best_cost)) best = s, best_cost = cost;
    }
  }
  if (best_split) *best_split = best;
  return best_cost;
}
// Output-tile width of a launch with few m-tiles (the 18x24 and 9x12 levels of the UNet: 432 / 108 rows): the widest
// divisor of N is not the best choice there -- narrower tiles give more CTAs for the weight stream and fewer idle SMs in
// the last wave.  The same model that picks the split count ranks the candidates (tools/linear_tune_sweep.py: within
// 0.5 % of the measured best over the UNet's linears at those levels, 10 % better than "widest divisor").
inline int pick_bn_model(int N, int m_tiles, int nk, double bytesA, bool can_pair, bool may_split) {
  const int wide = pick_bn(N);
  static const bool off = getenv("MDC_NO_BNMODEL") != nullptr;
  static const int max_mt = getenv("MDC_BNMODEL_MAXMT") ? atoi(getenv("MDC_BNMODEL_MAXMT")) : 8;
  if (off || g_tune().bn || m_tiles > max_mt) return wide;
  const int cs = (can_pair && m_tiles >= 2 && !getenv("MDC_NO_PAIR")) ? 2 : 1;
  auto cost = [&](int bn) { return gemm_cost_us(((m_tiles + cs - 1) / cs) * cs * ((N + bn - 1) / bn), nk, bytesA, bn, cs, may_split, nullptr); };
  const double wide_cost = cost(wide);
  int best = wide;
  double best_cost = wide_cost;
  for (int bn = 64; bn <= 256; bn += 32) {
    if (N % bn || bn == wide) continue;
    const double c = cost(bn);
    if (c < best_cost) best = bn, best_cost = c;
  }
  return best_cost < 0.97 * wide_cost ? best : wide;  // leave the widest divisor unless the model sees a real difference
}

// Cluster mode of a launch; must be decided BEFORE the B tensor map is encoded because each CTA's TMA box covers only
// its BN/cs rows.  pair: CTA-pair MMA (see GemmParams::pair); otherwise cs > 1 = B multicast over cs CTAs (consecutive
// m-tiles of one n-tile).  Both need a K-major B whose per-CTA slice is a whole number of 8-row swizzle atoms.
inline void decide_cluster(GemmParams& p) {
  static const bool no_pair = getenv("MDC_NO_PAIR") != nullptr;
  p.cs = 1, p.pair = 0;
  if (p.b_mn) return;
  const int t = g_tune().cs;  // 0 auto, 1 none, 2 pair, 3 multicast x2, 4 multicast x4
  if (t == 1) return;
  if (t == 3 || t == 4) {
    const int cs = t == 3 ? 2 : 4;
    if ((p.BN / cs) % 8 == 0) p.cs = cs;
    return;
  }
  if ((t == 2 || !no_pair) && p.m_tiles >= 2 && p.BN % 16 == 0) p.cs = 2, p.pair = 1;
}
inline void finish_plan(GemmPlan& g) {
  GemmParams& p = g.p;
  const int b_stage = (p.pair ? p.BN / 2 : p.BN) * 128 * (p.rowshare ? 3 : 1);
  p.a_stage = p.rowshare ? GEMM_A_STAGE_RS : GEMM_A_STAGE;
  const int gn_bytes = p.gn_cpg ? GEMM_GN_SMEM : 0;
  auto stages_for = [&](int npanel) {
    return std::min((GEMM_SMEM_BUDGET - GEMM_HEADER0 - npanel * GEMM_PANEL - 1024 - gn_bytes) / (p.a_stage + b_stage), GEMM_MAX_STAGES);
  };
  p.npanel = (stages_for(4) == stages_for(2)) ? 4 : 2;  // never trade a pipeline stage for store depth
  const int stages = std::max(2, stages_for(p.npanel));
  p.stages = stages;
  p.gn_smem_off = GEMM_HEADER0 + p.npanel * GEMM_PANEL + stages * (p.a_stage + b_stage);
  g.smem = GEMM_HEADER0 + p.npanel * GEMM_PANEL + 1024 + stages * (p.a_stage + b_stage) + gn_bytes;
  p.idesc = ptx::make_idesc_bf16(p.pair ? 2 * GEMM_BM : GEMM_BM, p.BN, p.a_mn, p.b_mn);
  if (p.cs < 1) p.cs = 1, p.pair = 0;  // set by the planners via decide_cluster() before the B map was built
  long long total = 1LL * ((p.m_tiles + p.cs - 1) / p.cs) * p.n_tiles * (p.ksplit > 1 ? p.ksplit : p.nb0) * p.nb1;
  g.grid = static_cast<int>(std::min<long long>(total * p.cs, (g_num_sms() / p.cs) * p.cs));
  const uintptr_t o = reinterpret_cast<uintptr_t>(p.out), r = reinterpret_cast<uintptr_t>(p.res);
  bool ok = (o % 16 == 0) && (p.ldc % 8 == 0) && (p.sc0 % 8 == 0) && (p.sc1 % 8 == 0) && (p.N % 8 == 0);
  if (p.res) ok = ok && (r % 16 == 0) && (p.ldr % 8 == 0) && (p.sr0 % 8 == 0) && (p.sr1 % 8 == 0);
  if (p.bias) ok = ok && (reinterpret_cast<uintptr_t>(p.bias) % 16 == 0);
  if (p.bias_img) ok = ok && (reinterpret_cast<uintptr_t>(p.bias_img) % 16 == 0) && (p.N % 4 == 0);
  p.vec_ok = ok ? 1 : 0;
  // residual prefetch into L2 (conv mode, single-phase, big images only: small ones live in L2 anyway)
  static const bool no_rp = getenv("MDC_NO_RESPREFETCH") != nullptr;
  p.res_prefetch = 0;
  if (!no_rp && p.res && p.conv == 1 && p.nphase == 1 && p.ksplit <= 1 && ok && p.m_tiles >= 512 && p.ldr % 8 == 0) {
    uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)std::max<long long>(1, p.m_tiles / (p.tiles_h * p.tiles_w))};
    uint64_t str[3] = {(uint64_t)p.ldr, (uint64_t)p.ldr * p.W, (uint64_t)p.sr1};
    uint32_t box[4] = {(uint32_t)std::min(p.BN, 256), (uint32_t)p.BW, (uint32_t)p.BH, 1};
    p.tmR = make_tmap_bf16(p.res, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    p.res_prefetch = 1;
  }
  // TMA-store epilogue: bf16 output, aligned, whole 32-column chunks per n-tile
  static const bool no_ts = getenv("MDC_NO_TMASTORE") != nullptr;
  p.tma_store = 0;
  if (!no_ts && ok && !p.out_f32 && p.BN % 32 == 0 && p.ksplit <= 1 && p.conv != 2 + 100) {
    __nv_bfloat16* base = static_cast<__nv_bfloat16*>(p.out);
    if (!p.conv) {
      uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.nb0, (uint64_t)p.nb1};
      uint64_t str[3] = {(uint64_t)p.ldc, (uint64_t)(p.sc0 ? p.sc0 : 8), (uint64_t)(p.sc1 ? p.sc1 : 8)};
      uint32_t box[4] = {32, GEMM_BM, 1, 1};
      p.tmC = make_tmap_bf16(base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
      p.tma_store = 1;
    } else {
      uint64_t dims[4] = {(uint64_t)p.N, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)std::max<long long>(1, p.m_tiles / (p.tiles_h * p.tiles_w))};
      uint64_t str[3] = {(uint64_t)p.out_sw, (uint64_t)p.out_sh, (uint64_t)p.sc1};
      uint32_t box[4] = {32, (uint32_t)p.BW, (uint32_t)p.BH, 1};
      CUtensorMap* maps[4] = {&p.tmC, &p.tmC1, &p.tmC2, &p.tmC3};
      const int nph = p.nphase > 1 ? 4 : 1;
      for (int ph = 0; ph < nph; ++ph)
        *maps[ph] = make_tmap_bf16(base + (ph >> 1) * p.out_pa + (ph & 1) * p.out_pb, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
      p.tma_store = 1;
    }
  }
}

inline void fill_epilogue(GemmParams& p, const Epilogue& e) {
  p.out = e.out, p.out_f32 = e.out_f32, p.ldc = e.ldc, p.sc0 = e.sc0, p.sc1 = e.sc1;
  p.bias = e.bias, p.bias_img = e.bias_img, p.res = e.res, p.ldr = e.ldr, p.sr0 = e.sr0, p.sr1 = e.sr1;
  p.alpha = e.alpha, p.relu = e.relu;
}

// Batched GEMM: for each (b0, b1): D = alpha * A . B^T (+...).  K need not be a multiple of 64 (TMA zero-fills).
inline GemmPlan plan_gemm(int M, int N, int K, const Operand& A, const Operand& B, const Epilogue& e, int nb0 = 1,
                          int nb1 = 1, int bn_override = 0) {
  GemmPlan g;
  memset(&g.p, 0, sizeof(g.p));
  GemmParams& p = g.p;
  p.M = M, p.N = N;
  p.nb0 = nb0, p.nb1 = nb1;
  p.a_mn = A.mn_major, p.b_mn = B.mn_major;
  p.BN = bn_override ? bn_override : pick_bn(N);
  if (!bn_override && !p.b_mn && !p.a_mn && nb0 == 1 && nb1 == 1)  // plain linears: model-ranked tile width on the small maps
    p.BN = pick_bn_model(N, (M + GEMM_BM - 1) / GEMM_BM, (K + GEMM_BK - 1) / GEMM_BK, GEMM_A_STAGE, true, !e.out_f32 && !e.bias_img);
  if (p.b_mn) p.BN = (N >= 256) ? 256 : ((N + 63) / 64) * 64;  // MN-major B is loaded in 64-wide chunks
  MDC_CHECK(p.BN % 16 == 0 && p.BN >= 16 && p.BN <= 256, "bad BN %d", p.BN);
  MDC_CHECK(!p.b_mn || p.BN % 64 == 0, "MN-major B needs BN %% 64 == 0");
  p.m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  p.n_tiles = (N + p.BN - 1) / p.BN;
  p.num_k_chunks = (K + GEMM_BK - 1) / GEMM_BK;
  p.bytesA = GEMM_A_STAGE;
  p.bytesB = p.BN * 128;
  auto mk = [&](const Operand& o, int MN, int rows_box) {
    uint64_t dims[4], str[3];
    uint32_t box[4];
    long long s0 = o.sb0 ? o.sb0 : 8, s1 = o.sb1 ? o.sb1 : 8;
    if (!o.mn_major) {
      dims[0] = K, dims[1] = MN, box[0] = 64, box[1] = rows_box;
    } else {
      dims[0] = MN, dims[1] = K, box[0] = 64, box[1] = 64;
    }
    dims[2] = nb0, dims[3] = nb1, box[2] = 1, box[3] = 1;
    str[0] = o.ld, str[1] = s0, str[2] = s1;
    return make_tmap_bf16(o.ptr, dims, str, box);
  };
  decide_cluster(p);
  p.tmA = mk(A, M, GEMM_BM);
  p.tmB = mk(B, N, p.BN / p.cs);
  fill_epilogue(p, e);
  finish_plan(g);
  g.flops = 2.0 * M * N * K * nb0 * nb1;
  g.rows = M;
  return g;
}

// Choose the rectangular pixel tile (BW x BH <= 128) that needs the fewest tiles for an H x W image.
inline void pick_conv_tile(int H, int W, int& BW, int& BH) {
  long long best = -1;
  BW = 1, BH = 1;
  for (int bw = 1; bw <= 128 && bw <= W; ++bw) {
    int bh = std::min(128 / bw, H);
    if (bh < 1) continue;
    long long tiles = 1LL * ((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    if (best < 0 || tiles < best || (tiles == best && bw > BW)) best = tiles, BW = bw, BH = bh;
  }
}

// 3x3 stride-1 pad-1 convolution over an NHWC bf16 tensor as an implicit GEMM.
//   x   : [NB, H, W, C] with pixel stride ldx (>= C, multiple of 8)
//   wpk : packed weights [Cout, 9 * Cp] bf16, Cp = roundup(C, 64), k = (r*3+s)*Cp + c
//   out : [NB, H, W, Cout] with pixel stride e.ldc; e.sc1 / e.sr1 are per-image strides.
inline GemmPlan plan_conv3x3(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const void* wpk,
                             Epilogue e) {
  GemmPlan g;
  memset(&g.p, 0, sizeof(g.p));
  GemmParams& p = g.p;
  const int Cp = ((C + 63) / 64) * 64;
  p.conv = 1;
  p.H = H, p.W = W;
  p.taps_w = 3, p.ntaps = 9, p.off_h0 = -1, p.off_w0 = -1, p.nphase = 1;
  p.out_sh = 1LL * W * e.ldc, p.out_sw = e.ldc;
  pick_conv_tile(H, W, p.BW, p.BH);
  p.tiles_w = (W + p.BW - 1) / p.BW;
  p.tiles_h = (H + p.BH - 1) / p.BH;
  p.m_tiles = NB * p.tiles_h * p.tiles_w;
  p.M = 0, p.N = Cout;
  p.nb0 = 1, p.nb1 = 1;
  p.BN = pick_bn(Cout);
  // few m-tiles (the 18x24 / 9x12 maps) are weight streaming: model-ranked tile width (tools/conv_tune_sweep.py)
  p.BN = pick_bn_model(Cout, p.m_tiles, 9 * (Cp / 64), 64.0 * p.BW * p.BH * 2.0, true, !e.bias_img && !e.relu);
  p.n_tiles = (Cout + p.BN - 1) / p.BN;
  p.chunks_per_tap = Cp / 64;
  p.num_k_chunks = 9 * p.chunks_per_tap;
  p.bytesA = 64u * p.BW * p.BH * 2u;
  p.bytesB = p.BN * 128;
  // row-shared taps: full-width row tiles, narrow N (the A operand dominates the L2 -> SM traffic), many tiles
  static const bool no_rs = getenv("MDC_NO_ROWSHARE") != nullptr;
  const int rs_tune = g_tune().rowshare;  // 0 auto, 1 off, 2 force when legal
  p.rowshare = (p.BW == 128 && p.BH == 1 && rs_tune != 1 && ((!no_rs && p.BN <= 128 && p.m_tiles >= 1024) || rs_tune >= 2)) ? 1 : 0;
  if (p.rowshare && rs_tune == 3) p.rowshare = 2;  // diagnostic variant (wrong on purpose): additionally sets the matrix base offset
  if (p.rowshare) {
    p.num_k_chunks = 3 * p.chunks_per_tap;  // one pipeline stage = one kernel row x one channel chunk = three taps
    p.bytesA = 64u * 130u * 2u;
  }
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldx, (uint64_t)ldx * W, (uint64_t)ldx * W * H};
    uint32_t box[4] = {64, (uint32_t)(p.rowshare ? 130 : p.BW), (uint32_t)p.BH, 1};
    p.tmA = make_tmap_bf16(x, dims, str, box);
  }
  {
    uint64_t dims[4] = {(uint64_t)9 * Cp, (uint64_t)Cout, 1, 1};
    uint64_t str[3] = {(uint64_t)9 * Cp, (uint64_t)9 * Cp * Cout, (uint64_t)9 * Cp * Cout};
    decide_cluster(p);
    uint32_t box[4] = {64, (uint32_t)(p.BN / p.cs), 1, 1};
    p.tmB = make_tmap_bf16(wpk, dims, str, box);
  }
  if (!e.sc1) e.sc1 = 1LL * H * W * e.ldc;
  if (e.res && !e.sr1) e.sr1 = 1LL * H * W * e.ldr;
  fill_epilogue(p, e);
  finish_plan(g);
  g.flops = 2.0 * NB * H * W * 9.0 * C * Cout;
  g.rows = 1LL * NB * H * W;
  return g;
}

// Kernel launch with the programmatic-stream-serialization attribute (PDL); MDC_NO_PDL=1 disables it.
inline bool g_use_pdl() {
  static int v = -1;
  if (v < 0) v = getenv("MDC_NO_PDL") ? 0 : 1;
  return v == 1;
}
// MDC_DEBUG_SYNC=1 (not under graph capture): synchronise after every launch and name the kernel that faulted.
inline void debug_sync(const void* func, cudaStream_t st) {
  static const bool on = getenv("MDC_DEBUG_SYNC") != nullptr;
  if (!on) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs != cudaStreamCaptureStatusNone) return;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    const char* name = "?";
    cudaFuncGetName(&name, func);
    throw HostError{std::string("kernel ") + name + " failed: " + cudaGetErrorString(e)};
  }
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_use_pdl() ? 1 : 0;
  MDC_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  debug_sync(reinterpret_cast<const void*>(kernel), st);
}

// Fused nearest-2x upsample + conv3x3 (pad 1) as four 2x2 "phase" convolutions on the LOW-resolution input:
//   y[2i+a, 2j+b] = sum_{tr,ts} Weff[a][b][tr][ts] . x[i + tr - 1 + a, j + ts - 1 + b]      (2.25x fewer FLOPs, no upsampled
// tensor).  x: [NB, H, W, C] (pixel stride ldx); wpk: [Cout, 16 * Cp] with k = ((phase*4 + tr*2 + ts) * Cp + c);
// out: [NB, 2H, 2W, Cout] (pixel stride ldc).
inline GemmPlan plan_upconv_fwd(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const void* wpk,
                                Epilogue e) {
  GemmPlan g;
  memset(&g.p, 0, sizeof(g.p));
  GemmParams& p = g.p;
  const int Cp = ((C + 63) / 64) * 64;
  p.conv = 1;
  p.H = H, p.W = W;
  p.taps_w = 2, p.ntaps = 4, p.off_h0 = -1, p.off_w0 = -1, p.nphase = 4;
  const long long Wf = 2LL * W, Hf = 2LL * H;
  p.out_sh = 2 * Wf * e.ldc, p.out_sw = 2 * e.ldc, p.out_pa = Wf * e.ldc, p.out_pb = e.ldc;
  pick_conv_tile(H, W, p.BW, p.BH);
  p.tiles_w = (W + p.BW - 1) / p.BW;
  p.tiles_h = (H + p.BH - 1) / p.BH;
  p.m_tiles = NB * p.tiles_h * p.tiles_w;
  p.M = 0, p.N = Cout;
  p.nb0 = 4, p.nb1 = 1;
  p.BN = pick_bn(Cout);
  p.n_tiles = (Cout + p.BN - 1) / p.BN;
  p.chunks_per_tap = Cp / 64;
  p.num_k_chunks = 4 * p.chunks_per_tap;
  p.bytesA = 64u * p.BW * p.BH * 2u;
  p.bytesB = p.BN * 128;
  {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldx, (uint64_t)ldx * W, (uint64_t)ldx * W * H};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    p.tmA = make_tmap_bf16(x, dims, str, box);
  }
  {
    uint64_t dims[4] = {(uint64_t)16 * Cp, (uint64_t)Cout, 1, 1};
    uint64_t str[3] = {(uint64_t)16 * Cp, (uint64_t)16 * Cp * Cout, (uint64_t)16 * Cp * Cout};
    decide_cluster(p);
    uint32_t box[4] = {64, (uint32_t)(p.BN / p.cs), 1, 1};
    p.tmB = make_tmap_bf16(wpk, dims, str, box);
  }
  e.sc1 = Hf * Wf * e.ldc;
  fill_epilogue(p, e);
  finish_plan(g);
  g.flops = 2.0 * NB * Hf * Wf * 9.0 * C * Cout;  // algorithmic FLOPs of the op it replaces
  g.rows = 1LL * NB * H * W;
  return g;
}
// Input gradient of the fused op: dx[i,j] = sum_{phase,tr,ts} Weff^T . dy_phase[i - (tr-1+a), j - (ts-1+b)].
//   dy: [NB, 2H, 2W, Cout] (pixel stride lddy); wpk: [C, 16 * Coutp], k = (phase*4 + tr*2 + ts) * Coutp + co; out: [NB,H,W,C].
inline GemmPlan plan_upconv_bwd(int NB, int H, int W, int C, int Cout, const void* dy, long long lddy, const void* wpk,
                                Epilogue e) {
  GemmPlan g;
  memset(&g.p, 0, sizeof(g.p));
  GemmParams& p = g.p;
  const int Cop = ((Cout + 63) / 64) * 64;
  p.conv = 2;
  p.H = H, p.W = W;
  p.taps_w = 2, p.ntaps = 16, p.nphase = 1;
  p.out_sh = 1LL * W * e.ldc, p.out_sw = e.ldc;
  pick_conv_tile(H, W, p.BW, p.BH);
  p.tiles_w = (W + p.BW - 1) / p.BW;
  p.tiles_h = (H + p.BH - 1) / p.BH;
  p.m_tiles = NB * p.tiles_h * p.tiles_w;
  p.M = 0, p.N = C;
  p.nb0 = 1, p.nb1 = 1;
  p.BN = pick_bn(C);
  p.n_tiles = (C + p.BN - 1) / p.BN;
  p.chunks_per_tap = Cop / 64;
  p.num_k_chunks = 16 * p.chunks_per_tap;
  p.bytesA = 64u * p.BW * p.BH * 2u;
  p.bytesB = p.BN * 128;
  const long long Wf = 2LL * W, Hf = 2LL * H;
  CUtensorMap* maps[4] = {&p.tmA, &p.tmA1, &p.tmA2, &p.tmA3};
  for (int ph = 0; ph < 4; ++ph) {
    const int a = ph >> 1, b = ph & 1;
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(dy) + (a * Wf + b) * lddy;
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)(2 * lddy), (uint64_t)(2 * Wf * lddy), (uint64_t)(Hf * Wf * lddy)};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    *maps[ph] = make_tmap_bf16(base, dims, str, box);
  }
  {
    uint64_t dims[4] = {(uint64_t)16 * Cop, (uint64_t)C, 1, 1};
    uint64_t str[3] = {(uint64_t)16 * Cop, (uint64_t)16 * Cop * C, (uint64_t)16 * Cop * C};
    decide_cluster(p);
    uint32_t box[4] = {64, (uint32_t)(p.BN / p.cs), 1, 1};
    p.tmB = make_tmap_bf16(wpk, dims, str, box);
  }
  if (!e.sc1) e.sc1 = 1LL * H * W * e.ldc;
  if (e.res && !e.sr1) e.sr1 = 1LL * H * W * e.ldr;
  fill_epilogue(p, e);
  finish_plan(g);
  g.flops = 2.0 * NB * Hf * Wf * 9.0 * C * Cout;
  g.rows = 1LL * NB * H * W;
  return g;
}

inline void launch_gemm_kernel(const GemmPlan& g, cudaStream_t st);
inline void gemm_set_smem_attr() {
  static bool done[64] = {false};
  if (first_use_on_device(done)) {
    MDC_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    MDC_CUDA(cudaFuncSetAttribute(umma_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
}

inline void launch_gemm_kernel(const GemmPlan& g, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g.grid), cfg.blockDim = dim3(GEMM_THREADS), cfg.dynamicSmemBytes = g.smem + 1024, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g.p.cs > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = g.p.cs, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr, cfg.numAttrs = na;
  MDC_CUDA(cudaLaunchKernelEx(&cfg, g.p.pair ? umma_gemm_pair_kernel : umma_gemm_kernel, g.p));
  if (getenv("MDC_DEBUG_SYNC")) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs == cudaStreamCaptureStatusNone) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        char b[400];
        snprintf(b, sizeof(b), "GEMM failed (%s): conv %d M %d N %d BN %d tiles %d x %d k-chunks %d H %d W %d pair %d ksplit %d nphase %d res %p out %p",
                 cudaGetErrorString(e), g.p.conv, g.p.M, g.p.N, g.p.BN, g.p.m_tiles, g.p.n_tiles, g.p.num_k_chunks, g.p.H, g.p.W, g.p.pair,
                 g.p.ksplit, g.p.nphase, (const void*)g.p.res, g.p.out);
        throw HostError{std::string(b)};
      }
    }
  }
}

// Decide on split-K for a finished plan: few output tiles, long K loop.  `ws` must hold ws_floats(plan) floats.
inline int choose_ksplit(const GemmPlan& g) {
  const GemmParams& p = g.p;
  if (p.nb0 != 1 || p.nb1 != 1 || p.bias_img || p.out_f32 || p.nphase > 1 || p.relu) return 1;
  if (g_tune().ksplit > 0) return g_tune().ksplit;
  const int tiles = ((p.m_tiles + p.cs - 1) / p.cs) * p.cs * p.n_tiles;
  int best = 1;
  gemm_cost_us(tiles, p.num_k_chunks, p.bytesA, p.BN, p.cs, true, &best);
  return best;
}
inline size_t enable_splitk(GemmPlan& g, int ksplit) {  // returns the workspace size in floats
  GemmParams& p = g.p;
  if (ksplit <= 1) return 0;
  p.kc_per_split = (p.num_k_chunks + ksplit - 1) / ksplit;
  p.ksplit = (p.num_k_chunks + p.kc_per_split - 1) / p.kc_per_split;
  if (p.ksplit <= 1) {
    p.ksplit = 0;
    return 0;
  }
  p.ws_ld = ((p.N + 15) / 16) * 16;
  p.ws_split_stride = g.rows * p.ws_ld;
  g.grid = std::min(((p.m_tiles + p.cs - 1) / p.cs) * p.cs * p.n_tiles * p.ksplit, (g_num_sms() / p.cs) * p.cs);
  return static_cast<size_t>(p.ksplit) * p.ws_split_stride;
}

inline void run_gemm(const GemmPlan& g, cudaStream_t st) {
  gemm_set_smem_attr();
  launch_gemm_kernel(g, st);
  if (g.p.ksplit > 1) {
    const GemmParams& p = g.p;
    const long long work = g.rows * ((p.N + 3) / 4);
    const int grid = static_cast<int>(std::min<long long>((work + 255) / 256, 148 * 8));
    launch_k(splitk_reduce_kernel, dim3(grid), dim3(256), 0, st, p.ws, p.ksplit, p.ws_split_stride, p.ws_ld, g.rows, p.N,
             p.out, p.out_f32, p.ldc, p.bias, p.res, p.ldr, p.alpha);
  }
}

}  // namespace mdc
