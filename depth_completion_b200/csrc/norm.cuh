// Host-side planning and launching of the GroupNorm(+SiLU) kernels (kernels.cuh), shared by the step engine and the
// kernel-level debug entry points so that the tests exercise exactly the selector the engine uses.
//
//   * tensors whose per-CTA slab fits in shared memory: ONE launch (gn_fused_fwd / gn_fused_bwd, grid barrier);
//   * larger tensors (the 56-226 MB decoder activations): two passes (statistics, then apply), the statistics pass
//     being skipped when the producing tcgen05 GEMM / conv already emitted them from its epilogue (GemmParams::gn_*).
#pragma once
#include <vector>

#include "gemm.cuh"
#include "kernels.cuh"

namespace mdc {

constexpr size_t GN_FUSED_SMEM_CAP = 200 * 1024;

struct GNPlan {
  GNShape s;    // two-pass tiling
  GNShape sfu;  // single-launch tiling (grid-barrier variant)
  int G = 0, threads = 0, threads_b = 0;
  bool fuse_f = false, fuse_b = false;  // grid-barrier single-launch kernels (tensors whose slab fits one SM each)
  size_t smem_f = 0, smem_b = 0;
  // cluster single-launch kernels (one cluster of K CTAs per (image, group), no grid barrier): concurrent handles
  bool cl_f = false, cl_b = false;
  GNClusterShape cs_f{}, cs_b{};
  size_t cl_smem_f = 0, cl_smem_b = 0;
  size_t partial_floats = 0;  // scratch the launchers need (per-block partial sums)
  bool single_f() const { return cl_f || fuse_f; }
  bool single_b() const { return cl_b || fuse_b; }
};
inline size_t gnc_smem_bytes(int pix_per_cta, int cpg, bool bwd) {
  const size_t slab = (static_cast<size_t>(pix_per_cta) * cpg * 2 + 15) & ~size_t(15);
  return slab + (bwd ? static_cast<size_t>(pix_per_cta) * cpg * 4 : 0) + 68 * sizeof(float);
}
// Smallest cluster size K in {1, 2, 4, 8} that gives the chip enough CTAs and whose pixel slice fits in shared memory.
inline bool plan_gn_cluster(int N, int HW, int C, int G, long long ld, bool bwd, GNClusterShape& cs, size_t& smem) {
  const int cpg = C / G;
  int kpar = 1;
  while (kpar < 8 && N * G * kpar < 96) kpar *= 2;
  for (int K = kpar; K <= 8; K *= 2) {
    const int ppc = (HW + K - 1) / K;
    const size_t b = gnc_smem_bytes(ppc, cpg, bwd);
    if (b <= GN_FUSED_SMEM_CAP) {
      cs.N = N, cs.HW = HW, cs.C = C, cs.G = G, cs.K = K, cs.ld = ld, cs.pix_per_cta = ppc;
      smem = b;
      return true;
    }
  }
  return false;
}


// Function attributes are per device: call once per device before the first launch there.
inline void gn_set_attrs() {
  static bool done[64] = {false};
  if (!first_use_on_device(done)) return;
  MDC_CUDA(cudaFuncSetAttribute(gn_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GN_FUSED_SMEM_CAP) + 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GN_FUSED_SMEM_CAP) + 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_stats_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_apply_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_bwd_stats_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_bwd_apply_s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
  MDC_CUDA(cudaFuncSetAttribute(gn_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GN_FUSED_SMEM_CAP)));
  MDC_CUDA(cudaFuncSetAttribute(gn_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(GN_FUSED_SMEM_CAP)));
}

// mode: 0 = automatic, 1 = force the two-pass kernels, 2 = require a single-launch kernel (cluster kernels preferred),
// 3 = require the grid-barrier single-launch kernels, 4 = automatic without grid-barrier kernels (concurrent handles).
inline GNPlan plan_groupnorm(int N, int HW, int C, int G, long long ld, int mode = 0) {
  MDC_CHECK(C % G == 0 && (C / G) % 2 == 0 && C % 8 == 0, "GroupNorm: C=%d G=%d unsupported", C, G);
  GNPlan p;
  p.G = G;
  const int CV = C / 8;
  MDC_CHECK(CV <= 384, "GroupNorm: C=%d too wide", C);
  const int R = std::max(1, 512 / CV);
  p.threads = CV * R;
  p.threads_b = CV * std::max(1, 384 / CV);
  GNShape s;
  s.N = N, s.HW = HW, s.C = C, s.G = G, s.ld = ld;
  const int sms = g_num_sms();
  // blocks per SM of the two-pass kernels (two are resident; more than two makes a second, self-balancing wave).
  // MDC_GN_BPS overrides it for sweeps (read once).
  static const int bps = [] {
    const char* e = getenv("MDC_GN_BPS");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 16 ? v : 2;
  }();
  int want_blocks = std::max(1, (bps * sms) / N);
  int ppb = std::max(R, (s.HW + want_blocks - 1) / want_blocks);
  ppb = ((ppb + R - 1) / R) * R;
  s.pix_per_block = ppb;
  s.blocks_per_img = (s.HW + ppb - 1) / ppb;
  // ring depth of the streamed one-tensor kernels: as many 16 KB stages as two resident blocks can hold under the
  // 112 KB opt-in (MDC_GNS_STAGES overrides it for sweeps, read once)
  static const int st1 = [] {
    const char* e = getenv("MDC_GNS_STAGES");
    const int v = e ? atoi(e) : 0;
    return v >= 2 && v <= 12 ? v : 6;  // 4 -> 6: 19.28 / 19.54 -> 19.22 / 19.31 ms per step in an interleaved sweep on one box
  }();
  s.stages1 = st1;
  while (s.stages1 > 2 && gns_smem_bytes(p.threads, 1, s.stages1) + ((p.threads + 31) / 32) * 2 * G * sizeof(float) > 112 * 1024) --s.stages1;
  p.s = s;
  // single-launch variant: one CTA per SM at most, its pixel slab (x, or x and dy) staged in shared memory
  static const bool no_fuse = getenv("MDC_NO_GNFUSE") != nullptr;
  GNShape f = s;
  const int bpi = std::max(1, std::min(sms / N, s.HW));
  f.pix_per_block = (s.HW + bpi - 1) / bpi;
  f.blocks_per_img = (s.HW + f.pix_per_block - 1) / f.pix_per_block;
  const size_t slab = static_cast<size_t>(f.pix_per_block) * C * 2;
  const size_t extra = (static_cast<size_t>((p.threads + 31) / 32) * 2 * G + 2 * G) * sizeof(float);
  const bool ok = !(no_fuse && mode != 2 && mode != 3) && mode != 1 && mode != 4 && p.threads >= 8 * G && N * f.blocks_per_img <= sms;
  p.sfu = f;
  p.smem_f = slab + extra, p.smem_b = 2 * slab + extra;
  p.fuse_f = ok && p.smem_f <= GN_FUSED_SMEM_CAP;
  p.fuse_b = ok && p.smem_b <= GN_FUSED_SMEM_CAP;
  // Measured on a B200 (profiles/r02_groupnorm_variants.md): timed alone the cluster kernels win on tensors up to ~3.5 MB
  // (7-11 us against 11-12 us) and lose on the 4-28 MB ones (20-160 byte channel segments; at 8 CTAs per group no
  // longer one wave), but inside the captured step -- where programmatic dependent launch overlaps neighbouring small
  // kernels -- a hybrid selection measured 20.80 ms per step against 20.70 ms for the grid-barrier kernels alone.  So
  // the default (mode 0) keeps the grid-barrier kernels; the cluster kernels serve handles created with
  // `mdc_config.concurrent` (mode 4), which must not spin on a grid barrier, and mode 2 (tests).
  const bool want_cluster = mode == 2 || mode == 4;
  if (want_cluster && !(no_fuse && mode != 2)) {
    p.cl_f = plan_gn_cluster(N, HW, C, G, ld, false, p.cs_f, p.cl_smem_f);
    p.cl_b = plan_gn_cluster(N, HW, C, G, ld, true, p.cs_b, p.cl_smem_b);
    if (mode == 4 && (p.cs_f.K * N * G > 2 * sms || p.cs_b.K * N * G > 2 * sms)) p.cl_f = p.cl_b = false;  // too many waves: two-pass instead
  }
  if (p.cl_f) p.fuse_f = false;
  if (p.cl_b) p.fuse_b = false;
  if (mode == 4) p.fuse_f = p.fuse_b = false;
  MDC_CHECK(mode != 2 || (p.single_f() && p.single_b()), "GroupNorm: the single-launch variant does not fit (N=%d HW=%d C=%d)", N, HW, C);
  MDC_CHECK(mode != 3 || (p.fuse_f && p.fuse_b), "GroupNorm: the grid-barrier variant does not fit (N=%d HW=%d C=%d)", N, HW, C);
  // the single-launch variant may use MORE blocks than the two-pass one on small maps (one pixel row per block)
  p.partial_floats = static_cast<size_t>(2) * G * N * std::max(s.blocks_per_img, f.blocks_per_img);
  return p;
}

// launch_k with a thread-block cluster of `cluster` CTAs along x (plus the programmatic-dependent-launch attribute)
template <typename... KArgs, typename... Args>
inline void launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, int cluster, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = cluster, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
  ++na;
  cfg.attrs = attr, cfg.numAttrs = na;
  MDC_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  debug_sync(reinterpret_cast<const void*>(kernel), st);
}

struct GNScratch {
  float* partial = nullptr;        // >= plan.partial_floats
  float* gstats = nullptr;         // 2 * G * N floats (backward group sums)
  unsigned int* ticket = nullptr;  // N counters, zero-initialised
  unsigned int* bar = nullptr;     // 3 words (arrivals, generation, sticky timeout flag), zero-initialised
};

// epi_partial / epi_parts: the producing GEMM's epilogue already emitted per-CTA (sum, sum of squares) rows for x
// (GemmParams::gn_partial; two-pass path only): the statistics pass is skipped and gn_apply reduces those rows.
inline void run_gn_fwd(const GNPlan& p, const bf16* x, bf16* y, long long ldy, const float* gamma, const float* beta, float eps,
                       int silu, float* stats, const GNScratch& sc, cudaStream_t st, const float* epi_partial = nullptr,
                       int epi_parts = 0) {
  const bool have_stats = epi_partial != nullptr;
  if (p.cl_f) {
    launch_cluster(gn_cluster_fwd_kernel, dim3(p.cs_f.N * p.cs_f.G * p.cs_f.K), dim3(256), p.cl_smem_f, p.cs_f.K, st, x, p.cs_f, eps, stats, gamma, beta,
                   silu, y, ldy);
    return;
  }
  if (p.fuse_f) {
    launch_k(gn_fused_fwd_kernel, dim3(p.sfu.N * p.sfu.blocks_per_img), dim3(p.threads), p.smem_f, st, x, p.sfu, sc.partial, eps, stats, sc.bar,
             gamma, beta, silu, y, ldy);
    return;
  }
  const int grid = p.s.N * p.s.blocks_per_img;
  // streamed variants (bulk-copy fed shared-memory ring) need the block's pixel slab to be contiguous memory
  static const bool no_stream = getenv("MDC_NO_GNSTREAM") != nullptr;
  const size_t wsum = ((p.threads + 31) / 32) * 2 * p.G * sizeof(float);
  if (!no_stream && p.s.ld == p.s.C) {
    if (!have_stats)
      launch_k(gn_stats_s_kernel, dim3(grid), dim3(p.threads), gns_smem_bytes(p.threads, 1, gns_stages1(p.s)) + wsum, st, x, p.s, sc.partial, eps, stats, sc.ticket);
    launch_k(gn_apply_s_kernel, dim3(grid), dim3(p.threads), gns_smem_bytes(p.threads, 1, gns_stages1(p.s)), st, x, p.s, stats, gamma, beta, silu, y, ldy,
             epi_partial, epi_parts, eps);
    return;
  }
  if (!have_stats) launch_k(gn_stats_kernel, dim3(grid), dim3(p.threads), wsum, st, x, p.s, sc.partial, eps, stats, sc.ticket);
  launch_k(gn_apply_kernel, dim3(grid), dim3(p.threads), 0, st, x, p.s, stats, gamma, beta, silu, y, ldy, epi_partial, epi_parts, eps);
}
// Forward statistics only (two-pass plans): (mean, rstd) per (image, group) into `stats`, no normalised output.  Used by
// the sparse output head of the decoder (head.cuh), which evaluates the normalisation at a few thousand pixels itself.
inline void run_gn_stats(const GNPlan& p, const bf16* x, float eps, float* stats, const GNScratch& sc, cudaStream_t st) {
  MDC_CHECK(!p.single_f(), "run_gn_stats: two-pass plans only");
  const int grid = p.s.N * p.s.blocks_per_img;
  static const bool no_stream = getenv("MDC_NO_GNSTREAM") != nullptr;
  const size_t wsum = ((p.threads + 31) / 32) * 2 * p.G * sizeof(float);
  if (!no_stream && p.s.ld == p.s.C)
    launch_k(gn_stats_s_kernel, dim3(grid), dim3(p.threads), gns_smem_bytes(p.threads, 1, gns_stages1(p.s)) + wsum, st, x, p.s, sc.partial, eps, stats, sc.ticket);
  else
    launch_k(gn_stats_kernel, dim3(grid), dim3(p.threads), wsum, st, x, p.s, sc.partial, eps, stats, sc.ticket);
}
inline void run_gn_bwd(const GNPlan& p, const bf16* x, const bf16* dy, long long lddy, const float* gamma, const float* beta, int silu,
                       const float* stats, bf16* dx, long long lddx, int acc, const GNScratch& sc, cudaStream_t st) {
  if (p.cl_b) {
    launch_cluster(gn_cluster_bwd_kernel, dim3(p.cs_b.N * p.cs_b.G * p.cs_b.K), dim3(256), p.cl_smem_b, p.cs_b.K, st, x, dy, lddy, p.cs_b, stats, gamma,
                   beta, silu, dx, lddx, acc);
    return;
  }
  if (p.fuse_b) {
    launch_k(gn_fused_bwd_kernel, dim3(p.sfu.N * p.sfu.blocks_per_img), dim3(p.threads), p.smem_b, st, x, dy, lddy, p.sfu, stats, gamma, beta, silu,
             sc.partial, sc.bar, dx, lddx, acc);
    return;
  }
  const int grid = p.s.N * p.s.blocks_per_img;
  static const bool no_stream = getenv("MDC_NO_GNSTREAM") != nullptr;
  if (!no_stream && p.s.ld == p.s.C && lddy == p.s.C) {
    const size_t wsum = ((p.threads_b + 31) / 32) * 2 * p.G * sizeof(float);
    launch_k(gn_bwd_stats_s_kernel, dim3(grid), dim3(p.threads_b), gns_smem_bytes(p.threads_b, 2) + wsum, st, x, dy, p.s, stats, gamma, beta, silu,
             sc.partial, sc.gstats, sc.ticket);
    launch_k(gn_bwd_apply_s_kernel, dim3(grid), dim3(p.threads_b), gns_smem_bytes(p.threads_b, 2), st, x, dy, p.s, stats,
             static_cast<const float*>(sc.gstats), gamma, beta, silu, dx, lddx, acc);
    return;
  }
  launch_k(gn_bwd_stats_kernel, dim3(grid), dim3(p.threads_b), ((p.threads_b + 31) / 32) * 2 * p.G * sizeof(float), st, x, dy, lddy, p.s, stats,
           gamma, beta, silu, sc.partial, sc.gstats, sc.ticket);
  launch_k(gn_bwd_apply_kernel, dim3(grid), dim3(p.threads_b), 0, st, x, dy, lddy, p.s, stats, static_cast<const float*>(sc.gstats), gamma, beta,
           silu, dx, lddx, acc);
}

}  // namespace mdc
