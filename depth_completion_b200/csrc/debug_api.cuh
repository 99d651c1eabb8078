// Kernel-level entry points used by tests/ and profiling only (declared in include/mdc_debug.h).
// They run a single kernel of the hot path on caller-provided device buffers so each kernel can be
// checked against the oracle / a torch fp32 reference in isolation.
#pragma once
#include "attn.cuh"
#include "gemm.cuh"
#include "pack.cuh"

namespace mdc {

inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}

template <typename F>
inline int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const HostError& e) {
    last_error() = e.msg;
    return 1;
  } catch (const std::exception& e) {
    last_error() = e.what();
    return 2;
  }
}

inline float time_plan(const GemmPlan& g, int iters, cudaStream_t st) {
  cudaEvent_t e0, e1;
  MDC_CUDA(cudaEventCreate(&e0));
  MDC_CUDA(cudaEventCreate(&e1));
  run_gemm(g, st);  // warm-up / correctness launch
  MDC_CUDA(cudaGetLastError());
  MDC_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  if (iters > 0) {
    MDC_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) run_gemm(g, st);
    MDC_CUDA(cudaEventRecord(e1, st));
    MDC_CUDA(cudaEventSynchronize(e1));
    MDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms;
}

}  // namespace mdc

extern "C" {

const char* mdc_last_error() { return mdc::last_error().c_str(); }

int mdc_dbg_gemm(int M, int N, int K, const void* A, int a_mn, long long lda, long long sa0, long long sa1,
                 const void* B, int b_mn, long long ldb, long long sb0, long long sb1, void* out, int out_f32,
                 long long ldc, long long sc0, long long sc1, const float* bias, const void* res, long long ldr,
                 long long sr0, long long sr1, float alpha, int nb0, int nb1, int bn_override, int iters,
                 float* ms_out) {
  return mdc::guarded([&] {
    mdc::Operand a{A, a_mn, lda, sa0, sa1}, b{B, b_mn, ldb, sb0, sb1};
    mdc::Epilogue e;
    e.out = out, e.out_f32 = out_f32, e.ldc = ldc, e.sc0 = sc0, e.sc1 = sc1, e.bias = bias;
    e.res = static_cast<const __nv_bfloat16*>(res), e.ldr = ldr, e.sr0 = sr0, e.sr1 = sr1, e.alpha = alpha;
    // mdc_dbg_tune: split-K (< 0: the engine's cost model) and, for a plain K-major B, several weight copies used
    // round-robin in the timed loop (weights then stream from HBM as in the real step instead of sitting in L2)
    const int ncopy = (!b_mn && nb0 == 1 && nb1 == 1) ? std::max(1, mdc::g_tune().wcopies) : 1;
    const size_t wbytes = static_cast<size_t>(N) * ldb * 2;
    std::vector<const void*> copies(ncopy, B);
    std::vector<void*> owned;
    for (int i = 1; i < ncopy; ++i) {
      void* c = nullptr;
      MDC_CUDA(cudaMalloc(&c, wbytes));
      MDC_CUDA(cudaMemcpy(c, B, wbytes, cudaMemcpyDeviceToDevice));
      owned.push_back(c), copies[i] = c;
    }
    std::vector<mdc::GemmPlan> plans;
    float* ws = nullptr;
    for (int i = 0; i < ncopy; ++i) {
      mdc::Operand bi{copies[i], b_mn, ldb, sb0, sb1};
      mdc::GemmPlan g = mdc::plan_gemm(M, N, K, a, bi, e, nb0, nb1, bn_override);
      if (mdc::g_tune().ksplit) {
        size_t fl = mdc::enable_splitk(g, mdc::choose_ksplit(g));
        if (fl && !ws) MDC_CUDA(cudaMalloc(&ws, fl * 4 + 256));
        g.p.ws = ws;
      }
      plans.push_back(g);
    }
    if (ms_out && mdc::g_tune().ksplit < 0)
      fprintf(stderr, "[dbg] auto ksplit = %d (tiles %d x %d, k-chunks %d, BN %d, cs %d)\n", plans[0].p.ksplit, plans[0].p.m_tiles,
              plans[0].p.n_tiles, plans[0].p.num_k_chunks, plans[0].p.BN, plans[0].p.cs);
    float ms = 0.f;
    if (ncopy == 1) {
      ms = mdc::time_plan(plans[0], iters, 0);
    } else {
      cudaEvent_t e0, e1;
      MDC_CUDA(cudaEventCreate(&e0));
      MDC_CUDA(cudaEventCreate(&e1));
      for (int i = 0; i < ncopy; ++i) mdc::run_gemm(plans[i], 0);
      MDC_CUDA(cudaStreamSynchronize(0));
      if (iters > 0) {
        MDC_CUDA(cudaEventRecord(e0, 0));
        for (int i = 0; i < iters; ++i) mdc::run_gemm(plans[i % ncopy], 0);
        MDC_CUDA(cudaEventRecord(e1, 0));
        MDC_CUDA(cudaEventSynchronize(e1));
        MDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (ms_out) *ms_out = ms;
    MDC_CUDA(cudaDeviceSynchronize());
    for (void* c : owned) cudaFree(c);
    if (ws) cudaFree(ws);
  });
}

// x: NHWC bf16 (pixel stride ldx); w: OIHW fp32 [Cout][C][3][3]; mode 0 = forward conv, 1 = input gradient
// (x is then dy with Cout channels and the result has C channels).
int mdc_dbg_conv3x3(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const float* w_oihw,
                    int dgrad, const float* bias, const float* bias_img, const void* res, long long ldr, void* out,
                    long long ldc, int iters, float* ms_out) {
  return mdc::guarded([&] {
    const int Cin_g = dgrad ? Cout : C;   // channels of the tensor being convolved
    const int Cout_g = dgrad ? C : Cout;  // channels produced
    const int Kp = ((Cin_g + 63) / 64) * 64;
    __nv_bfloat16* wpk = nullptr;
    MDC_CUDA(cudaMalloc(&wpk, sizeof(__nv_bfloat16) * 9ull * Kp * Cout_g));
    if (!dgrad)
      mdc::pack_conv3x3_fwd_kernel<float><<<592, 256>>>(w_oihw, wpk, Cout, C, Kp);
    else
      mdc::pack_conv3x3_dgrad_kernel<float><<<592, 256>>>(w_oihw, wpk, Cout, C, Kp);
    MDC_CUDA(cudaGetLastError());
    mdc::Epilogue e;
    e.out = out, e.ldc = ldc, e.bias = bias, e.bias_img = bias_img;
    e.res = static_cast<const __nv_bfloat16*>(res), e.ldr = ldr;
    // optional: several weight copies used round-robin (so the timed loop streams weights from HBM like the real step
    // does) and split-K with its workspace, both selected through mdc_dbg_tune.
    const int ncopy = std::max(1, mdc::g_tune().wcopies);
    const size_t wbytes = sizeof(__nv_bfloat16) * 9ull * Kp * Cout_g;
    std::vector<__nv_bfloat16*> copies(ncopy, wpk);
    for (int i = 1; i < ncopy; ++i) {
      MDC_CUDA(cudaMalloc(&copies[i], wbytes));
      MDC_CUDA(cudaMemcpy(copies[i], wpk, wbytes, cudaMemcpyDeviceToDevice));
    }
    std::vector<mdc::GemmPlan> plans;
    float* ws = nullptr;
    for (int i = 0; i < ncopy; ++i) {
      mdc::GemmPlan g = mdc::plan_conv3x3(NB, H, W, Cin_g, Cout_g, x, ldx, copies[i], e);
      if (mdc::g_tune().ksplit) {
        size_t fl = mdc::enable_splitk(g, mdc::choose_ksplit(g));
        if (fl && !ws) MDC_CUDA(cudaMalloc(&ws, fl * 4 + 256));
        g.p.ws = ws;
      }
      plans.push_back(g);
    }
    if (ms_out && mdc::g_tune().ksplit < 0) fprintf(stderr, "[dbg] auto ksplit = %d (tiles %d x %d, k-chunks %d, BN %d, cs %d)\n",
                                                     plans[0].p.ksplit, plans[0].p.m_tiles, plans[0].p.n_tiles, plans[0].p.num_k_chunks, plans[0].p.BN, plans[0].p.cs);
    float ms = 0.f;
    if (ncopy == 1) {
      ms = mdc::time_plan(plans[0], iters, 0);
    } else {
      cudaEvent_t e0, e1;
      MDC_CUDA(cudaEventCreate(&e0));
      MDC_CUDA(cudaEventCreate(&e1));
      for (int i = 0; i < ncopy; ++i) mdc::run_gemm(plans[i], 0);
      MDC_CUDA(cudaStreamSynchronize(0));
      if (iters > 0) {
        MDC_CUDA(cudaEventRecord(e0, 0));
        for (int i = 0; i < iters; ++i) mdc::run_gemm(plans[i % ncopy], 0);
        MDC_CUDA(cudaEventRecord(e1, 0));
        MDC_CUDA(cudaEventSynchronize(e1));
        MDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
      }
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    if (ms_out) *ms_out = ms;
    MDC_CUDA(cudaDeviceSynchronize());
    for (int i = 1; i < ncopy; ++i) cudaFree(copies[i]);
    if (ws) cudaFree(ws);
    cudaFree(wpk);
  });
}

// Self-attention exactly as the engine plans it (attn.cuh): head_dim 64 -> the fused tcgen05 flash kernels, otherwise
// GEMM + softmax + GEMM.  qkv: [n, T, 3 * heads * dh] bf16 (q | k | v column blocks); o / dout: [n, T, heads * dh];
// dqkv like qkv.  dout == NULL: forward only.  ms_out[0] / [1] = forward / backward time per call over `iters`.
int mdc_dbg_attention(int n, int T, int heads, int dh, const void* qkv, void* o, const void* dout, void* dqkv, int iters,
                      float* ms_out) {
  return mdc::guarded([&] {
    using namespace mdc;
    MDC_CHECK(qkv && o && n >= 1 && T >= 1 && heads >= 1, "bad argument");
    MDC_CHECK((dout == nullptr) == (dqkv == nullptr), "dout and dqkv must be given together");
    set_kernel_attrs_for_device();
    const int d = heads * dh;
    const long long ldq = 3LL * d, ldo = d;
    float *lse2 = nullptr, *delta = nullptr, *S = nullptr;
    bf16 *P = nullptr, *dummy_do = nullptr, *dummy_dq = nullptr;
    MDC_CUDA(cudaMalloc(&lse2, flash_stat_floats(n, heads, T) * 4));
    MDC_CUDA(cudaMalloc(&delta, flash_stat_floats(n, heads, T) * 4));
    const bool bwd = dout != nullptr;
    if (!bwd) {  // the planners encode tensor maps for the gradient operands too
      MDC_CUDA(cudaMalloc(&dummy_do, 1ull * n * T * ldo * 2 + 256));
      MDC_CUDA(cudaMalloc(&dummy_dq, 1ull * n * T * ldq * 2 + 256));
    }
    const bool flash = dh == 64 && !getenv("MDC_NO_FLASH");
    if (!flash) {
      const size_t pel = static_cast<size_t>(n) * heads * T * (((T + 7) / 8) * 8) + 64;
      MDC_CUDA(cudaMalloc(&P, pel * 2));
      MDC_CUDA(cudaMalloc(&S, pel * 4));
    }
    AttnPlan a = plan_attention(n, T, heads, dh, static_cast<const bf16*>(qkv), ldq, static_cast<bf16*>(o),
                                bwd ? const_cast<bf16*>(static_cast<const bf16*>(dout)) : dummy_do, ldo,
                                bwd ? static_cast<bf16*>(dqkv) : dummy_dq, ldq, lse2, delta, P, S, true);
    cudaEvent_t e0, e1;
    MDC_CUDA(cudaEventCreate(&e0));
    MDC_CUDA(cudaEventCreate(&e1));
    float ms[2] = {0.f, 0.f};
    run_attention_fwd(a, 0);
    MDC_CUDA(cudaStreamSynchronize(0));
    if (iters > 0) {
      MDC_CUDA(cudaEventRecord(e0, 0));
      for (int i = 0; i < iters; ++i) run_attention_fwd(a, 0);
      MDC_CUDA(cudaEventRecord(e1, 0));
      MDC_CUDA(cudaEventSynchronize(e1));
      MDC_CUDA(cudaEventElapsedTime(&ms[0], e0, e1));
      ms[0] /= iters;
    }
    if (bwd) {
      run_attention_bwd(a, 0);  // (the unfused path overwrites P with dS: run the forward again before every backward)
      MDC_CUDA(cudaStreamSynchronize(0));
      if (iters > 0) {
        float tot = 0.f;
        for (int i = 0; i < iters; ++i) {
          run_attention_fwd(a, 0);
          MDC_CUDA(cudaEventRecord(e0, 0));
          run_attention_bwd(a, 0);
          MDC_CUDA(cudaEventRecord(e1, 0));
          MDC_CUDA(cudaEventSynchronize(e1));
          float t = 0.f;
          MDC_CUDA(cudaEventElapsedTime(&t, e0, e1));
          tot += t;
        }
        ms[1] = tot / iters;
      }
    }
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaDeviceSynchronize());
    if (ms_out) ms_out[0] = ms[0], ms_out[1] = ms[1];
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaFree(lse2), cudaFree(delta), cudaFree(P), cudaFree(S), cudaFree(dummy_do), cudaFree(dummy_dq);
  });
}

// GroupNorm (+SiLU) forward and, with dy != NULL, backward on an NHWC bf16 tensor x [n, HW, C] (pixel stride C), with the
// engine's kernel selection (norm.cuh).  mode: 0 automatic, 1 force the two-pass kernels, 2 require the single launch.
// stats_out (optional): [n, groups, 2] (mean, rstd).  ms_out[0] / [1]: forward / backward time per call.
int mdc_dbg_groupnorm(int n, int HW, int C, int groups, float eps, int silu, const void* x, const float* gamma,
                      const float* beta, void* y, const void* dy, void* dx, int acc, int mode, float* stats_out, int iters,
                      float* ms_out) {
  return mdc::guarded([&] {
    using namespace mdc;
    MDC_CHECK(x && gamma && beta && y, "null argument");
    set_kernel_attrs_for_device();
    GNPlan p = plan_groupnorm(n, HW, C, groups, C, mode);
    GNScratch sc;
    float* stats = nullptr;
    MDC_CUDA(cudaMalloc(&sc.partial, (p.partial_floats + 64) * 4));
    MDC_CUDA(cudaMalloc(&sc.gstats, (2ull * groups * n + 64) * 4));
    MDC_CUDA(cudaMalloc(&sc.ticket, (n + 16) * 4));
    MDC_CUDA(cudaMalloc(&sc.bar, 64));
    MDC_CUDA(cudaMalloc(&stats, (2ull * groups * n + 64) * 4));
    MDC_CUDA(cudaMemset(sc.ticket, 0, (n + 16) * 4));
    MDC_CUDA(cudaMemset(sc.bar, 0, 64));
    cudaEvent_t e0, e1;
    MDC_CUDA(cudaEventCreate(&e0));
    MDC_CUDA(cudaEventCreate(&e1));
    float ms[2] = {0.f, 0.f};
    auto fwd = [&] { run_gn_fwd(p, static_cast<const bf16*>(x), static_cast<bf16*>(y), C, gamma, beta, eps, silu, stats, sc, 0); };
    auto bwd = [&](int a) {
      run_gn_bwd(p, static_cast<const bf16*>(x), static_cast<const bf16*>(dy), C, gamma, beta, silu, stats, static_cast<bf16*>(dx), C, a, sc, 0);
    };
    fwd();
    MDC_CUDA(cudaStreamSynchronize(0));
    if (iters > 0) {
      MDC_CUDA(cudaEventRecord(e0, 0));
      for (int i = 0; i < iters; ++i) fwd();
      MDC_CUDA(cudaEventRecord(e1, 0));
      MDC_CUDA(cudaEventSynchronize(e1));
      MDC_CUDA(cudaEventElapsedTime(&ms[0], e0, e1));
      ms[0] /= iters;
    }
    if (dy && dx) {
      bwd(acc);
      MDC_CUDA(cudaStreamSynchronize(0));
      if (iters > 0 && !acc) {
        MDC_CUDA(cudaEventRecord(e0, 0));
        for (int i = 0; i < iters; ++i) bwd(0);
        MDC_CUDA(cudaEventRecord(e1, 0));
        MDC_CUDA(cudaEventSynchronize(e1));
        MDC_CUDA(cudaEventElapsedTime(&ms[1], e0, e1));
        ms[1] /= iters;
      }
    }
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaDeviceSynchronize());
    unsigned int flag = 0;
    MDC_CUDA(cudaMemcpy(&flag, sc.bar + 2, 4, cudaMemcpyDeviceToHost));
    if (stats_out) MDC_CUDA(cudaMemcpy(stats_out, stats, 2ull * groups * n * 4, cudaMemcpyDeviceToDevice));
    if (ms_out) ms_out[0] = ms[0], ms_out[1] = ms[1];
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaFree(sc.partial), cudaFree(sc.gstats), cudaFree(sc.ticket), cudaFree(sc.bar), cudaFree(stats);
    MDC_CHECK(flag == 0, "GroupNorm grid barrier timed out");
    MDC_CHECK(mode != 1 || (!p.single_f() && !p.single_b()), "mode 1 did not select the two-pass kernels");
  });
}

// Fused nearest-2x upsample + conv3x3 (four 2x2 phase convolutions on the low-resolution input, gemm.cuh) and its input
// gradient.  dgrad = 0: x [NB, H, W, C] -> out [NB, 2H, 2W, Cout] (+bias);  dgrad = 1: x is dy [NB, 2H, 2W, Cout] ->
// out [NB, H, W, C].  w_oihw: [Cout][C][3][3] fp32.
int mdc_dbg_upconv(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const float* w_oihw, int dgrad,
                   const float* bias, void* out, long long ldc, int iters, float* ms_out) {
  return mdc::guarded([&] {
    using namespace mdc;
    set_kernel_attrs_for_device();
    const int Kp = (((dgrad ? Cout : C) + 63) / 64) * 64, rows = dgrad ? C : Cout;
    __nv_bfloat16* wpk = nullptr;
    MDC_CUDA(cudaMalloc(&wpk, sizeof(__nv_bfloat16) * (16ull * Kp * rows + 64)));
    if (!dgrad)
      pack_upconv_fwd_kernel<float><<<592, 256>>>(w_oihw, wpk, Cout, C, Kp);
    else
      pack_upconv_bwd_kernel<float><<<592, 256>>>(w_oihw, wpk, Cout, C, Kp);
    MDC_CUDA(cudaGetLastError());
    Epilogue e;
    e.out = out, e.ldc = ldc, e.bias = bias;
    GemmPlan g = dgrad ? plan_upconv_bwd(NB, H, W, C, Cout, x, ldx, wpk, e) : plan_upconv_fwd(NB, H, W, C, Cout, x, ldx, wpk, e);
    float* ws = nullptr;
    if (g_tune().ksplit) {  // forced (> 0) or cost-model (< 0) split-K, as the engine applies it to the input gradient
      const size_t fl = enable_splitk(g, choose_ksplit(g));
      if (fl) MDC_CUDA(cudaMalloc(&ws, fl * 4 + 256));
      g.p.ws = ws;
      if (ms_out) fprintf(stderr, "[dbg] upconv ksplit = %d (tiles %d x %d, k-chunks %d)\n", g.p.ksplit, g.p.m_tiles, g.p.n_tiles, g.p.num_k_chunks);
    }
    float ms = time_plan(g, iters, 0);
    if (ws) {
      MDC_CUDA(cudaDeviceSynchronize());
      cudaFree(ws);
    }
    if (ms_out) *ms_out = ms;
    MDC_CUDA(cudaDeviceSynchronize());
    cudaFree(wpk);
  });
}

// Test / tuning overrides for the planners (0 = automatic): output-tile width, cluster size, split-K factor (-1 = use
// the engine's cost model also in the debug entry points), number of weight copies rotated in timed loops.
int mdc_dbg_tune(int bn, int cs, int ksplit, int wcopies) {
  mdc::g_tune().bn = bn, mdc::g_tune().cs = cs, mdc::g_tune().ksplit = ksplit, mdc::g_tune().wcopies = wcopies > 0 ? wcopies : 1;
  return 0;
}
// Row-shared-taps mode of the 3x3 convolution (GemmParams::rowshare): 0 automatic, 1 off, 2 on whenever legal.
int mdc_dbg_tune_rowshare(int mode) {
  mdc::g_tune().rowshare = mode;
  return 0;
}

}  // extern "C"
