// Self-attention over the tokens of each image, as the step engine plans it (shared with the kernel-level debug entry
// points so the tests run exactly the engine's selection):
//   * head_dim 64 (the UNet's attn1 blocks): the fused tcgen05 flash kernels of flash.cuh;
//   * other head dims (the VAE mid block: one head of 512): S = QK^T and O = PV through the tcgen05 GEMM with a
//     softmax kernel in between (probabilities saved for the backward), reference F.scaled_dot_product_attention.
// q, k, v are column blocks of one [n, T, 3 * heads * dh] tensor (fused projection), gradients likewise.
#pragma once
#include "flash.cuh"
#include "norm.cuh"

namespace mdc {

struct AttnPlan {
  int n = 0, T = 0, heads = 0, dh = 0;
  long long ldS = 0;
  bool use_flash = false;
  FlashPlan fp;
  GemmPlan p_s, p_o, p_dv, p_dp, p_dq, p_dk;
  bf16* P = nullptr;   // saved probabilities [n, heads, T, ldS] (unfused path)
  float* S = nullptr;  // fp32 score / dP scratch of the same shape (unfused path; may be patched later, see set_scratch)
  size_t scratch_floats() const { return use_flash ? 0 : static_cast<size_t>(n) * heads * T * ldS + 64; }
  size_t prob_elems() const { return scratch_floats(); }
};

// lse2 / delta: [n * heads * T + 64] floats each (flash path); P: prob_elems() bf16, S: scratch_floats() floats (unfused).
inline AttnPlan plan_attention(int n, int T, int heads, int dh, const bf16* qkv, long long ldq, bf16* o, bf16* dout, long long ldo,
                               bf16* dqkv, long long lddq, float* lse2, float* delta, bf16* P, float* S, bool with_backward = true) {
  MDC_CHECK(dh % 64 == 0 && dh <= 512, "attention head_dim %d unsupported (need a multiple of 64)", dh);
  AttnPlan a;
  a.n = n, a.T = T, a.heads = heads, a.dh = dh, a.ldS = ((T + 7) / 8) * 8;
  const int d = heads * dh;
  static const bool no_flash = getenv("MDC_NO_FLASH") != nullptr;
  if (dh == 64 && !no_flash) {
    a.use_flash = true;
    a.fp = plan_flash(n, T, heads, qkv, qkv + d, qkv + 2 * d, ldq, o, dout, ldo, dqkv, dqkv ? dqkv + d : nullptr, dqkv ? dqkv + 2 * d : nullptr,
                      lddq, lse2, delta);
    return a;
  }
  a.P = P, a.S = S;
  const long long ldS = a.ldS, tok = 1LL * T * ldq, sP0 = 1LL * T * ldS, sP1 = 1LL * heads * T * ldS;
  const bf16 *q = qkv, *k = qkv + d, *v = qkv + 2 * d;
  {  // S = scale * Q K^T -> fp32 scratch
    Operand A{q, 0, ldq, dh, tok}, B{k, 0, ldq, dh, tok};
    Epilogue e;
    e.out = S, e.out_f32 = 1, e.ldc = ldS, e.sc0 = sP0, e.sc1 = sP1, e.alpha = 1.f / sqrtf(static_cast<float>(dh));
    a.p_s = plan_gemm(T, T, dh, A, B, e, heads, n);
  }
  {  // O = P V : B = V^T is N-major
    Operand A{P, 0, ldS, sP0, sP1}, B{v, 1, ldq, dh, tok};
    Epilogue e;
    e.out = o, e.ldc = ldo, e.sc0 = dh, e.sc1 = 1LL * T * ldo;
    a.p_o = plan_gemm(T, dh, T, A, B, e, heads, n);
  }
  if (!with_backward || !dqkv) return a;
  const long long gtok = 1LL * T * lddq;
  bf16 *dq = dqkv, *dk = dqkv + d, *dv = dqkv + 2 * d;
  {  // dV[key, c] = sum_q P[q, key] dO[q, c]     A = P^T (M-major), B = dO^T (N-major)
    Operand A{P, 1, ldS, sP0, sP1}, B{dout, 1, ldo, dh, 1LL * T * ldo};
    Epilogue e;
    e.out = dv, e.ldc = lddq, e.sc0 = dh, e.sc1 = gtok;
    a.p_dv = plan_gemm(T, dh, T, A, B, e, heads, n);
  }
  {  // dP[q, key] = sum_c dO[q, c] V[key, c]      -> fp32 scratch
    Operand A{dout, 0, ldo, dh, 1LL * T * ldo}, B{v, 0, ldq, dh, tok};
    Epilogue e;
    e.out = S, e.out_f32 = 1, e.ldc = ldS, e.sc0 = sP0, e.sc1 = sP1;
    a.p_dp = plan_gemm(T, T, dh, A, B, e, heads, n);
  }
  {  // dQ[q, c] = sum_key dS[q, key] K[key, c]    A = dS (K-major), B = K^T (N-major)
    Operand A{P, 0, ldS, sP0, sP1}, B{k, 1, ldq, dh, tok};
    Epilogue e;
    e.out = dq, e.ldc = lddq, e.sc0 = dh, e.sc1 = gtok;
    a.p_dq = plan_gemm(T, dh, T, A, B, e, heads, n);
  }
  {  // dK[key, c] = sum_q dS[q, key] Q[q, c]      A = dS^T (M-major), B = Q^T (N-major)
    Operand A{P, 1, ldS, sP0, sP1}, B{q, 1, ldq, dh, tok};
    Epilogue e;
    e.out = dk, e.ldc = lddq, e.sc0 = dh, e.sc1 = gtok;
    a.p_dk = plan_gemm(T, dh, T, A, B, e, heads, n);
  }
  return a;
}
// The engine allocates one fp32 scratch for all unfused attentions after the tapes are built.
inline void attention_set_scratch(AttnPlan& a, float* S, bool with_backward) {
  if (a.use_flash) return;
  a.S = S;
  a.p_s.p.out = S;
  finish_plan(a.p_s);
  if (with_backward) {
    a.p_dp.p.out = S;
    finish_plan(a.p_dp);
  }
}
// Function attributes (opt-in shared memory sizes) are per device: set them for the current device once.
inline void set_kernel_attrs_for_device() {
  static bool done[64] = {false};
  if (!first_use_on_device(done)) return;
  gemm_set_smem_attr();
  flash_set_attrs();
  gn_set_attrs();
  MDC_CUDA(cudaFuncSetAttribute(softmax_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  MDC_CUDA(cudaFuncSetAttribute(softmax_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  // collapsed cross-attention: the backward stages three fp32 copies of its rows (<= 8 x 640 or 4 x 1280 channels)
#define MDC_XB_ATTR(RB, K4)                                                                                                          \
  MDC_CUDA(cudaFuncSetAttribute(xattn_block_bwd_kernel<RB, K4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); \
  MDC_CUDA(cudaFuncSetAttribute(xattn_block_fwd_kernel<RB, K4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  MDC_XB_ATTR(8, 3) MDC_XB_ATTR(8, 5) MDC_XB_ATTR(8, 10) MDC_XB_ATTR(4, 3) MDC_XB_ATTR(4, 5) MDC_XB_ATTR(4, 10)
  MDC_XB_ATTR(2, 3) MDC_XB_ATTR(2, 5) MDC_XB_ATTR(2, 10)
#undef MDC_XB_ATTR
}

inline void run_attention_fwd(const AttnPlan& a, cudaStream_t st) {
  if (a.use_flash) {
    run_flash_fwd(a.fp, st);
    return;
  }
  run_gemm(a.p_s, st);
  const int rows = a.n * a.heads * a.T;
  launch_k(softmax_fwd_kernel, dim3(rows), dim3(256), (((a.T + 3) & ~3) + 32) * sizeof(float), st, static_cast<const float*>(a.S), a.P, a.T, a.ldS);
  run_gemm(a.p_o, st);
}
inline void run_attention_bwd(const AttnPlan& a, cudaStream_t st, const SideBranch* sb = nullptr) {
  if (a.use_flash) {
    run_flash_bwd(a.fp, st, sb);
    return;
  }
  run_gemm(a.p_dv, st);
  run_gemm(a.p_dp, st);
  const int rows = a.n * a.heads * a.T;
  launch_k(softmax_bwd_kernel, dim3(rows), dim3(256), (((a.T + 3) & ~3) + 32) * sizeof(float), st, static_cast<const float*>(a.S), a.P, a.T, a.ldS,
           1.f / sqrtf(static_cast<float>(a.dh)));
  run_gemm(a.p_dq, st);
  run_gemm(a.p_dk, st);
}

}  // namespace mdc
