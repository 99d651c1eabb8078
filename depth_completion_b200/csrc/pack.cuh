// Weight re-packing kernels (run once, at mdc_set_weight time; not on the hot path).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mdc {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

// OIHW [Cout][C][3][3] -> forward pack [Cout][9*Cp], k = (r*3+s)*Cp + c  (zero padded to Cp).
template <typename T>
__global__ void pack_conv3x3_fwd_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int C,
                                        int Cp) {
  long long total = 1LL * Cout * 9 * Cp;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int c = i % Cp;
    int tap = (i / Cp) % 9;
    int co = i / (9LL * Cp);
    float v = 0.f;
    if (c < C) v = to_f32(w[((1LL * co * C + c) * 9) + tap]);
    out[i] = __float2bfloat16(v);
  }
}

// OIHW [Cout][C][3][3] -> input-gradient pack [C][9*Cop]: the dgrad of a stride-1 pad-1 3x3 conv is a 3x3
// conv of dy with the taps flipped and the channel roles swapped.
template <typename T>
__global__ void pack_conv3x3_dgrad_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int C,
                                          int Cop) {
  long long total = 1LL * C * 9 * Cop;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int co = i % Cop;
    int tap = (i / Cop) % 9;
    int ci = i / (9LL * Cop);
    float v = 0.f;
    if (co < Cout) v = to_f32(w[((1LL * co * C + ci) * 9) + (8 - tap)]);
    out[i] = __float2bfloat16(v);
  }
}

// Effective 2x2 weights of the fused nearest-2x-upsample + conv3x3: rows {r} of the 3x3 kernel that land on the same
// low-resolution input row are summed.  phase a = 0: tap 0 <- r{0}, tap 1 <- r{1,2};  a = 1: tap 0 <- r{0,1}, tap 1 <- r{2}.
template <typename T>
__device__ __forceinline__ float upconv_weff(const T* __restrict__ w, int co, int c, int C, int a, int b, int tr, int ts) {
  const int r0 = (a == 0) ? (tr == 0 ? 0 : 1) : (tr == 0 ? 0 : 2), r1 = (a == 0) ? (tr == 0 ? 0 : 2) : (tr == 0 ? 1 : 2);
  const int s0 = (b == 0) ? (ts == 0 ? 0 : 1) : (ts == 0 ? 0 : 2), s1 = (b == 0) ? (ts == 0 ? 0 : 2) : (ts == 0 ? 1 : 2);
  float v = 0.f;
  for (int r = r0; r <= r1; ++r)
    for (int s = s0; s <= s1; ++s) v += to_f32(w[((1LL * co * C + c) * 9) + r * 3 + s]);
  return v;
}
// forward pack [Cout][16*Cp], k = (phase*4 + tr*2 + ts)*Cp + c
template <typename T>
__global__ void pack_upconv_fwd_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int C, int Cp) {
  long long total = 1LL * Cout * 16 * Cp;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int c = i % Cp;
    int tap = (i / Cp) % 16;
    int co = i / (16LL * Cp);
    float v = 0.f;
    if (c < C) v = upconv_weff(w, co, c, C, tap >> 3, (tap >> 2) & 1, (tap >> 1) & 1, tap & 1);
    out[i] = __float2bfloat16(v);
  }
}
// input-gradient pack [C][16*Cop], k = (phase*4 + tr*2 + ts)*Cop + co
template <typename T>
__global__ void pack_upconv_bwd_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int C, int Cop) {
  long long total = 1LL * C * 16 * Cop;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int co = i % Cop;
    int tap = (i / Cop) % 16;
    int ci = i / (16LL * Cop);
    float v = 0.f;
    if (co < Cout) v = upconv_weff(w, co, ci, C, tap >> 3, (tap >> 2) & 1, (tap >> 1) & 1, tap & 1);
    out[i] = __float2bfloat16(v);
  }
}

// [rows][cols] source -> bf16: out[r*ld + c] (transpose = 0) or out[c*ld + r] (transpose = 1).
template <typename T>
__global__ void pack_matrix_kernel(const T* __restrict__ w, __nv_bfloat16* __restrict__ out, int rows, int cols,
                                   long long ld, int transpose) {
  long long total = 1LL * rows * cols;
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
    int c = i % cols;
    int r = i / cols;
    float v = to_f32(w[i]);
    if (transpose)
      out[1LL * c * ld + r] = __float2bfloat16(v);
    else
      out[1LL * r * ld + c] = __float2bfloat16(v);
  }
}

template <typename T>
__global__ void to_f32_kernel(const T* __restrict__ w, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x)
    out[i] = to_f32(w[i]);
}

}  // namespace mdc
