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


// ---------------------------------------------------------------------------------------------- batched re-packing
// One launch re-packs ANY number of parameters (mdc_set_weights): blockIdx.y selects the job, the blocks of a row
// grid-stride over its elements.  A model is ~1100 parameters and ~2300 destination layouts; one launch per layout
// would bury the hot kernels of a short run under thousands of pack launches in any launch-list profile.
enum PackKind { PK_CONV_FWD = 0, PK_CONV_DGRAD = 1, PK_UPCONV_FWD = 2, PK_UPCONV_BWD = 3, PK_MATRIX = 4, PK_MATRIX_T = 5, PK_VEC = 6 };
struct PackJob {
  const void* src;
  void* dst;
  int kind, src_bf16;
  int out, in, pad;   // conv kinds: Cout, C, padded K-channel count; matrix: rows, cols; vec: n
  long long ld;       // matrix kinds: destination row stride
  float vscale, vshift;
};
template <typename T>
__device__ __forceinline__ void pack_job_body(const PackJob& j) {
  const T* w = static_cast<const T*>(j.src);
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(j.dst);
  const long long stride = 1LL * gridDim.x * blockDim.x, first = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  const int Cout = j.out, C = j.in, P = j.pad;
  switch (j.kind) {
    case PK_CONV_FWD:
      for (long long i = first; i < 1LL * Cout * 9 * P; i += stride) {
        const int c = i % P, tap = (i / P) % 9, co = i / (9LL * P);
        out[i] = __float2bfloat16(c < C ? to_f32(w[((1LL * co * C + c) * 9) + tap]) : 0.f);
      }
      break;
    case PK_CONV_DGRAD:
      for (long long i = first; i < 1LL * C * 9 * P; i += stride) {
        const int co = i % P, tap = (i / P) % 9, ci = i / (9LL * P);
        out[i] = __float2bfloat16(co < Cout ? to_f32(w[((1LL * co * C + ci) * 9) + (8 - tap)]) : 0.f);
      }
      break;
    case PK_UPCONV_FWD:
      for (long long i = first; i < 1LL * Cout * 16 * P; i += stride) {
        const int c = i % P, tap = (i / P) % 16, co = i / (16LL * P);
        out[i] = __float2bfloat16(c < C ? upconv_weff(w, co, c, C, tap >> 3, (tap >> 2) & 1, (tap >> 1) & 1, tap & 1) : 0.f);
      }
      break;
    case PK_UPCONV_BWD:
      for (long long i = first; i < 1LL * C * 16 * P; i += stride) {
        const int co = i % P, tap = (i / P) % 16, ci = i / (16LL * P);
        out[i] = __float2bfloat16(co < Cout ? upconv_weff(w, co, ci, C, tap >> 3, (tap >> 2) & 1, (tap >> 1) & 1, tap & 1) : 0.f);
      }
      break;
    case PK_MATRIX:
    case PK_MATRIX_T:
      for (long long i = first; i < 1LL * j.out * j.in; i += stride) {
        const int c = i % j.in, r = i / j.in;
        out[j.kind == PK_MATRIX ? 1LL * r * j.ld + c : 1LL * c * j.ld + r] = __float2bfloat16(to_f32(w[i]));
      }
      break;
    default:  // PK_VEC: fp32 destination, optionally stored as vscale * v + vshift
      for (long long i = first; i < j.out; i += stride) static_cast<float*>(j.dst)[i] = j.vscale * to_f32(w[i]) + j.vshift;
      break;
  }
}
__global__ void pack_jobs_kernel(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  if (j.src_bf16)
    pack_job_body<__nv_bfloat16>(j);
  else
    pack_job_body<float>(j);
}

}  // namespace mdc
