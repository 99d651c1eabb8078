// The small, bandwidth-bound "tail" of one guided step (marigold_dc.py:813-904, :970-984):
// UNet input assembly, DDIM x0-prediction, the sparse-mask L1+L2 loss with its gradient, the
// grad-norm rescale, the fused Adam update of latent/scale/shift and the DDIM step.
//
// Everything step-dependent (sqrt(alpha_bar_t) ..., Adam bias corrections, the per-resnet
// time-embedding biases) is read from device memory selected by a device-side step counter, so one
// guided step is a fixed launch sequence (CUDA-graph friendly, no host sync inside the loop).
//
// bf16 mode of the reference computes this algebra with bf16 tensors and 0-dim fp32 scalars; every
// torch op rounds its result to bf16.  The kernels reproduce those rounding points (bf16r).
#pragma once
#include "kernels.cuh"

namespace mdc {

struct StepTables {            // device arrays of length `steps`
  const float* sqrt_a;         // sqrt(alpha_bar_t)
  const float* sqrt_1ma;       // sqrt(1 - alpha_bar_t)
  const float* sqrt_ap;        // sqrt(alpha_bar_prev)
  const float* sqrt_1map;      // sqrt(1 - alpha_bar_prev)
  const float* temb_bias;      // [steps][temb_total]
  int temb_total;
  int steps;
};
struct StepCur {               // scalars of the step being executed (device memory)
  float sqrt_a, sqrt_1ma, sqrt_ap, sqrt_1map;
  float adam_step_size_x, adam_step_size_s, adam_bc2_sqrt;
  int step;
};
constexpr int MAXN = 16;       // max frames per handle
struct StepAccum {             // per-sample accumulators (device memory)
  float loss[MAXN], s_grad[MAXN], t_grad[MAXN];
  float scale[MAXN], shift[MAXN];
  float s_m[MAXN], s_v[MAXN], t_m[MAXN], t_v[MAXN];
};

// Per-call options of the non-default branches of marigold_dc.py:467-493 (device memory, written by mdc_set_options /
// mdc_begin so that the captured step graph stays valid when they change between calls).
struct TailOpts {
  int projection;    // 0 linear, 1 log, 2 log10 (get_projection_fn, marigold_dc.py:23-50)
  int inv;           // 1: inverse depth (:743-749, :848-862)
  int opt;           // 0 adam, 1 sgd, 2 adagrad (:776-789; torch defaults otherwise)
  int kld_mode;      // 0 off, 1 simple, 2 strict (utils.py:28-86)
  float kld_weight;
  float w_l1, w_l2;  // how many times "l1" / "l2" appear in loss_funcs (:177-193 loops over the list)
  float w_edge, w_smooth;  // likewise for "edge" / "smooth" (:195-236)
  float lr_x, lr_s;  // learning rates of the latent and of scale / shift (:649, :776-783)
  int closed_form;   // 1: scale / shift refitted by least squares every step instead of learned (:332-336)
};
__device__ __forceinline__ float project_depth(float d, int projection) {
  return projection == 1 ? logf(d) : projection == 2 ? log10f(d) : d;
}

// Select the step: fill StepCur and the current time-embedding biases.  One block.
__global__ void begin_step_kernel(StepTables tb, int* __restrict__ counter, StepCur* __restrict__ cur,
                                  float* __restrict__ temb_cur, const TailOpts* __restrict__ opts) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int s = min(*counter, tb.steps - 1);
  if (threadIdx.x == 0) {
    const float lr_x = opts->lr_x, lr_s = opts->lr_s;
    cur->sqrt_a = tb.sqrt_a[s], cur->sqrt_1ma = tb.sqrt_1ma[s];
    cur->sqrt_ap = tb.sqrt_ap[s], cur->sqrt_1map = tb.sqrt_1map[s];
    const double t = s + 1;
    const double bc1 = 1.0 - pow(0.9, t), bc2 = 1.0 - pow(0.999, t);
    cur->adam_step_size_x = static_cast<float>(lr_x / bc1);
    cur->adam_step_size_s = static_cast<float>(lr_s / bc1);
    cur->adam_bc2_sqrt = static_cast<float>(sqrt(bc2));
    cur->step = s;
  }
  for (int i = threadIdx.x; i < tb.temb_total; i += blockDim.x) temb_cur[i] = tb.temb_bias[1LL * s * tb.temb_total + i];
}

// cat([img_latents, x], dim=1) as NHWC [N, h, w, 8]  (marigold_dc.py:459)
__global__ void unet_input_kernel(const bf16* __restrict__ img_lat, const bf16* __restrict__ x, int N, int hw,
                                  bf16* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * hw) return;
  long long n = i / hw, p = i % hw;
  BF8 o;
  bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ob[c] = img_lat[(n * 4 + c) * hw + p];
    ob[4 + c] = x[(n * 4 + c) * hw + p];
  }
  *reinterpret_cast<BF8*>(out + i * 8) = o;
}

// x0 = sqrt(a) x - sqrt(1-a) v ; eps = sqrt(a) v + sqrt(1-a) x  (marigold_dc.py:813-826); z = x0 / scaling -> NHWC
// eps_part[n*bpi + b] = partial sum of eps^2 (fixed-order reduction later).
__global__ void x0_kernel(const bf16* __restrict__ v_nhwc, const bf16* __restrict__ x, const StepCur* __restrict__ cur,
                          int N, int hw, float scaling, bf16* __restrict__ z_nhwc, float* __restrict__ eps_part,
                          float* __restrict__ x1_part, float* __restrict__ x2_part) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float red[32];
  const int bpi = gridDim.x / N, n = blockIdx.x / bpi, b = blockIdx.x % bpi;
  const float sa = cur->sqrt_a, sb = cur->sqrt_1ma;
  float acc = 0.f, xs1 = 0.f, xs2 = 0.f;
  for (int p = b * blockDim.x + threadIdx.x; p < hw; p += bpi * blockDim.x) {
    BF8 vv = *reinterpret_cast<const BF8*>(v_nhwc + (1LL * n * hw + p) * 8);
    const bf16* vb = reinterpret_cast<const bf16*>(&vv);
    BF8 o;
    bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float xv = __bfloat162float(x[(1LL * n * 4 + c) * hw + p]), vf = __bfloat162float(vb[c]);
      float x0 = bf16r(bf16r(sa * xv) - bf16r(sb * vf));
      float e = bf16r(bf16r(sa * vf) + bf16r(sb * xv));
      acc += e * e;
      xs1 += xv, xs2 += xv * xv;
      ob[c] = __float2bfloat16(x0 / scaling);
      ob[4 + c] = __float2bfloat16(0.f);
    }
    *reinterpret_cast<BF8*>(z_nhwc + (1LL * n * hw + p) * 8) = o;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) eps_part[blockIdx.x] = acc;
  if (x1_part) {  // sums of x and x^2 for the kld penalty (utils.py:69-77)
    xs1 = block_sum(xs1, red);
    xs2 = block_sum(xs2, red);
    if (threadIdx.x == 0) x1_part[blockIdx.x] = xs1, x2_part[blockIdx.x] = xs2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Geometry of the decoder output -> input-resolution map (marigold_dc.py:366-370):
// decoder output [N, PPH, PPW, 3]; unpad to [ph, pw]; bilinear (align_corners=False, no antialias) to [H, W].
struct TailGeom {
  int N, H, W, ph, pw, PPH, PPW;
  long long ld_dec;  // pixel stride of the decoder output (8)
  int nearest;       // interp_mode: 0 "bilinear", 1 "nearest" (marigold_dc.py:343, :366-370; predict.py:200-206)
};
// Source taps of one output index for F.interpolate(size=...): bilinear (align_corners=False) or, with `nearest`, the
// single tap min(floor(dst * in / out), in - 1) with weight 1.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1, int nearest) {
  if (nearest) {
    i0 = i1 = min(static_cast<int>(floorf(dst * scale)), in_size - 1);
    l1 = 0.f;
    return;
  }
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = min(static_cast<int>(src), in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = src - i0;
}
// affine-invariant prediction at one decoder pixel: (clip(mean_c dec, -1, 1) + 1) / 2, with the bf16 rounding points
__device__ __forceinline__ float affine_at(const bf16* __restrict__ dec, long long ld, long long pix, bool& inside) {
  const bf16* p = dec + pix * ld;
  float m = bf16r((__bfloat162float(p[0]) + __bfloat162float(p[1]) + __bfloat162float(p[2])) / 3.f);
  inside = (m >= -1.f && m <= 1.f);
  m = fminf(fmaxf(m, -1.f), 1.f);
  return bf16r((m + 1.f) * 0.5f);
}

// Masked L1+L2 loss at the valid points of one sample and its gradient (marigold_dc.py:829-840, :181-193, :877).
// pts: compacted valid pixels of all samples, sample n owns [pt_off[n], pt_off[n+1]).  One block per sample.
// dmean[N, PPH, PPW] (fp32, zero on entry) receives d loss / d mean_c(dec) by atomics (4 taps per point).
__global__ void loss_points_kernel(const bf16* __restrict__ dec, TailGeom g, const int* __restrict__ pt_idx,
                                   const float* __restrict__ pt_val, const int* __restrict__ pt_off,
                                   const float* __restrict__ gminmax, const float* __restrict__ depth_minmax,
                                   const TailOpts* __restrict__ opts, StepAccum* __restrict__ acc,
                                   float* __restrict__ dmean) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float red[32];
  if (opts->closed_form) return;  // loss_points_cf_kernel does this step's loss
  const int n = blockIdx.x;
  const int p0 = pt_off[n], p1 = pt_off[n + 1];
  const float cnt = static_cast<float>(p1 - p0);
  const float s = acc->scale[n], t = acc->shift[n];
  const float gmin = gminmax[2 * n], gmax = gminmax[2 * n + 1], range = gmax - gmin;
  // depth-space conversion of the prediction (marigold_dc.py:842-862): metric range and its projected ends
  const int projection = opts->projection, inv = opts->inv;
  const bool convert = projection != 0 || inv != 0;
  const float w_l1 = opts->w_l1, w_l2 = opts->w_l2;
  const float dlo = depth_minmax[2 * n], dhi = depth_minmax[2 * n + 1];
  float plo = project_depth(dlo, projection), phi = project_depth(dhi, projection);
  if (inv) {
    const float a = 1.f / phi, b = 1.f / plo;
    plo = a, phi = b;
  }
  const float sy = static_cast<float>(g.ph) / g.H, sx = static_cast<float>(g.pw) / g.W;
  float l_sum = 0.f, ds = 0.f, dt = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const int pix = pt_idx[i];
    const int Y = pix / g.W, X = pix % g.W;
    const float guide = pt_val[i];
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(Y, sy, g.ph, y0, y1, ly, g.nearest);
    bilinear_src(X, sx, g.pw, x0, x1, lx, g.nearest);
    const long long base = 1LL * n * g.PPH * g.PPW;
    const long long q00 = base + 1LL * y0 * g.PPW + x0, q01 = base + 1LL * y0 * g.PPW + x1;
    const long long q10 = base + 1LL * y1 * g.PPW + x0, q11 = base + 1LL * y1 * g.PPW + x1;
    bool in00, in01, in10, in11;
    const float a00 = affine_at(dec, g.ld_dec, q00, in00), a01 = affine_at(dec, g.ld_dec, q01, in01);
    const float a10 = affine_at(dec, g.ld_dec, q10, in10), a11 = affine_at(dec, g.ld_dec, q11, in11);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const float aff = bf16r(w00 * a00 + w01 * a01 + w10 * a10 + w11 * a11);
    const float pre = s * s * range * aff + t * t * gmin;
    float dense = fminf(fmaxf(pre, 0.f), 1.f);
    float chain = 1.f;  // d(guide-space value) / d(normalised linear depth)
    if (convert) {
      const float metric = dense * (dhi - dlo) + dlo;
      float pr = project_depth(metric, projection);
      chain = (dhi - dlo) * (projection == 1 ? 1.f / metric : projection == 2 ? 0.4342944819f / metric : 1.f);
      if (inv) {
        pr = 1.f / pr;
        chain *= -pr * pr;
      }
      dense = (pr - plo) / (phi - plo);
      chain /= (phi - plo);
    }
    const float diff = dense - guide;
    l_sum += (w_l1 * fabsf(diff) + w_l2 * diff * diff) / cnt;
    float dd = (w_l1 * ((diff > 0.f) - (diff < 0.f)) + w_l2 * 2.f * diff) / cnt * chain;
    if (pre < 0.f || pre > 1.f) dd = 0.f;  // clamp backward
    ds += dd * 2.f * s * range * aff;
    dt += dd * 2.f * t * gmin;
    const float da = dd * s * s * range * 0.5f;  // through (m + 1) / 2
    if (in00) atomicAdd(&dmean[q00], da * w00);
    if (in01) atomicAdd(&dmean[q01], da * w01);
    if (in10) atomicAdd(&dmean[q10], da * w10);
    if (in11) atomicAdd(&dmean[q11], da * w11);
  }
  l_sum = block_sum(l_sum, red);
  ds = block_sum(ds, red);
  dt = block_sum(dt, red);
  if (threadIdx.x == 0) acc->loss[n] = l_sum, acc->s_grad[n] = ds, acc->t_grad[n] = dt;
}
// ---- "edge" / "smooth" terms (marigold_dc.py:195-236): dense losses on the whole H x W map.  Both kernels return at
// once unless loss_funcs lists one of them (opts->w_edge / w_smooth), so the default path only pays two empty launches.
// Prediction in the guide's space at one output pixel and d(that) / d(pre-clamp value); shared by both kernels.
struct DensePix {
  long long q[4];
  float w[4];
  bool in[4];
  float aff, value, chain;
};
__device__ __forceinline__ DensePix dense_pixel(const bf16* __restrict__ dec, const TailGeom& g, int n, int Y, int X, float s,
                                                float t, float gmin, float range, float dlo, float dhi, float plo,
                                                float phi, int projection, int inv) {
  DensePix r;
  const float sy = static_cast<float>(g.ph) / g.H, sx = static_cast<float>(g.pw) / g.W;
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(Y, sy, g.ph, y0, y1, ly, g.nearest);
  bilinear_src(X, sx, g.pw, x0, x1, lx, g.nearest);
  const long long base = 1LL * n * g.PPH * g.PPW;
  r.q[0] = base + 1LL * y0 * g.PPW + x0, r.q[1] = base + 1LL * y0 * g.PPW + x1;
  r.q[2] = base + 1LL * y1 * g.PPW + x0, r.q[3] = base + 1LL * y1 * g.PPW + x1;
  r.w[0] = (1.f - ly) * (1.f - lx), r.w[1] = (1.f - ly) * lx, r.w[2] = ly * (1.f - lx), r.w[3] = ly * lx;
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) a += r.w[k] * affine_at(dec, g.ld_dec, r.q[k], r.in[k]);
  r.aff = bf16r(a);
  const float pre = s * s * range * r.aff + t * t * gmin;
  float v = fminf(fmaxf(pre, 0.f), 1.f), chain = (pre < 0.f || pre > 1.f) ? 0.f : 1.f;
  if (projection != 0 || inv != 0) {
    const float metric = v * (dhi - dlo) + dlo;
    float pr = project_depth(metric, projection);
    chain *= (dhi - dlo) * (projection == 1 ? 1.f / metric : projection == 2 ? 0.4342944819f / metric : 1.f);
    if (inv) {
      pr = 1.f / pr;
      chain *= -pr * pr;
    }
    v = (pr - plo) / (phi - plo);
    chain /= (phi - plo);
  }
  r.value = v, r.chain = chain;
  return r;
}
__device__ __forceinline__ void projected_ends(const float* __restrict__ depth_minmax, int n, int projection, int inv,
                                               float& dlo, float& dhi, float& plo, float& phi) {
  dlo = depth_minmax[2 * n], dhi = depth_minmax[2 * n + 1];
  plo = project_depth(dlo, projection), phi = project_depth(dhi, projection);
  if (inv) {
    const float a = 1.f / phi, b = 1.f / plo;
    plo = a, phi = b;
  }
}
// closed_form=True while the latent is trained (marigold_dc.py:332-336 inside the loop): every step refits
// scale = cov(a, g) / (var(a) + 1e-7), shift = mean(g) - scale mean(a) over the valid points (:53-128) and the loss
// gradient flows THROUGH the fit:  dL/da_k = s G_k + ds/da_k (sum G a - mean(a) sum G) - s sum G / n  with
// G_i = dL/d dense_i and ds/da_k = ((g_k - mean g)(var + eps) - 2 cov (a_k - mean a)) / (var + eps)^2.
// One block per sample; pt_a / pt_G are per-point scratch.  Returns at once unless opts->closed_form.
__global__ void loss_points_cf_kernel(const bf16* __restrict__ dec, TailGeom g, const int* __restrict__ pt_idx,
                                      const float* __restrict__ pt_val, const int* __restrict__ pt_off,
                                      const float* __restrict__ depth_minmax, const TailOpts* __restrict__ opts,
                                      StepAccum* __restrict__ acc, float* __restrict__ dmean, float* __restrict__ pt_a,
                                      float* __restrict__ pt_G) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  if (!opts->closed_form) return;
  __shared__ float red[32];
  __shared__ float sh[6];
  const int n = blockIdx.x, p0 = pt_off[n], p1 = pt_off[n + 1];
  const float cnt = static_cast<float>(p1 - p0);
  const int projection = opts->projection, inv = opts->inv;
  const float w_l1 = opts->w_l1, w_l2 = opts->w_l2;
  float dlo, dhi, plo, phi;
  projected_ends(depth_minmax, n, projection, inv, dlo, dhi, plo, phi);
  float sa = 0.f, sg = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const int pix = pt_idx[i];
    const float a = dense_pixel(dec, g, n, pix / g.W, pix % g.W, 1.f, 0.f, 0.f, 1.f, 0.f, 1.f, 0.f, 1.f, 0, 0).aff;
    pt_a[i] = a;
    sa += a, sg += pt_val[i];
  }
  sa = block_sum(sa, red);
  sg = block_sum(sg, red);
  if (threadIdx.x == 0) sh[0] = sa / cnt, sh[1] = sg / cnt;
  __syncthreads();
  const float am = sh[0], gm = sh[1];
  float var = 0.f, cov = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const float ac = pt_a[i] - am;
    var += ac * ac, cov += ac * (pt_val[i] - gm);
  }
  var = block_sum(var, red);
  cov = block_sum(cov, red);
  if (threadIdx.x == 0) sh[2] = var + 1e-7f, sh[3] = cov;
  __syncthreads();
  const float vpe = sh[2], cv = sh[3], sc = cv / vpe, sf = gm - sc * am;
  float l_sum = 0.f, SG = 0.f, SGa = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const float a = pt_a[i], pre = sc * a + sf;
    float dense = fminf(fmaxf(pre, 0.f), 1.f), chain = (pre < 0.f || pre > 1.f) ? 0.f : 1.f;
    if (projection != 0 || inv != 0) {
      const float metric = dense * (dhi - dlo) + dlo;
      float pr = project_depth(metric, projection);
      chain *= (dhi - dlo) * (projection == 1 ? 1.f / metric : projection == 2 ? 0.4342944819f / metric : 1.f);
      if (inv) {
        pr = 1.f / pr;
        chain *= -pr * pr;
      }
      dense = (pr - plo) / (phi - plo);
      chain /= (phi - plo);
    }
    const float diff = dense - pt_val[i];
    l_sum += (w_l1 * fabsf(diff) + w_l2 * diff * diff) / cnt;
    const float G = (w_l1 * ((diff > 0.f) - (diff < 0.f)) + w_l2 * 2.f * diff) / cnt * chain;
    pt_G[i] = G;
    SG += G, SGa += G * a;
  }
  l_sum = block_sum(l_sum, red);
  SG = block_sum(SG, red);
  SGa = block_sum(SGa, red);
  if (threadIdx.x == 0) sh[4] = SG, sh[5] = SGa;
  __syncthreads();
  SG = sh[4], SGa = sh[5];
  const float k1 = SGa - am * SG, k2 = sc * SG / cnt;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const float a = pt_a[i];
    const float ds = ((pt_val[i] - gm) * vpe - 2.f * cv * (a - am)) / (vpe * vpe);
    const float da = (sc * pt_G[i] + ds * k1 - k2) * 0.5f;  // 0.5: through (m + 1) / 2
    const int pix = pt_idx[i];
    const DensePix px = dense_pixel(dec, g, n, pix / g.W, pix % g.W, 1.f, 0.f, 0.f, 1.f, 0.f, 1.f, 0.f, 1.f, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px.in[k] && px.w[k] != 0.f) atomicAdd(&dmean[px.q[k]], da * px.w[k]);
  }
  if (threadIdx.x == 0) {
    acc->loss[n] = l_sum, acc->s_grad[n] = 0.f, acc->t_grad[n] = 0.f;  // no learned affine parameters (:764-775)
    acc->scale[n] = sc, acc->shift[n] = sf;
  }
}

// dn[N, H, W] fp32: the dense prediction in the guide's space (:829-862).
__global__ void dense_map_kernel(const bf16* __restrict__ dec, TailGeom g, const float* __restrict__ gminmax,
                                 const float* __restrict__ depth_minmax, const TailOpts* __restrict__ opts,
                                 const StepAccum* __restrict__ acc, float* __restrict__ dn) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  if (opts->w_edge == 0.f && opts->w_smooth == 0.f) return;
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * g.N * g.H * g.W) return;
  const int X = i % g.W, Y = (i / g.W) % g.H, n = i / (1LL * g.W * g.H);
  float dlo, dhi, plo, phi;
  projected_ends(depth_minmax, n, opts->projection, opts->inv, dlo, dhi, plo, phi);
  const float gmin = gminmax[2 * n], gmax = gminmax[2 * n + 1];
  dn[i] = dense_pixel(dec, g, n, Y, X, acc->scale[n], acc->shift[n], gmin, gmax - gmin, dlo, dhi, plo, phi, opts->projection,
                      opts->inv).value;
}
// Loss and gradient of w_smooth * (mean|d_y dn| + mean|d_x dn|) + w_edge * (mean||d_x dn| - gx| + mean||d_y dn| - gy|)
// with gx / gy the absolute forward differences of the grey image (gray_grad_kernel).  Pixel (Y, X) owns the pairs to
// its right and below for the loss, and gathers the gradient of its four pairs.  grid = (blocks per image, N).
__global__ void dense_loss_kernel(const bf16* __restrict__ dec, TailGeom g, const float* __restrict__ gminmax,
                                  const float* __restrict__ depth_minmax, const TailOpts* __restrict__ opts,
                                  const float* __restrict__ dn, const float* __restrict__ gx, const float* __restrict__ gy,
                                  StepAccum* __restrict__ acc, float* __restrict__ dmean) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const float we = opts->w_edge, ws = opts->w_smooth;
  if (we == 0.f && ws == 0.f) return;
  __shared__ float red[32];
  const int n = blockIdx.y, HW = g.H * g.W;
  const float s = acc->scale[n], t = acc->shift[n];
  const float gmin = gminmax[2 * n], gmax = gminmax[2 * n + 1], range = gmax - gmin;
  float dlo, dhi, plo, phi;
  projected_ends(depth_minmax, n, opts->projection, opts->inv, dlo, dhi, plo, phi);
  const float cx = 1.f / (1.f * g.H * (g.W - 1)), cy = 1.f / (1.f * (g.H - 1) * g.W);
  const float* d = dn + 1LL * n * HW;
  const float* ex = gx + 1LL * n * HW;
  const float* ey = gy + 1LL * n * HW;
  auto sgn = [](float v) { return static_cast<float>((v > 0.f) - (v < 0.f)); };
  // d loss / d (a - b) of one pair with image gradient e and mean factor c
  auto pair_grad = [&](float diff, float e, float c) { return c * sgn(diff) * (ws + we * sgn(fabsf(diff) - e)); };
  float l_sum = 0.f, ds = 0.f, dt = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int Y = p / g.W, X = p % g.W;
    const float v = d[p];
    float grad = 0.f;
    if (X + 1 < g.W) {
      const float diff = v - d[p + 1];
      l_sum += cx * (ws * fabsf(diff) + we * fabsf(fabsf(diff) - ex[p]));
      grad += pair_grad(diff, ex[p], cx);
    }
    if (X > 0) grad -= pair_grad(d[p - 1] - v, ex[p - 1], cx);
    if (Y + 1 < g.H) {
      const float diff = v - d[p + g.W];
      l_sum += cy * (ws * fabsf(diff) + we * fabsf(fabsf(diff) - ey[p]));
      grad += pair_grad(diff, ey[p], cy);
    }
    if (Y > 0) grad -= pair_grad(d[p - g.W] - v, ey[p - g.W], cy);
    if (grad != 0.f) {
      const DensePix px = dense_pixel(dec, g, n, Y, X, s, t, gmin, range, dlo, dhi, plo, phi, opts->projection, opts->inv);
      const float dd = grad * px.chain;
      ds += dd * 2.f * s * range * px.aff;
      dt += dd * 2.f * t * gmin;
      const float da = dd * s * s * range * 0.5f;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (px.in[k] && px.w[k] != 0.f) atomicAdd(&dmean[px.q[k]], da * px.w[k]);
    }
  }
  l_sum = block_sum(l_sum, red);
  ds = block_sum(ds, red);
  dt = block_sum(dt, red);
  if (threadIdx.x == 0) atomicAdd(&acc->loss[n], l_sum), atomicAdd(&acc->s_grad[n], ds), atomicAdd(&acc->t_grad[n], dt);
}
// Absolute forward differences of the grey image the edge loss compares with (marigold_dc.py:199-216), computed on the
// RAW images exactly as the reference does: 0.299 R + 0.587 G + 0.114 B in fp32 (0..255 for uint8 input), or the single
// channel itself -- for a uint8 single channel the subtraction wraps modulo 256 like torch's uint8 arithmetic.
__global__ void gray_grad_kernel(const void* __restrict__ img, int is_u8, int C, int N, int H, int W, float* __restrict__ gx,
                                 float* __restrict__ gy) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * H * W) return;
  const int X = i % W, Y = (i / W) % H, n = i / (1LL * W * H);
  const long long HW = 1LL * H * W;
  auto px = [&](int c, int y, int x) -> float {
    const long long o = (1LL * n * C + c) * HW + 1LL * y * W + x;
    return is_u8 ? static_cast<float>(static_cast<const uint8_t*>(img)[o]) : static_cast<const float*>(img)[o];
  };
  auto gray = [&](int y, int x) -> float {
    if (C == 1) return px(0, y, x);
    return __fadd_rn(__fadd_rn(__fmul_rn(0.299f, px(0, y, x)), __fmul_rn(0.587f, px(1, y, x))), __fmul_rn(0.114f, px(2, y, x)));
  };
  auto adiff = [&](float a, float b) -> float {
    if (C == 1 && is_u8) return static_cast<float>(static_cast<uint8_t>(static_cast<int>(a) - static_cast<int>(b)));
    return fabsf(a - b);
  };
  const float v = gray(Y, X);
  gx[i] = X + 1 < W ? adiff(v, gray(Y, X + 1)) : 0.f;
  gy[i] = Y + 1 < H ? adiff(v, gray(Y + 1, X)) : 0.f;
}

// d dec[n, y, x, c] = dmean / 3 for c < 3 (bf16); clears dmean for the next step.
__global__ void dec_grad_kernel(float* __restrict__ dmean, long long npix, bf16* __restrict__ ddec) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  float d = dmean[i] * (1.f / 3.f);
  dmean[i] = 0.f;
  BF8 o;
  bf16* ob = reinterpret_cast<bf16*>(&o);
  bf16 dv = __float2bfloat16(d), z = __float2bfloat16(0.f);
  ob[0] = dv, ob[1] = dv, ob[2] = dv, ob[3] = z, ob[4] = z, ob[5] = z, ob[6] = z, ob[7] = z;
  *reinterpret_cast<BF8*>(ddec + i * 8) = o;
}

// dz (decoder-input gradient, NHWC) -> d x0 = dz / scaling; dv = -sqrt(1-a) d x0 (UNet output gradient, NHWC);
// direct part of dx = sqrt(a) d x0 (NCHW fp32 holding bf16-rounded values).
__global__ void dx0_kernel(const bf16* __restrict__ dz_nhwc, const StepCur* __restrict__ cur, int N, int hw,
                           float scaling, bf16* __restrict__ dv_nhwc, float* __restrict__ dx_direct) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * hw) return;
  long long n = i / hw, p = i % hw;
  const float sa = cur->sqrt_a, sb = cur->sqrt_1ma;
  BF8 in = *reinterpret_cast<const BF8*>(dz_nhwc + i * 8);
  const bf16* ib = reinterpret_cast<const bf16*>(&in);
  BF8 o;
  bf16* ob = reinterpret_cast<bf16*>(&o);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float d0 = bf16r(__bfloat162float(ib[c]) / scaling);
    ob[c] = __float2bfloat16(-sb * d0);
    ob[4 + c] = __float2bfloat16(0.f);
    dx_direct[(n * 4 + c) * hw + p] = bf16r(sa * d0);
  }
  *reinterpret_cast<BF8*>(dv_nhwc + i * 8) = o;
}

// total latent gradient g = direct + UNet path (channels 4..7 of the UNet-input gradient); partial sums of g^2.
// With the kld penalty (marigold_dc.py:238-241, utils.py:69-77) the latent also receives kld_weight * d dist / d x:
// simple: dist = mean(x^2); strict: dist = (mu^2 + var - log(var + eps) - 1) / 2 with eps = finfo(bf16).eps.
__global__ void grad_total_kernel(const float* __restrict__ dx_direct, const bf16* __restrict__ din_nhwc, int N, int hw,
                                  float* __restrict__ gbuf, float* __restrict__ g_part, const bf16* __restrict__ x,
                                  const float* __restrict__ x1_part, const float* __restrict__ x2_part,
                                  const TailOpts* __restrict__ opts, StepAccum* __restrict__ accum) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float red[32];
  const int bpi = gridDim.x / N, n = blockIdx.x / bpi, b = blockIdx.x % bpi;
  const int kld = opts->kld_mode;
  float kc = 0.f, kb = 0.f;  // kld gradient = kc * x + kb
  if (kld) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < bpi; ++k) s1 += x1_part[n * bpi + k], s2 += x2_part[n * bpi + k];
    const float M = 4.f * hw, kw = bf16r(opts->kld_weight);
    float dist;
    if (kld == 1) {
      dist = bf16r(s2 / M);
      kc = 2.f * bf16r(kw / M);
    } else {
      const float mu = bf16r(s1 / M), var = bf16r(fmaxf(s2 / M - (s1 / M) * (s1 / M), 0.f)), eps = 0.0078125f;
      dist = 0.5f * (mu * mu + var - logf(var + eps) - 1.f);
      const float r = 1.f - 1.f / (var + eps);
      kc = kw / M * r;
      kb = kw / M * mu * (1.f - r);
    }
    if (b == 0 && threadIdx.x == 0) accum->loss[n] += opts->kld_weight * dist;
  }
  float acc = 0.f;
  for (int p = b * blockDim.x + threadIdx.x; p < hw; p += bpi * blockDim.x) {
    BF8 in = *reinterpret_cast<const BF8*>(din_nhwc + (1LL * n * hw + p) * 8);
    const bf16* ib = reinterpret_cast<const bf16*>(&in);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      long long j = (1LL * n * 4 + c) * hw + p;
      float gg = bf16r(dx_direct[j] + __bfloat162float(ib[4 + c]));
      if (kld) gg = bf16r(gg + bf16r(kc * __bfloat162float(x[j]) + kb));
      gbuf[j] = gg;
      acc += gg * gg;
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) g_part[blockIdx.x] = acc;
}

// grad-norm rescale (marigold_dc.py:881-894) + torch.optim.Adam (or SGD / Adagrad, :776-789) foreach update in the
// parameter dtype (bf16 latent, fp32 scale/shift; :897) + DDIM prev_sample with the stale v and the updated x (:901-904).  Advances the step counter.
__global__ void adam_ddim_kernel(const float* __restrict__ gbuf, const float* __restrict__ eps_part,
                                 const float* __restrict__ g_part, int parts_per_img, const bf16* __restrict__ v_nhwc,
                                 const StepCur* __restrict__ cur, int N, int hw, bf16* __restrict__ x,
                                 bf16* __restrict__ m1, bf16* __restrict__ m2, StepAccum* __restrict__ acc,
                                 int* __restrict__ counter, bf16* __restrict__ x_adam_dbg,
                                 const TailOpts* __restrict__ opts) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const int bpi = gridDim.x / N, n = blockIdx.x / bpi, b = blockIdx.x % bpi;
  const int opt = opts->opt;
  const float lr_x = opts->lr_x, lr_s = opts->lr_s;
  float e2 = 0.f, g2 = 0.f;
  for (int k = 0; k < parts_per_img; ++k) e2 += eps_part[n * parts_per_img + k], g2 += g_part[n * parts_per_img + k];
  const float en = bf16r(sqrtf(e2)), gn = bf16r(sqrtf(g2));
  const float factor = bf16r(en / fmaxf(gn, 1e-7f));
  const float sa = cur->sqrt_a, sb = cur->sqrt_1ma, sap = cur->sqrt_ap, sbp = cur->sqrt_1map;
  const float step_size = cur->adam_step_size_x, bc2s = cur->adam_bc2_sqrt;
  for (int p = b * blockDim.x + threadIdx.x; p < hw; p += bpi * blockDim.x) {
    BF8 vv = *reinterpret_cast<const BF8*>(v_nhwc + (1LL * n * hw + p) * 8);
    const bf16* vb = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long j = (1LL * n * 4 + c) * hw + p;
      const float g = bf16r(gbuf[j] * factor);
      float m = __bfloat162float(m1[j]), v2 = __bfloat162float(m2[j]), xv = __bfloat162float(x[j]);
      if (opt == 0) {
        m = bf16r(m + 0.1f * (g - m));                      // exp_avg.lerp_(grad, 1 - beta1)
        v2 = bf16r(v2 * 0.999f);                            // exp_avg_sq.mul_(beta2)
        v2 = bf16r(v2 + 0.001f * g * g);                    //           .addcmul_(grad, grad, 1 - beta2)
        float den = bf16r(sqrtf(v2));
        den = bf16r(den / bc2s);
        den = bf16r(den + 1e-8f);
        xv = bf16r(xv - step_size * (m / den));             // param.addcdiv_(exp_avg, denom, -step_size)
      } else if (opt == 1) {                                // torch.optim.SGD, no momentum: param.add_(grad, alpha=-lr)
        xv = bf16r(xv - lr_x * g);
      } else {                                              // torch.optim.Adagrad (foreach): lr_decay 0, eps 1e-10
        v2 = bf16r(v2 + g * g);                             // state_sum.addcmul_(grad, grad, value=1)
        const float sd = bf16r(bf16r(sqrtf(v2)) + 1e-10f);  // std = state_sum.sqrt().add_(eps)
        xv = bf16r(xv + bf16r(-lr_x * g) / sd);             // param.addcdiv_(grad * -clr, std)
      }
      m1[j] = __float2bfloat16(m), m2[j] = __float2bfloat16(v2);
      if (x_adam_dbg) x_adam_dbg[j] = __float2bfloat16(xv);
      const float vf = __bfloat162float(vb[c]);
      const float x0 = bf16r(bf16r(sa * xv) - bf16r(sb * vf));
      const float e = bf16r(bf16r(sa * vf) + bf16r(sb * xv));
      x[j] = __float2bfloat16(bf16r(sap * x0) + bf16r(sbp * e));
    }
  }
  if (b == 0 && threadIdx.x == 0) {  // fp32 update of scale / shift of sample n
    const float ss = cur->adam_step_size_s;
    float gs = acc->s_grad[n], gt = acc->t_grad[n];
    if (opt == 0) {
      float sm = acc->s_m[n] + 0.1f * (gs - acc->s_m[n]), sv = acc->s_v[n] * 0.999f + 0.001f * gs * gs;
      float tm = acc->t_m[n] + 0.1f * (gt - acc->t_m[n]), tv = acc->t_v[n] * 0.999f + 0.001f * gt * gt;
      acc->scale[n] -= ss * (sm / (sqrtf(sv) / bc2s + 1e-8f));
      acc->shift[n] -= ss * (tm / (sqrtf(tv) / bc2s + 1e-8f));
      acc->s_m[n] = sm, acc->s_v[n] = sv, acc->t_m[n] = tm, acc->t_v[n] = tv;
    } else if (opt == 1) {
      acc->scale[n] -= lr_s * gs;
      acc->shift[n] -= lr_s * gt;
    } else {
      const float sv = acc->s_v[n] + gs * gs, tv = acc->t_v[n] + gt * gt;
      acc->scale[n] += (-lr_s * gs) / (sqrtf(sv) + 1e-10f);
      acc->shift[n] += (-lr_s * gt) / (sqrtf(tv) + 1e-10f);
      acc->s_v[n] = sv, acc->t_v[n] = tv;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = *counter + 1;
}

// Plain DDIM step of the no-grad branch (train_latents=False, marigold_dc.py:905-909): prev_sample from v and x, no
// guidance.  Same bf16 rounding points as the guided kernel's tail.  Advances the step counter.
__global__ void ddim_only_kernel(const bf16* __restrict__ v_nhwc, const StepCur* __restrict__ cur, int N, int hw,
                                 bf16* __restrict__ x, int* __restrict__ counter) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  const float sa = cur->sqrt_a, sb = cur->sqrt_1ma, sap = cur->sqrt_ap, sbp = cur->sqrt_1map;
  const long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i < 1LL * N * hw) {
    const long long n = i / hw, p = i % hw;
    BF8 vv = *reinterpret_cast<const BF8*>(v_nhwc + i * 8);
    const bf16* vb = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long j = (n * 4 + c) * hw + p;
      const float xv = __bfloat162float(x[j]), vf = __bfloat162float(vb[c]);
      const float x0 = bf16r(bf16r(sa * xv) - bf16r(sb * vf));
      const float e = bf16r(bf16r(sa * vf) + bf16r(sb * xv));
      x[j] = __float2bfloat16(bf16r(sap * x0) + bf16r(sbp * e));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *counter = *counter + 1;
}

// Closed-form masked least-squares scale / shift (compute_affine_params, marigold_dc.py:53-128) over the valid points of
// one sample: scale = cov(a, g) / (var(a) + 1e-7), shift = mean(g) - scale * mean(a); the affine-side sums round to
// bf16 like the reference's bf16 tensors.  Writes them to acc->scale / acc->shift.  One block per sample.
__global__ void affine_lsq_kernel(const bf16* __restrict__ dec, TailGeom g, const int* __restrict__ pt_idx,
                                  const float* __restrict__ pt_val, const int* __restrict__ pt_off,
                                  StepAccum* __restrict__ acc) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float red[32];
  __shared__ float s_am, s_gm;
  const int n = blockIdx.x, p0 = pt_off[n], p1 = pt_off[n + 1];
  const float cnt = static_cast<float>(p1 - p0);
  auto aff_at = [&](int i) {
    const int pix = pt_idx[i];
    return dense_pixel(dec, g, n, pix / g.W, pix % g.W, 1.f, 0.f, 0.f, 1.f, 0.f, 1.f, 0.f, 1.f, 0, 0).aff;
  };
  float sa = 0.f, sg = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) sa += aff_at(i), sg += pt_val[i];
  sa = block_sum(sa, red);
  sg = block_sum(sg, red);
  if (threadIdx.x == 0) s_am = bf16r(bf16r(sa) / cnt), s_gm = sg / cnt;
  __syncthreads();
  const float am = s_am, gm = s_gm;
  float var = 0.f, cov = 0.f;
  for (int i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
    const float ac = bf16r(aff_at(i) - am);
    var += bf16r(ac * ac);
    cov += ac * (pt_val[i] - gm);
  }
  var = block_sum(var, red);
  cov = block_sum(cov, red);
  if (threadIdx.x == 0) {
    const float sc = cov / (bf16r(var) + 1e-7f);
    acc->scale[n] = sc;
    acc->shift[n] = gm - sc * am;
  }
}

// Final dense map (marigold_dc.py:970-984): clamp(s^2 range aff + t^2 gmin, 0, 1) * (max - min) + min -> fp32 [N,1,H,W]
__global__ void dense_out_kernel(const bf16* __restrict__ dec, TailGeom g, const float* __restrict__ gminmax,
                                 const float* __restrict__ depth_minmax, const StepAccum* __restrict__ acc,
                                 float* __restrict__ out, int closed_form) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * g.N * g.H * g.W) return;
  const int X = i % g.W, Y = (i / g.W) % g.H, n = i / (1LL * g.W * g.H);
  const float sy = static_cast<float>(g.ph) / g.H, sx = static_cast<float>(g.pw) / g.W;
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src(Y, sy, g.ph, y0, y1, ly, g.nearest);
  bilinear_src(X, sx, g.pw, x0, x1, lx, g.nearest);
  const long long base = 1LL * n * g.PPH * g.PPW;
  bool in;
  const float a00 = affine_at(dec, g.ld_dec, base + 1LL * y0 * g.PPW + x0, in);
  const float a01 = affine_at(dec, g.ld_dec, base + 1LL * y0 * g.PPW + x1, in);
  const float a10 = affine_at(dec, g.ld_dec, base + 1LL * y1 * g.PPW + x0, in);
  const float a11 = affine_at(dec, g.ld_dec, base + 1LL * y1 * g.PPW + x1, in);
  const float aff = bf16r((1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11));
  const float s = acc->scale[n], t = acc->shift[n];
  const float gmin = gminmax[2 * n], gmax = gminmax[2 * n + 1];
  float d = closed_form ? s * aff + t                           // :332-336, closed-form scale / shift
                        : s * s * (gmax - gmin) * aff + t * t * gmin;  // :320-331, learned scale / shift
  d = fminf(fmaxf(d, 0.f), 1.f);
  const float dmin = depth_minmax[2 * n], dmax = depth_minmax[2 * n + 1];
  out[i] = d * (dmax - dmin) + dmin;
}

// ---------------------------------------------------------------------------------------------- per-frame prologue
// MarigoldImageProcessor.preprocess (called at marigold_dc.py:687-694): integer images / 255,
// * 2 - 1 (both rounded to bf16 like the reference's bf16 pipeline), antialiased bilinear resize to ph x pw
// (F.interpolate(mode="bilinear", antialias=True): triangle filter whose support grows with the down-scaling factor,
// weights normalised over the taps inside the image), replicate padding to PPH x PPW.  Output NHWC bf16, 3 channels.
struct PreGeom {
  int N, C, H, W, ph, pw, PPH, PPW;
  int is_u8;
  long long ld_out;
};
__device__ __forceinline__ void aa_span(int i, int in_size, float scale, int& lo, int& size, float& center, float& inv) {
  const float support = scale >= 1.f ? scale : 1.f;
  center = scale * (i + 0.5f);
  inv = scale >= 1.f ? 1.f / scale : 1.f;
  lo = max(static_cast<int>(center - support + 0.5f), 0);
  size = min(static_cast<int>(center + support + 0.5f), in_size) - lo;
}
__device__ __forceinline__ float aa_tri(float x) {
  x = fabsf(x);
  return x < 1.f ? 1.f - x : 0.f;
}
__global__ void preprocess_image_kernel(const void* __restrict__ img, PreGeom g, bf16* __restrict__ out) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * g.N * g.PPH * g.PPW) return;
  const int ox = i % g.PPW, oy = (i / g.PPW) % g.PPH, n = i / (1LL * g.PPW * g.PPH);
  const int ry = min(oy, g.ph - 1), rx = min(ox, g.pw - 1);  // replicate padding (bottom / right)
  const float sh = static_cast<float>(g.H) / g.ph, sw = static_cast<float>(g.W) / g.pw;
  int y0, ny, x0, nx;
  float cy, cx, iy, ix;
  aa_span(ry, g.H, sh, y0, ny, cy, iy);
  aa_span(rx, g.W, sw, x0, nx, cx, ix);
  float wy_tot = 0.f, wx_tot = 0.f;
  for (int j = 0; j < ny; ++j) wy_tot += aa_tri((j + y0 - cy + 0.5f) * iy);
  for (int k = 0; k < nx; ++k) wx_tot += aa_tri((k + x0 - cx + 0.5f) * ix);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < ny; ++j) {
    float wy = aa_tri((j + y0 - cy + 0.5f) * iy);
    if (wy_tot != 0.f) wy /= wy_tot;
    float row[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < nx; ++k) {
      float wx = aa_tri((k + x0 - cx + 0.5f) * ix);
      if (wx_tot != 0.f) wx /= wx_tot;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const long long idx = ((1LL * n * g.C + (g.C == 1 ? 0 : c)) * g.H + (y0 + j)) * g.W + (x0 + k);
        float v = g.is_u8 ? bf16r(static_cast<float>(static_cast<const uint8_t*>(img)[idx]) / 255.f)
                          : bf16r(static_cast<const float*>(img)[idx]);
        v = bf16r(v * 2.f - 1.f);
        row[c] += wx * v;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] += wy * row[c];
  }
  bf16* o = out + i * g.ld_out;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = __float2bfloat16(acc[c]);
}
// Sparse-depth normalisation (marigold_dc.py:707-756, linear projection): mask = sparse > 0; "minmax": (lo, hi) =
// masked min / max of the sample, values clamped to it, then lo = max(lo, min_depth), hi = min(hi, max_depth);
// "const": (lo, hi) = (min_depth, max_depth); guide = (clamp(sparse) - lo) / (hi - lo).  Also the masked min / max of
// the guide that _affine_to_metric recomputes every step (:326).  One block per sample, fp32 like the reference.
// out_stats[n] = {lo, hi, gmin, gmax, n_valid}.
// norm_mode 2 ("percentile", :715-728): (lo, hi) come from range_in[2n], range_in[2n+1] (quantile_range_kernel) and
// are then treated like the min / max of "minmax".  projection / inv (:737-756): lo, hi and the clamped values go
// through log / log10 and 1 / x (swapping the ends) before the normalisation; out_stats keeps the metric lo / hi.
__global__ void sparse_norm_kernel(const float* __restrict__ sparse, int HW, float min_depth, float max_depth,
                                   int norm_mode, const float* __restrict__ range_in, int projection, int inv,
                                   float* __restrict__ guide, uint8_t* __restrict__ mask,
                                   float* __restrict__ out_stats) {
  const int norm_const = norm_mode == 1;
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float s_lo[32], s_hi[32];
  __shared__ int s_cnt[32];
  const int n = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* sp = sparse + 1LL * n * HW;
  auto block_minmax = [&](float lo, float hi, int cnt, float& olo, float& ohi, int& ocnt) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __syncthreads();
    if (lane == 0) s_lo[w] = lo, s_hi[w] = hi, s_cnt[w] = cnt;
    __syncthreads();
    olo = INFINITY, ohi = -INFINITY, ocnt = 0;
    for (int i = 0; i < nw; ++i) olo = fminf(olo, s_lo[i]), ohi = fmaxf(ohi, s_hi[i]), ocnt += s_cnt[i];
  };
  float lo = INFINITY, hi = -INFINITY;
  int cnt = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = sp[i];
    if (v > 0.f) lo = fminf(lo, v), hi = fmaxf(hi, v), ++cnt;
  }
  float mlo, mhi;
  int total;
  block_minmax(lo, hi, cnt, mlo, mhi, total);
  float clo = norm_const ? min_depth : mlo, chi = norm_const ? max_depth : mhi;  // clamp range of the raw values
  if (norm_mode == 2) clo = range_in[2 * n], chi = range_in[2 * n + 1];
  float nlo = clo, nhi = chi;                                                      // normalisation range
  if (!norm_const) nlo = fmaxf(clo, min_depth), nhi = fminf(chi, max_depth);
  float plo = project_depth(nlo, projection), phi = project_depth(nhi, projection);
  if (inv) {
    const float a = 1.f / phi, b = 1.f / plo;
    plo = a, phi = b;
  }
  float glo = INFINITY, ghi = -INFINITY;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = sp[i];
    float pv = project_depth(fminf(fmaxf(v, clo), chi), projection);
    if (inv) pv = 1.f / pv;
    const float g = (pv - plo) / (phi - plo);
    guide[1LL * n * HW + i] = g;
    const bool m = v > 0.f;
    mask[1LL * n * HW + i] = m ? 1 : 0;
    if (m) glo = fminf(glo, g), ghi = fmaxf(ghi, g);
  }
  float ogl, ogh;
  int dummy;
  block_minmax(glo, ghi, 0, ogl, ogh, dummy);
  if (threadIdx.x == 0) {
    float* o = out_stats + 5 * n;
    o[0] = nlo, o[1] = nhi, o[2] = ogl, o[3] = ogh, o[4] = static_cast<float>(total);
  }
}

// norm="percentile" (marigold_dc.py:715-728): the positive values of each sample, compacted (any order) into
// vals[n * HW ...]; counts[n] = how many.  One block per sample.
__global__ void compact_positive_kernel(const float* __restrict__ sparse, int HW, float* __restrict__ vals,
                                        int* __restrict__ counts) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ int s_count;
  const int n = blockIdx.x;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const float v = sparse[1LL * n * HW + i];
    if (v > 0.f) vals[1LL * n * HW + atomicAdd(&s_count, 1)] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) counts[n] = s_count;
}
// torch.quantile(sorted, q) with linear interpolation (aten Sorting.cpp quantile_compute): rank = q * (n - 1) in
// fp32, lerp between the neighbours with torch's lerp formula (w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)).
__global__ void quantile_range_kernel(const float* __restrict__ sorted, int count, float q_lo, float q_hi,
                                      float* __restrict__ range_out) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  if (threadIdx.x >= 2 || blockIdx.x != 0) return;
  const float q = threadIdx.x == 0 ? q_lo : q_hi;
  const float rank = q * static_cast<float>(count - 1);
  const float below = floorf(rank), w = rank - below;
  const int i0 = static_cast<int>(below), i1 = static_cast<int>(ceilf(rank));
  const float a = sorted[i0], b = sorted[i1];
  range_out[threadIdx.x] = w < 0.5f ? a + w * (b - a) : b - (b - a) * (1.f - w);
}

// img_latent = mode * scaling (marigold_dc.py:696-698): first `C` channels of the NHWC moments -> NCHW bf16
__global__ void latent_out_kernel(const bf16* __restrict__ mom, long long ld, int N, int HW, int C, float scaling,
                                  bf16* __restrict__ out) {
  ptx::pdl_wait();  // launched with programmatic stream serialisation: without this the kernel would not wait
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * C * HW) return;
  const int p = i % HW, c = (i / HW) % C, n = i / (1LL * HW * C);
  out[i] = __float2bfloat16(__bfloat162float(mom[(1LL * n * HW + p) * ld + c]) * scaling);
}

// Valid points per sample (mask != 0): one block per sample.
__global__ void count_mask_kernel(const uint8_t* __restrict__ mask, int HW, int* __restrict__ counts) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  __shared__ float red[32];
  int c = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) c += mask[1LL * blockIdx.x * HW + i] != 0;
  const float t = block_sum(static_cast<float>(c), red);  // exact: a frame has far fewer than 2^24 pixels
  if (threadIdx.x == 0) counts[blockIdx.x] = static_cast<int>(t + 0.5f);
}
// Per-call device state derived from the point counts: pt_off = exclusive prefix sum over the samples.  With the
// 5-float frame statistics of sparse_norm_kernel ({lo, hi, gmin, gmax, n_valid} per sample) it also fills the guide /
// metric ranges the loss kernels read, so mdc_begin_frame needs no host round trip for them.
__global__ void frame_state_kernel(const int* __restrict__ counts, const float* __restrict__ stats5, int N, int* __restrict__ pt_off,
                                   float* __restrict__ gminmax, float* __restrict__ depth_minmax) {
  ptx::pdl_wait();
  ptx::pdl_launch();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int acc = 0;
  pt_off[0] = 0;
  for (int n = 0; n < N; ++n) {
    int c;
    if (stats5) {
      const float* s = stats5 + 5 * n;
      c = static_cast<int>(s[4] + 0.5f);
      depth_minmax[2 * n] = s[0], depth_minmax[2 * n + 1] = s[1];
      gminmax[2 * n] = s[2], gminmax[2 * n + 1] = s[3];
    } else {
      c = counts[n];
    }
    acc += c;
    pt_off[n + 1] = acc;
  }
}

// Ordered compaction of the valid pixels of one sample (mask != 0): one block per sample, run once per call.
__global__ void compact_points_kernel(const float* __restrict__ guide, const uint8_t* __restrict__ mask, int HW,
                                      const int* __restrict__ pt_off, int* __restrict__ pt_idx,
                                      float* __restrict__ pt_val) {
  ptx::pdl_wait();  // every kernel launched through launch_k must order itself after its predecessor
  ptx::pdl_launch();
  __shared__ int wcount[32];
  __shared__ int base;
  const int n = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) base = pt_off[n];
  __syncthreads();
  for (int p0 = 0; p0 < HW; p0 += blockDim.x) {
    const int p = p0 + threadIdx.x;
    const bool valid = p < HW && mask[1LL * n * HW + p] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) wcount[w] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < w; ++k) before += wcount[k];
    if (valid) {
      const int dst = base + before + __popc(bal & ((1u << lane) - 1));
      pt_idx[dst] = p;
      pt_val[dst] = guide[1LL * n * HW + p];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < nw; ++k) tot += wcount[k];
      base += tot;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// prepare-time helpers (not on the hot path): sinusoidal timestep embedding and small dense layers
// out[s][o] = act_out( sum_i W[o][i] * act_in(in[s][i]) + b[o] ), every stage rounded to bf16 like torch's bf16 ops.
__global__ void timestep_embedding_kernel(const int* __restrict__ timesteps, int steps, int dim, float* __restrict__ out) {
  ptx::pdl_wait();  // every kernel launched through launch_k must order itself after its predecessor
  ptx::pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= steps * dim) return;
  int s = i / dim, j = i % dim, half = dim / 2;
  int k = j < half ? j : j - half;
  float freq = expf(-logf(10000.f) * k / half);
  float arg = static_cast<float>(timesteps[s]) * freq;
  out[i] = bf16r(j < half ? cosf(arg) : sinf(arg));  // flip_sin_to_cos: [cos | sin]
}
__global__ void small_linear_kernel(const bf16* __restrict__ W, long long ldw, const float* __restrict__ bias,
                                    const float* __restrict__ in, long long ldin, int S, int In, int Out, int silu_in,
                                    float* __restrict__ out, long long ldout) {
  ptx::pdl_wait();  // every kernel launched through launch_k must order itself after its predecessor
  ptx::pdl_launch();
  const long long wid = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= 1LL * S * Out) return;
  const int s = wid / Out, o = wid % Out;
  float acc = 0.f;
  for (int i = lane; i < In; i += 32) {
    float a = in[s * ldin + i];
    if (silu_in) a = bf16r(siluf_(a));
    acc += __bfloat162float(W[o * ldw + i]) * a;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[s * ldout + o] = bf16r(acc + (bias ? bias[o] : 0.f));
}

}  // namespace mdc
