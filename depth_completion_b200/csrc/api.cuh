// extern "C" surface of libmdc_b200.so (include/mdc.h, include/mdc_debug.h) over mdc::Engine.
#pragma once
#include "debug_api.cuh"
#include "engine.cuh"

struct mdc_handle {
  mdc::Engine* e = nullptr;
  std::vector<std::string> keys, tnames;
};

namespace mdc {
__global__ void nchw_f32_to_nhwc_kernel(const float* __restrict__ src, int N, int HW, int C, bf16* __restrict__ dst,
                                        long long ld) {
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * HW * ld) return;
  int c = i % ld;
  long long p = i / ld;
  int n = p / HW, q = p % HW;
  dst[i] = __float2bfloat16(c < C ? src[(1LL * n * C + c) * HW + q] : 0.f);
}
inline void load_nchw(Tensor* t, bf16* dst, const float* src, cudaStream_t st) {
  long long tot = t->rows() * t->ld;
  nchw_f32_to_nhwc_kernel<<<static_cast<int>((tot + 255) / 256), 256, 0, st>>>(src, t->n, t->h * t->w, t->c, dst, t->ld);
}
inline void store_nchw(Tensor* t, const bf16* src, float* dst, cudaStream_t st) {
  long long tot = t->rows() * t->c;
  nhwc_to_nchw_f32_kernel<<<static_cast<int>((tot + 255) / 256), 256, 0, st>>>(src, t->ld, t->n, t->h * t->w, t->c, dst);
}
}  // namespace mdc

extern "C" {

static int create_impl(const mdc_config* cfg, mdc_handle* share, mdc_handle** out) {
  return mdc::guarded([&] {
    MDC_CHECK(cfg && out, "null argument");
    auto* h = new mdc_handle();
    try {
      h->e = new mdc::Engine(*cfg, share ? share->e->bank : nullptr);
    } catch (...) {
      delete h;
      throw;
    }
    for (auto& kv : h->e->wmap()) h->keys.push_back(kv.first);
    for (auto& kv : h->e->named) h->tnames.push_back(kv.first);
    *out = h;
  });
}
int mdc_create(const mdc_config* cfg, mdc_handle** out) { return create_impl(cfg, nullptr, out); }
int mdc_create_shared(const mdc_config* cfg, mdc_handle* share_weights_with, mdc_handle** out) {
  return create_impl(cfg, share_weights_with, out);
}
int mdc_set_stream(mdc_handle* h, void* cuda_stream) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->set_stream(static_cast<cudaStream_t>(cuda_stream));
  });
}
int mdc_release_workspace(mdc_handle* h) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    if (h->e->released) return;
    MDC_CUDA(cudaSetDevice(h->e->cfg.device));
    h->e->release_workspace();
    h->tnames.clear();
  });
}
int mdc_weights_loaded(mdc_handle* h) { return (h && h->e->weights_loaded()) ? 1 : 0; }
int mdc_weight_is_loaded(mdc_handle* h, int i) {
  if (!h || i < 0 || i >= static_cast<int>(h->keys.size())) return 0;
  for (const mdc::WeightSlot* s : h->e->wmap().at(h->keys[i]))
    if (!s->loaded) return 0;
  return 1;
}
int mdc_set_weights(mdc_handle* h, int n, const char* const* keys, const void* const* dev_ptrs, const long long* shapes4_host,
                    const int* ndims_host, const int* dtypes_host) {
  return mdc::guarded([&] {
    MDC_CHECK(h && keys && dev_ptrs && shapes4_host && ndims_host && dtypes_host && n >= 0, "bad argument");
    h->e->activate();
    h->e->set_weights(n, keys, dev_ptrs, shapes4_host, ndims_host, dtypes_host);
  });
}
void mdc_destroy(mdc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->e->cfg.device);
  cudaDeviceSynchronize();
  delete h->e;
  delete h;
}
int mdc_num_weights(mdc_handle* h) { return h ? static_cast<int>(h->keys.size()) : 0; }
const char* mdc_weight_key(mdc_handle* h, int i) {
  return (h && i >= 0 && i < static_cast<int>(h->keys.size())) ? h->keys[i].c_str() : nullptr;
}
int mdc_weight_shape(mdc_handle* h, int i, long long* shape4_host, int* ndim_host) {
  return mdc::guarded([&] {
    MDC_CHECK(h && shape4_host && ndim_host && i >= 0 && i < static_cast<int>(h->keys.size()), "bad argument");
    const mdc::WeightSlot* s = h->e->wmap().at(h->keys[i]).front();
    if (s->kind == mdc::W_CONV3 || s->kind == mdc::W_UPCONV) {
      shape4_host[0] = s->out, shape4_host[1] = s->in, shape4_host[2] = 3, shape4_host[3] = 3, *ndim_host = 4;
    } else if (s->kind == mdc::W_LIN) {
      shape4_host[0] = s->out, shape4_host[1] = s->in, *ndim_host = 2;
    } else {
      shape4_host[0] = s->out, *ndim_host = 1;
    }
  });
}
int mdc_set_weight(mdc_handle* h, const char* key, const void* dev_ptr, const long long* shape_host, int ndim, int dtype) {
  return mdc::guarded([&] {
    MDC_CHECK(h && key && dev_ptr && shape_host, "null argument");
    h->e->activate();
    h->e->set_weight(key, dev_ptr, shape_host, ndim, dtype);
  });
}
int mdc_prepare(mdc_handle* h, const void* ctx_bf16, const float* alphas_cumprod_host, const int* timesteps_host,
                int n_steps) {
  return mdc::guarded([&] {
    MDC_CHECK(h && ctx_bf16 && alphas_cumprod_host && timesteps_host, "null argument");
    h->e->activate();
    h->e->prepare(ctx_bf16, alphas_cumprod_host, timesteps_host, n_steps);
  });
}
int mdc_begin(mdc_handle* h, const void* img_latents_bf16, const void* x_bf16, const float* guide, const uint8_t* mask,
              const float* guide_minmax_host, const float* depth_minmax_host, float lr_latent, float lr_scaling) {
  return mdc::guarded([&] {
    MDC_CHECK(h && img_latents_bf16 && x_bf16 && guide && mask && guide_minmax_host && depth_minmax_host, "null argument");
    h->e->activate();
    h->e->begin(img_latents_bf16, x_bf16, guide, mask, guide_minmax_host, depth_minmax_host, lr_latent, lr_scaling);
  });
}
int mdc_begin_frame(mdc_handle* h, const void* imgs, int img_dtype, int channels, const float* sparse, const void* x_bf16,
                    float max_depth, float min_depth, int norm_mode, float lr_latent, float lr_scaling) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->activate();
    h->e->begin_frame(imgs, img_dtype, channels, sparse, x_bf16, max_depth, min_depth, norm_mode, lr_latent, lr_scaling);
  });
}
int mdc_begin_frame_encoded(mdc_handle* h, const void* img_latents_bf16, const void* imgs, int img_dtype, int channels,
                            const float* sparse, const void* x_bf16, float max_depth, float min_depth, int norm_mode, float lr_latent,
                            float lr_scaling) {
  return mdc::guarded([&] {
    MDC_CHECK(h && img_latents_bf16, "null argument");
    h->e->activate();
    h->e->begin_frame(imgs, img_dtype, channels, sparse, x_bf16, max_depth, min_depth, norm_mode, lr_latent, lr_scaling,
                      img_latents_bf16);
  });
}
int mdc_set_options(mdc_handle* h, int projection, int inv, int opt, const float* loss_weights4_host, int kld_mode,
                    float kld_weight, float percentile_lo, float percentile_hi, int closed_form,
                    int interp_nearest) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->set_options(projection, inv, opt, loss_weights4_host, kld_mode, kld_weight, percentile_lo, percentile_hi, closed_form,
                      interp_nearest);
  });
}
int mdc_run(mdc_handle* h, int n_steps) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->activate();
    for (int i = 0; i < n_steps; ++i) h->e->step();
    MDC_CUDA(cudaGetLastError());
  });
}
int mdc_get_state(mdc_handle* h, void* x_out_bf16, float* scale_host, float* shift_host, float* loss_host) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->activate();
    mdc::Engine* e = h->e;
    if (x_out_bf16)
      MDC_CUDA(cudaMemcpyAsync(x_out_bf16, e->x, 8ull * e->N * e->lh * e->lw, cudaMemcpyDeviceToDevice, e->stream));
    mdc::StepAccum a;
    MDC_CUDA(cudaMemcpyAsync(&a, e->accum, sizeof(a), cudaMemcpyDeviceToHost, e->stream));
    MDC_CUDA(cudaStreamSynchronize(e->stream));
    MDC_CUDA(cudaGetLastError());
    e->check_barrier_flag();
    for (int i = 0; i < e->N; ++i) {
      if (scale_host) scale_host[i] = a.scale[i];
      if (shift_host) shift_host[i] = a.shift[i];
      if (loss_host) loss_host[i] = a.loss[i];
    }
  });
}
int mdc_encode(mdc_handle* h, const void* imgs, int dtype, int channels, void* latents_out_bf16) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->activate();
    h->e->encode(imgs, dtype, channels, latents_out_bf16);
  });
}
int mdc_sample(mdc_handle* h, int n_steps) {
  return mdc::guarded([&] {
    MDC_CHECK(h, "null handle");
    h->e->activate();
    for (int i = 0; i < n_steps; ++i) h->e->sample_step();
    MDC_CUDA(cudaGetLastError());
  });
}
int mdc_decode_final_closed_form(mdc_handle* h, float* dense_out) {
  return mdc::guarded([&] {
    MDC_CHECK(h && dense_out, "null pointer");
    h->e->activate();
    h->e->decode_final(dense_out, true);
  });
}
int mdc_decode_final(mdc_handle* h, float* dense_out) {
  return mdc::guarded([&] {
    MDC_CHECK(h && dense_out, "null argument");
    h->e->activate();
    h->e->decode_final(dense_out);
  });
}
long long mdc_launch_count(mdc_handle* h) { return h ? h->e->launches : 0; }
long long mdc_device_bytes(mdc_handle* h) { return h ? static_cast<long long>(h->e->arena.total + h->e->bank->arena.total) : 0; }

// ---------------------------------------------------------------------------------------------- debug
int mdc_dbg_forward(mdc_handle* h, int which, int step, const float* in_nchw, float* out_nchw) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CHECK(e->prepared, "prepare first");
    mdc::Tensor* in = which == 0 ? e->unet_in : e->dec_in;
    mdc::Tensor* out = which == 0 ? e->unet_out : e->dec_out;
    int s = step;
    e->copy_sync(e->counter, &s, 4, cudaMemcpyHostToDevice);
    mdc::begin_step_kernel<<<1, 1024, 0, e->stream>>>(e->tables, e->counter, e->cur, e->temb_cur, e->opts);
    mdc::load_nchw(in, in->d, in_nchw, e->stream);
    e->run_ops(which == 0 ? e->unet_ops : e->dec_ops, false);
    mdc::store_nchw(out, out->d, out_nchw, e->stream);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaStreamSynchronize(e->stream));
  });
}
int mdc_dbg_backward(mdc_handle* h, int which, const float* dout_nchw, float* din_nchw) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    mdc::Tensor* in = which == 0 ? e->unet_in : e->dec_in;
    mdc::Tensor* out = which == 0 ? e->unet_out : e->dec_out;
    mdc::load_nchw(out, out->g, dout_nchw, e->stream);
    e->run_ops(which == 0 ? e->unet_ops : e->dec_ops, true);
    mdc::store_nchw(in, in->g, din_nchw, e->stream);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaStreamSynchronize(e->stream));
  });
}
int mdc_dbg_read_tensor(mdc_handle* h, const char* name, int which, float* out_nchw) {
  return mdc::guarded([&] { h->e->activate(); h->e->read_tensor(name, which, out_nchw); });
}
int mdc_dbg_tensor_shape(mdc_handle* h, const char* name, int* nchw_host) {
  return mdc::guarded([&] {
    auto it = h->e->named.find(name);
    MDC_CHECK(it != h->e->named.end(), "unknown tensor '%s'", name);
    nchw_host[0] = it->second->n, nchw_host[1] = it->second->c, nchw_host[2] = it->second->h, nchw_host[3] = it->second->w;
  });
}
int mdc_dbg_num_tensors(mdc_handle* h) { return static_cast<int>(h->tnames.size()); }
const char* mdc_dbg_tensor_name(mdc_handle* h, int i) {
  return (i >= 0 && i < static_cast<int>(h->tnames.size())) ? h->tnames[i].c_str() : nullptr;
}
int mdc_dbg_read_x_adam(mdc_handle* h, void* x_out_bf16) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CUDA(cudaStreamSynchronize(e->stream));
    e->copy_sync(x_out_bf16, e->x_adam_dbg, 8ull * e->N * e->lh * e->lw, cudaMemcpyDeviceToDevice);
  });
}
int mdc_dbg_read_buffer(mdc_handle* h, const char* which, float* out_dev) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    const std::string w(which);
    const size_t lat = 4ull * e->N * e->lh * e->lw;
    MDC_CUDA(cudaStreamSynchronize(e->stream));
    if (w == "grad")
      e->copy_sync(out_dev, e->gbuf, lat * 4, cudaMemcpyDeviceToDevice);
    else if (w == "dx_direct")
      e->copy_sync(out_dev, e->dx_direct, lat * 4, cudaMemcpyDeviceToDevice);
    else
      MDC_CHECK(false, "unknown buffer '%s'", which);
  });
}
int mdc_dbg_frame_state(mdc_handle* h, float* guide_dev, unsigned char* mask_dev, float* stats_host) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CHECK(e->fr_guide != nullptr, "mdc_begin_frame has not been called");
    const size_t n = 1ull * e->N * e->H * e->W;
    MDC_CUDA(cudaStreamSynchronize(e->stream));
    if (guide_dev) e->copy_sync(guide_dev, e->fr_guide, n * 4, cudaMemcpyDeviceToDevice);
    if (mask_dev) e->copy_sync(mask_dev, e->fr_mask, n, cudaMemcpyDeviceToDevice);
    if (stats_host) e->copy_sync(stats_host, e->fr_stats, 5ull * e->N * 4, cudaMemcpyDeviceToHost);
  });
}
int mdc_dbg_loss(mdc_handle* h, const float* dec_nchw, float* ddec_nchw, float* loss_host, float* sgrad_host,
                 float* tgrad_host) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CHECK(e->begun, "mdc_begin first");
    mdc::load_nchw(e->dec_out, e->dec_out->d, dec_nchw, e->stream);
    mdc::TailGeom g{e->N, e->H, e->W, e->ph, e->pw, e->PPH, e->PPW, e->dec_out->ld, e->interp_nearest};
    mdc::loss_points_kernel<<<e->N, 512, 0, e->stream>>>(e->dec_out->d, g, e->pt_idx, e->pt_val, e->pt_off, e->gminmax,
                                                         e->depth_minmax, e->opts, e->accum, e->dmean);
    mdc::loss_points_cf_kernel<<<e->N, 512, 0, e->stream>>>(e->dec_out->d, g, e->pt_idx, e->pt_val, e->pt_off, e->depth_minmax, e->opts,
                                                            e->accum, e->dmean, e->pt_a, e->pt_G);
    const long long npix = 1LL * e->N * e->PPH * e->PPW, opix = 1LL * e->N * e->H * e->W;
    mdc::dense_map_kernel<<<static_cast<int>((opix + 255) / 256), 256, 0, e->stream>>>(e->dec_out->d, g, e->gminmax, e->depth_minmax,
                                                                                      e->opts, e->accum, e->dn_map);
    mdc::dense_loss_kernel<<<dim3(std::max(1, std::min(148, (e->H * e->W + 255) / 256)), e->N), 256, 0, e->stream>>>(
        e->dec_out->d, g, e->gminmax, e->depth_minmax, e->opts, e->dn_map, e->gray_gx, e->gray_gy, e->accum, e->dmean);
    mdc::dec_grad_kernel<<<static_cast<int>((npix + 255) / 256), 256, 0, e->stream>>>(e->dmean, npix, e->dec_out->g);
    mdc::store_nchw(e->dec_out, e->dec_out->g, ddec_nchw, e->stream);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaStreamSynchronize(e->stream));
    mdc::StepAccum a;
    e->copy_sync(&a, e->accum, sizeof(a), cudaMemcpyDeviceToHost);
    for (int i = 0; i < e->N; ++i) loss_host[i] = a.loss[i], sgrad_host[i] = a.s_grad[i], tgrad_host[i] = a.t_grad[i];
  });
}
int mdc_dbg_update(mdc_handle* h, const float* v_nchw, const float* dz_nchw, const float* dunet_in_nchw) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CHECK(e->begun, "mdc_begin first");
    const int hw = e->lh * e->lw, lat_pix = e->N * hw, pgrid = e->N * e->parts_per_img;
    mdc::begin_step_kernel<<<1, 1024, 0, e->stream>>>(e->tables, e->counter, e->cur, e->temb_cur, e->opts);
    mdc::load_nchw(e->unet_out, e->unet_out->d, v_nchw, e->stream);
    mdc::x0_kernel<<<pgrid, 256, 0, e->stream>>>(e->unet_out->d, e->x, e->cur, e->N, hw, e->cfg.vae_scaling, e->dec_in->d,
                                                e->eps_part, e->x1_part, e->x2_part);
    mdc::load_nchw(e->dec_in, e->dec_in->g, dz_nchw, e->stream);
    mdc::dx0_kernel<<<(lat_pix + 255) / 256, 256, 0, e->stream>>>(e->dec_in->g, e->cur, e->N, hw, e->cfg.vae_scaling,
                                                                  e->unet_out->g, e->dx_direct);
    mdc::load_nchw(e->unet_in, e->unet_in->g, dunet_in_nchw, e->stream);
    mdc::grad_total_kernel<<<pgrid, 256, 0, e->stream>>>(e->dx_direct, e->unet_in->g, e->N, hw, e->gbuf, e->g_part, e->x,
                                                         e->x1_part, e->x2_part, e->opts, e->accum);
    mdc::adam_ddim_kernel<<<pgrid, 256, 0, e->stream>>>(e->gbuf, e->eps_part, e->g_part, e->parts_per_img, e->unet_out->d,
                                                       e->cur, e->N, hw, e->x, e->m1, e->m2, e->accum, e->counter,
                                                       e->x_adam_dbg, e->opts);
    MDC_CUDA(cudaGetLastError());
    MDC_CUDA(cudaStreamSynchronize(e->stream));
  });
}
// Teacher forcing: puts the handle (after mdc_begin / mdc_begin_frame) into the state of guided step `step`: latent x and
// its Adam moments [N,4,EH,EW] bf16 (device), per-sample scale / shift and their fp32 Adam moments (host, N floats each,
// order: scale, shift, s_m, s_v, t_m, t_v; NULL keeps the current values).  The next mdc_run(h, 1) executes step `step`.
int mdc_dbg_set_state(mdc_handle* h, int step, const void* x_bf16, const void* m1_bf16, const void* m2_bf16,
                      const float* affine6_host) {
  return mdc::guarded([&] {
    h->e->activate();
    mdc::Engine* e = h->e;
    MDC_CHECK(e->begun, "mdc_begin first");
    MDC_CHECK(step >= 0 && step < e->cfg.steps, "step %d out of range", step);
    const size_t lat = 8ull * e->N * e->lh * e->lw;
    if (x_bf16) MDC_CUDA(cudaMemcpyAsync(e->x, x_bf16, lat, cudaMemcpyDeviceToDevice, e->stream));
    if (m1_bf16) MDC_CUDA(cudaMemcpyAsync(e->m1, m1_bf16, lat, cudaMemcpyDeviceToDevice, e->stream));
    if (m2_bf16) MDC_CUDA(cudaMemcpyAsync(e->m2, m2_bf16, lat, cudaMemcpyDeviceToDevice, e->stream));
    if (affine6_host) {
      mdc::StepAccum a;
      e->copy_sync(&a, e->accum, sizeof(a), cudaMemcpyDeviceToHost);
      for (int i = 0; i < e->N; ++i) {
        a.scale[i] = affine6_host[0 * e->N + i], a.shift[i] = affine6_host[1 * e->N + i];
        a.s_m[i] = affine6_host[2 * e->N + i], a.s_v[i] = affine6_host[3 * e->N + i];
        a.t_m[i] = affine6_host[4 * e->N + i], a.t_v[i] = affine6_host[5 * e->N + i];
      }
      e->copy_sync(e->accum, &a, sizeof(a), cudaMemcpyHostToDevice);
    }
    e->copy_sync(e->counter, &step, 4, cudaMemcpyHostToDevice);
    e->steps_done = step;
  });
}
int mdc_dbg_profile_gemm_step(mdc_handle* h, float* ms_host, double* flops_host, int* launches_host) {
  return mdc::guarded([&] { h->e->activate(); h->e->profile_gemm_step(ms_host, flops_host, launches_host); });
}
int mdc_dbg_profile_ops(mdc_handle* h, const char* csv_path, int iters) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    FILE* f = fopen(csv_path, "w");
    MDC_CHECK(f != nullptr, "cannot open %s", csv_path);
    fprintf(f, "tape,idx,name,fwd_us,bwd_us,gemm_gflop_fwd,gemm_gflop_bwd,launches_fwd,launches_bwd\n");
    cudaEvent_t e0, e1;
    MDC_CUDA(cudaEventCreate(&e0));
    MDC_CUDA(cudaEventCreate(&e1));
    auto time_it = [&](mdc::Op* op, bool bwd) {
      bwd ? op->bwd(e->stream) : op->fwd(e->stream);
      MDC_CUDA(cudaEventRecord(e0, e->stream));
      for (int i = 0; i < iters; ++i) bwd ? op->bwd(e->stream) : op->fwd(e->stream);
      MDC_CUDA(cudaEventRecord(e1, e->stream));
      MDC_CUDA(cudaEventSynchronize(e1));
      MDC_CUDA(cudaGetLastError());
      float ms = 0;
      MDC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      return ms * 1e3f / iters;
    };
    int t = 0;
    for (auto* ops : {&e->unet_ops, &e->dec_ops}) {
      int idx = 0;
      for (auto& op : *ops) {
        std::vector<const mdc::GemmPlan*> gf, gb;
        op->gemm_plans(gf, gb);
        double ff = 0, fb = 0;
        for (auto* g : gf) ff += g->flops;
        for (auto* g : gb) fb += g->flops;
        float tf = op->n_fwd() ? time_it(op.get(), false) : 0.f;
        float tb = op->n_bwd() ? time_it(op.get(), true) : 0.f;
        fprintf(f, "%s,%d,%s,%.2f,%.2f,%.3f,%.3f,%d,%d\n", t == 0 ? "unet" : "dec", idx++, op->name.c_str(), tf, tb,
                ff / 1e9, fb / 1e9, op->n_fwd(), op->n_bwd());
      }
      ++t;
    }
    fclose(f);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
  });
}
int mdc_dbg_time_tapes(mdc_handle* h, int iters, float* ms_host) {
  return mdc::guarded([&] { h->e->activate();
    mdc::Engine* e = h->e;
    cudaEvent_t e0, e1;
    MDC_CUDA(cudaEventCreate(&e0));
    MDC_CUDA(cudaEventCreate(&e1));
    for (int k = 0; k < 4; ++k) {
      auto& ops = k < 2 ? e->unet_ops : e->dec_ops;
      const bool bwd = k & 1;
      e->run_ops(ops, bwd);
      MDC_CUDA(cudaStreamSynchronize(e->stream));
      MDC_CUDA(cudaEventRecord(e0, e->stream));
      for (int i = 0; i < iters; ++i) e->run_ops(ops, bwd);
      MDC_CUDA(cudaEventRecord(e1, e->stream));
      MDC_CUDA(cudaEventSynchronize(e1));
      MDC_CUDA(cudaGetLastError());
      MDC_CUDA(cudaEventElapsedTime(&ms_host[k], e0, e1));
      ms_host[k] /= iters;
    }
    cudaEventDestroy(e0), cudaEventDestroy(e1);
  });
}

}  // extern "C"
