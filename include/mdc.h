/* libmdc_b200.so -- C ABI of the B200-native Marigold-DC guided denoising loop.
 *
 * Drop-in boundary for the hot path of tier4/depth_completion:
 *   MarigoldDepthCompletionPipeline.__call__            /root/reference/marigold_dc.py:467-985
 *     the 50-iteration guided DDIM loop                 marigold_dc.py:799-909
 *     the final decode + de-normalisation               marigold_dc.py:970-985
 * The reference has no FFI of its own (it is pure Python over diffusers/torch); the Python class
 * depth_completion_b200.MarigoldDepthCompletionPipeline keeps the reference's call signature and binds these
 * entry points with ctypes.  Plain pointers and sizes only; no torch types.
 *
 * Conventions: every function returns 0 on success, non-zero on failure with the message available from
 * mdc_last_error() (thread-local).  Pointers are DEVICE pointers unless the name ends in _host.  One handle
 * owns one device and all workspace; it is not thread-safe.  Every entry point makes the handle's device current
 * (cudaSetDevice) and enqueues its work on the handle's stream: a private blocking stream by default (implicitly ordered
 * with the legacy default stream), or the caller's stream after mdc_set_stream -- the reference runs on torch's current
 * stream (marigold_dc.py has no stream handling of its own), and the Python binding passes exactly that on every call.
 * The library never frees or keeps caller memory beyond the call (weights are re-packed into library-owned layouts).
 */
#ifndef MDC_H_
#define MDC_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDC_DTYPE_F32 0
#define MDC_DTYPE_BF16 1
#define MDC_MAX_BLOCKS 8

typedef struct mdc_config {
  int device;                 /* CUDA device ordinal */
  int n_batch;                /* frames per call (N of imgs [N,C,H,W]), 1..16 */
  int height, width;          /* input resolution H, W (marigold_dc.py:595) */
  int proc_h, proc_w;         /* processed size before padding: H*res//max, W*res//max (image processor) */
  int pad_h, pad_w;           /* replicate padding up to a multiple of 8 (bottom, right) */
  int steps;                  /* number of DDIM steps (scheduler.set_timesteps, marigold_dc.py:800) */
  /* UNet2DConditionModel config */
  int unet_in_ch, unet_out_ch, unet_nblocks, unet_layers_per_block, unet_groups, cross_dim;
  int unet_block_ch[MDC_MAX_BLOCKS];
  int unet_heads[MDC_MAX_BLOCKS];
  int unet_down_attn[MDC_MAX_BLOCKS];
  /* AutoencoderKL decoder config (Appendix A.2) */
  int vae_nblocks, vae_layers_per_block, vae_groups, vae_latent_ch;
  int vae_block_ch[MDC_MAX_BLOCKS];
  float vae_scaling;          /* 0.18215 (AutoencoderKL) / 1.0 (AutoencoderTiny) */
  /* vae_kind 0: AutoencoderKL (fields above).  1: AutoencoderTiny, the reference CLI's default VAE (predict.py:44-52,
   * :484-488): vae_nblocks stages of width vae_block_ch[i] (all equal), tiny_dec_blocks[i] / tiny_enc_blocks[i] residual
   * blocks per stage, latents clamped with tanh(z / tiny_magnitude) * tiny_magnitude; vae_layers_per_block / vae_groups
   * are ignored. */
  int vae_kind;
  int tiny_enc_blocks[MDC_MAX_BLOCKS];
  int tiny_dec_blocks[MDC_MAX_BLOCKS];
  float tiny_magnitude;
  /* != 0: this handle will run CONCURRENTLY with other handles on the same GPU (several frames in flight on separate
   * streams).  Kernels that need all of their CTAs co-resident -- the single-launch GroupNorm with its grid barrier -- are
   * then replaced by their two-pass forms, because two such kernels from different streams can starve each other. */
  int concurrent;
} mdc_config;

typedef struct mdc_handle mdc_handle;

const char* mdc_last_error(void);

/* Builds the UNet + VAE-decoder tapes, all launch plans and all workspace for the configured shapes. */
int mdc_create(const mdc_config* cfg, mdc_handle** out);
/* Same, but the new handle SHARES the packed parameters of `share_weights_with` (same device and model configuration;
 * frame geometry, batch and step count may differ): a sequence whose last batch is short, or a second resolution, costs
 * workspace and launch plans but no re-packing of the 1.8 GB of weights (the reference keeps one set of modules for
 * every call shape, predict.py:463-481, :585-700).  The parameters live until the last sharing handle is destroyed. */
int mdc_create_shared(const mdc_config* cfg, mdc_handle* share_weights_with, mdc_handle** out);
void mdc_destroy(mdc_handle* h);
/* Frees the tapes, launch plans and workspace of `h` but keeps the handle (and thereby the packed parameters) alive, so
 * that it can still be passed to mdc_create_shared; every other call on it fails afterwards. */
int mdc_release_workspace(mdc_handle* h);
/* All later calls on `h` enqueue their work on `cuda_stream` (a cudaStream_t; NULL = back to the handle's own stream).
 * Inputs must be ready in stream order on that stream; outputs are ready in stream order on it (calls documented as
 * synchronising synchronise that stream). */
int mdc_set_stream(mdc_handle* h, void* cuda_stream);

/* Expected parameters, named by their diffusers state-dict key with a "unet." / "vae." prefix
 * (SURVEY.md Appendix A.5), e.g. "unet.down_blocks.0.resnets.0.conv1.weight". */
int mdc_num_weights(mdc_handle* h);
const char* mdc_weight_key(mdc_handle* h, int i);
/* Logical shape of parameter i as the reference's modules hold it: [out,in,3,3] (conv3x3), [out,in] (linear and 1x1
 * conv; a [out,in,1,1] tensor is accepted too) or [n] (bias / norm vectors).  shape4_host has room for 4 entries. */
int mdc_weight_shape(mdc_handle* h, int i, long long* shape4_host, int* ndim_host);
/* Re-packs one parameter (device pointer, contiguous, dtype MDC_DTYPE_*) into the library's layouts. */
int mdc_set_weight(mdc_handle* h, const char* key, const void* dev_ptr, const long long* shape_host, int ndim,
                   int dtype);
/* Batched form: n parameters re-packed by ONE kernel launch.  shapes4_host holds 4 entries per parameter (unused
 * trailing dimensions ignored).  The sources have been read when the call returns. */
int mdc_set_weights(mdc_handle* h, int n, const char* const* keys, const void* const* dev_ptrs, const long long* shapes4_host,
                    const int* ndims_host, const int* dtypes_host);
/* 1 when every parameter of the handle (or of the bank it shares) has been set; mdc_weight_is_loaded: parameter i in
 * every layout the sharing handles need (a handle with another frame geometry may add a layout of an already loaded
 * parameter, e.g. the plain 3x3 form of an upsampler weight for odd latent sizes: only those need setting again). */
int mdc_weights_loaded(mdc_handle* h);
int mdc_weight_is_loaded(mdc_handle* h, int i);

/* Step-invariant precomputation: DDIM scalars for `timesteps`, the time embedding of every step pushed through
 * every resnet's time_emb_proj, and the cross-attention K/V of the empty-prompt embedding ctx [1,2,cross_dim] bf16
 * (marigold_dc.py:664-674, :800).  alphas_cumprod_host has 1000 entries. */
int mdc_prepare(mdc_handle* h, const void* ctx_bf16, const float* alphas_cumprod_host, const int* timesteps_host,
                int n_steps);

/* Per-frame prologue (marigold_dc.py:687-698): MarigoldImageProcessor.preprocess (integer / 255,
 * * 2 - 1, antialiased bilinear resize to the processing resolution, replicate padding to a multiple of 8) followed by
 * AutoencoderKL.encode(...).latent_dist.mode() * scaling_factor.  imgs: device [N, channels, H, W], uint8 (dtype 0) or
 * fp32 in [0, 1] (dtype 1), channels 1 or 3 (a single channel is repeated).  latents_out: device [N,4,EH,EW] bf16,
 * ready to be passed to mdc_begin as img_latents. */
int mdc_encode(mdc_handle* h, const void* imgs, int dtype, int channels, void* latents_out_bf16);

/* Per-call state (marigold_dc.py:696-789): image latents and initial depth latent [N,4,EH,EW] bf16 NCHW, normalised
 * sparse depth `guide` [N,1,H,W] fp32 with `mask` [N,1,H,W] uint8, per-sample (min,max) of the masked guide and of the
 * metric depth range (host, 2 floats per sample), learning rates of the latent and of scale/shift.
 * Fails if a sample has an empty mask (reference: ValueError from utils.py:132). */
int mdc_begin(mdc_handle* h, const void* img_latents_bf16, const void* x_bf16, const float* guide, const uint8_t* mask,
              const float* guide_minmax_host, const float* depth_minmax_host, float lr_latent, float lr_scaling);

/* Per-call options of the non-default branches of the reference call (marigold_dc.py:467-493); they apply to the next
 * mdc_begin / mdc_begin_frame and stay until changed.  Defaults: linear, no inverse, adam, l1 + l2, no kld, (0.01, 0.99).
 *   projection   0 "linear", 1 "log", 2 "log10" (get_projection_fn, :23-50), inv != 0: inverse depth (:743-749, :842-862)
 *   opt          0 "adam", 1 "sgd", 2 "adagrad" with torch's default hyper-parameters (:776-789)
 *   loss_weights4_host  how many times "l1", "l2", "edge", "smooth" appear in loss_funcs (compute_loss, :171-236, sums
 *                the listed terms); "edge" needs the images, i.e. mdc_begin_frame rather than mdc_begin
 *   kld_mode     0 off, 1 "simple", 2 "strict" (utils.py:28-86), added as kld_weight * kld (:238-241)
 *   percentile_lo / hi   the quantiles of norm = "percentile" (:715-728)
 *   closed_form  != 0: scale / shift are refitted by masked least squares every guided step (compute_affine_params,
 *                :53-128, :332-336) and the loss gradient flows through the fit; finish with
 *                mdc_decode_final_closed_form.  Not available together with edge / smooth.
 *   interp_nearest  != 0: interp_mode = "nearest" for the resize of the prediction to the input resolution (:366-370)
 *                instead of "bilinear" (the two modes predict.py:200-206 offers) */
int mdc_set_options(mdc_handle* h, int projection, int inv, int opt, const float* loss_weights4_host, int kld_mode,
                    float kld_weight, float percentile_lo, float percentile_hi, int closed_form, int interp_nearest);

/* The whole per-frame prologue in one call (marigold_dc.py:687-756): mdc_encode on `imgs`, then the
 * sparse-depth normalisation on the device -- mask = sparse > 0, per-sample masked min / max ("minmax", norm_mode =
 * 0), the constant range [min_depth, max_depth] (norm_mode = 1) or per-sample quantiles of the valid values
 * ("percentile", norm_mode = 2), clamp, projection / inverse as set by mdc_set_options, guide = (d - lo) / (hi - lo),
 * masked min / max of the guide -- then mdc_begin with those.  sparse: device [N,1,H,W] fp32 metres (0 = missing); x_bf16:
 * the initial depth latent [N,4,EH,EW] (the reference draws it from torch's seeded generator, :677-684).
 * Fails with the reference's "No valid values found in mask ..." message (utils.py:132-136) for an empty sample. */
int mdc_begin_frame(mdc_handle* h, const void* imgs, int img_dtype, int channels, const float* sparse, const void* x_bf16,
                    float max_depth, float min_depth, int norm_mode, float lr_latent, float lr_scaling);

/* mdc_begin_frame with the image latents already encoded ("pipeline frame k+1's encoder under frame
 * k's loop", predict.py:599-700): img_latents_bf16 [N,4,EH,EW] is the output of an earlier mdc_encode -- typically of a
 * second handle of the same geometry (mdc_create_shared, concurrent = 1) that ran on another stream while this handle
 * was still in frame k's guided loop -- so only the sparse-depth normalisation and the per-call state are left on this
 * handle's critical path.  imgs may be NULL unless the edge loss is selected (it reads the raw image, :195-236). */
int mdc_begin_frame_encoded(mdc_handle* h, const void* img_latents_bf16, const void* imgs, int img_dtype, int channels,
                            const float* sparse, const void* x_bf16, float max_depth, float min_depth, int norm_mode,
                            float lr_latent, float lr_scaling);

/* n guided steps (marigold_dc.py:801-904 each), asynchronous, no host synchronisation inside. */
int mdc_run(mdc_handle* h, int n_steps);

/* Synchronises and copies out the current latent [N,4,EH,EW] bf16 and, per sample, scale, shift and the last loss
 * (any pointer may be NULL).  x_out is a device pointer; the float arrays are host pointers. */
int mdc_get_state(mdc_handle* h, void* x_out_bf16, float* scale_host, float* shift_host, float* loss_host);

/* Final decode + affine + clamp + de-normalisation -> dense [N,1,H,W] fp32 metric depth (marigold_dc.py:970-984). */
int mdc_decode_final(mdc_handle* h, float* dense_out);

/* The no-grad branch of the call (train_latents = False, hence closed_form = True; marigold_dc.py:605-613): n plain DDIM
 * steps -- UNet forward + scheduler.step(...).prev_sample, no decoder, no guidance (:805-809, :905-909) -- and the final
 * decode with the closed-form masked least-squares scale / shift (compute_affine_params, :53-128, :332-336) instead of
 * the learned ones.  mdc_get_state then returns that scale / shift. */
int mdc_sample(mdc_handle* h, int n_steps);
int mdc_decode_final_closed_form(mdc_handle* h, float* dense_out);

/* Number of kernel launches issued by the handle so far, and bytes of device memory it owns. */
long long mdc_launch_count(mdc_handle* h);
long long mdc_device_bytes(mdc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* MDC_H_ */
