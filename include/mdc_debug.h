/* Kernel- and tape-level entry points of libmdc_b200.so used by tests/ and profiling only.
 * They let each CUDA kernel / sub-graph of the hot path be compared with the oracle in isolation. */
#ifndef MDC_DEBUG_H_
#define MDC_DEBUG_H_
#include "mdc.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Batched D = alpha * A . B^T (+bias) (+res) with the tcgen05 GEMM; *_mn = 1 selects MN-major operands. */
int mdc_dbg_gemm(int M, int N, int K, const void* A, int a_mn, long long lda, long long sa0, long long sa1,
                 const void* B, int b_mn, long long ldb, long long sb0, long long sb1, void* out, int out_f32,
                 long long ldc, long long sc0, long long sc1, const float* bias, const void* res, long long ldr,
                 long long sr0, long long sr1, float alpha, int nb0, int nb1, int bn_override, int iters,
                 float* ms_out);
/* 3x3/s1/p1 convolution (dgrad = 0) or its input gradient (dgrad = 1) on an NHWC bf16 tensor. */
int mdc_dbg_conv3x3(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const float* w_oihw,
                    int dgrad, const float* bias, const float* bias_img, const void* res, long long ldr, void* out,
                    long long ldc, int iters, float* ms_out);

/* Fused nearest-2x upsample + conv3x3 (dgrad = 0: x [NB,H,W,C] -> out [NB,2H,2W,Cout]) or its input gradient (dgrad = 1:
 * x is dy [NB,2H,2W,Cout] -> out [NB,H,W,C]); w_oihw [Cout][C][3][3] fp32. */
int mdc_dbg_upconv(int NB, int H, int W, int C, int Cout, const void* x, long long ldx, const float* w_oihw, int dgrad,
                   const float* bias, void* out, long long ldc, int iters, float* ms_out);
/* Self-attention as the engine plans it: head_dim 64 = the fused tcgen05 flash kernels (UNet attn1), otherwise GEMM +
 * softmax + GEMM (VAE mid block, head_dim 512).  qkv [n,T,3*heads*dh] bf16 (q | k | v column blocks), o / dout
 * [n,T,heads*dh], dqkv like qkv; dout = dqkv = NULL runs the forward only.  ms_out[2]: forward / backward ms per call. */
int mdc_dbg_attention(int n, int T, int heads, int dh, const void* qkv, void* o, const void* dout, void* dqkv, int iters,
                      float* ms_out);
/* GroupNorm (+SiLU when silu != 0) forward and (dy != NULL) input-gradient backward on x [n,HW,C] bf16 NHWC with the
 * engine's kernel selection: mode 0 automatic, 1 force the two-pass statistics / apply kernels (the 56-226 MB decoder
 * tensors), 2 require the single-launch grid-barrier kernels.  acc != 0: dx += (gradient accumulation at fan-out
 * points).  stats_out (device, optional): [n,groups,2] (mean, rstd).  ms_out[2]: forward / backward ms per call. */
int mdc_dbg_groupnorm(int n, int HW, int C, int groups, float eps, int silu, const void* x, const float* gamma,
                      const float* beta, void* y, const void* dy, void* dx, int acc, int mode, float* stats_out, int iters,
                      float* ms_out);

/* Tape-level: which = 0 UNet, 1 VAE decoder.  Inputs/outputs are NCHW fp32 device buffers.
 * forward: copies `in` into the tape input, runs the forward tape (step index selects the time embedding),
 * writes the tape output to `out`.  backward: seeds the output gradient with `dout`, runs the backward tape,
 * writes the input gradient to `din`. */
int mdc_dbg_forward(mdc_handle* h, int which, int step, const float* in_nchw, float* out_nchw);
int mdc_dbg_backward(mdc_handle* h, int which, const float* dout_nchw, float* din_nchw);
/* Copies a named intermediate (diffusers module path, e.g. "unet.down_blocks.0.resnets.0") as NCHW fp32;
 * which = 0 activation, 1 gradient.  mdc_dbg_tensor_shape fills {n, c, h, w}. */
int mdc_dbg_read_tensor(mdc_handle* h, const char* name, int which, float* out_nchw);
int mdc_dbg_tensor_shape(mdc_handle* h, const char* name, int* nchw_host);
int mdc_dbg_num_tensors(mdc_handle* h);
const char* mdc_dbg_tensor_name(mdc_handle* h, int i);
/* Latent after the Adam update but before the DDIM step of the last guided step ([N,4,EH,EW] bf16). */
int mdc_dbg_read_x_adam(mdc_handle* h, void* x_out_bf16);
/* Step-level scratch as fp32 [N,4,EH,EW] device buffers: "grad" = total latent gradient before the norm rescale
 * (marigold_dc.py:877), "dx_direct" = its part that does not pass through the UNet. */
int mdc_dbg_read_buffer(mdc_handle* h, const char* which, float* out_dev);
/* Tail kernels in isolation (after mdc_begin).  mdc_dbg_loss: masked L1+L2 loss and its gradient for a given decoder
 * output [N,3,PPH,PPW] fp32 (device) -> d/d dec (device), per-sample loss / d scale / d shift (host).
 * mdc_dbg_update: given the UNet output v [N,4,EH,EW], the decoder-input gradient dz [N,4,EH,EW] and the UNet-input
 * gradient [N,8,EH,EW] (fp32, device) runs x0/eps, the gradient assembly, the norm rescale, Adam and the DDIM step of
 * the current step index; read the results with mdc_get_state / mdc_dbg_read_x_adam / mdc_dbg_read_buffer. */
int mdc_dbg_loss(mdc_handle* h, const float* dec_nchw, float* ddec_nchw, float* loss_host, float* sgrad_host,
                 float* tgrad_host);
int mdc_dbg_update(mdc_handle* h, const float* v_nchw, const float* dz_nchw, const float* dunet_in_nchw);
/* Teacher forcing at any step (after mdc_begin / mdc_begin_frame): latent x and its Adam moments exp_avg / exp_avg_sq
 * ([N,4,EH,EW] bf16, device; NULL keeps the current one), per-sample scale, shift and their fp32 Adam moments (host,
 * 6 x N floats: scale, shift, s_m, s_v, t_m, t_v; NULL keeps them), and the step index the next mdc_run(h, 1) executes. */
int mdc_dbg_set_state(mdc_handle* h, int step, const void* x_bf16, const void* m1_bf16, const void* m2_bf16,
                      const float* affine6_host);
/* In-situ timing of the dominant kernel: replays the launch sequence of one guided step (state is NOT advanced
 * meaningfully: call after mdc_begin, before/after mdc_run) with CUDA events around every tcgen05 GEMM / conv launch;
 * returns the summed kernel time (ms), the summed algorithmic FLOPs and the number of such launches. */
int mdc_dbg_profile_gemm_step(mdc_handle* h, float* ms_host, double* flops_host, int* launches_host);
/* Per-op timing report (CSV: tape, index, op name, fwd us, bwd us, GEMM GFLOP, launches): every op of both tapes is
 * replayed `iters` times back to back (warm caches: use for ranking, not for absolute step time). */
int mdc_dbg_profile_ops(mdc_handle* h, const char* csv_path, int iters);
/* Per-tape timing: runs the forward (and backward) tapes `iters` times, returns ms per pass. */
int mdc_dbg_time_tapes(mdc_handle* h, int iters, float* ms_host /* [4]: unet fwd, unet bwd, dec fwd, dec bwd */);
/* What mdc_begin_frame derived from the sparse depth: guide [N,1,H,W] fp32 and mask [N,1,H,W] uint8 (device pointers,
 * may be NULL) and per sample {lo, hi, guide min, guide max, number of valid points} (host, 5 floats each). */
int mdc_dbg_frame_state(mdc_handle* h, float* guide_dev, unsigned char* mask_dev, float* stats_host);
/* Planner overrides for tests and tuning sweeps (0 = automatic): output-tile width BN, cluster size (B multicast),
 * split-K factor (> 0 forced, < 0 = the engine's cost model also in mdc_dbg_conv3x3), and the number of weight copies
 * mdc_dbg_conv3x3 rotates through in its timed loop (so weights stream from HBM as in the real step). */
int mdc_dbg_tune(int bn, int cs, int ksplit, int wcopies);
/* Row-shared-taps mode of the 3x3 convolution (one 130-pixel A box per kernel row instead of one box per tap):
 * 0 automatic (wide images, <= 128 output channels per tile), 1 off, 2 on whenever the tile shape allows it. */
int mdc_dbg_tune_rowshare(int mode);

#ifdef __cplusplus
}
#endif
#endif
