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
/* Per-tape timing: runs the forward (and backward) tapes `iters` times, returns ms per pass. */
int mdc_dbg_time_tapes(mdc_handle* h, int iters, float* ms_host /* [4]: unet fwd, unet bwd, dec fwd, dec bwd */);

#ifdef __cplusplus
}
#endif
#endif
