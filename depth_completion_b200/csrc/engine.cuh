// The step engine: a static tape of layer ops for the SD2-derived depth UNet and the SD2 VAE decoder, each with a
// hand-written forward and input-gradient backward (weights are frozen, so no weight gradients exist -- SURVEY.md G8).
//
// Design (B200-first, not a translation of autograd):
//   * every activation is NHWC bf16 and lives for the whole handle lifetime in HBM (180 GB: no recompute, no
//     allocator on the hot path); all GEMM/conv launch plans (TMA descriptors included) are built once;
//   * skip-connection concats are zero-copy: producers write straight into channel slices of the concat buffer;
//   * residual adds, biases and time-embedding adds ride in the GEMM/conv epilogue; gradient accumulation at
//     fan-out points rides in the consumer kernels (`acc` flag resolved at plan time);
//   * only the saved tensors the input-gradient path needs are kept: GroupNorm/LayerNorm inputs + statistics,
//     attention probabilities, GEGLU pre-activations.
#pragma once
#include <cmath>
#include <deque>
#include <functional>
#include <map>
#include <memory>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/mdc.h"
#include "attn.cuh"
#include "gemm.cuh"
#include "head.cuh"
#include "norm.cuh"
#include "pack.cuh"
#include "tail.cuh"

namespace mdc {

// ------------------------------------------------------------------------------------------------ memory
struct Arena {
  std::vector<void*> blocks;
  size_t total = 0;
  void* alloc(size_t bytes) {
    void* p = nullptr;
    bytes = (bytes + 255) & ~size_t(255);
    MDC_CUDA(cudaMalloc(&p, bytes));
    MDC_CUDA(cudaMemset(p, 0, bytes));
    blocks.push_back(p);
    total += bytes;
    return p;
  }
  template <typename T>
  T* make(size_t n) {
    return static_cast<T*>(alloc(n * sizeof(T)));
  }
  void free_all() {
    for (void* p : blocks) cudaFree(p);
    blocks.clear();
    total = 0;
  }
  ~Arena() { free_all(); }
};

struct Tensor {
  bf16* d = nullptr;
  bf16* g = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  long long ld = 0;
  bool grad_set = false;
  bool is_view = false;   // channel slice of a concat buffer: its gradient must stay inside the parent's
  bool relu_out = false;  // produced by a conv + ReLU epilogue: its gradient is masked by (value > 0) once complete
  bool mask_owned = false;  // ... by the backward of its first forward consumer (the last one to contribute)
  struct GemmPlan* producer = nullptr;  // forward plan of the conv that writes this tensor (GroupNorm statistics in its epilogue)
  bool producer_gn = false;             // ... already claimed by a GroupNorm
  std::string name;
  long long rows() const { return 1LL * n * h * w; }
};

struct Op {
  std::string name;
  // A resnet's 1x1 conv_shortcut only depends on the block input: it runs on the engine's side stream beside norm1 ->
  // conv1 -> norm2 (forward) and beside norm2 <- conv1 <- ... (backward).  `side` marks that op; the block's norm1 carries
  // `fork_fwd` (record the fork event before it is launched) and `join_bwd` (its backward accumulates into the gradient
  // the shortcut's backward writes first: wait for the side stream before launching it).
  bool side = false, fork_fwd = false, join_bwd = false;
  virtual ~Op() {}
  virtual void plan_bwd() {}
  virtual void fwd(cudaStream_t) = 0;
  virtual void bwd(cudaStream_t) {}
  virtual int n_fwd() const { return 1; }  // kernel launches per forward / backward call
  virtual int n_bwd() const { return 1; }
  virtual void gemm_plans(std::vector<const GemmPlan*>& f, std::vector<const GemmPlan*>& b) const {}
};

enum WKind { W_CONV3 = 0, W_LIN = 1, W_VEC = 2, W_UPCONV = 3 };
struct WeightSlot {
  WKind kind;
  int out = 0, in = 0;          // logical dims ([out, in, 3, 3] / [out, in] / [out])
  bf16* w = nullptr;            // CONV3: fwd pack; LIN: [out rows at w][ld_w]
  long long ld_w = 0;
  bf16* wt = nullptr;           // CONV3: dgrad pack; LIN: transposed pack (column offset applied), ld_wt
  long long ld_wt = 0;
  float* vec = nullptr;         // VEC destination
  float vscale = 1.f, vshift = 0.f;  // VEC: stored as vscale * value + vshift (bias of a scaled / shifted epilogue)
  bool loaded = false;
};

// Packed weights live in a bank that several engines (same model configuration and device, different N / H / W /
// resolution / steps) share: a new frame geometry builds new tapes and workspace but never re-packs the 1.8 GB of
// parameters (mdc_create_shared).  Engines request their slots in the same deterministic order, so the i-th request for
// a key maps to the i-th slot the first engine created for it.
struct WeightBank {
  Arena arena;
  std::deque<WeightSlot> slots;
  std::map<std::string, std::vector<WeightSlot*>> wmap;
  std::map<std::string, std::pair<void*, size_t>> blobs;  // raw named buffers (fused q/k/v matrices)
  int device = 0;
  int model_sig[64];  // the mdc_config fields that determine the parameter set
  void* blob(const std::string& name, size_t bytes) {
    auto it = blobs.find(name);
    if (it != blobs.end()) {
      MDC_CHECK(it->second.second == bytes, "weight bank: blob '%s' has %zu bytes, requested %zu", name.c_str(), it->second.second, bytes);
      return it->second.first;
    }
    void* p = arena.alloc(bytes);
    blobs[name] = {p, bytes};
    return p;
  }
};

struct Engine;

// ------------------------------------------------------------------------------------------------ ops
struct ConvOp : Op {  // 3x3 stride-1 pad-1 convolution (+bias, +residual)
  Engine* E;
  Tensor *x, *y, *res;
  WeightSlot* W;
  const float* bias;
  GemmPlan pf, pb;
  bool acc_res = false, alias_res = false;
  float alpha = 1.f;     // y = alpha * conv(x) + bias (+ res); the input gradient carries the same factor
  bool mask_x = false;   // x is a ReLU output and this op is its first consumer: mask x's finished gradient by (x > 0)
  void plan_bwd() override;
  void fwd(cudaStream_t st) override;
  void bwd(cudaStream_t st) override;
  int n_bwd() const override { return ((res && !alias_res) ? 2 : 1) + (mask_x ? 1 : 0); }
  void gemm_plans(std::vector<const GemmPlan*>& f, std::vector<const GemmPlan*>& b) const override {
    f.push_back(&pf), b.push_back(&pb);
  }
};
struct LinearOp : Op {  // y[rows, out] = x[rows, in] W^T (+bias, +residual)
  Engine* E;
  Tensor *x, *y, *res;
  bf16 *w, *wt;
  long long ld_w, ld_wt;
  const float* bias;
  GemmPlan pf, pb;
  bool acc_res = false, alias_res = false;
  void plan_bwd() override;
  void fwd(cudaStream_t st) override { run_gemm(pf, st); }
  void bwd(cudaStream_t st) override;
  int n_bwd() const override { return (res && !alias_res) ? 2 : 1; }
  void gemm_plans(std::vector<const GemmPlan*>& f, std::vector<const GemmPlan*>& b) const override {
    f.push_back(&pf), b.push_back(&pb);
  }
};
struct GroupNormOp : Op {
  Engine* E;
  Tensor *x, *y;
  const float *gamma, *beta;
  float eps;
  int groups, silu;
  float* stats;
  GNPlan plan;          // two-pass kernels, or the single-launch variants when the tensor is small enough (norm.cuh)
  bool acc = false;
  GemmPlan* epi = nullptr;  // the producing conv whose epilogue emits this op's statistics partials (two-pass path only)
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override;
  void bwd(cudaStream_t st) override;
  int n_fwd() const override { return plan.single_f() ? 1 : (epi ? 1 : 2); }
  int n_bwd() const override { return plan.single_b() ? 1 : 2; }
};
struct LayerNormOp : Op {
  Tensor *x, *y;
  const float *gamma, *beta;
  float* stats;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    int rows = static_cast<int>(x->rows());
    launch_k(ln_fwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, x->d, x->ld, rows, x->c, gamma, beta, 1e-5f, y->d, y->ld, stats);
  }
  void bwd(cudaStream_t st) override {
    int rows = static_cast<int>(x->rows());
    launch_k(ln_bwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, x->d, x->ld, y->g, y->ld, rows, x->c, gamma, stats, x->g, x->ld, acc);
  }
};
struct GegluOp : Op {
  Tensor *x, *y;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    launch_k(geglu_fwd_kernel, dim3(ew_grid(x->rows() * (y->c / 8))), dim3(256), 0, st, x->d, x->ld, x->rows(), y->c, y->d, y->ld);
  }
  void bwd(cudaStream_t st) override {
    launch_k(geglu_bwd_kernel, dim3(ew_grid(x->rows() * (y->c / 8))), dim3(256), 0, st, x->d, x->ld, y->g, y->ld, x->rows(), y->c, x->g,
                                                                     x->ld, acc);
  }
};
struct SelfAttnOp : Op {  // softmax(q k^T / sqrt(dh)) v over the tokens of each image; qkv [rows, 3*d] fused (attn.cuh)
  Engine* E;
  Tensor *qkv, *o;
  AttnPlan a;
  void plan_bwd() override { qkv->grad_set = true; }  // q, k, v gradient slices are each written exactly once
  void fwd(cudaStream_t st) override { run_attention_fwd(a, st); }
  void bwd(cudaStream_t st) override;
  int n_fwd() const override { return a.use_flash ? 1 : 3; }
  int n_bwd() const override { return a.use_flash ? 3 : 5; }
  void gemm_plans(std::vector<const GemmPlan*>& f, std::vector<const GemmPlan*>& b) const override {
    if (a.use_flash) return;
    f.push_back(&a.p_s), f.push_back(&a.p_o);
    if (qkv->g) b.push_back(&a.p_dv), b.push_back(&a.p_dp), b.push_back(&a.p_dq), b.push_back(&a.p_dk);
  }
};
struct CrossAttn2Op : Op {  // attention over the 2 tokens of the empty-prompt embedding (step-invariant K, V)
  Tensor *q, *o;
  int heads;
  const float *kc, *vc;
  bool acc = false;
  void plan_bwd() override {
    acc = q->grad_set;
    q->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    long long warps = q->rows() * heads;
    launch_k(xattn2_fwd_kernel, dim3(static_cast<int>((warps * 32 + 255) / 256)), dim3(256), 0, st, q->d, q->ld, q->rows(), heads, kc, vc,
                                                                                 0.125f, o->d, o->ld);
  }
  void bwd(cudaStream_t st) override {
    long long warps = q->rows() * heads;
    launch_k(xattn2_bwd_kernel, dim3(static_cast<int>((warps * 32 + 255) / 256)), dim3(256), 0, st, q->d, q->ld, o->g, o->ld, q->rows(),
                                                                                 heads, kc, vc, 0.125f, q->g, q->ld, acc);
  }
};
struct CrossAttnFusedOp : Op {  // LN2 + attn2 (2 constant key tokens) + to_out + residual, algebraically collapsed: ONE kernel each way
  Tensor *x, *y;
  int heads;
  const float *gamma, *beta, *At, *U, *bo;
  float* stats;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  // rows per CTA: enough CTAs to fill the chip on the small maps, longer reuse of At / U on the large ones
  int rows_per_block() const {
    const long long rows = x->rows();
    return rows >= 1500 ? 8 : (rows >= 400 ? 4 : 2);
  }
  template <int RB, int K4>
  void run(cudaStream_t st, bool backward) {
    const int rows = static_cast<int>(x->rows()), d = x->c;
    const dim3 grid(static_cast<unsigned>((rows + RB - 1) / RB)), block(XB_THREADS);
    if (!backward) {
      const size_t smem = (static_cast<size_t>(RB) * d + RB * XB_LDS) * sizeof(float);
      launch_k(xattn_block_fwd_kernel<RB, K4>, grid, block, smem, st, x->d, x->ld, rows, d, 2 * heads, gamma, beta, At, U, bo, y->d, y->ld,
               stats);
    } else {
      const size_t smem = (3 * static_cast<size_t>(RB) * d + 2 * RB * XB_LDS) * sizeof(float);
      launch_k(xattn_block_bwd_kernel<RB, K4>, grid, block, smem, st, x->d, x->ld, y->g, y->ld, rows, d, 2 * heads, gamma, beta, At, U,
               static_cast<const float*>(stats), x->g, x->ld, static_cast<int>(acc));
    }
  }
  template <int RB>
  void run_k(cudaStream_t st, bool backward) {  // loop bounds sized for the width class: d <= 384 / 640 / 1280
    if (x->c <= 384)
      run<RB, 3>(st, backward);
    else if (x->c <= 640)
      run<RB, 5>(st, backward);
    else
      run<RB, 10>(st, backward);
  }
  void dispatch(cudaStream_t st, bool backward) {
    switch (rows_per_block()) {
      case 8: run_k<8>(st, backward); break;
      case 4: run_k<4>(st, backward); break;
      default: run_k<2>(st, backward); break;
    }
  }
  void fwd(cudaStream_t st) override { dispatch(st, false); }
  void bwd(cudaStream_t st) override { dispatch(st, true); }
};
struct UpsampleOp : Op {
  Tensor *x, *y;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    launch_k(upsample_nearest_fwd_kernel, dim3(ew_grid(y->rows() * (y->c / 8))), dim3(256), 0, st, x->d, x->ld, x->n, x->h, x->w, x->c,
                                                                                y->d, y->ld, y->h, y->w);
  }
  void bwd(cudaStream_t st) override {
    launch_k(upsample_nearest_bwd_kernel, dim3(ew_grid(x->rows() * (x->c / 8))), dim3(256), 0, st, y->g, y->ld, x->n, x->h, x->w, x->c,
                                                                                x->g, x->ld, y->h, y->w, acc);
  }
};
struct UpConvOp : Op {  // fused nearest-2x upsample + conv3x3 (four 2x2 phase convolutions on the low-res input)
  Engine* E;
  Tensor *x, *y;
  WeightSlot* W;
  const float* bias;
  GemmPlan pf, pb;
  bool mask_x = false;  // see ConvOp::mask_x
  void plan_bwd() override;
  void fwd(cudaStream_t st) override { run_gemm(pf, st); }
  void bwd(cudaStream_t st) override;
  int n_bwd() const override { return mask_x ? 2 : 1; }
  void gemm_plans(std::vector<const GemmPlan*>& f, std::vector<const GemmPlan*>& b) const override {
    f.push_back(&pf), b.push_back(&pb);
  }
};
struct TanhClampOp : Op {  // DecoderTiny: y = tanh(x / m) * m
  Tensor *x, *y;
  float m;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    const long long tot = x->rows() * x->c;
    launch_k(tanh_clamp_fwd_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, st, x->d, x->ld, y->d, y->ld, x->rows(), x->c, m);
  }
  void bwd(cudaStream_t st) override {
    const long long tot = x->rows() * x->c;
    launch_k(tanh_clamp_bwd_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, st, x->d, x->ld, y->g, y->ld, x->g, x->ld, x->rows(),
             x->c, m, acc ? 1 : 0);
  }
};
struct UnitRangeOp : Op {  // EncoderTiny: y = (x + 1) / 2 (forward only)
  Tensor *x, *y;
  void fwd(cudaStream_t st) override {
    const long long tot = x->rows() * x->c;
    launch_k(unit_range_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, st, x->d, x->ld, y->d, y->ld, x->rows(), x->c);
  }
  void bwd(cudaStream_t) override {}
  int n_bwd() const override { return 0; }
};
struct SubsampleOp : Op {
  Tensor *x, *y;
  int off;
  bool acc = false;
  void plan_bwd() override {
    acc = x->grad_set;
    x->grad_set = true;
  }
  void fwd(cudaStream_t st) override {
    launch_k(subsample2_fwd_kernel, dim3(ew_grid(y->rows() * (y->c / 8))), dim3(256), 0, st, x->d, x->ld, x->n, x->h, x->w, x->c, off, y->d,
                                                                          y->ld, y->h, y->w);
  }
  void bwd(cudaStream_t st) override {
    launch_k(subsample2_bwd_kernel, dim3(ew_grid(x->rows() * (x->c / 8))), dim3(256), 0, st, y->g, y->ld, x->n, x->h, x->w, x->c, off, x->g,
                                                                          x->ld, y->h, y->w, acc);
  }
};
struct ConcatOp : Op {  // zero-copy: a and b are channel-slice views of cat; only propagates the plan-time flags
  Tensor *a, *b, *cat;
  void plan_bwd() override { a->grad_set = b->grad_set = cat->grad_set; }
  void fwd(cudaStream_t) override {}
  int n_fwd() const override { return 0; }
  int n_bwd() const override { return 0; }
};

// ------------------------------------------------------------------------------------------------ engine
struct Engine {
  mdc_config cfg;
  Arena arena;
  std::deque<Tensor> tensors;
  std::map<std::string, Tensor*> named;
  std::shared_ptr<WeightBank> bank;              // packed parameters, possibly shared with other engines
  std::map<std::string, std::vector<WeightSlot*>>& wmap() { return bank->wmap; }
  std::map<std::string, size_t> slot_cursor;     // how many slots of a key this engine has requested so far
  WeightSlot* bank_slot(const std::string& key, WKind kind, int out, int in, bool& fresh);
  std::vector<std::unique_ptr<Op>> unet_ops, dec_ops, enc_ops;  // enc_ops: VAE encoder, forward only (per-frame prologue)
  bool alloc_grads = true;
  // Sparse output head (head.cuh): inside a guided step the KL decoder's conv_norm_out / conv_out are evaluated only
  // where the loss looks.  `sparse_head` is raised around the decoder tape of step_launches() and read by the two ops.
  struct GroupNormOp* head_gn = nullptr;
  struct ConvOp* head_conv = nullptr;
  unsigned char* head_act = nullptr;  // [N, PPH*PPW]: pixels that receive a gradient from conv_out in the current step
  bool sparse_head = false;
  bool head_points_ok = true;  // set per frame in begin_state(): the taps' 3x3 neighbourhoods are a small part of the map
  bool sparse_head_allowed() const;
  void head_fwd(cudaStream_t st);
  void head_bwd(cudaStream_t st);
  size_t n_split_step = 0;
  Tensor *enc_in = nullptr, *enc_out = nullptr;
  std::vector<std::unique_ptr<Op>>* cur_ops = nullptr;
  // `stream` is where every launch of the public entry points is enqueued: the engine's own stream by default, or the
  // caller's (mdc_set_stream; the Python binding passes torch's current stream on every call).  Graphs are captured
  // on the private `cap_stream` (capture is not allowed on the legacy default stream) and launched on `stream`.
  cudaStream_t stream = 0, own_stream = 0, cap_stream = 0;
  cudaStream_t side_stream = 0;  // second branch of the step graph (Op::side); joins `stream` before anything reads its results
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;    // resnet shortcut branch
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;  // attention backward (dQ beside dK/dV)
  float* split_ws_side = nullptr;  // split-K workspace of the side-stream GEMMs (they overlap main-stream split-K GEMMs)
  void set_stream(cudaStream_t s) { stream = s ? s : own_stream; }
  // stream-ordered copy + synchronisation (a plain cudaMemcpy runs on the legacy default stream, which is NOT ordered
  // with a caller-provided non-blocking stream)
  void copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    MDC_CUDA(cudaMemcpyAsync(dst, src, bytes, kind, stream));
    MDC_CUDA(cudaStreamSynchronize(stream));
  }
  void activate() {
    MDC_CHECK(!released, "this handle's workspace was released (mdc_release_workspace): it only keeps the packed weights alive");
    MDC_CUDA(cudaSetDevice(cfg.device));
  }  // every extern "C" entry: the handle's device becomes current
  void check_barrier_flag();

  // geometry
  int N, H, W, ph, pw, PPH, PPW, lh, lw;
  // graph endpoints
  Tensor *unet_in = nullptr, *unet_out = nullptr, *dec_in = nullptr, *dec_out = nullptr;
  // shared scratch
  float* gn_partial = nullptr;
  size_t gn_partial_floats = 0;
  float* gn_gstats = nullptr;
  unsigned int* gn_ticket = nullptr;
  unsigned int* gn_bar = nullptr;  // grid barrier of the single-launch GroupNorm kernels (count, generation, timeout flag)
  GNScratch gn_scratch() const { return GNScratch{gn_partial, gn_gstats, gn_ticket, gn_bar}; }
  float* gn_epi_partial = nullptr;       // statistics rows written by conv epilogues (GemmParams::gn_partial), one scratch for all
  size_t gn_epi_floats = 0;
  std::vector<GemmPlan*> gn_epi_plans;
  float* attn_S = nullptr;
  size_t attn_S_floats = 0;
  // time embedding
  struct TembUse {
    WeightSlot* proj_w;
    WeightSlot* proj_b;
    WeightSlot* conv_b;
    int off, cout;
  };
  std::vector<TembUse> temb_uses;
  int temb_total = 0;
  float* temb_cur = nullptr;
  float* temb_table = nullptr;
  WeightSlot *te_l1w = nullptr, *te_l1b = nullptr, *te_l2w = nullptr, *te_l2b = nullptr;
  // cross attention K/V
  struct XUse {
    WeightSlot *wk, *wv;
    float *kc, *vc;
    int d;
    // collapsed form (CrossAttnFusedOp): At / U built from to_q / to_out.0 in prepare()
    WeightSlot *wq = nullptr, *wo = nullptr;
    float *At = nullptr, *U = nullptr;
    int heads = 0;
  };
  std::vector<XUse> xuses;
  // step state
  StepTables tables{};
  StepCur* cur = nullptr;
  StepAccum* accum = nullptr;
  int* counter = nullptr;
  float *d_sqrt_a = nullptr, *d_sqrt_1ma = nullptr, *d_sqrt_ap = nullptr, *d_sqrt_1map = nullptr;
  bf16 *x = nullptr, *m1 = nullptr, *m2 = nullptr, *img_lat = nullptr, *x_adam_dbg = nullptr;
  float *dx_direct = nullptr, *gbuf = nullptr, *eps_part = nullptr, *g_part = nullptr, *dmean = nullptr;
  int parts_per_img = 1;
  int* pt_idx = nullptr;
  float* pt_val = nullptr;
  int* pt_off = nullptr;
  int* pt_cnt = nullptr;            // valid points per sample (device)
  StepCur* final_cur = nullptr;     // identity DDIM scalars for the final decode (x0 = x)
  float* final_scratch = nullptr;
  float *gminmax = nullptr, *depth_minmax = nullptr;
  float lr_x = 0.05f, lr_s = 0.005f;
  TailOpts h_opts{0, 0, 0, 0, 0.1f, 1.f, 1.f, 0.f, 0.f, 0.05f, 0.005f, 0};  // defaults of marigold_dc.py:467-493; mdc_set_options
  TailOpts* opts = nullptr;                                               // device copy, refreshed by begin()
  float *x1_part = nullptr, *x2_part = nullptr;                           // partial sums of x, x^2 (kld)
  float *dn_map = nullptr, *gray_gx = nullptr, *gray_gy = nullptr;        // edge / smooth losses: dense map, image gradients
  float *pt_a = nullptr, *pt_G = nullptr;                                 // closed-form loss: per-point scratch
  bool have_gray = false;
  int interp_nearest = 0;  // baked into the step graph through TailGeom: changing it drops the captured graph
  float q_lo = 0.01f, q_hi = 0.99f;                                       // norm="percentile"
  float *pc_vals = nullptr, *pc_sorted = nullptr, *pc_range = nullptr;
  int* pc_counts = nullptr;
  void* pc_tmp = nullptr;
  size_t pc_tmp_bytes = 0;
  void percentile_ranges(const float* sparse);
  void set_options(int projection, int inv, int opt, const float* loss_weights4, int kld_mode, float kld_weight, float qlo, float qhi,
                   int closed_form, int nearest);
  bool prepared = false, begun = false;
  bool released = false;  // mdc_release_workspace: tapes and workspace freed, the handle only keeps the weight bank alive
  void release_workspace();
  int steps_done = 0;
  long long launches = 0;

  explicit Engine(const mdc_config& c, std::shared_ptr<WeightBank> share = nullptr);
  // builders
  Tensor* new_tensor(int n, int h, int w, int c, const std::string& name, bool with_grad = true, long long ld = 0);
  Tensor* view(Tensor* parent, int c0, int c, const std::string& name);
  WeightSlot* slot(const std::string& key, WKind kind, int out, int in);
  WeightSlot* slot_lin_into(const std::string& key, int out, int in, bf16* w, long long ld_w, bf16* wt, long long ld_wt);
  WeightSlot* slot_vec_into(const std::string& key, int n, float* dst);
  void push(Op* op, const std::string& name) {
    op->name = name;
    cur_ops->emplace_back(op);
  }
  Tensor* conv3x3(Tensor* x, int cout, const std::string& key, Tensor* res = nullptr, Tensor* out = nullptr,
                  const std::string& temb_key = "");
  Tensor* linear(Tensor* x, int cout, const std::string& key, bool has_bias, Tensor* res = nullptr, Tensor* out = nullptr);
  Tensor* group_norm(Tensor* x, const std::string& key, float eps, bool silu);
  Tensor* layer_norm(Tensor* x, const std::string& key);
  Tensor* resnet(Tensor* x, int cout, const std::string& key, bool temb, float eps, Tensor* out = nullptr);
  Tensor* transformer(Tensor* x, int heads, const std::string& key, Tensor* out = nullptr);
  Tensor* self_attention(Tensor* qkv, int heads, const std::string& name);
  Tensor* upsample_conv(Tensor* x, int H2, int W2, const std::string& key, Tensor* out = nullptr);
  Tensor* downsample_conv(Tensor* x, const std::string& key, Tensor* out = nullptr);
  void build_unet();
  void build_decoder();
  void finalize_plans();
  // runtime
  void set_weight(const std::string& key, const void* src, const long long* shape, int ndim, int dtype);
  void set_weights(int n, const char* const* keys, const void* const* srcs, const long long* shapes4, const int* ndims, const int* dtypes);
  bool weights_loaded();
  void prepare(const void* ctx_bf16, const float* alphas_cumprod, const int* timesteps, int n_steps);
  void begin(const void* img_latents, const void* x0, const float* guide, const uint8_t* mask, const float* gmm,
             const float* dmm, float lrx, float lrs);
  void begin_state(const void* img_latents, const void* x0, const float* guide, const uint8_t* mask, float lrx, float lrs,
                   const float* stats5_dev);
  void run_ops(std::vector<std::unique_ptr<Op>>& ops, bool backward);
  void step_launches();  // the fixed launch sequence of one guided step
  void step();           // replays it as a CUDA graph (captured once; every step-dependent value lives in device memory)
  cudaGraphExec_t step_graph = nullptr;
  cudaGraphExec_t sample_graph = nullptr;  // the no-grad DDIM step (train_latents=False)
  void sample_launches();
  void sample_step();
  bool use_graph = true;
  ~Engine();
  void decode_final(float* dense_out, bool closed_form = false);
  void build_encoder();
  void build_tiny_decoder();
  void build_tiny_encoder();
  Tensor* tiny_conv(Tensor* x, int cout, const std::string& key, bool has_bias, bool relu, Tensor* res = nullptr, float alpha = 1.f,
                    float bias_shift = 0.f);
  Tensor* tiny_block(Tensor* x, const std::string& key);
  Tensor* tiny_upconv(Tensor* x, const std::string& key);
  Tensor* vae_mid_attention(Tensor* h, const std::string& A);
  void encode(const void* imgs, int dtype, int channels, void* latents_out);
  void begin_frame(const void* imgs, int dtype, int channels, const float* sparse, const void* x0, float max_depth,
                   float min_depth, int norm_mode, float lrx, float lrs, const void* img_latents = nullptr);
  float* fr_guide = nullptr;     // normalised sparse depth of the current frame
  uint8_t* fr_mask = nullptr;
  float* fr_stats = nullptr;     // per sample: lo, hi, guide min, guide max, valid count
  void read_tensor(const std::string& name, int which, float* out_nchw);
  long long launches_per_step = 0;
  long long launches_per_sample_step = 3;  // no-grad step: step tables + UNet input + UNet forward ops + DDIM update
  // split-K: plans with few output tiles and a long K loop share one fp32 partial-sum workspace
  std::vector<GemmPlan*> split_plans;
  size_t split_ws_floats = 0;
  float* split_ws = nullptr;
  void maybe_split(GemmPlan& g) {
    if (getenv("MDC_NO_SPLITK")) return;
    size_t fl = enable_splitk(g, choose_ksplit(g));
    if (fl) {
      split_ws_floats = std::max(split_ws_floats, fl);
      split_plans.push_back(&g);
    }
  }
  // Times every tcgen05 GEMM / conv launch of one guided step in situ (CUDA events around each launch).
  void profile_gemm_step(float* ms_out, double* flops_out, int* launches_out);
};

// ================================================================================================ op bodies
inline void ConvOp::fwd(cudaStream_t st) {
  if (E->sparse_head && this == E->head_conv) {
    E->head_fwd(st);
    return;
  }
  run_gemm(pf, st);
}
inline void ConvOp::plan_bwd() {
  Epilogue e;
  e.out = x->g, e.ldc = x->ld, e.alpha = alpha;
  if (x->grad_set) e.res = x->g, e.ldr = x->ld;
  x->grad_set = true;
  pb = plan_conv3x3(y->n, y->h, y->w, y->c, x->c, y->g, y->ld, W->wt, e);
  E->maybe_split(pb);
  if (res) {
    acc_res = res->grad_set;
    res->grad_set = true;
    // First gradient contribution of a stand-alone residual tensor: d(res) == dy, so let res share dy's buffer instead
    // of copying it (dy is dead once this op's own input gradient has been computed, which happens first).
    if (!acc_res && !res->is_view && !y->is_view && res->ld == y->ld && res->c == y->c && !getenv("MDC_NO_ALIAS")) {
      res->g = y->g;
      alias_res = true;
    }
  }
}
inline void relu_mask(Tensor* x, cudaStream_t st) {
  launch_k(relu_mask_kernel, dim3(ew_grid(x->rows() * (x->c / 8))), dim3(256), 0, st, x->g, x->ld, x->d, x->ld, x->rows(), x->c);
}
inline void ConvOp::bwd(cudaStream_t st) {
  if (E->sparse_head && this == E->head_conv) return;  // folded into the GroupNorm backward of the sparse head
  run_gemm(pb, st);
  if (res && !alias_res)
    launch_k(add_rows_kernel, dim3(ew_grid(res->rows() * (res->c / 8))), dim3(256), 0, st, y->g, y->ld, res->g, res->ld, res->rows(), res->c,
                                                                        acc_res);
  if (mask_x) relu_mask(x, st);
}
inline void UpConvOp::plan_bwd() {
  Epilogue e;
  e.out = x->g, e.ldc = x->ld;
  if (x->grad_set) e.res = x->g, e.ldr = x->ld;
  x->grad_set = true;
  pb = plan_upconv_bwd(x->n, x->h, x->w, x->c, y->c, y->g, y->ld, W->wt, e);
  E->maybe_split(pb);  // 16 taps x C/64 k-chunks on a low-resolution map: few tiles, a very long K loop
}
inline void UpConvOp::bwd(cudaStream_t st) {
  run_gemm(pb, st);
  if (mask_x) relu_mask(x, st);
}
inline void LinearOp::plan_bwd() {
  Epilogue e;
  e.out = x->g, e.ldc = x->ld;
  if (x->grad_set) e.res = x->g, e.ldr = x->ld;
  x->grad_set = true;
  Operand A{y->g, 0, y->ld, 0, 0}, B{wt, 0, ld_wt, 0, 0};
  pb = plan_gemm(static_cast<int>(x->rows()), x->c, y->c, A, B, e);
  E->maybe_split(pb);
  if (res) {
    acc_res = res->grad_set;
    res->grad_set = true;
    // First gradient contribution of a stand-alone residual tensor: d(res) == dy, so let res share dy's buffer instead
    // of copying it (dy is dead once this op's own input gradient has been computed, which happens first).
    if (!acc_res && !res->is_view && !y->is_view && res->ld == y->ld && res->c == y->c && !getenv("MDC_NO_ALIAS")) {
      res->g = y->g;
      alias_res = true;
    }
  }
}
inline void LinearOp::bwd(cudaStream_t st) {
  run_gemm(pb, st);
  if (res && !alias_res)
    launch_k(add_rows_kernel, dim3(ew_grid(res->rows() * (res->c / 8))), dim3(256), 0, st, y->g, y->ld, res->g, res->ld, res->rows(), res->c,
                                                                        acc_res);
}
inline void SelfAttnOp::bwd(cudaStream_t st) {
  static const bool no_side = getenv("MDC_NO_SIDE") != nullptr;
  SideBranch sb{E->side_stream, E->ev_fork2, E->ev_join2};
  run_attention_bwd(a, st, no_side || st == E->side_stream ? nullptr : &sb);
}
inline void GroupNormOp::fwd(cudaStream_t st) {
  if (E->sparse_head && this == E->head_gn) {  // statistics only; head_points_fwd_kernel normalises the pixels it needs
    run_gn_stats(plan, x->d, eps, stats, E->gn_scratch(), st);
    return;
  }
  run_gn_fwd(plan, x->d, y->d, y->ld, gamma, beta, eps, silu, stats, E->gn_scratch(), st, epi ? E->gn_epi_partial : nullptr,
             epi ? epi->grid : 0);
}
inline void GroupNormOp::bwd(cudaStream_t st) {
  if (E->sparse_head && this == E->head_gn) {
    E->head_bwd(st);
    return;
  }
  run_gn_bwd(plan, x->d, y->g, y->ld, gamma, beta, silu, stats, x->g, x->ld, acc, E->gn_scratch(), st);
}

// ================================================================================================ builders
inline Tensor* Engine::new_tensor(int n, int h, int w, int c, const std::string& name, bool with_grad, long long ld) {
  tensors.emplace_back();
  Tensor* t = &tensors.back();
  t->n = n, t->h = h, t->w = w, t->c = c;
  t->ld = ld ? ld : ((c + 7) / 8) * 8;
  t->name = name;
  size_t el = static_cast<size_t>(t->rows()) * t->ld + 64;
  t->d = arena.make<bf16>(el);
  if (with_grad && alloc_grads) t->g = arena.make<bf16>(el);
  if (!name.empty()) named[name] = t;
  return t;
}
inline Tensor* Engine::view(Tensor* parent, int c0, int c, const std::string& name) {
  tensors.emplace_back();
  Tensor* t = &tensors.back();
  *t = *parent;
  t->c = c;
  t->d = parent->d + c0;
  t->g = parent->g ? parent->g + c0 : nullptr;
  t->grad_set = false;
  t->is_view = true;
  t->name = name;
  if (!name.empty()) named[name] = t;
  return t;
}
// The i-th request of this engine for (`key`, kind) maps to the i-th slot of that kind the bank holds for the key
// (created on first use).  One parameter may be held in more than one layout: an upsampler's 3x3 weight is packed as
// four 2x2 phase kernels for an exact 2x upsample (W_UPCONV) and as a plain 3x3 (W_CONV3) for the explicit-size
// upsampling of odd latent sizes, depending on the frame geometry of the engine that asks.
inline WeightSlot* Engine::bank_slot(const std::string& key, WKind kind, int out, int in, bool& fresh) {
  auto& list = bank->wmap[key];
  const size_t want = slot_cursor[key + "#" + std::to_string(static_cast<int>(kind))]++;
  size_t seen = 0;
  for (WeightSlot* s : list) {
    if (s->kind != kind) continue;
    if (seen++ == want) {
      MDC_CHECK(s->out == out && s->in == in, "weight bank: '%s' was built as [%d, %d], now requested as [%d, %d]", key.c_str(), s->out,
                s->in, out, in);
      fresh = false;
      return s;
    }
  }
  fresh = true;
  bank->slots.emplace_back();
  WeightSlot* s = &bank->slots.back();
  s->kind = kind, s->out = out, s->in = in;
  list.push_back(s);
  return s;
}
inline WeightSlot* Engine::slot(const std::string& key, WKind kind, int out, int in) {
  bool fresh;
  WeightSlot* s = bank_slot(key, kind, out, in, fresh);
  if (!fresh) return s;
  Arena& wa = bank->arena;
  if (kind == W_CONV3 || kind == W_UPCONV) {
    const int inp = ((in + 63) / 64) * 64, outp = ((out + 63) / 64) * 64;
    const size_t taps = kind == W_CONV3 ? 9 : 16;
    s->w = wa.make<bf16>(taps * inp * out + 64);
    s->wt = wa.make<bf16>(taps * outp * in + 64);
  } else if (kind == W_LIN) {
    s->ld_w = ((in + 7) / 8) * 8, s->ld_wt = ((out + 7) / 8) * 8;  // TMA needs 16-byte row strides
    s->w = wa.make<bf16>(1ull * out * s->ld_w + 64);
    s->wt = wa.make<bf16>(1ull * in * s->ld_wt + 64);
  } else {
    s->vec = wa.make<float>(out + 16);
  }
  return s;
}
inline WeightSlot* Engine::slot_lin_into(const std::string& key, int out, int in, bf16* w, long long ld_w, bf16* wt,
                                         long long ld_wt) {
  bool fresh;
  WeightSlot* s = bank_slot(key, W_LIN, out, in, fresh);
  if (fresh) s->w = w, s->ld_w = ld_w, s->wt = wt, s->ld_wt = ld_wt;
  MDC_CHECK(s->w == w && s->wt == wt, "weight bank: '%s' points at another buffer", key.c_str());
  return s;
}
inline WeightSlot* Engine::slot_vec_into(const std::string& key, int n, float* dst) {
  bool fresh;
  WeightSlot* s = bank_slot(key, W_VEC, n, 0, fresh);
  if (fresh) s->vec = dst;
  MDC_CHECK(s->vec == dst, "weight bank: '%s' points at another buffer", key.c_str());
  return s;
}

inline Tensor* Engine::conv3x3(Tensor* x, int cout, const std::string& key, Tensor* res, Tensor* out,
                               const std::string& temb_key) {
  WeightSlot* W = slot(key + ".weight", W_CONV3, cout, x->c);
  WeightSlot* B = slot(key + ".bias", W_VEC, cout, 0);
  Tensor* y = out ? out : new_tensor(x->n, x->h, x->w, cout, "");
  auto* op = new ConvOp();
  op->E = this, op->x = x, op->y = y, op->res = res, op->W = W;
  op->bias = B->vec;
  if (!temb_key.empty()) {  // conv bias + time_emb_proj(silu(temb)) of the current step (filled by begin_step_kernel)
    TembUse u;
    u.proj_w = slot(temb_key + ".weight", W_LIN, cout, cfg.unet_block_ch[0] * 4);
    u.proj_b = slot(temb_key + ".bias", W_VEC, cout, 0);
    u.conv_b = B, u.off = temb_total, u.cout = cout;
    temb_total += ((cout + 3) / 4) * 4;
    temb_uses.push_back(u);
    op->bias = reinterpret_cast<const float*>(static_cast<uintptr_t>(u.off) + 1);  // patched in finalize_plans()
  }
  Epilogue e;
  e.out = y->d, e.ldc = y->ld, e.bias = op->bias;
  if (res) e.res = res->d, e.ldr = res->ld;
  op->pf = plan_conv3x3(x->n, x->h, x->w, x->c, cout, x->d, x->ld, W->w, e);
  maybe_split(op->pf);
  if (!y->is_view) y->producer = &op->pf;
  push(op, key);
  return y;
}
inline Tensor* Engine::linear(Tensor* x, int cout, const std::string& key, bool has_bias, Tensor* res, Tensor* out) {
  WeightSlot* W = slot(key + ".weight", W_LIN, cout, x->c);
  WeightSlot* B = has_bias ? slot(key + ".bias", W_VEC, cout, 0) : nullptr;
  Tensor* y = out ? out : new_tensor(x->n, x->h, x->w, cout, "");
  auto* op = new LinearOp();
  op->E = this, op->x = x, op->y = y, op->res = res;
  op->w = W->w, op->wt = W->wt, op->ld_w = W->ld_w, op->ld_wt = W->ld_wt;
  op->bias = B ? B->vec : nullptr;
  Epilogue e;
  e.out = y->d, e.ldc = y->ld, e.bias = op->bias;
  if (res) e.res = res->d, e.ldr = res->ld;
  Operand A{x->d, 0, x->ld, 0, 0}, Bm{W->w, 0, W->ld_w, 0, 0};
  op->pf = plan_gemm(static_cast<int>(x->rows()), cout, x->c, A, Bm, e);
  maybe_split(op->pf);
  push(op, key);
  return y;
}
inline Tensor* Engine::group_norm(Tensor* x, const std::string& key, float eps, bool silu) {
  const int G = cur_ops == &unet_ops ? cfg.unet_groups : cfg.vae_groups;
  WeightSlot* ga = slot(key + ".weight", W_VEC, x->c, 0);
  WeightSlot* be = slot(key + ".bias", W_VEC, x->c, 0);
  Tensor* y = new_tensor(x->n, x->h, x->w, x->c, "");
  auto* op = new GroupNormOp();
  op->E = this, op->x = x, op->y = y, op->gamma = ga->vec, op->beta = be->vec, op->eps = eps, op->groups = G;
  op->silu = silu ? 1 : 0;
  op->plan = plan_groupnorm(x->n, x->h * x->w, x->c, G, x->ld, cfg.concurrent ? 4 : 0);
  op->stats = arena.make<float>(2ull * x->n * G);
  gn_partial_floats = std::max<size_t>(gn_partial_floats, op->plan.partial_floats);
  // Large tensors (two-pass path): let the conv that produces x emit the statistics from its epilogue
  // (measured slower than the separate statistics kernel -- 22.10 vs 21.73 ms per step: the four epilogue warps are the
  // scarce resource of the 128-channel convolutions, DESIGN.md section 4 -- so it is opt-in: MDC_GNEPI=1)
  static const bool no_epi = getenv("MDC_GNEPI") == nullptr;
  const int cpg = x->c / G;
  if (!no_epi && !op->plan.single_f() && x->producer && !x->producer_gn && G == 32 && 32 % cpg == 0 && x->c % 32 == 0) {
    GemmPlan* g = x->producer;
    if (g->p.conv == 1 && g->p.ksplit <= 1 && g->p.N == x->c && g->p.BN % 32 == 0 && !g->p.out_f32) {
      g->p.gn_cpg = cpg, g->p.gn_nimg = x->n;
      finish_plan(*g);  // shared-memory layout changes (accumulator columns), possibly one pipeline stage fewer
      if (g->p.tma_store) {
        x->producer_gn = true;
        op->epi = g;
        gn_epi_floats = std::max<size_t>(gn_epi_floats, static_cast<size_t>(x->n) * g->grid * 64);
        gn_epi_plans.push_back(g);
      } else {  // the statistics ride on the TMA-store epilogue path only
        g->p.gn_cpg = 0;
        finish_plan(*g);
      }
    }
  }
  push(op, key);
  return y;
}
inline Tensor* Engine::layer_norm(Tensor* x, const std::string& key) {
  MDC_CHECK(x->c % 8 == 0 && x->c <= 32 * 8 * LN_MAXV, "LayerNorm: d=%d unsupported", x->c);
  WeightSlot* ga = slot(key + ".weight", W_VEC, x->c, 0);
  WeightSlot* be = slot(key + ".bias", W_VEC, x->c, 0);
  Tensor* y = new_tensor(x->n, x->h, x->w, x->c, "");
  auto* op = new LayerNormOp();
  op->x = x, op->y = y, op->gamma = ga->vec, op->beta = be->vec;
  op->stats = arena.make<float>(2ull * x->rows());
  push(op, key);
  return y;
}
inline Tensor* Engine::resnet(Tensor* x, int cout, const std::string& key, bool temb, float eps, Tensor* out) {
  Tensor* h = group_norm(x, key + ".norm1", eps, true);
  Op* norm1 = cur_ops->back().get();
  h = conv3x3(h, cout, key + ".conv1", nullptr, nullptr, temb ? key + ".time_emb_proj" : "");
  h = group_norm(h, key + ".norm2", eps, true);
  Tensor* sc = x;
  if (x->c != cout) {
    sc = linear(x, cout, key + ".conv_shortcut", true);
    static const bool no_side = getenv("MDC_NO_SIDE") != nullptr;
    if (!no_side) cur_ops->back()->side = true, norm1->fork_fwd = norm1->join_bwd = true;
  }
  Tensor* y = conv3x3(h, cout, key + ".conv2", sc, out);
  if (y->name.empty()) y->name = key, named[key] = y;
  return y;
}
inline Tensor* Engine::self_attention(Tensor* qkv, int heads, const std::string& name) {
  const int d = qkv->c / 3, dh = d / heads, T = qkv->h * qkv->w, n = qkv->n;
  Tensor* o = new_tensor(n, qkv->h, qkv->w, d, "");
  auto* op = new SelfAttnOp();
  op->E = this, op->qkv = qkv, op->o = o;
  const bool flash = dh == 64 && !getenv("MDC_NO_FLASH");
  const size_t pel = static_cast<size_t>(n) * heads * T * (((T + 7) / 8) * 8) + 64;
  float* lse2 = flash ? arena.make<float>(flash_stat_floats(n, heads, T)) : nullptr;
  float* delta = flash ? arena.make<float>(flash_stat_floats(n, heads, T)) : nullptr;
  bf16* P = flash ? nullptr : arena.make<bf16>(pel);
  if (!flash) attn_S_floats = std::max<size_t>(attn_S_floats, pel);
  // the fp32 score scratch is shared by all unfused attentions: allocated after the tapes are built (finalize_plans)
  op->a = plan_attention(n, T, heads, dh, qkv->d, qkv->ld, o->d, o->g, o->ld, qkv->g, qkv->ld, lse2, delta, P, nullptr, qkv->g != nullptr);
  push(op, name);
  return o;
}
inline Tensor* Engine::transformer(Tensor* x, int heads, const std::string& key, Tensor* out) {
  const int d = x->c;
  const std::string tb = key + ".transformer_blocks.0";
  Tensor* h = group_norm(x, key + ".norm", 1e-6f, false);
  h = linear(h, d, key + ".proj_in", true);
  // --- self attention (fused q/k/v projection, no bias)
  Tensor* n1 = layer_norm(h, tb + ".norm1");
  Tensor* qkv = new_tensor(x->n, x->h, x->w, 3 * d, "");
  {
    bf16* w = static_cast<bf16*>(bank->blob(tb + ".attn1.qkv.w", (3ull * d * d + 64) * 2));
    bf16* wt = static_cast<bf16*>(bank->blob(tb + ".attn1.qkv.wt", (3ull * d * d + 64) * 2));
    const char* nm[3] = {".attn1.to_q.weight", ".attn1.to_k.weight", ".attn1.to_v.weight"};
    for (int i = 0; i < 3; ++i) slot_lin_into(tb + nm[i], d, d, w + 1ull * i * d * d, d, wt + i * d, 3 * d);
    auto* op = new LinearOp();
    op->E = this, op->x = n1, op->y = qkv, op->res = nullptr, op->w = w, op->wt = wt, op->ld_w = d, op->ld_wt = 3 * d;
    op->bias = nullptr;
    Epilogue e;
    e.out = qkv->d, e.ldc = qkv->ld;
    Operand A{n1->d, 0, n1->ld, 0, 0}, Bm{w, 0, d, 0, 0};
    op->pf = plan_gemm(static_cast<int>(n1->rows()), 3 * d, d, A, Bm, e);
    maybe_split(op->pf);
    push(op, tb + ".attn1.to_qkv");
  }
  Tensor* ao = self_attention(qkv, heads, tb + ".attn1");
  h = linear(ao, d, tb + ".attn1.to_out.0", true, h);
  // --- cross attention over the 2 empty-prompt tokens (K, V precomputed in prepare())
  MDC_CHECK(d / heads == 64 && 2 * heads <= XA_MAXC, "cross-attention needs head_dim 64 and <= %d heads", XA_MAXC / 2);
  MDC_CHECK(d % 8 == 0 && d <= 32 * 8 * LN_MAXV, "cross-attention width %d unsupported", d);
  if (!getenv("MDC_NO_XFUSE")) {  // LN2 + to_q + 2-token attention + to_out + residual collapsed into one kernel (kernels.cuh)
    XUse u;
    u.wk = slot(tb + ".attn2.to_k.weight", W_LIN, d, cfg.cross_dim);
    u.wv = slot(tb + ".attn2.to_v.weight", W_LIN, d, cfg.cross_dim);
    u.wq = slot(tb + ".attn2.to_q.weight", W_LIN, d, d);
    u.wo = slot(tb + ".attn2.to_out.0.weight", W_LIN, d, d);
    WeightSlot* bo = slot(tb + ".attn2.to_out.0.bias", W_VEC, d, 0);
    WeightSlot* ga = slot(tb + ".norm2.weight", W_VEC, d, 0);
    WeightSlot* be = slot(tb + ".norm2.bias", W_VEC, d, 0);
    u.kc = arena.make<float>(2ull * d), u.vc = arena.make<float>(2ull * d);
    u.At = arena.make<float>(2ull * heads * d), u.U = arena.make<float>(2ull * heads * d);
    u.d = d, u.heads = heads;
    xuses.push_back(u);
    Tensor* h2 = new_tensor(x->n, x->h, x->w, d, "");
    auto* op = new CrossAttnFusedOp();
    op->x = h, op->y = h2, op->heads = heads, op->gamma = ga->vec, op->beta = be->vec, op->At = u.At, op->U = u.U;
    op->bo = bo->vec;
    op->stats = arena.make<float>(2ull * x->rows());
    push(op, tb + ".attn2");
    h = h2;
  } else {
    Tensor* n2 = layer_norm(h, tb + ".norm2");
    Tensor* q2 = linear(n2, d, tb + ".attn2.to_q", false);
    Tensor* o2 = new_tensor(x->n, x->h, x->w, d, "");
    XUse u;
    u.wk = slot(tb + ".attn2.to_k.weight", W_LIN, d, cfg.cross_dim);
    u.wv = slot(tb + ".attn2.to_v.weight", W_LIN, d, cfg.cross_dim);
    u.kc = arena.make<float>(2ull * d);
    u.vc = arena.make<float>(2ull * d);
    u.d = d;
    xuses.push_back(u);
    auto* op = new CrossAttn2Op();
    op->q = q2, op->o = o2, op->heads = heads, op->kc = u.kc, op->vc = u.vc;
    push(op, tb + ".attn2");
    h = linear(o2, d, tb + ".attn2.to_out.0", true, h);
  }
  // --- GEGLU feed-forward
  Tensor* n3 = layer_norm(h, tb + ".norm3");
  Tensor* pr = linear(n3, 8 * d, tb + ".ff.net.0.proj", true);
  Tensor* gg = new_tensor(x->n, x->h, x->w, 4 * d, "");
  {
    auto* op = new GegluOp();
    op->x = pr, op->y = gg;
    push(op, tb + ".ff.net.0");
  }
  h = linear(gg, d, tb + ".ff.net.2", true, h);
  Tensor* y = linear(h, d, key + ".proj_out", true, x, out);
  if (y->name.empty()) y->name = key, named[key] = y;
  return y;
}
inline Tensor* Engine::upsample_conv(Tensor* x, int H2, int W2, const std::string& key, Tensor* out) {
  if (H2 == 2 * x->h && W2 == 2 * x->w && !getenv("MDC_NO_UPCONV")) {
    WeightSlot* W = slot(key + ".conv.weight", W_UPCONV, x->c, x->c);
    WeightSlot* B = slot(key + ".conv.bias", W_VEC, x->c, 0);
    Tensor* y = out ? out : new_tensor(x->n, H2, W2, x->c, "");
    auto* op = new UpConvOp();
    op->E = this, op->x = x, op->y = y, op->W = W, op->bias = B->vec;
    Epilogue e;
    e.out = y->d, e.ldc = y->ld, e.bias = B->vec;
    op->pf = plan_upconv_fwd(x->n, x->h, x->w, x->c, x->c, x->d, x->ld, W->w, e);
    if (!y->is_view) y->producer = &op->pf;
    push(op, key + ".conv");
    if (y->name.empty()) y->name = key, named[key] = y;
    return y;
  }
  Tensor* up = new_tensor(x->n, H2, W2, x->c, "");
  auto* op = new UpsampleOp();
  op->x = x, op->y = up;
  push(op, key + ".nearest");
  Tensor* y = conv3x3(up, x->c, key + ".conv", nullptr, out);
  if (y->name.empty()) y->name = key, named[key] = y;
  return y;
}
inline Tensor* Engine::downsample_conv(Tensor* x, const std::string& key, Tensor* out) {
  Tensor* full = conv3x3(x, x->c, key + ".conv");
  const int Ho = (x->h - 1) / 2 + 1, Wo = (x->w - 1) / 2 + 1;
  Tensor* y = out ? out : new_tensor(x->n, Ho, Wo, x->c, "");
  MDC_CHECK(y->h == Ho && y->w == Wo, "downsample: bad output size");
  auto* op = new SubsampleOp();
  op->x = full, op->y = y, op->off = 0;
  push(op, key);
  if (y->name.empty()) y->name = key, named[key] = y;
  return y;
}

// ------------------------------------------------------------------------------------------------ UNet graph
inline void Engine::build_unet() {
  cur_ops = &unet_ops;
  const int nb = cfg.unet_nblocks, L = cfg.unet_layers_per_block;
  const int* boc = cfg.unet_block_ch;
  const std::string U = "unet.";
  unet_in = new_tensor(N, lh, lw, cfg.unet_in_ch, "unet.in");
  // Plan of the up path's concat buffers (consumption order), so that skip producers can write into them directly.
  struct CatPlan {
    int rin, skipc;
  };
  std::vector<CatPlan> plan;
  {
    int cout = boc[nb - 1];
    for (int i = 0; i < nb; ++i) {
      int prev = cout;
      cout = boc[nb - 1 - i];
      int cin = boc[std::max(nb - 2 - i, 0)];
      for (int j = 0; j < L + 1; ++j) plan.push_back({j == 0 ? prev : cout, j == L ? cin : cout});
    }
  }
  const int n_skips = static_cast<int>(plan.size());
  std::vector<Tensor*> cats(n_skips, nullptr), skip_views(n_skips, nullptr);
  int skip_idx = 0;  // push order
  auto skip_target = [&](int h, int w) -> Tensor* {
    const int ci = n_skips - 1 - skip_idx;
    const CatPlan& cp = plan[ci];
    Tensor* cat = new_tensor(N, h, w, cp.rin + cp.skipc, "unet.cat" + std::to_string(ci));
    cats[ci] = cat;
    ++skip_idx;
    skip_views[ci] = view(cat, cp.rin, cp.skipc, "");
    return skip_views[ci];
  };
  Tensor* h = conv3x3(unet_in, boc[0], U + "conv_in", nullptr, skip_target(lh, lw));
  named["unet.conv_in"] = h;
  for (int i = 0; i < nb; ++i) {
    const std::string B = U + "down_blocks." + std::to_string(i);
    for (int j = 0; j < L; ++j) {
      const bool attn = cfg.unet_down_attn[i] != 0;
      Tensor* tgt = skip_target(h->h, h->w);
      h = resnet(h, boc[i], B + ".resnets." + std::to_string(j), true, 1e-5f, attn ? nullptr : tgt);
      if (attn) h = transformer(h, cfg.unet_heads[i], B + ".attentions." + std::to_string(j), tgt);
    }
    if (i != nb - 1) {
      const int Ho = (h->h - 1) / 2 + 1, Wo = (h->w - 1) / 2 + 1;
      h = downsample_conv(h, B + ".downsamplers.0", skip_target(Ho, Wo));
    }
  }
  MDC_CHECK(skip_idx == n_skips, "skip bookkeeping mismatch");
  const int cm = boc[nb - 1];
  h = resnet(h, cm, U + "mid_block.resnets.0", true, 1e-5f);
  h = transformer(h, cfg.unet_heads[nb - 1], U + "mid_block.attentions.0");
  h = resnet(h, cm, U + "mid_block.resnets.1", true, 1e-5f, view(cats[0], 0, plan[0].rin, ""));
  int ci = 0;
  for (int i = 0; i < nb; ++i) {
    const std::string B = U + "up_blocks." + std::to_string(i);
    const int cout = boc[nb - 1 - i];
    const bool attn = cfg.unet_down_attn[nb - 1 - i] != 0;
    for (int j = 0; j < L + 1; ++j, ++ci) {
      Tensor* cat = cats[ci];
      {
        auto* op = new ConcatOp();
        op->a = h, op->b = skip_views[ci], op->cat = cat;  // the very Tensor objects the producers wrote
        MDC_CHECK(h->d == cat->d && h->c == plan[ci].rin && h->ld == cat->ld, "concat %d: producer did not write in place", ci);
        push(op, "cat" + std::to_string(ci));
      }
      const bool last_in_block = (j == L);
      const bool has_up = last_in_block && (i != nb - 1);
      Tensor* next_tgt = nullptr;
      if (!has_up && ci + 1 < n_skips) next_tgt = view(cats[ci + 1], 0, plan[ci + 1].rin, "");
      h = resnet(cat, cout, B + ".resnets." + std::to_string(j), true, 1e-5f, attn ? nullptr : next_tgt);
      if (attn) h = transformer(h, cfg.unet_heads[nb - 1 - i], B + ".attentions." + std::to_string(j), next_tgt);
      if (has_up) {
        Tensor* nxt = cats[ci + 1];
        h = upsample_conv(h, nxt->h, nxt->w, B + ".upsamplers.0", view(nxt, 0, plan[ci + 1].rin, ""));
      }
    }
  }
  h = group_norm(h, U + "conv_norm_out", 1e-5f, true);
  unet_out = conv3x3(h, cfg.unet_out_ch, U + "conv_out");
  named["unet.out"] = unet_out;
  // time embedding MLP (evaluated for all steps in prepare())
  te_l1w = slot(U + "time_embedding.linear_1.weight", W_LIN, boc[0] * 4, boc[0]);
  te_l1b = slot(U + "time_embedding.linear_1.bias", W_VEC, boc[0] * 4, 0);
  te_l2w = slot(U + "time_embedding.linear_2.weight", W_LIN, boc[0] * 4, boc[0] * 4);
  te_l2b = slot(U + "time_embedding.linear_2.bias", W_VEC, boc[0] * 4, 0);
}

// single-head attention of the VAE mid blocks (bias on q/k/v/out, residual): GroupNorm -> fused qkv linear -> SDPA -> out
inline Tensor* Engine::vae_mid_attention(Tensor* h, const std::string& A) {
  const int c0 = h->c;
  Tensor* gn = group_norm(h, A + ".group_norm", 1e-6f, false);
  Tensor* qkv = new_tensor(N, h->h, h->w, 3 * c0, "");
  bf16* w = static_cast<bf16*>(bank->blob(A + ".qkv.w", (3ull * c0 * c0 + 64) * 2));
  bf16* wt = static_cast<bf16*>(bank->blob(A + ".qkv.wt", (3ull * c0 * c0 + 64) * 2));
  float* b = static_cast<float*>(bank->blob(A + ".qkv.b", (3ull * c0 + 16) * 4));
  const char* nm[3] = {".to_q", ".to_k", ".to_v"};
  for (int i = 0; i < 3; ++i) {
    slot_lin_into(A + nm[i] + ".weight", c0, c0, w + 1ull * i * c0 * c0, c0, wt + i * c0, 3 * c0);
    slot_vec_into(A + nm[i] + ".bias", c0, b + i * c0);
  }
  auto* op = new LinearOp();
  op->E = this, op->x = gn, op->y = qkv, op->res = nullptr, op->w = w, op->wt = wt, op->ld_w = c0, op->ld_wt = 3 * c0;
  op->bias = b;
  Epilogue e;
  e.out = qkv->d, e.ldc = qkv->ld, e.bias = b;
  Operand Aop{gn->d, 0, gn->ld, 0, 0}, Bm{w, 0, c0, 0, 0};
  op->pf = plan_gemm(static_cast<int>(gn->rows()), 3 * c0, c0, Aop, Bm, e);
  push(op, A + ".to_qkv");
  Tensor* ao = self_attention(qkv, 1, A + ".sdpa");
  Tensor* y = linear(ao, c0, A + ".to_out.0", true, h);
  y->name = A, named[A] = y;
  return y;
}

// ------------------------------------------------------------------------------------------------ VAE encoder graph
// Forward only (marigold_dc.py:687-698 -> AutoencoderKL.encode(...).latent_dist.mode()).
// Down-sampling is F.pad(x, (0,1,0,1)) + conv3x3 stride 2 pad 0 = the stride-1 pad-1 conv sampled at ODD positions.
inline void Engine::build_encoder() {
  cur_ops = &enc_ops;
  alloc_grads = false;
  const int nb = cfg.vae_nblocks, L = cfg.vae_layers_per_block;
  const int* boc = cfg.vae_block_ch;
  const std::string V = "vae.";
  enc_in = new_tensor(N, PPH, PPW, 3, "vae.enc_in");
  Tensor* h = conv3x3(enc_in, boc[0], V + "encoder.conv_in");
  for (int i = 0; i < nb; ++i) {
    const std::string B = V + "encoder.down_blocks." + std::to_string(i);
    for (int j = 0; j < L; ++j) h = resnet(h, boc[i], B + ".resnets." + std::to_string(j), false, 1e-6f);
    if (i != nb - 1) {
      MDC_CHECK(h->h % 2 == 0 && h->w % 2 == 0, "encoder: odd feature map %dx%d", h->h, h->w);
      Tensor* full = conv3x3(h, h->c, B + ".downsamplers.0.conv");
      Tensor* y = new_tensor(N, h->h / 2, h->w / 2, h->c, "");
      auto* op = new SubsampleOp();
      op->x = full, op->y = y, op->off = 1;
      push(op, B + ".downsamplers.0");
      h = y;
    }
  }
  h = resnet(h, h->c, V + "encoder.mid_block.resnets.0", false, 1e-6f);
  h = vae_mid_attention(h, V + "encoder.mid_block.attentions.0");
  h = resnet(h, h->c, V + "encoder.mid_block.resnets.1", false, 1e-6f);
  h = group_norm(h, V + "encoder.conv_norm_out", 1e-6f, true);
  h = conv3x3(h, 2 * cfg.vae_latent_ch, V + "encoder.conv_out");
  enc_out = linear(h, 2 * cfg.vae_latent_ch, V + "quant_conv", true);
  named["vae.enc_out"] = enc_out;
  MDC_CHECK(enc_out->h == lh && enc_out->w == lw, "encoder output %dx%d != latent %dx%d", enc_out->h, enc_out->w, lh, lw);
  alloc_grads = true;
}

// ------------------------------------------------------------------------------------------------ AutoencoderTiny graphs
// The VAE the reference CLI uses by default (predict.py:44-52, 484-488): plain 64-channel
// conv3x3 + ReLU residual blocks, nearest x2 upsampling, no normalisation, no attention.  ReLU rides in the conv
// epilogue; its backward is a mask applied to a tensor's gradient once the first forward consumer (the last backward
// contributor) has added its part.
inline Tensor* Engine::tiny_conv(Tensor* x, int cout, const std::string& key, bool has_bias, bool relu, Tensor* res, float alpha,
                                 float bias_shift) {
  WeightSlot* W = slot(key + ".weight", W_CONV3, cout, x->c);
  WeightSlot* B = has_bias ? slot(key + ".bias", W_VEC, cout, 0) : nullptr;
  if (B) B->vscale = alpha, B->vshift = bias_shift;
  Tensor* y = new_tensor(x->n, x->h, x->w, cout, "");
  auto* op = new ConvOp();
  op->E = this, op->x = x, op->y = y, op->res = res, op->W = W, op->alpha = alpha;
  op->bias = B ? B->vec : nullptr;
  if (x->relu_out && !x->mask_owned && alloc_grads) op->mask_x = true, x->mask_owned = true;
  Epilogue e;
  e.out = y->d, e.ldc = y->ld, e.bias = op->bias, e.alpha = alpha, e.relu = relu ? 1 : 0;
  if (res) e.res = res->d, e.ldr = res->ld;
  op->pf = plan_conv3x3(x->n, x->h, x->w, x->c, cout, x->d, x->ld, W->w, e);
  y->relu_out = relu;
  push(op, key);
  y->name = key, named[key] = y;
  return y;
}
inline Tensor* Engine::tiny_block(Tensor* x, const std::string& key) {
  Tensor* a = tiny_conv(x, x->c, key + ".conv.0", true, true);
  a = tiny_conv(a, x->c, key + ".conv.2", true, true);
  return tiny_conv(a, x->c, key + ".conv.4", true, true, x);  // relu(conv(a) + x)
}
inline Tensor* Engine::tiny_upconv(Tensor* x, const std::string& key) {  // nn.Upsample(x2, nearest) + conv3x3 without bias
  WeightSlot* W = slot(key + ".weight", W_UPCONV, x->c, x->c);
  Tensor* y = new_tensor(x->n, 2 * x->h, 2 * x->w, x->c, "");
  auto* op = new UpConvOp();
  op->E = this, op->x = x, op->y = y, op->W = W, op->bias = nullptr;
  if (x->relu_out && !x->mask_owned && alloc_grads) op->mask_x = true, x->mask_owned = true;
  Epilogue e;
  e.out = y->d, e.ldc = y->ld;
  op->pf = plan_upconv_fwd(x->n, x->h, x->w, x->c, x->c, x->d, x->ld, W->w, e);
  push(op, key);
  y->name = key, named[key] = y;
  return y;
}
inline void Engine::build_tiny_decoder() {
  cur_ops = &dec_ops;
  const std::string L = "vae.decoder.layers.";
  dec_in = new_tensor(N, lh, lw, cfg.vae_latent_ch, "vae.in");
  Tensor* z = new_tensor(N, lh, lw, cfg.vae_latent_ch, "vae.decoder.clamped");
  {
    auto* op = new TanhClampOp();
    op->x = dec_in, op->y = z, op->m = cfg.tiny_magnitude;
    push(op, "vae.decoder.tanh");
  }
  Tensor* h = tiny_conv(z, cfg.vae_block_ch[0], L + "0", true, true);
  int idx = 2;  // index 1 is the ReLU module
  for (int i = 0; i < cfg.vae_nblocks; ++i) {
    const bool last = i == cfg.vae_nblocks - 1;
    MDC_CHECK(cfg.vae_block_ch[i] == h->c, "AutoencoderTiny: decoder widths must be equal (got %d after %d)", cfg.vae_block_ch[i], h->c);
    for (int j = 0; j < cfg.tiny_dec_blocks[i]; ++j) h = tiny_block(h, L + std::to_string(idx++));
    if (!last) {
      ++idx;  // the nn.Upsample module
      h = tiny_upconv(h, L + std::to_string(idx++));
    } else {
      // sample = conv(h) * 2 - 1: alpha = 2 with the bias stored as 2 b - 1
      dec_out = tiny_conv(h, 3, L + std::to_string(idx++), true, false, nullptr, 2.f, -1.f);
    }
  }
  named["vae.out"] = dec_out;
  MDC_CHECK(dec_out->h == PPH && dec_out->w == PPW, "AutoencoderTiny decoder output %dx%d != %dx%d", dec_out->h, dec_out->w, PPH, PPW);
}
inline void Engine::build_tiny_encoder() {
  cur_ops = &enc_ops;
  alloc_grads = false;
  const std::string L = "vae.encoder.layers.";
  enc_in = new_tensor(N, PPH, PPW, 3, "vae.enc_in");
  Tensor* u = new_tensor(N, PPH, PPW, 3, "");
  {
    auto* op = new UnitRangeOp();
    op->x = enc_in, op->y = u;
    push(op, "vae.encoder.unit_range");
  }
  Tensor* h = nullptr;
  int idx = 0;
  for (int i = 0; i < cfg.vae_nblocks; ++i) {
    const int c = cfg.vae_block_ch[i];
    if (i == 0) {
      h = tiny_conv(u, c, L + std::to_string(idx++), true, false);
    } else {  // conv3x3 stride 2 pad 1 without bias = the stride-1 conv sampled at even positions
      MDC_CHECK(h->h % 2 == 0 && h->w % 2 == 0, "AutoencoderTiny encoder: odd feature map %dx%d", h->h, h->w);
      Tensor* full = tiny_conv(h, c, L + std::to_string(idx++), false, false);
      Tensor* y = new_tensor(N, h->h / 2, h->w / 2, c, "");
      auto* op = new SubsampleOp();
      op->x = full, op->y = y, op->off = 0;
      push(op, full->name + ".stride2");
      h = y;
    }
    for (int j = 0; j < cfg.tiny_enc_blocks[i]; ++j) h = tiny_block(h, L + std::to_string(idx++));
  }
  enc_out = tiny_conv(h, cfg.vae_latent_ch, L + std::to_string(idx++), true, false);
  named["vae.enc_out"] = enc_out;
  MDC_CHECK(enc_out->h == lh && enc_out->w == lw, "encoder output %dx%d != latent %dx%d", enc_out->h, enc_out->w, lh, lw);
  alloc_grads = true;
}

// ------------------------------------------------------------------------------------------------ VAE decoder graph
inline void Engine::build_decoder() {
  cur_ops = &dec_ops;
  const int nb = cfg.vae_nblocks, L = cfg.vae_layers_per_block;
  const int* boc = cfg.vae_block_ch;
  const std::string V = "vae.";
  dec_in = new_tensor(N, lh, lw, cfg.vae_latent_ch, "vae.in");
  Tensor* h = linear(dec_in, cfg.vae_latent_ch, V + "post_quant_conv", true);
  named["vae.post_quant_conv"] = h;
  const int c0 = boc[nb - 1];
  h = conv3x3(h, c0, V + "decoder.conv_in");
  named["vae.decoder.conv_in"] = h;
  h = resnet(h, c0, V + "decoder.mid_block.resnets.0", false, 1e-6f);
  h = vae_mid_attention(h, V + "decoder.mid_block.attentions.0");
  h = resnet(h, c0, V + "decoder.mid_block.resnets.1", false, 1e-6f);
  for (int i = 0; i < nb; ++i) {
    const std::string B = V + "decoder.up_blocks." + std::to_string(i);
    const int cout = boc[nb - 1 - i];
    for (int j = 0; j < L + 1; ++j) h = resnet(h, cout, B + ".resnets." + std::to_string(j), false, 1e-6f);
    if (i != nb - 1) h = upsample_conv(h, h->h * 2, h->w * 2, B + ".upsamplers.0");
  }
  h = group_norm(h, V + "decoder.conv_norm_out", 1e-6f, true);
  head_gn = static_cast<GroupNormOp*>(cur_ops->back().get());
  dec_out = conv3x3(h, 3, V + "decoder.conv_out");
  head_conv = static_cast<ConvOp*>(cur_ops->back().get());
  head_act = arena.make<unsigned char>(static_cast<size_t>(N) * PPH * PPW + 64);
  named["vae.out"] = dec_out;
  MDC_CHECK(dec_out->h == PPH && dec_out->w == PPW, "decoder output %dx%d != padded size %dx%d", dec_out->h, dec_out->w,
            PPH, PPW);
}

inline void Engine::finalize_plans() {
  gn_partial = arena.make<float>(gn_partial_floats + 64);
  gn_gstats = arena.make<float>(2ull * 64 * MAXN + 64);
  gn_ticket = arena.make<unsigned int>(MAXN + 16);
  gn_bar = arena.make<unsigned int>(16);
  gn_epi_partial = arena.make<float>(gn_epi_floats + 64);
  for (GemmPlan* g : gn_epi_plans) g->p.gn_partial = gn_epi_partial;
  attn_S = arena.make<float>(attn_S_floats + 64);
  temb_cur = arena.make<float>(temb_total + 64);
  for (auto* ops : {&unet_ops, &dec_ops, &enc_ops}) {
    for (auto& op : *ops) {
      if (auto* c = dynamic_cast<ConvOp*>(op.get())) {
        uintptr_t tag = reinterpret_cast<uintptr_t>(c->bias);
        if (tag & 1) {  // tagged time-embedding offset
          c->bias = temb_cur + (tag - 1);
          c->pf.p.bias = c->bias;
          finish_plan(c->pf);
        }
      } else if (auto* at = dynamic_cast<SelfAttnOp*>(op.get())) {
        attention_set_scratch(at->a, attn_S, at->qkv->g != nullptr);
      }
    }
    if (ops == &enc_ops) continue;  // forward only, not part of the guided step
    for (auto it = ops->rbegin(); it != ops->rend(); ++it) (*it)->plan_bwd();
    for (auto& op : *ops) launches_per_step += op->n_fwd() + op->n_bwd();
    if (ops == &unet_ops)
      for (auto& op : *ops) launches_per_sample_step += op->n_fwd();
  }
  split_ws = arena.make<float>(split_ws_floats + 64);
  for (GemmPlan* g : split_plans) g->p.ws = split_ws;
  {  // side-stream GEMMs get their own partial-sum workspace
    size_t side_fl = 0;
    std::vector<LinearOp*> side_ops;
    for (auto* ops : {&unet_ops, &dec_ops})
      for (auto& op : *ops)
        if (op->side)
          if (auto* l = dynamic_cast<LinearOp*>(op.get())) {
            side_ops.push_back(l);
            for (GemmPlan* g : {&l->pf, &l->pb})
              if (g->p.ksplit > 1) side_fl = std::max(side_fl, static_cast<size_t>(g->p.ksplit) * static_cast<size_t>(g->p.ws_split_stride));
          }
    if (side_fl) {
      split_ws_side = arena.make<float>(side_fl + 64);
      for (LinearOp* l : side_ops)
        for (GemmPlan* g : {&l->pf, &l->pb})
          if (g->p.ksplit > 1) g->p.ws = split_ws_side;
    }
  }
  for (auto* ops : {&unet_ops, &dec_ops}) {  // one reduction launch per split-K GEMM
    std::vector<const GemmPlan*> f, b;
    for (auto& op : *ops) op->gemm_plans(f, b);
    for (auto* g : f) launches_per_step += g->p.ksplit > 1;
    for (auto* g : b) launches_per_step += g->p.ksplit > 1;
  }
  launches_per_step += 12;  // tail kernels of step(): 9 of the default path + closed-form loss + dense map / dense loss (the
                            // last three return at once unless their option is set)
}

// The fields of mdc_config that determine the parameter set (everything but geometry / batch / steps / device).
inline void model_signature(const mdc_config& c, int (&sig)[64]) {
  memset(sig, 0, sizeof(sig));
  int k = 0;
  auto put = [&](int v) { sig[k++] = v; };
  put(c.unet_in_ch), put(c.unet_out_ch), put(c.unet_nblocks), put(c.unet_layers_per_block), put(c.unet_groups), put(c.cross_dim);
  for (int i = 0; i < MDC_MAX_BLOCKS; ++i) put(c.unet_block_ch[i]), put(c.unet_heads[i]), put(c.unet_down_attn[i]);
  put(c.vae_nblocks), put(c.vae_layers_per_block), put(c.vae_groups), put(c.vae_latent_ch), put(c.vae_kind);
  for (int i = 0; i < MDC_MAX_BLOCKS; ++i) put(c.vae_block_ch[i]), put(c.tiny_enc_blocks[i]), put(c.tiny_dec_blocks[i]);
}

inline Engine::Engine(const mdc_config& c, std::shared_ptr<WeightBank> share) : cfg(c) {
  N = c.n_batch, H = c.height, W = c.width, ph = c.proc_h, pw = c.proc_w;
  PPH = ph + c.pad_h, PPW = pw + c.pad_w;
  MDC_CHECK(N >= 1 && N <= MAXN, "n_batch %d out of range (1..%d)", N, MAXN);
  MDC_CHECK(PPH % 8 == 0 && PPW % 8 == 0, "padded processing size %dx%d is not a multiple of 8", PPH, PPW);
  MDC_CHECK(c.unet_in_ch == 8 && c.unet_out_ch == 4 && c.vae_latent_ch == 4, "unsupported channel configuration");
  MDC_CHECK(c.unet_nblocks >= 2 && c.unet_nblocks <= 8 && c.vae_nblocks >= 2 && c.vae_nblocks <= 8, "bad block count");
  MDC_CHECK(c.steps >= 1 && c.steps <= 1000, "bad step count %d", c.steps);
  lh = PPH / 8, lw = PPW / 8;
  int expect = 1;
  for (int i = 1; i < c.vae_nblocks; ++i) expect *= 2;
  MDC_CHECK(expect == 8, "VAE must upsample by 8 (got %d)", expect);
  MDC_CUDA(cudaSetDevice(c.device));
  int sig[64];
  model_signature(c, sig);
  if (share) {
    MDC_CHECK(share->device == c.device, "mdc_create_shared: the handles are on different devices (%d vs %d)", share->device, c.device);
    MDC_CHECK(memcmp(sig, share->model_sig, sizeof(sig)) == 0, "mdc_create_shared: the model configurations differ");
    bank = share;
  } else {
    bank = std::make_shared<WeightBank>();
    bank->device = c.device;
    memcpy(bank->model_sig, sig, sizeof(sig));
  }
  MDC_CUDA(cudaStreamCreate(&own_stream));  // blocking: implicitly ordered with the legacy default stream for plain C callers
  MDC_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  MDC_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
  MDC_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  MDC_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  MDC_CUDA(cudaEventCreateWithFlags(&ev_fork2, cudaEventDisableTiming));
  MDC_CUDA(cudaEventCreateWithFlags(&ev_join2, cudaEventDisableTiming));
  stream = own_stream;
  use_graph = getenv("MDC_NO_GRAPH") == nullptr;
  set_kernel_attrs_for_device();
  build_unet();
  if (cfg.vae_kind == 1) build_tiny_decoder(); else build_decoder();
  n_split_step = split_plans.size();
  if (cfg.vae_kind == 1) build_tiny_encoder(); else build_encoder();
  finalize_plans();
  // step state
  const size_t lat = 4ull * N * lh * lw;
  x = arena.make<bf16>(lat), m1 = arena.make<bf16>(lat), m2 = arena.make<bf16>(lat), img_lat = arena.make<bf16>(lat);
  x_adam_dbg = arena.make<bf16>(lat);
  dx_direct = arena.make<float>(lat), gbuf = arena.make<float>(lat);
  parts_per_img = std::max(1, std::min(64, (lh * lw + 255) / 256));
  eps_part = arena.make<float>(1ull * N * parts_per_img), g_part = arena.make<float>(1ull * N * parts_per_img);
  x1_part = arena.make<float>(1ull * N * parts_per_img), x2_part = arena.make<float>(1ull * N * parts_per_img);
  opts = arena.make<TailOpts>(1);
  pt_a = arena.make<float>(1ull * N * H * W), pt_G = arena.make<float>(1ull * N * H * W);
  dn_map = arena.make<float>(1ull * N * H * W), gray_gx = arena.make<float>(1ull * N * H * W), gray_gy = arena.make<float>(1ull * N * H * W);
  dmean = arena.make<float>(1ull * N * PPH * PPW);
  pt_idx = arena.make<int>(1ull * N * H * W), pt_val = arena.make<float>(1ull * N * H * W);
  pt_off = arena.make<int>(N + 1), pt_cnt = arena.make<int>(MAXN);
  gminmax = arena.make<float>(2 * MAXN), depth_minmax = arena.make<float>(2 * MAXN);
  cur = arena.make<StepCur>(1), accum = arena.make<StepAccum>(1), counter = arena.make<int>(1);
  final_cur = arena.make<StepCur>(1), final_scratch = arena.make<float>(1ull * N * parts_per_img + 16);
  {
    StepCur one;
    memset(&one, 0, sizeof(one));
    one.sqrt_a = 1.f, one.sqrt_1ma = 0.f;
    copy_sync(final_cur, &one, sizeof(one), cudaMemcpyHostToDevice);
  }
  fr_guide = arena.make<float>(1ull * N * H * W + 64);
  fr_mask = arena.make<uint8_t>(1ull * N * H * W + 64);
  fr_stats = arena.make<float>(5ull * MAXN);
  d_sqrt_a = arena.make<float>(c.steps), d_sqrt_1ma = arena.make<float>(c.steps);
  d_sqrt_ap = arena.make<float>(c.steps), d_sqrt_1map = arena.make<float>(c.steps);
  temb_table = arena.make<float>(1ull * c.steps * temb_total + 64);
  MDC_CUDA(cudaDeviceSynchronize());
}

// ================================================================================================ runtime
// Re-packs `n` parameters with ONE kernel launch (pack_jobs_kernel): the job table is built on the host, validated
// against the slots' logical shapes and uploaded.  The sources are read before this function returns.
inline void Engine::set_weights(int n, const char* const* keys, const void* const* srcs, const long long* shapes4, const int* ndims,
                                const int* dtypes) {
  std::vector<PackJob> jobs;
  std::vector<WeightSlot*> touched;
  for (int k = 0; k < n; ++k) {
    const std::string key(keys[k]);
    const long long* shape = shapes4 + 4 * k;
    const int ndim = ndims[k], dtype = dtypes[k];
    auto it = wmap().find(key);
    MDC_CHECK(it != wmap().end(), "unknown weight key '%s'", key.c_str());
    MDC_CHECK(dtype == MDC_DTYPE_F32 || dtype == MDC_DTYPE_BF16, "weight dtype %d unsupported", dtype);
    MDC_CHECK(srcs[k] != nullptr && ndim >= 1 && ndim <= 4, "weight '%s': bad pointer / rank", key.c_str());
    long long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    for (WeightSlot* s : it->second) {
      PackJob j;
      memset(&j, 0, sizeof(j));
      j.src = srcs[k], j.src_bf16 = dtype == MDC_DTYPE_BF16 ? 1 : 0, j.out = s->out, j.in = s->in, j.vscale = 1.f;
      if (s->kind == W_CONV3 || s->kind == W_UPCONV) {
        MDC_CHECK(ndim == 4 && shape[0] == s->out && shape[1] == s->in && shape[2] == 3 && shape[3] == 3,
                  "weight '%s': expected [%d,%d,3,3]", key.c_str(), s->out, s->in);
        const bool up = s->kind == W_UPCONV;
        j.kind = up ? PK_UPCONV_FWD : PK_CONV_FWD, j.dst = s->w, j.pad = ((s->in + 63) / 64) * 64;
        jobs.push_back(j);
        j.kind = up ? PK_UPCONV_BWD : PK_CONV_DGRAD, j.dst = s->wt, j.pad = ((s->out + 63) / 64) * 64;
        jobs.push_back(j);
      } else if (s->kind == W_LIN) {
        MDC_CHECK(numel == 1LL * s->out * s->in && shape[0] == s->out, "weight '%s': expected [%d,%d]", key.c_str(), s->out, s->in);
        j.kind = PK_MATRIX, j.dst = s->w, j.ld = s->ld_w;
        jobs.push_back(j);
        j.kind = PK_MATRIX_T, j.dst = s->wt, j.ld = s->ld_wt;
        jobs.push_back(j);
      } else {  // bf16 parameters are used at bf16 precision, like the reference's bf16 modules
        MDC_CHECK(numel == s->out, "weight '%s': expected %d elements, got %lld", key.c_str(), s->out, numel);
        j.kind = PK_VEC, j.dst = s->vec, j.vscale = s->vscale, j.vshift = s->vshift;
        jobs.push_back(j);
      }
      touched.push_back(s);
    }
  }
  if (jobs.empty()) return;
  PackJob* d_jobs = nullptr;
  MDC_CUDA(cudaMalloc(&d_jobs, jobs.size() * sizeof(PackJob)));
  MDC_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, stream));
  for (size_t first = 0; first < jobs.size(); first += 65535) {  // gridDim.y limit
    const unsigned cnt = static_cast<unsigned>(std::min<size_t>(65535, jobs.size() - first));
    pack_jobs_kernel<<<dim3(64, cnt), 256, 0, stream>>>(d_jobs + first);
  }
  MDC_CUDA(cudaGetLastError());
  MDC_CUDA(cudaStreamSynchronize(stream));  // the caller may release its tensors as soon as we return
  cudaFree(d_jobs);
  for (WeightSlot* s : touched) s->loaded = true;
}
inline void Engine::set_weight(const std::string& key, const void* src, const long long* shape, int ndim, int dtype) {
  long long s4[4] = {1, 1, 1, 1};
  MDC_CHECK(ndim >= 1 && ndim <= 4, "weight '%s': rank %d", key.c_str(), ndim);
  for (int i = 0; i < ndim; ++i) s4[i] = shape[i];
  const char* k = key.c_str();
  set_weights(1, &k, &src, s4, &ndim, &dtype);
}
inline bool Engine::weights_loaded() {
  for (auto& kv : wmap())
    for (WeightSlot* s : kv.second)
      if (!s->loaded) return false;
  return true;
}

inline void Engine::prepare(const void* ctx_bf16, const float* alphas_cumprod, const int* timesteps, int n_steps) {
  for (auto& kv : wmap())
    for (WeightSlot* s : kv.second) MDC_CHECK(s->loaded, "weight '%s' has not been set", kv.first.c_str());
  MDC_CHECK(n_steps == cfg.steps, "prepare: n_steps %d != configured steps %d", n_steps, cfg.steps);
  // DDIM scalars (host, double -> float like torch's 0-dim fp32 tensors)
  std::vector<float> sa(n_steps), sb(n_steps), sap(n_steps), sbp(n_steps);
  const int stride = 1000 / n_steps;
  for (int i = 0; i < n_steps; ++i) {
    const int t = timesteps[i], tp = t - stride;
    MDC_CHECK(t >= 0 && t < 1000, "timestep %d out of range", t);
    const float a = alphas_cumprod[t], ap = tp >= 0 ? alphas_cumprod[tp] : alphas_cumprod[0];
    sa[i] = sqrtf(a), sb[i] = sqrtf(1.f - a), sap[i] = sqrtf(ap), sbp[i] = sqrtf(1.f - ap);
  }
  copy_sync(d_sqrt_a, sa.data(), n_steps * 4, cudaMemcpyHostToDevice);
  copy_sync(d_sqrt_1ma, sb.data(), n_steps * 4, cudaMemcpyHostToDevice);
  copy_sync(d_sqrt_ap, sap.data(), n_steps * 4, cudaMemcpyHostToDevice);
  copy_sync(d_sqrt_1map, sbp.data(), n_steps * 4, cudaMemcpyHostToDevice);
  // time embedding for every step: sinusoid -> linear_1 -> SiLU -> linear_2, then per resnet time_emb_proj(SiLU(.)) + conv bias
  const int c0 = cfg.unet_block_ch[0], tc = c0 * 4;
  int* d_ts = nullptr;
  float *emb = nullptr, *h1 = nullptr, *h2 = nullptr;
  MDC_CUDA(cudaMalloc(&d_ts, n_steps * 4));
  MDC_CUDA(cudaMalloc(&emb, 4ull * n_steps * c0));
  MDC_CUDA(cudaMalloc(&h1, 4ull * n_steps * tc));
  MDC_CUDA(cudaMalloc(&h2, 4ull * n_steps * tc));
  copy_sync(d_ts, timesteps, n_steps * 4, cudaMemcpyHostToDevice);
  launch_k(timestep_embedding_kernel, dim3((n_steps * c0 + 255) / 256), dim3(256), 0, stream, d_ts, n_steps, c0, emb);
  auto lin = [&](const bf16* Wm, long long ldw, const float* b, const float* in, long long ldin, int In, int Out, int silu,
                 float* out, long long ldout, int S) {
    long long warps = 1LL * S * Out;
    launch_k(small_linear_kernel, dim3(static_cast<int>((warps * 32 + 255) / 256)), dim3(256), 0, stream, Wm, ldw, b, in, ldin, S, In, Out,
                                                                                        silu, out, ldout);
  };
  lin(te_l1w->w, te_l1w->ld_w, te_l1b->vec, emb, c0, c0, tc, 0, h1, tc, n_steps);
  lin(te_l2w->w, te_l2w->ld_w, te_l2b->vec, h1, tc, tc, tc, 1, h2, tc, n_steps);
  MDC_CUDA(cudaMemsetAsync(temb_table, 0, 4ull * n_steps * temb_total, stream));
  for (auto& u : temb_uses) {
    lin(u.proj_w->w, u.proj_w->ld_w, u.proj_b->vec, h2, tc, tc, u.cout, 1, temb_table + u.off, temb_total, n_steps);
  }
  // add the conv bias: table[s][off + c] += conv_b[c]   (small host loop over uses, one kernel each would be overkill)
  MDC_CUDA(cudaStreamSynchronize(stream));
  {
    std::vector<float> tab(1ull * n_steps * temb_total), cb;
    copy_sync(tab.data(), temb_table, tab.size() * 4, cudaMemcpyDeviceToHost);
    for (auto& u : temb_uses) {
      cb.resize(u.cout);
      copy_sync(cb.data(), u.conv_b->vec, u.cout * 4, cudaMemcpyDeviceToHost);
      for (int s = 0; s < n_steps; ++s)
        for (int c = 0; c < u.cout; ++c) tab[1ull * s * temb_total + u.off + c] += cb[c];
    }
    copy_sync(temb_table, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice);
  }
  // cross-attention K / V of the 2 empty-prompt tokens: [2, d] = ctx [2, cross] . W^T
  {
    float* ctxf = nullptr;
    MDC_CUDA(cudaMalloc(&ctxf, 4ull * 2 * cfg.cross_dim));
    to_f32_kernel<bf16><<<8, 256, 0, stream>>>(static_cast<const bf16*>(ctx_bf16), ctxf, 2LL * cfg.cross_dim);
    for (auto& u : xuses) {
      lin(u.wk->w, u.wk->ld_w, nullptr, ctxf, cfg.cross_dim, cfg.cross_dim, u.d, 0, u.kc, u.d, 2);
      lin(u.wv->w, u.wv->ld_w, nullptr, ctxf, cfg.cross_dim, cfg.cross_dim, u.d, 0, u.vc, u.d, 2);
      if (u.At) {
        const int tot = 2 * u.heads * u.d;
        xattn_collapse_kernel<<<(tot + 255) / 256, 256, 0, stream>>>(u.wq->w, u.wq->ld_w, u.wo->w, u.wo->ld_w, u.kc, u.vc, u.d,
                                                                     u.heads, 0.125f, u.At, u.U);
      }
    }
    MDC_CUDA(cudaStreamSynchronize(stream));
    cudaFree(ctxf);
  }
  cudaFree(d_ts), cudaFree(emb), cudaFree(h1), cudaFree(h2);
  tables.sqrt_a = d_sqrt_a, tables.sqrt_1ma = d_sqrt_1ma, tables.sqrt_ap = d_sqrt_ap, tables.sqrt_1map = d_sqrt_1map;
  tables.temb_bias = temb_table, tables.temb_total = temb_total, tables.steps = n_steps;
  copy_sync(opts, &h_opts, sizeof(h_opts), cudaMemcpyHostToDevice);
  MDC_CUDA(cudaGetLastError());
  prepared = true;
}

// Per-call state shared by mdc_begin and mdc_begin_frame.  Everything is enqueued on `stream`; the valid points are
// counted and compacted on the device (pt_cnt -> pt_off -> pt_idx / pt_val), the host only reads back N + 1 offsets
// at the end to raise the reference's empty-mask error.
inline void Engine::begin_state(const void* img_latents, const void* x0, const float* guide, const uint8_t* mask, float lrx, float lrs,
                                const float* stats5_dev) {
  MDC_CHECK(prepared, "mdc_begin called before mdc_prepare");
  const size_t lat = 4ull * N * lh * lw;
  if (img_latents != img_lat) MDC_CUDA(cudaMemcpyAsync(img_lat, img_latents, lat * 2, cudaMemcpyDeviceToDevice, stream));
  MDC_CUDA(cudaMemcpyAsync(x, x0, lat * 2, cudaMemcpyDeviceToDevice, stream));
  MDC_CUDA(cudaMemsetAsync(m1, 0, lat * 2, stream));
  MDC_CUDA(cudaMemsetAsync(m2, 0, lat * 2, stream));
  MDC_CUDA(cudaMemsetAsync(dmean, 0, 4ull * N * PPH * PPW, stream));
  MDC_CUDA(cudaMemsetAsync(counter, 0, 4, stream));
  StepAccum a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < MAXN; ++i) a.scale[i] = 1.f;
  MDC_CUDA(cudaMemcpyAsync(accum, &a, sizeof(a), cudaMemcpyHostToDevice, stream));
  MDC_CHECK(h_opts.w_edge == 0.f || have_gray, "image must be provided for edge loss (use mdc_begin_frame)");
  have_gray = false;
  h_opts.lr_x = lrx, h_opts.lr_s = lrs;
  MDC_CUDA(cudaMemcpyAsync(opts, &h_opts, sizeof(h_opts), cudaMemcpyHostToDevice, stream));
  if (!stats5_dev) launch_k(count_mask_kernel, dim3(N), dim3(1024), 0, stream, mask, H * W, pt_cnt);
  launch_k(frame_state_kernel, dim3(1), dim3(32), 0, stream, static_cast<const int*>(pt_cnt), stats5_dev, N, pt_off, gminmax, depth_minmax);
  launch_k(compact_points_kernel, dim3(N), dim3(1024), 0, stream, guide, mask, H * W, pt_off, pt_idx, pt_val);
  MDC_CUDA(cudaGetLastError());
  std::vector<int> off(N + 1, 0);
  MDC_CUDA(cudaMemcpyAsync(off.data(), pt_off, (N + 1) * 4, cudaMemcpyDeviceToHost, stream));
  MDC_CUDA(cudaStreamSynchronize(stream));
  for (int n = 0; n < N; ++n)
    MDC_CHECK(off[n + 1] > off[n], "No valid values found in mask for some positions. Ensure that mask has at least one True value "
                                   "along the specified dimensions. (sample %d has no valid sparse-depth point: empty mask)", n);
  {
    // Sparse output head (head.cuh): its backward does the full GroupNorm arithmetic per pixel in the 3x3 neighbourhoods
    // of the taps (at most 4 taps x 9 pixels per point) and streams elsewhere.  A/B-measured with 500 scattered points
    // (4 % of the 576x768 map active: -0.17 ms per step); the KITTI-shaped frame (21 k points on 64 scan lines, about a
    // third of the map active) was benchmarked and parity-tested with it on, without an A/B of its own.  Once the valid
    // points exceed ~11 % of the pixels every decoder pixel is active and the dense kernels cannot lose: such a frame
    // keeps the dense head.
    // MDC_SPARSEHEAD_MAXFRAC overrides the bound on 36 x points / pixels.
    static const double max_frac = [] {
      const char* e = getenv("MDC_SPARSEHEAD_MAXFRAC");
      return e ? atof(e) : 4.0;
    }();
    const bool before = sparse_head_allowed();
    head_points_ok = 36.0 * off[N] <= max_frac * N * PPH * PPW;
    if (before != sparse_head_allowed() && step_graph) {  // the captured step holds the other head: capture again
      cudaGraphExecDestroy(step_graph);
      step_graph = nullptr;
    }
  }
  if (getenv("MDC_DEBUG_SYNC")) {  // consistency of the compacted point list
    std::vector<int> idx(off[N]);
    copy_sync(idx.data(), pt_idx, idx.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < idx.size(); ++i) MDC_CHECK(idx[i] >= 0 && idx[i] < H * W, "pt_idx[%zu] = %d out of range (H*W = %d)", i, idx[i], H * W);
    fprintf(stderr, "[mdc debug] begin: %d points, geometry H %d W %d ph %d pw %d PPH %d PPW %d dec ld %lld\n", off[N], H, W, ph, pw, PPH, PPW,
            dec_out->ld);
  }
  lr_x = lrx, lr_s = lrs;
  steps_done = 0;
  begun = true;
}
inline void Engine::begin(const void* img_latents, const void* x0, const float* guide, const uint8_t* mask,
                          const float* gmm, const float* dmm, float lrx, float lrs) {
  MDC_CHECK(prepared, "mdc_begin called before mdc_prepare");
  // host (min, max) pairs: staged through pageable memory, i.e. copied before cudaMemcpyAsync returns
  MDC_CUDA(cudaMemcpyAsync(gminmax, gmm, 8ull * N, cudaMemcpyHostToDevice, stream));
  MDC_CUDA(cudaMemcpyAsync(depth_minmax, dmm, 8ull * N, cudaMemcpyHostToDevice, stream));
  begin_state(img_latents, x0, guide, mask, lrx, lrs, nullptr);
}

// Keeps the stream busy for ~`ns` nanoseconds so the host can enqueue a long launch sequence ahead of the GPU.
__global__ void hold_stream_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}
inline void Engine::profile_gemm_step(float* ms_out, double* flops_out, int* launches_out) {
  MDC_CHECK(begun, "profile_gemm_step: call mdc_begin first");
  std::vector<cudaEvent_t> ev;
  std::vector<double> fl;
  // pre-created events + a GPU-side hold: everything below is enqueued while the GPU waits, so the event pairs
  // bracket kernel time, not the host's launch rate
  std::vector<cudaEvent_t> pool(2048);
  for (auto& e : pool) MDC_CUDA(cudaEventCreate(&e));
  size_t next_ev = 0;
  hold_stream_kernel<<<1, 1, 0, stream>>>(30ull * 1000 * 1000);
  auto timed = [&](const GemmPlan* g) {
    MDC_CHECK(next_ev + 2 <= pool.size(), "profile_gemm_step: event pool exhausted");
    cudaEvent_t a = pool[next_ev++], b = pool[next_ev++];
    MDC_CUDA(cudaEventRecord(a, stream));
    run_gemm(*g, stream);
    MDC_CUDA(cudaEventRecord(b, stream));
    ev.push_back(a), ev.push_back(b);
    fl.push_back(g->flops);
  };
  // Same launch order as step(), but only the GEMM-class kernels are bracketed; the others run unbracketed in between
  // so caches and clocks see the real mix.
  auto run_instrumented = [&](std::vector<std::unique_ptr<Op>>& ops, bool backward) {
    std::vector<const GemmPlan*> f, b;
    auto one = [&](Op* op) {
      f.clear(), b.clear();
      op->gemm_plans(f, b);
      auto& list = backward ? b : f;
      if (list.empty()) {
        backward ? op->bwd(stream) : op->fwd(stream);
        return;
      }
      // ops with GEMMs: replay the op normally (keeps non-GEMM kernels in order), then time its GEMMs again in place
      backward ? op->bwd(stream) : op->fwd(stream);
      for (const GemmPlan* g : list) timed(g);
    };
    if (!backward)
      for (auto& op : ops) one(op.get());
    else
      for (auto it = ops.rbegin(); it != ops.rend(); ++it) one(it->get());
  };
  run_instrumented(unet_ops, false);
  run_instrumented(dec_ops, false);
  run_instrumented(dec_ops, true);
  run_instrumented(unet_ops, true);
  MDC_CUDA(cudaStreamSynchronize(stream));
  double ms = 0, flops = 0;
  for (size_t i = 0; i < fl.size(); ++i) {
    float t = 0;
    MDC_CUDA(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
    ms += t, flops += fl[i];
  }
  for (auto& e : pool) cudaEventDestroy(e);
  *ms_out = static_cast<float>(ms), *flops_out = flops, *launches_out = static_cast<int>(fl.size());
}

inline void Engine::run_ops(std::vector<std::unique_ptr<Op>>& ops, bool backward) {
  static const bool dbg = getenv("MDC_DEBUG_SYNC") != nullptr;
  auto fork = [&]() {  // the side stream may start once everything launched on `stream` so far has finished
    MDC_CUDA(cudaEventRecord(ev_fork, stream));
    MDC_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
  };
  if (!backward)
    for (auto& op : ops) {
      if (op->fork_fwd) MDC_CUDA(cudaEventRecord(ev_fork, stream));
      if (op->side) {  // depends on the block input only (event recorded before norm1); its consumer is the next op
        MDC_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
        op->fwd(side_stream);
        MDC_CUDA(cudaEventRecord(ev_join, side_stream));
        MDC_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
      } else {
        op->fwd(stream);
      }
      if (dbg && begun && pt_off) {  // which op scribbles over the point list?
        int o2[2] = {0, 0};
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cs);
        if (cs == cudaStreamCaptureStatusNone) {
          copy_sync(o2, pt_off, 8, cudaMemcpyDeviceToHost);
          MDC_CHECK(o2[0] == 0 && o2[1] > 0 && o2[1] <= H * W, "pt_off corrupted (%d, %d) right after the forward of op '%s'", o2[0], o2[1],
                    op->name.c_str());
        }
      }
    }
  else {
    bool pending = false;
    for (auto it = ops.rbegin(); it != ops.rend(); ++it) {
      Op* op = it->get();
      if (op->side) {  // reads the block's output gradient (complete by now), writes the first contribution to dx
        fork();
        op->bwd(side_stream);
        MDC_CUDA(cudaEventRecord(ev_join, side_stream));
        pending = true;
        continue;
      }
      if (op->join_bwd && pending) {
        MDC_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
        pending = false;
      }
      op->bwd(stream);
    }
    if (pending) MDC_CUDA(cudaStreamWaitEvent(stream, ev_join, 0));
  }
}

inline Engine::~Engine() {
  if (step_graph) cudaGraphExecDestroy(step_graph);
  if (sample_graph) cudaGraphExecDestroy(sample_graph);
  if (own_stream) cudaStreamDestroy(own_stream);
  if (cap_stream) cudaStreamDestroy(cap_stream);
  if (side_stream) cudaStreamDestroy(side_stream);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (ev_fork2) cudaEventDestroy(ev_fork2);
  if (ev_join2) cudaEventDestroy(ev_join2);
}
inline void Engine::release_workspace() {
  MDC_CUDA(cudaStreamSynchronize(stream));
  if (step_graph) cudaGraphExecDestroy(step_graph);
  if (sample_graph) cudaGraphExecDestroy(sample_graph);
  step_graph = sample_graph = nullptr;
  unet_ops.clear(), dec_ops.clear(), enc_ops.clear();
  head_gn = nullptr, head_conv = nullptr, head_act = nullptr, sparse_head = false;  // they pointed into dec_ops / the arena
  split_plans.clear();
  tensors.clear(), named.clear();
  arena.free_all();
  prepared = begun = false;
  released = true;
}
// The single-launch GroupNorm kernels meet at a hand-rolled grid barrier; if its CTAs were ever not co-resident (GPU
// shared through MPS, preemption) the barrier times out after ~2 s, sets a sticky flag and the results are garbage.
// Called (after a stream synchronisation) by every entry point that hands results to the caller.
inline void Engine::check_barrier_flag() {
  unsigned int f = 0;
  copy_sync(&f, gn_bar + 2, 4, cudaMemcpyDeviceToHost);
  if (f) {
    MDC_CUDA(cudaMemsetAsync(gn_bar, 0, 16, stream));
    MDC_CHECK(false, "GroupNorm grid barrier timed out (the kernel's CTAs were not co-resident: is the GPU shared?); results of this "
                     "call are invalid.  Set MDC_NO_GNFUSE=1 to use the two-pass kernels.");
  }
}
// Captures `launches()` on the private capture stream with `stream` temporarily pointing at it.
template <typename F>
inline cudaGraphExec_t capture_graph(Engine* e, F&& launches) {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t saved = e->stream;
  e->stream = e->cap_stream;
  MDC_CUDA(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
  try {
    launches();
  } catch (...) {
    cudaStreamEndCapture(e->cap_stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    e->stream = saved;
    throw;
  }
  e->stream = saved;
  MDC_CUDA(cudaStreamEndCapture(e->cap_stream, &graph));
  MDC_CUDA(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  return exec;
}
// ---- sparse output head (head.cuh)
inline bool Engine::sparse_head_allowed() const {
  static const bool off = getenv("MDC_NO_SPARSEHEAD") != nullptr;
  if (off || !head_gn || !head_conv || !head_points_ok) return false;
  const GroupNormOp& gn = *head_gn;
  const ConvOp& cv = *head_conv;
  return !gn.plan.single_f() && !gn.plan.single_b() && !gn.epi && gn.silu && gn.x->ld == gn.x->c && gn.x->c % 64 == 0 && cv.y->c == 3 &&
         !cv.res && cv.alpha == 1.f && cv.x == gn.y && h_opts.w_edge == 0.f && h_opts.w_smooth == 0.f &&
         9 * gn.x->c * 4 + gn.plan.s.pix_per_block + 4096 <= 48 * 1024;  // wsum + flag bytes + warp sums in default shared memory
}
inline void Engine::head_fwd(cudaStream_t st) {
  const GroupNormOp& gn = *head_gn;
  const ConvOp& cv = *head_conv;
  const Tensor* x = gn.x;
  TailGeom g{N, H, W, ph, pw, PPH, PPW, dec_out->ld, interp_nearest};
  const int Cp = ((x->c + 63) / 64) * 64;
  launch_k(head_points_fwd_kernel, dim3(2 * g_num_sms()), dim3(256), 0, st, x->d, x->ld, x->c, gn.groups, Cp, gn.stats, gn.gamma, gn.beta, gn.silu,
           cv.W->w, cv.bias, g, pt_idx, pt_off, dec_out->d);
}
inline void Engine::head_bwd(cudaStream_t st) {
  const GroupNormOp& gn = *head_gn;
  const ConvOp& cv = *head_conv;
  const Tensor* x = gn.x;
  const GNPlan& p = gn.plan;
  const int Cp = ((x->c + 63) / 64) * 64;
  SparseDy sp{dec_out->g, dec_out->ld, PPH, PPW};
  const int grid = p.s.N * p.s.blocks_per_img;
  const size_t wsum_b = static_cast<size_t>(9) * x->c * sizeof(float) + ((p.s.pix_per_block + 15) & ~15);  // + the block's flag bytes
  const size_t warp_b = static_cast<size_t>((((p.threads_b + 31) / 32) * 2 * p.G + 3) & ~3) * sizeof(float);
  const GNScratch sc = gn_scratch();
  launch_k(gn_bwd_stats_sp_kernel, dim3(grid), dim3(p.threads_b), warp_b + wsum_b, st, x->d, sp, cv.W->w, Cp, p.s, gn.stats, gn.gamma, gn.beta,
           gn.silu, sc.partial, sc.gstats, sc.ticket, head_act);
  if (gn.acc)
    launch_k(gn_bwd_apply_sp_kernel<true>, dim3(grid), dim3(p.threads_b), wsum_b, st, x->d, sp, cv.W->w, Cp, p.s, gn.stats,
             static_cast<const float*>(sc.gstats), gn.gamma, gn.beta, gn.silu, static_cast<const unsigned char*>(head_act), x->g, x->ld);
  else
    launch_k(gn_bwd_apply_sp_kernel<false>, dim3(grid), dim3(p.threads_b), wsum_b, st, x->d, sp, cv.W->w, Cp, p.s, gn.stats,
             static_cast<const float*>(sc.gstats), gn.gamma, gn.beta, gn.silu, static_cast<const unsigned char*>(head_act), x->g, x->ld);
}
inline void Engine::step() {
  MDC_CHECK(begun, "mdc_step called before mdc_begin");
  MDC_CHECK(steps_done < cfg.steps, "all %d steps already done", cfg.steps);
  if (!use_graph) {
    step_launches();
  } else {
    if (!step_graph) step_graph = capture_graph(this, [&] { step_launches(); });
    MDC_CUDA(cudaGraphLaunch(step_graph, stream));
  }
  ++steps_done;
  launches += launches_per_step - (sparse_head_allowed() ? 2 : 0);  // sparse head: no gn_apply forward, no conv_out dgrad
}
inline void Engine::step_launches() {
  const int hw = lh * lw, lat_pix = N * hw;
  const int pgrid = N * parts_per_img;
  TailGeom g{N, H, W, ph, pw, PPH, PPW, dec_out->ld, interp_nearest};
  launch_k(begin_step_kernel, dim3(1), dim3(1024), 0, stream, tables, counter, cur, temb_cur, opts);
  launch_k(unet_input_kernel, dim3((lat_pix + 255) / 256), dim3(256), 0, stream, img_lat, x, N, hw, unet_in->d);
  auto dbg_points = [&](const char* tag) {  // MDC_DEBUG_SYNC: has anything scribbled over the compacted point list?
    static const bool on = getenv("MDC_DEBUG_SYNC") != nullptr;
    if (!on) return;
    std::vector<int> o2(N + 1);
    copy_sync(o2.data(), pt_off, o2.size() * 4, cudaMemcpyDeviceToHost);
    MDC_CHECK(o2[0] == 0 && o2[N] > 0 && o2[N] <= N * H * W, "%s: pt_off corrupted (%d .. %d)", tag, o2[0], o2[N]);
    std::vector<int> idx(o2[N]);
    copy_sync(idx.data(), pt_idx, idx.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < idx.size(); ++i) MDC_CHECK(idx[i] >= 0 && idx[i] < H * W, "%s: pt_idx[%zu] = %d corrupted", tag, i, idx[i]);
    StepAccum a;
    copy_sync(&a, accum, sizeof(a), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[mdc debug] %s: points ok, scale %g shift %g, dmean %p dec %p acc %p\n", tag, a.scale[0], a.shift[0], (void*)dmean,
            (void*)dec_out->d, (void*)accum);
  };
  run_ops(unet_ops, false);
  dbg_points("after unet fwd");
  launch_k(x0_kernel, dim3(pgrid), dim3(256), 0, stream, unet_out->d, x, cur, N, hw, cfg.vae_scaling, dec_in->d, eps_part, x1_part, x2_part);
  struct HeadGuard {  // never leave the flag raised (an exception from a launch must not turn later dense decodes sparse)
    bool& f;
    ~HeadGuard() { f = false; }
  } head_guard{sparse_head};
  sparse_head = sparse_head_allowed();  // the loss below reads the decoded map at the tap pixels of the valid points only
  run_ops(dec_ops, false);
  dbg_points("after decoder fwd");
  launch_k(loss_points_kernel, dim3(N), dim3(512), 0, stream, dec_out->d, g, pt_idx, pt_val, pt_off, gminmax, depth_minmax, opts, accum, dmean);
  launch_k(loss_points_cf_kernel, dim3(N), dim3(512), 0, stream, dec_out->d, g, pt_idx, pt_val, pt_off, depth_minmax, opts, accum, dmean,
           pt_a, pt_G);  // returns at once unless closed_form
  const long long npix = 1LL * N * PPH * PPW, opix = 1LL * N * H * W;
  // "edge" / "smooth" (marigold_dc.py:195-236); both return immediately unless loss_funcs lists them
  launch_k(dense_map_kernel, dim3(static_cast<int>((opix + 255) / 256)), dim3(256), 0, stream, dec_out->d, g, gminmax, depth_minmax, opts,
           accum, dn_map);
  launch_k(dense_loss_kernel, dim3(std::max(1, std::min(148, (H * W + 255) / 256)), N), dim3(256), 0, stream, dec_out->d, g, gminmax,
           depth_minmax, opts, dn_map, gray_gx, gray_gy, accum, dmean);
  launch_k(dec_grad_kernel, dim3(static_cast<int>((npix + 255) / 256)), dim3(256), 0, stream, dmean, npix, dec_out->g);
  run_ops(dec_ops, true);
  sparse_head = false;
  launch_k(dx0_kernel, dim3((lat_pix + 255) / 256), dim3(256), 0, stream, dec_in->g, cur, N, hw, cfg.vae_scaling, unet_out->g, dx_direct);
  run_ops(unet_ops, true);
  launch_k(grad_total_kernel, dim3(pgrid), dim3(256), 0, stream, dx_direct, unet_in->g, N, hw, gbuf, g_part, x, x1_part, x2_part, opts, accum);
  launch_k(adam_ddim_kernel, dim3(pgrid), dim3(256), 0, stream, gbuf, eps_part, g_part, parts_per_img, unet_out->d, cur, N, hw, x, m1, m2,
                                              accum, counter, x_adam_dbg, opts);
}

// One step of the no-grad branch (marigold_dc.py:805-809, :905-909): UNet forward + DDIM prev_sample.
inline void Engine::sample_launches() {
  const int hw = lh * lw, lat_pix = N * hw;
  launch_k(begin_step_kernel, dim3(1), dim3(1024), 0, stream, tables, counter, cur, temb_cur, opts);
  launch_k(unet_input_kernel, dim3((lat_pix + 255) / 256), dim3(256), 0, stream, img_lat, x, N, hw, unet_in->d);
  run_ops(unet_ops, false);
  launch_k(ddim_only_kernel, dim3((lat_pix + 255) / 256), dim3(256), 0, stream, unet_out->d, cur, N, hw, x, counter);
}
inline void Engine::sample_step() {
  MDC_CHECK(begun, "mdc_sample called before mdc_begin");
  MDC_CHECK(steps_done < cfg.steps, "all %d steps already done", cfg.steps);
  if (!use_graph) {
    sample_launches();
  } else {
    if (!sample_graph) sample_graph = capture_graph(this, [&] { sample_launches(); });
    MDC_CUDA(cudaGraphLaunch(sample_graph, stream));
  }
  launches += launches_per_sample_step;
  ++steps_done;
}

inline void Engine::decode_final(float* dense_out, bool closed_form) {
  MDC_CHECK(begun, "mdc_decode_final called before mdc_begin");
  const int hw = lh * lw;
  // z = x / scaling as NHWC: x0_kernel with the identity scalars (sqrt_a = 1, sqrt_1ma = 0) and v = 0
  MDC_CUDA(cudaMemsetAsync(unet_out->d, 0, static_cast<size_t>(unet_out->rows()) * unet_out->ld * 2, stream));
  launch_k(x0_kernel, dim3(N * parts_per_img), dim3(256), 0, stream, unet_out->d, x, final_cur, N, hw, cfg.vae_scaling, dec_in->d, final_scratch,
           static_cast<float*>(nullptr), static_cast<float*>(nullptr));
  run_ops(dec_ops, false);
  TailGeom g{N, H, W, ph, pw, PPH, PPW, dec_out->ld, interp_nearest};
  const long long tot = 1LL * N * H * W;
  if (closed_form) launch_k(affine_lsq_kernel, dim3(N), dim3(512), 0, stream, dec_out->d, g, pt_idx, pt_val, pt_off, accum);
  launch_k(dense_out_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, stream, dec_out->d, g, gminmax, depth_minmax, accum,
                                                                           dense_out, closed_form ? 1 : 0);
  MDC_CUDA(cudaGetLastError());
  MDC_CUDA(cudaStreamSynchronize(stream));
  check_barrier_flag();
}

// Per-frame prologue behind the C ABI: image normalise / resize / pad + VAE encoder forward.
// imgs: [N, channels, H, W] uint8 (dtype 0) or fp32 in [0, 1] (dtype 1); latents_out: [N, 4, lh, lw] bf16 NCHW.
inline void Engine::encode(const void* imgs, int dtype, int channels, void* latents_out) {
  MDC_CHECK(dtype == 0 || dtype == 1, "mdc_encode: dtype %d (0 = uint8, 1 = fp32)", dtype);
  MDC_CHECK(channels == 1 || channels == 3, "Input image is not 1- or 3-channel: %d", channels);
  MDC_CHECK(imgs && latents_out, "mdc_encode: null pointer");
  PreGeom g{N, channels, H, W, ph, pw, PPH, PPW, dtype == 0 ? 1 : 0, enc_in->ld};
  const long long tot = 1LL * N * PPH * PPW;
  launch_k(preprocess_image_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, stream, imgs, g, enc_in->d);
  run_ops(enc_ops, false);
  const long long lt = 1LL * N * cfg.vae_latent_ch * lh * lw;
  launch_k(latent_out_kernel, dim3(static_cast<int>((lt + 255) / 256)), dim3(256), 0, stream, enc_out->d, enc_out->ld, N, lh * lw,
           cfg.vae_latent_ch, cfg.vae_scaling, static_cast<bf16*>(latents_out));
  MDC_CUDA(cudaGetLastError());  // stream-ordered: the caller's next use of `latents_out` on the handle's stream sees the result
}

// norm="percentile" (marigold_dc.py:715-728): per sample, torch.quantile of the positive sparse values at (q_lo, q_hi):
// compaction, one radix sort per sample (cub), linear-interpolated ranks.  Prologue only (once per frame).
inline void Engine::percentile_ranges(const float* sparse) {
  const size_t HW = 1ull * H * W;
  if (!pc_vals) {
    pc_vals = arena.make<float>(N * HW + 64), pc_sorted = arena.make<float>(HW + 64);
    pc_range = arena.make<float>(2ull * MAXN), pc_counts = arena.make<int>(MAXN);
    MDC_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, pc_tmp_bytes, pc_vals, pc_sorted, static_cast<int>(HW), 0, 32, stream));
    pc_tmp = arena.make<uint8_t>(pc_tmp_bytes + 256);
    MDC_CUDA(cudaDeviceSynchronize());  // Arena zero-fills on the legacy stream, which is not ordered with a caller's stream
  }
  launch_k(compact_positive_kernel, dim3(N), dim3(1024), 0, stream, sparse, static_cast<int>(HW), pc_vals, pc_counts);
  std::vector<int> cnt(N);
  MDC_CUDA(cudaMemcpyAsync(cnt.data(), pc_counts, 4ull * N, cudaMemcpyDeviceToHost, stream));
  MDC_CUDA(cudaStreamSynchronize(stream));
  for (int n = 0; n < N; ++n) {
    // the reference fails inside torch.quantile for an empty selection ("quantile() input tensor must be non-empty")
    MDC_CHECK(cnt[n] > 0, "quantile() input tensor must be non-empty (sample %d has no valid sparse-depth point)", n);
    size_t tb = pc_tmp_bytes;
    MDC_CUDA(cub::DeviceRadixSort::SortKeys(pc_tmp, tb, pc_vals + n * HW, pc_sorted, cnt[n], 0, 32, stream));
    launch_k(quantile_range_kernel, dim3(1), dim3(32), 0, stream, pc_sorted, cnt[n], q_lo, q_hi, pc_range + 2 * n);
  }
  MDC_CUDA(cudaGetLastError());
}
inline void Engine::set_options(int projection, int inv, int opt, const float* w4, int kld_mode, float kld_weight, float qlo, float qhi,
                                int closed_form, int nearest) {
  if ((nearest ? 1 : 0) != interp_nearest) {  // kernel argument of the captured step: re-capture on the next step
    if (step_graph) cudaGraphExecDestroy(step_graph);
    step_graph = nullptr;
    interp_nearest = nearest ? 1 : 0;
  }
  MDC_CHECK(!(closed_form && w4 && (w4[2] != 0.f || w4[3] != 0.f)),
            "closed_form with edge / smooth losses is not implemented (gradient of the dense terms through the least-squares fit)");
  MDC_CHECK(projection >= 0 && projection <= 2, "Unknown projection method: %d (0 linear, 1 log, 2 log10)", projection);
  MDC_CHECK(opt >= 0 && opt <= 2, "Unknown optimizer: %d (0 adam, 1 sgd, 2 adagrad)", opt);
  MDC_CHECK(kld_mode >= 0 && kld_mode <= 2, "Unknown mode: %d (0 off, 1 simple, 2 strict)", kld_mode);
  MDC_CHECK(w4 && w4[0] + w4[1] + w4[2] + w4[3] > 0.f, "loss_funcs must contain at least one loss function");
  MDC_CHECK(qlo >= 0.f && qlo <= 1.f && qhi >= 0.f && qhi <= 1.f, "percentile must be in [0, 1], but got (%g, %g)", qlo, qhi);
  h_opts.projection = projection, h_opts.inv = inv ? 1 : 0, h_opts.opt = opt, h_opts.kld_mode = kld_mode, h_opts.kld_weight = kld_weight;
  const bool sparse_before = sparse_head_allowed();
  h_opts.w_l1 = w4[0], h_opts.w_l2 = w4[1], h_opts.w_edge = w4[2], h_opts.w_smooth = w4[3];
  if (sparse_before != sparse_head_allowed() && step_graph) {  // the dense losses need the dense head: re-capture
    cudaGraphExecDestroy(step_graph);
    step_graph = nullptr;
  }
  q_lo = qlo, q_hi = qhi;
  h_opts.closed_form = closed_form ? 1 : 0;
}

// One call per frame: image prologue + sparse-depth normalisation + per-call state.
// Raises "No valid values found in mask ..." for a sample without a positive sparse value, like utils.py:132-136.
// `img_latents` != nullptr (mdc_begin_frame_encoded): the image latents were produced earlier by mdc_encode -- e.g. by a
// second handle on another stream while this handle was still running the previous frame --
// and the encoder is skipped; `imgs` is then only needed for the edge loss.
inline void Engine::begin_frame(const void* imgs, int dtype, int channels, const float* sparse, const void* x0,
                                float max_depth, float min_depth, int norm_mode, float lrx, float lrs, const void* img_latents) {
  MDC_CHECK(sparse && x0, "mdc_begin_frame: null pointer");
  MDC_CHECK(norm_mode >= 0 && norm_mode <= 2, "Unknown norm method: %d (0 minmax, 1 const, 2 percentile)", norm_mode);
  MDC_CHECK(!((h_opts.projection != 0 || h_opts.inv) && min_depth <= 1e-7f),
            "min_depth must be > 1e-07 when projection is 'log' or 'log10' or inv is True, but got %g", min_depth);
  if (img_latents) {
    MDC_CHECK(imgs || h_opts.w_edge == 0.f, "image must be provided for edge loss");
    MDC_CHECK(!imgs || ((dtype == 0 || dtype == 1) && (channels == 1 || channels == 3)), "mdc_begin_frame_encoded: bad image format");
  } else {
    encode(imgs, dtype, channels, img_lat);
  }
  if (norm_mode == 2) percentile_ranges(sparse);
  if (h_opts.w_edge != 0.f) {
    const long long opix = 1LL * N * H * W;
    launch_k(gray_grad_kernel, dim3(static_cast<int>((opix + 255) / 256)), dim3(256), 0, stream, imgs, dtype == 0 ? 1 : 0, channels, N, H, W,
             gray_gx, gray_gy);
    have_gray = true;
  }
  launch_k(sparse_norm_kernel, dim3(N), dim3(1024), 0, stream, sparse, H * W, min_depth, max_depth, norm_mode, pc_range, h_opts.projection,
           h_opts.inv, fr_guide, fr_mask, fr_stats);
  MDC_CUDA(cudaGetLastError());
  // ranges, point offsets and the compaction stay on the device; one synchronisation at the end raises the empty-mask error
  begin_state(img_latents ? img_latents : img_lat, x0, fr_guide, fr_mask, lrx, lrs, fr_stats);
}

// NHWC bf16 -> NCHW fp32 copy of a named tensor (which = 0 data, 1 gradient); debug / tests only.
__global__ void nhwc_to_nchw_f32_kernel(const bf16* __restrict__ src, long long ld, int N, int HW, int C,
                                        float* __restrict__ out) {
  ptx::pdl_wait();  // every kernel launched through launch_k must order itself after its predecessor
  ptx::pdl_launch();
  long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x;
  if (i >= 1LL * N * HW * C) return;
  int p = i % HW, c = (i / HW) % C, n = i / (1LL * HW * C);
  out[i] = __bfloat162float(src[(1LL * n * HW + p) * ld + c]);
}
inline void Engine::read_tensor(const std::string& name, int which, float* out_nchw) {
  auto it = named.find(name);
  MDC_CHECK(it != named.end(), "unknown tensor '%s'", name.c_str());
  Tensor* t = it->second;
  const bf16* src = which ? t->g : t->d;
  MDC_CHECK(src != nullptr, "tensor '%s' has no gradient buffer", name.c_str());
  long long tot = t->rows() * t->c;
  launch_k(nhwc_to_nchw_f32_kernel, dim3(static_cast<int>((tot + 255) / 256)), dim3(256), 0, stream, src, t->ld, t->n, t->h * t->w, t->c,
                                                                                  out_nchw);
  MDC_CUDA(cudaGetLastError());
  MDC_CUDA(cudaStreamSynchronize(stream));
}

}  // namespace mdc
