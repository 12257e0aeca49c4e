// C ABI (include/iic.h): engine handle, weight table, forward orchestration.  Host-only logic; every kernel is
// launched through kernels.h.  One handle = one GPU = one CUDA stream user at a time.
#include "../../include/iic.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gemm_sm100.cuh"
#include "kernels.h"

using namespace iic;

namespace iic {
thread_local int g_pdl = 0;
}

namespace {

thread_local std::string g_create_error;

struct Tensor {
  const void* ptr = nullptr;
};

struct LoraSlot {
  const float* a = nullptr;          // f32 [in, r4]   = scaling * lora_A
  const void* bt = nullptr;          // 16-bit [out, lora_pad] = lora_B^T
  int rank = 0, r4 = 0, r_pad = 0;
  // training only
  const void* a16 = nullptr;         // 16-bit [in, lora_pad]  = scaling * lora_A (operand of the dX LoRA k-step)
  const float* bt32 = nullptr;       // f32 [out, r4]          = lora_B^T (operand of dP = dY . B^T)
  float scaling = 1.f;
  float* grad_a = nullptr;           // f32 [in, rank]   d(loss)/d(lora_A)
  float* grad_b = nullptr;           // f32 [rank, out]  d(loss)/d(lora_B)
  // rank > 4: the down-projections run on the tcgen05 GEMM (N = lora_pad) with 16-bit operands (iic_set_lora_operands16)
  const void* at16 = nullptr;        // 16-bit [lora_pad, in]  = (scaling * lora_A)^T, zero-padded rows
  const void* b16 = nullptr;         // 16-bit [lora_pad, out] = lora_B, zero-padded rows (training: dP = dY . B^T)
  // iic_set_lora_source: the fp32 parameters themselves, so that iic_refresh_lora can rebuild every operand above in place
  const float* src_a = nullptr;      // f32 [in, rank]
  const float* src_b = nullptr;      // f32 [rank, out]
};

struct Block {
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  const void *w_qkv = nullptr, *w_out = nullptr, *w_fc = nullptr, *w_proj = nullptr;
  // transposed copies [in, out] for the dX GEMMs of the backward pass (training only)
  const void *w_qkv_t = nullptr, *w_out_t = nullptr, *w_fc_t = nullptr, *w_proj_t = nullptr;
  const float *b_qkv = nullptr, *b_out = nullptr, *b_fc = nullptr, *b_proj = nullptr;
  LoraSlot lora[4];
};

struct Workspace {
  uint16_t *xln, *qkv, *attn, *hid, *p_a, *p_b;
  float *x, *xpre, *down_part, *head_small;
  size_t total;
};

}  // namespace

namespace {
// Optional per-kernel-class CUDA-event timing on the caller's stream (bench.py's roofline line) + launch counter.
enum KClass { kGemm = 0, kLayerNorm = 1, kAttention = 2, kLoraDown = 3, kHead = 4, kPreprocess = 5, kMisc = 6,
              kGemmQkv = 7, kGemmOut = 8, kGemmFc = 9, kGemmProj = 10, kGemmOther = 11, kNumClasses = 12 };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t used = 0;
  long long launches[kNumClasses] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  ~Profiler() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};
struct Scope {  // brackets one launch
  Profiler& p; int cls; cudaStream_t s; cudaEvent_t a = nullptr;
  Scope(Profiler& p_, int cls_, cudaStream_t s_) : p(p_), cls(cls_), s(s_) {
    p.launches[cls]++;
    if (p.on) { a = p.get(); cudaEventRecord(a, s); }
  }
  ~Scope() {
    if (p.on) { cudaEvent_t b = p.get(); cudaEventRecord(b, s); p.spans.push_back({cls, a, b}); }
  }
};
}  // namespace

struct iic_handle {
  Profiler prof;
  iic_config cfg;
  int T, g, patch_k, patch_kpad, lora_pad;
  int num_sms = 0;
  int ctas = 2;
  std::string err;
  const void* conv_w = nullptr;
  int f16 = 0;  // 16-bit operand format of activations and matmul weights: 0 = bf16, 1 = fp16
  int attn_impl = 0;  // 0 auto (whole-row tcgen05 kernel for T <= 208 unmasked, else the block-wise tcgen05 kernel inside its
                      // envelope), 1 mma.sync kernel, 2 block-wise tcgen05 kernel, 3 whole-row tcgen05 kernel (IIC_ATTN_IMPL)
  int train_fused = 1;    // 1: the training forward keeps the c_fc pre-activation (dual-output epilogue) and the c_proj dX GEMM
                          // applies act'(u) in its epilogue; 0 (IIC_TRAIN_FUSED=0): recompute u in the backward + act_bwd kernel
  int lora_bwd_fused = 1; // 1: dB and dP of a LoRA pair from one pass over the output gradient (IIC_LORA_BWD_FUSED=0: GEMM + reduction)
  int pdl_max_batch = 16;  // batches up to this size launch their kernels with programmatic dependent launch (IIC_PDL_MAX_BATCH; 0 = off)
  int attn_bwd_impl = 0;  // 0 auto (tcgen05: one-pass kernel for T <= 256, two-pass kernel up to 592), 1 mma.sync, 2 two-pass tcgen05 (IIC_ATTN_BWD_IMPL)
  float grad_unscale = 1.f;  // 1 / loss scale: folded into the LoRA gradient reductions (iic_train_set_loss_scale)
  const float *cls = nullptr, *pos = nullptr, *lnpre_g = nullptr, *lnpre_b = nullptr, *lnpost_g = nullptr,
              *lnpost_b = nullptr, *proj = nullptr;
  std::vector<Block> blocks;
  // labels
  const float* text = nullptr;
  int L = 0, G = 0, topk = 0;
  float logit_scale = 100.f;
  int *d_group_off = nullptr, *d_group_split = nullptr;
  bool has_split = false;
  PreprocessPlan* pre = nullptr;
};

namespace {

// programmatic dependent launch for the duration of one small-batch call on this thread (kernels.h: g_pdl)
struct PdlScope {
  PdlScope(const iic_handle* h, int B) { g_pdl = (h && h->pdl_max_batch > 0 && B > 0 && B <= h->pdl_max_batch) ? 1 : 0; }
  ~PdlScope() { g_pdl = 0; }
};

int fail(iic_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

Workspace carve(const iic_handle* h, int B, void* base) {
  const size_t M = size_t(B) * h->T;
  const size_t d = h->cfg.width, mlp = h->cfg.mlp_dim;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  Workspace w;
  w.x = static_cast<float*>(take(M * d * 4));
  w.xln = static_cast<uint16_t*>(take(M * d * 2));
  w.qkv = static_cast<uint16_t*>(take(M * 3 * d * 2));
  w.attn = static_cast<uint16_t*>(take(M * d * 2));
  w.hid = static_cast<uint16_t*>(take(M * mlp * 2));  // also hosts x_pre (f32 [M, d]) before ln_pre: mlp*2 >= d*4
  w.p_a = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
  w.p_b = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
  {
    const int act_epi = h->cfg.activation == IIC_ACT_GELU_ERF ? kEpiGeluExactBf16 : kEpiBiasGeluBf16;
    const int parts_wide = 2 * int((mlp + 255) / 256), parts = gemm_down_parts(int(M), int(mlp), act_epi, h->ctas, h->num_sms);
    w.down_part = static_cast<float*>(take(size_t(parts > parts_wide ? parts : parts_wide) * M * 16));
  }
  // small batches: scratch of the multi-CTA head path (launch_head) - embedding, logits and a probability working copy
  {
    const int Bs = B < 16 ? B : 16, Lcap = h->L > 1024 ? h->L : 1024;
    w.head_small = static_cast<float*>(take(head_small_scratch_bytes(Bs, h->cfg.embed_dim, Lcap)));
  }
  w.xpre = reinterpret_cast<float*>(w.hid);
  w.total = off;
  return w;
}

bool check_ready(iic_handle* h) {
  if (!h->lnpost_g || !h->lnpost_b || !h->proj) return false;
  // sequence (text-tower) handles have no patch embedding / class token / ln_pre: the caller supplies the embedded tokens
  if (h->cfg.seq_tokens <= 0 && (!h->conv_w || !h->cls || !h->pos || !h->lnpre_g || !h->lnpre_b)) return false;
  for (const Block& b : h->blocks)
    if (!b.ln1_g || !b.ln1_b || !b.ln2_g || !b.ln2_b || !b.w_qkv || !b.w_out || !b.w_fc || !b.w_proj || !b.b_qkv ||
        !b.b_out || !b.b_fc || !b.b_proj)
      return false;
  return true;
}

// attention forward: tcgen05/TMEM kernel whenever the shape is inside its envelope, mma.sync kernel otherwise.
// lse (nullable): log2-domain log-sum-exp per (image, head, query), kept by the training forward for the backward pass.
int run_attention(iic_handle* h, const void* qkv, void* out, float* lse, int B, int T, int H, int hd, int impl, cudaStream_t s) {
  if (impl == 0) impl = h->attn_impl;
  // impl: 0 auto, 1 mma.sync kernel, 2 block-wise tcgen05 kernel (online softmax), 3 whole-row tcgen05 kernel (T <= 208, no mask)
  if ((impl == 0 || impl == 3) && !h->cfg.causal) {
    int rc = launch_attention_row_sm100(qkv, out, lse, B, T, H, hd, h->f16, h->num_sms, s);
    if (rc != -3 || impl == 3) return rc == -3 ? -1 : rc;
  }
  if (impl == 3) return -1;
  if (impl != 1) {
    int rc = launch_attention_sm100(qkv, out, lse, B, T, H, hd, h->f16, h->cfg.causal != 0, h->num_sms, s);
    if (rc != -3 || impl == 2) return rc == -3 ? -1 : rc;
  }
  if (h->cfg.causal) return -1;   // the mma.sync kernel has no causal mask (text sequences always fit the tcgen05 kernel)
  return launch_attention(qkv, out, lse, B, T, H, hd, h->f16, s);
}

// attention backward: tcgen05 kernel for T <= 592 (needs a [B*H*T] f32 scratch for D), mma.sync kernel otherwise
int run_attention_bwd(iic_handle* h, const void* qkv, const void* out, const void* d_out, const float* lse, float* dsum,
                      void* dqkv, int B, int T, int H, int hd, cudaStream_t s) {
  if ((h->attn_bwd_impl == 0 || h->attn_bwd_impl == 3) && dsum != nullptr && T <= 256) {   // one pass per key tile (attention_bwd_fused_sm100.cu)
    int rc = launch_attention_bwd_dsum(d_out, out, dsum, B, T, H, h->f16, s);
    if (rc == 0) rc = launch_attention_bwd_fused_sm100(qkv, d_out, lse, dsum, dqkv, B, T, H, hd, h->f16, h->cfg.causal, h->num_sms, s);
    if (rc != -3) return rc;
  }
  if (h->attn_bwd_impl != 1 && dsum != nullptr) {
    int rc = launch_attention_bwd_sm100(qkv, out, d_out, lse, dsum, dqkv, B, T, H, hd, h->f16, h->cfg.causal, h->num_sms, s);
    if (rc != -3) return rc;
  }
  if (h->cfg.causal) return -1;   // the mma.sync fallback has no causal mask (text sequences are 77 tokens: never needed)
  return launch_attention_bwd(qkv, out, d_out, lse, dqkv, B, T, H, hd, h->f16, s);
}

int run_gemm(iic_handle* h, const void* a, int lda, const void* w, int M, int N, int K,
             const LoraSlot* lora, const void* p, int epi, const float* bias, const float* residual, void* out,
             int ldc, int group, cudaStream_t s, const float* down_a = nullptr, float* down_part = nullptr,
             int prof_class = kGemm, void* out2 = nullptr) {
  GemmProblem g;
  g.out2 = out2;
  g.down_a = down_a;
  g.down_part = down_part;
  g.a = a; g.lda = lda; g.w = w; g.ldw = K; g.M = M; g.N = N; g.K = K;
  const bool use_lora = lora != nullptr && lora->rank > 0;
  g.lora_p = use_lora ? p : nullptr;
  g.lora_bt = use_lora ? lora->bt : nullptr;
  g.r_pad = use_lora ? lora->r_pad : 0;
  g.lora_ld = h->lora_pad;
  g.epilogue = epi; g.bias = bias; g.residual = residual; g.out = out; g.ldc = ldc; g.group = group;
  g.f16 = h->f16;
  const char* e = nullptr;
  // sub-class by shape (forward: qkv N=3d, out N=K=d, fc N=mlp, proj K=mlp); class 0 stays the total over all GEMMs
  const int d_ = h->cfg.width, mlp_ = h->cfg.mlp_dim;
  const int sub = (epi == kEpiPosF32) ? kGemmOther : (N == 3 * d_ && K == d_) ? kGemmQkv : (N == d_ && K == d_) ? kGemmOut
                  : (N == mlp_ && K == d_) ? kGemmFc : (N == d_ && K == mlp_) ? kGemmProj : kGemmOther;
  // the skinny LoRA down-projection GEMMs are booked under their own class, not under the encoder GEMMs
  int rc;
  if (prof_class == kGemm) {
    Scope sc(h->prof, kGemm, s);
    Scope sc2(h->prof, sub, s);
    rc = launch_gemm(g, h->ctas, h->num_sms, s, &e);
  } else {
    Scope sc(h->prof, prof_class, s);
    rc = launch_gemm(g, h->ctas, h->num_sms, s, &e);
  }
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, e ? e : "gemm failed");
  return 0;
}

template <class F>
int timed(iic_handle* h, int cls, cudaStream_t s, F&& f) {
  Scope sc(h->prof, cls, s);
  return f();
}

// P[M, lora_pad] (16-bit) = x[M, K] . A  for one LoRA slot: rank <= 4 -> the fp32-A row kernel; larger ranks -> the tcgen05
// GEMM with N = lora_pad and the 16-bit transposed operand w16 [lora_pad, K] (memory-bound on reading x either way).
int run_lora_down(iic_handle* h, const void* x, int K, int M, const float* a32, const void* w16, int r4, void* p_out,
                  cudaStream_t s, bool prefer_gemm = false) {
  // prefer_gemm: the backward's dP = dY . B^T has no producer kernel to ride in, and the GEMM (one 256-wide tile column, the
  // rest of the weight tile zero-filled by TMA) reads dY at several TB/s where the row kernel manages a fraction of that
  if ((r4 > 4 || prefer_gemm) && w16 != nullptr)
    return run_gemm(h, x, K, w16, M, h->lora_pad, K, nullptr, nullptr, kEpiBiasBf16, nullptr, nullptr, p_out, h->lora_pad, 1, s,
                    nullptr, nullptr, kLoraDown);
  return timed(h, kLoraDown, s, [&] { return launch_lora_down_bf16(x, K, M, a32, r4, p_out, h->lora_pad, h->f16, s); });
}

#define IIC_TRY(expr)                                                  \
  do {                                                                 \
    int _rc = (expr);                                                  \
    if (_rc != 0) {                                                    \
      if (h->err.empty()) h->err = std::string("failed: ") + #expr;    \
      return _rc < -2 ? _rc : (_rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA); \
    }                                                                  \
  } while (0)

int run_blocks(iic_handle* h, int B, const Workspace& w, cudaStream_t s);

// patches -> residual stream after the last block (x f32 [M, d] in the workspace)
int run_encoder(iic_handle* h, const void* patches, int B, const Workspace& w, cudaStream_t s) {
  const int d = h->cfg.width, T = h->T, M = B * T;
  const int gg = h->g * h->g;
  const float eps = 1e-5f;
  h->err.clear();
  // conv1 as GEMM; epilogue adds positional_embedding[1 + p] and scatters to token rows 1..g*g of each image
  IIC_TRY(run_gemm(h, patches, h->patch_kpad, h->conv_w, B * gg, d, h->patch_kpad, nullptr, nullptr, kEpiPosF32,
                   nullptr, h->pos, w.xpre, d, gg, s));
  IIC_TRY(timed(h, kMisc, s, [&] { return launch_fill_cls(w.xpre, h->cls, h->pos, B, T, d, s); }));
  IIC_TRY(timed(h, kLayerNorm, s, [&] {
    return launch_layernorm(w.xpre, d, h->lnpre_g, h->lnpre_b, nullptr, w.x, d, M, d, eps, nullptr, 0, nullptr, 0, h->f16, s);
  }));
  return run_blocks(h, B, w, s);
}

// the residual blocks on the stream x f32 [M, d] of the workspace (vision tower after ln_pre; text tower after the
// token + positional embedding, with the causal mask of iic_config.causal)
int run_blocks(iic_handle* h, int B, const Workspace& w, cudaStream_t s) {
  const int d = h->cfg.width, T = h->T, M = B * T, mlp = h->cfg.mlp_dim, H = h->cfg.heads;
  const float eps = 1e-5f;
  const int act_epi = h->cfg.activation == IIC_ACT_GELU_ERF ? kEpiGeluExactBf16 : kEpiBiasGeluBf16;
  for (size_t bi = 0; bi < h->blocks.size(); ++bi) {
    Block& b = h->blocks[bi];
    const LoraSlot& l_in = b.lora[IIC_LORA_IN_PROJ];
    const LoraSlot& l_out = b.lora[IIC_LORA_OUT_PROJ];
    const LoraSlot& l_fc = b.lora[IIC_LORA_C_FC];
    const LoraSlot& l_pr = b.lora[IIC_LORA_C_PROJ];
    // x = x + attn(ln_1(x))
    // LoRA down-projections of rank <= 4 ride inside the LayerNorm; larger ranks go through the GEMM (run_lora_down)
    const bool in_gemm = l_in.rank > 0 && l_in.r4 > 4 && l_in.at16 != nullptr;
    IIC_TRY(timed(h, kLayerNorm, s, [&] {
        return launch_layernorm(w.x, d, b.ln1_g, b.ln1_b, w.xln, nullptr, d, M, d, eps, (l_in.rank && !in_gemm) ? l_in.a : nullptr,
                                l_in.r4, w.p_a, h->lora_pad, h->f16, s);
      }));
    if (in_gemm) IIC_TRY(run_lora_down(h, w.xln, d, M, l_in.a, l_in.at16, l_in.r4, w.p_a, s));
    IIC_TRY(run_gemm(h, w.xln, d, b.w_qkv, M, 3 * d, d, &l_in, w.p_a, kEpiBiasBf16, b.b_qkv, nullptr, w.qkv, 3 * d, 1, s));
    IIC_TRY(timed(h, kAttention, s, [&] { return run_attention(h, w.qkv, w.attn, nullptr, B, T, H, d / H, 0, s); }));
    if (l_out.rank) IIC_TRY(run_lora_down(h, w.attn, d, M, l_out.a, l_out.at16, l_out.r4, w.p_b, s));
    // x = x + c_proj(act(c_fc(ln_2(x))))   -- LoRALinear on both (main.py:42-43)
    const bool fc_gemm = l_fc.rank > 0 && l_fc.r4 > 4 && l_fc.at16 != nullptr;
    IIC_TRY(run_gemm(h, w.attn, d, b.w_out, M, d, d, &l_out, w.p_b, kEpiBiasResF32, b.b_out, w.x, w.x, d, 1, s));
    IIC_TRY(timed(h, kLayerNorm, s, [&] {
      return launch_layernorm(w.x, d, b.ln2_g, b.ln2_b, w.xln, nullptr, d, M, d, eps, (l_fc.rank && !fc_gemm) ? l_fc.a : nullptr,
                              l_fc.r4, w.p_a, h->lora_pad, h->f16, s);
    }));
    if (fc_gemm) IIC_TRY(run_lora_down(h, w.xln, d, M, l_fc.a, l_fc.at16, l_fc.r4, w.p_a, s));
    // c_proj's LoRA down-projection (h . A2) rides in the c_fc epilogue while h is still in registers (rank <= 4)
    const bool fuse_down = l_pr.rank > 0 && l_pr.r4 == 4;
    IIC_TRY(run_gemm(h, w.xln, d, b.w_fc, M, mlp, d, &l_fc, w.p_a, act_epi, b.b_fc, nullptr, w.hid, mlp, 1, s,
                     fuse_down ? l_pr.a : nullptr, fuse_down ? w.down_part : nullptr));
    if (fuse_down)
      IIC_TRY(timed(h, kLoraDown, s, [&] {
        return launch_lora_reduce(w.down_part, gemm_down_parts(M, mlp, act_epi, h->ctas, h->num_sms), M, w.p_b, h->lora_pad, h->f16, s);
      }));
    else if (l_pr.rank)
      IIC_TRY(run_lora_down(h, w.hid, mlp, M, l_pr.a, l_pr.at16, l_pr.r4, w.p_b, s));
    IIC_TRY(run_gemm(h, w.hid, mlp, b.w_proj, M, d, mlp, &l_pr, w.p_b, kEpiBiasResF32, b.b_proj, w.x, w.x, d, 1, s));
  }
  return 0;
}

int run_head(iic_handle* h, const float* x_cls, long long x_img_stride, const float* emb_in, int B, float* emb_out,
             const iic_head_out* out, cudaStream_t s, float* small_scratch = nullptr) {
  const bool scores = out != nullptr;
  if (scores && (h->text == nullptr || h->G <= 0)) return fail(h, IIC_ERR_STATE, "iic_set_labels has not been called");
  if (scores && (out->topk_val == nullptr || out->topk_idx == nullptr))
    return fail(h, IIC_ERR_ARG, "iic_head_out.topk_val / topk_idx are required");
  Scope sc(h->prof, kHead, s);
  int rc = launch_head(x_cls, x_img_stride, h->lnpost_g, h->lnpost_b, 1e-5f, h->proj, h->cfg.width, h->cfg.embed_dim,
                       scores ? h->text : nullptr, scores ? h->L : 0, h->d_group_off,
                       h->has_split ? h->d_group_split : nullptr, scores ? h->G : 0, h->topk, h->logit_scale, B, emb_out,
                       scores ? out->logits : nullptr, scores ? out->probs : nullptr, scores ? out->topk_val : nullptr,
                       scores ? out->topk_idx : nullptr, scores ? out->split_sum : nullptr, emb_in,
                       B <= 16 ? small_scratch : nullptr, s);
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "head kernel launch failed");
  return 0;
}


// ------------------------------------------------------------------------------------------------------------------
// training step (LoRA-only gradients): forward that keeps what the backward needs, backward through the frozen blocks
// ------------------------------------------------------------------------------------------------------------------
struct TrainLayer {
  float *x_in, *x_mid;            // f32 [M, d]: inputs of ln_1 / ln_2
  uint16_t *qkv, *attn, *y2;      // 16-bit [M, 3d], [M, d], [M, d] (ln_2 output = c_fc operand)
  uint16_t *p1, *p2;              // 16-bit [M, lora_pad]: s1*(y2.A1), s2*(h.A2)
  uint16_t* u;                    // 16-bit [M, mlp]: c_fc pre-activation (train_fused; else recomputed into the shared buffer)
  float* lse;                     // f32 [B*H, T]
};
struct TrainWorkspace {
  std::vector<TrainLayer> layers;
  float *x, *xpre, *dx, *down_part, *outer_scratch, *attn_d;
  uint16_t *xln, *hid, *u, *dh, *g16, *dy, *dqkv, *da, *dp1, *dp2;
  size_t total;
};

TrainWorkspace carve_train(const iic_handle* h, int B, void* base) {
  const size_t M = size_t(B) * h->T;
  const size_t d = h->cfg.width, mlp = h->cfg.mlp_dim, H = h->cfg.heads;
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = p ? p + off : nullptr;
    off += align_up(bytes, 1024);
    return r;
  };
  TrainWorkspace w;
  w.x = static_cast<float*>(take(M * d * 4));
  w.dx = static_cast<float*>(take(M * d * 4));
  w.xln = static_cast<uint16_t*>(take(M * d * 2));
  w.hid = static_cast<uint16_t*>(take(M * mlp * 2));
  w.u = static_cast<uint16_t*>(take(M * mlp * 2));
  w.dh = static_cast<uint16_t*>(take(M * mlp * 2));
  w.g16 = static_cast<uint16_t*>(take(M * d * 2));
  w.dy = static_cast<uint16_t*>(take(M * d * 2));
  w.dqkv = static_cast<uint16_t*>(take(M * 3 * d * 2));
  w.da = static_cast<uint16_t*>(take(M * d * 2));
  w.dp1 = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
  w.dp2 = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
  {
    const int act_epi = h->cfg.activation == IIC_ACT_GELU_ERF ? kEpiGeluExactBf16 : kEpiBiasGeluBf16;
    const int parts_wide = 2 * int((mlp + 255) / 256), parts = gemm_down_parts(int(M), int(mlp), act_epi, h->ctas, h->num_sms);
    w.down_part = static_cast<float*>(take(size_t(parts > parts_wide ? parts : parts_wide) * M * 16));
  }
  {   // one scratch for every LoRA gradient reduction of the step: the CTA shapes (and so the partial counts) depend on N
    size_t sc = lora_outer_scratch_bytes(int(mlp), int(M));
    for (size_t n : {d, 3 * d})
      if (lora_outer_scratch_bytes(int(n), int(M)) > sc) sc = lora_outer_scratch_bytes(int(n), int(M));
    w.outer_scratch = static_cast<float*>(take(sc));
  }
  w.attn_d = static_cast<float*>(take(size_t(B) * H * h->T * 4));   // D = rowsum(dO o O) of the attention backward
  w.xpre = reinterpret_cast<float*>(w.hid);
  w.layers.resize(h->blocks.size());
  for (TrainLayer& l : w.layers) {
    l.x_in = static_cast<float*>(take(M * d * 4));
    l.x_mid = static_cast<float*>(take(M * d * 4));
    l.qkv = static_cast<uint16_t*>(take(M * 3 * d * 2));
    l.attn = static_cast<uint16_t*>(take(M * d * 2));
    l.y2 = static_cast<uint16_t*>(take(M * d * 2));
    l.p1 = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
    l.p2 = static_cast<uint16_t*>(take(M * h->lora_pad * 2 + 4096));
    l.lse = static_cast<float*>(take(size_t(B) * H * h->T * 4));
    l.u = h->train_fused ? static_cast<uint16_t*>(take(M * mlp * 2)) : w.u;
  }
  w.total = off;
  return w;
}

bool check_train_ready(iic_handle* h) {
  for (const Block& b : h->blocks) {
    if (!b.w_qkv_t || !b.w_out_t || !b.w_fc_t || !b.w_proj_t) return false;
    for (int w : {IIC_LORA_C_FC, IIC_LORA_C_PROJ})
      if (b.lora[w].rank > 0 && (!b.lora[w].a16 || !b.lora[w].bt32 || !b.lora[w].grad_a || !b.lora[w].grad_b)) return false;
  }
  return true;
}

int copy_rows_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows, cudaStream_t s) {
  return cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? 0 : -2;
}

// forward in training mode: same kernels as inference, per-layer activations kept
// vision tower: `patches` -> class-token rows.  Sequence (text) handles: x_tokens f32 [B*T, d] (token + positional embedding)
// -> the rows row_index[b] (EOT) of the final residual stream.
int run_train_forward(iic_handle* h, const void* patches, const float* x_tokens, const int32_t* row_index, int B,
                      TrainWorkspace& w, float* x_cls_out, cudaStream_t s) {
  const int d = h->cfg.width, T = h->T, M = B * T, mlp = h->cfg.mlp_dim, H = h->cfg.heads;
  const int gg = h->g * h->g;
  const float eps = 1e-5f;
  h->err.clear();
  // The residual stream is never copied: every residual epilogue writes straight into the buffer the backward pass will read
  // (block l: x_in(l) -> out_proj -> x_mid(l) -> c_proj -> x_in(l+1); the last block writes the workspace stream w.x).
  if (x_tokens != nullptr) {
    if (cudaMemcpyAsync(w.layers[0].x_in, x_tokens, size_t(M) * d * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return fail(h, IIC_ERR_CUDA, "copy failed");
  } else {
    IIC_TRY(run_gemm(h, patches, h->patch_kpad, h->conv_w, B * gg, d, h->patch_kpad, nullptr, nullptr, kEpiPosF32,
                     nullptr, h->pos, w.xpre, d, gg, s));
    IIC_TRY(timed(h, kMisc, s, [&] { return launch_fill_cls(w.xpre, h->cls, h->pos, B, T, d, s); }));
    IIC_TRY(timed(h, kLayerNorm, s, [&] {
      return launch_layernorm(w.xpre, d, h->lnpre_g, h->lnpre_b, nullptr, w.layers[0].x_in, d, M, d, eps, nullptr, 0, nullptr, 0, h->f16, s);
    }));
  }
  const int act_epi = h->cfg.activation == IIC_ACT_GELU_ERF ? kEpiGeluExactBf16 : kEpiBiasGeluBf16;
  for (size_t li = 0; li < h->blocks.size(); ++li) {
    Block& b = h->blocks[li];
    TrainLayer& t = w.layers[li];
    float* x_next = li + 1 < h->blocks.size() ? w.layers[li + 1].x_in : w.x;
    const LoraSlot& l_fc = b.lora[IIC_LORA_C_FC];
    const LoraSlot& l_pr = b.lora[IIC_LORA_C_PROJ];
    if (b.lora[IIC_LORA_IN_PROJ].rank || b.lora[IIC_LORA_OUT_PROJ].rank)
      return fail(h, IIC_ERR_ARG, "training supports LoRA on mlp.c_fc / mlp.c_proj (what the reference's wrap makes effective)");
    IIC_TRY(timed(h, kLayerNorm, s, [&] {
      return launch_layernorm(t.x_in, d, b.ln1_g, b.ln1_b, w.xln, nullptr, d, M, d, eps, nullptr, 0, nullptr, 0, h->f16, s);
    }));
    IIC_TRY(run_gemm(h, w.xln, d, b.w_qkv, M, 3 * d, d, nullptr, nullptr, kEpiBiasBf16, b.b_qkv, nullptr, t.qkv, 3 * d, 1, s));
    IIC_TRY(timed(h, kAttention, s, [&] { return run_attention(h, t.qkv, t.attn, t.lse, B, T, H, d / H, 0, s); }));
    IIC_TRY(run_gemm(h, t.attn, d, b.w_out, M, d, d, nullptr, nullptr, kEpiBiasResF32, b.b_out, t.x_in, t.x_mid, d, 1, s));
    const bool fc_gemm = l_fc.rank > 0 && l_fc.r4 > 4 && l_fc.at16 != nullptr;
    IIC_TRY(timed(h, kLayerNorm, s, [&] {
      return launch_layernorm(t.x_mid, d, b.ln2_g, b.ln2_b, t.y2, nullptr, d, M, d, eps, (l_fc.rank && !fc_gemm) ? l_fc.a : nullptr,
                              l_fc.r4, t.p1, h->lora_pad, h->f16, s);
    }));
    if (fc_gemm) IIC_TRY(run_lora_down(h, t.y2, d, M, l_fc.a, l_fc.at16, l_fc.r4, t.p1, s));
    const bool fuse_down = l_pr.rank > 0 && l_pr.r4 == 4;
    if (h->train_fused)   // h -> hid (c_proj operand), u -> kept for the backward
      IIC_TRY(run_gemm(h, t.y2, d, b.w_fc, M, mlp, d, &l_fc, t.p1, kEpiBiasActDualBf16, b.b_fc, nullptr, w.hid, mlp,
                       act_epi == kEpiGeluExactBf16 ? 2 : 1, s, fuse_down ? l_pr.a : nullptr, fuse_down ? w.down_part : nullptr,
                       kGemm, t.u));
    else
      IIC_TRY(run_gemm(h, t.y2, d, b.w_fc, M, mlp, d, &l_fc, t.p1, act_epi, b.b_fc, nullptr, w.hid, mlp, 1, s,
                       fuse_down ? l_pr.a : nullptr, fuse_down ? w.down_part : nullptr));
    if (fuse_down)
      IIC_TRY(timed(h, kLoraDown, s, [&] {
        return launch_lora_reduce(w.down_part, gemm_down_parts(M, mlp, h->train_fused ? int(kEpiBiasActDualBf16) : act_epi, h->ctas, h->num_sms), M,
                                  t.p2, h->lora_pad, h->f16, s);
      }));
    else if (l_pr.rank)
      IIC_TRY(run_lora_down(h, w.hid, mlp, M, l_pr.a, l_pr.at16, l_pr.r4, t.p2, s));
    IIC_TRY(run_gemm(h, w.hid, mlp, b.w_proj, M, d, mlp, &l_pr, t.p2, kEpiBiasResF32, b.b_proj, t.x_mid, x_next, d, 1, s));
  }
  if (row_index != nullptr) {   // the EOT rows (input of ln_final)
    Scope sc(h->prof, kMisc, s);
    return launch_gather_rows(w.x, row_index, T, d, B, x_cls_out, s) ? fail(h, IIC_ERR_CUDA, "row gather launch failed") : 0;
  }
  // class-token rows of the final residual stream (input of ln_post), [B, d] contiguous
  return copy_rows_async(x_cls_out, size_t(d) * 4, w.x, size_t(T) * d * 4, size_t(d) * 4, size_t(B), s) ? fail(h, IIC_ERR_CUDA, "copy failed") : 0;
}

int run_train_backward_begin(iic_handle* h, int B, TrainWorkspace& w, const float* dx_cls, const int32_t* row_index, cudaStream_t s) {
  const int d = h->cfg.width, T = h->T, M = B * T;
  h->err.clear();
  // dx = 0 except the class-token rows (sequence handles: the rows row_index[b])
  if (cudaMemsetAsync(w.dx, 0, size_t(M) * d * 4, s) != cudaSuccess) return fail(h, IIC_ERR_CUDA, "memset failed");
  if (row_index != nullptr) {
    if (launch_scatter_rows(dx_cls, row_index, T, d, B, w.dx, s)) return fail(h, IIC_ERR_CUDA, "row scatter launch failed");
  } else if (copy_rows_async(w.dx, size_t(T) * d * 4, dx_cls, size_t(d) * 4, size_t(d) * 4, size_t(B), s)) {
    return fail(h, IIC_ERR_CUDA, "copy failed");
  }
  IIC_TRY(timed(h, kMisc, s, [&] { return launch_cast16(w.dx, w.g16, (long long)M * d, h->f16, s); }));
  return 0;
}

// backward through block `li` (call for li = layers-1 ... 0 after run_train_backward_begin); its LoRA gradients are final on return
int run_train_backward_layer(iic_handle* h, int B, TrainWorkspace& w, int li, cudaStream_t s) {
  const int d = h->cfg.width, T = h->T, M = B * T, mlp = h->cfg.mlp_dim, H = h->cfg.heads;
  const float eps = 1e-5f;
  const int act = h->cfg.activation == IIC_ACT_GELU_ERF ? 2 : 1;
  LoraSlot none;
  {
    Block& b = h->blocks[li];
    TrainLayer& t = w.layers[li];
    const LoraSlot& l_fc = b.lora[IIC_LORA_C_FC];
    const LoraSlot& l_pr = b.lora[IIC_LORA_C_PROJ];
    // ---- c_proj:  x_out = x_mid + h W2^T + b2 + P2 B2   (P2 = s2 h A2) ----
    LoraSlot bw_pr;   // LoRA k-step of the dX GEMM: dh += dP2 . (s2 A2)^T
    // dP = dOut . B^T and dB = P^T . dOut of one LoRA pair: one pass over dOut (lora_bwd_kernel), else skinny GEMM + reduction
    auto lora_dp_db = [&](const LoraSlot& l, const void* p_fwd, const void* dout, int n_out, void* dp_out) -> int {
      int rc = -3;
      if (h->lora_bwd_fused && l.b16 != nullptr)
        rc = timed(h, kLoraDown, s, [&] {
          return launch_lora_bwd(p_fwd, h->lora_pad, dout, n_out, M, l.b16, l.rank, h->grad_unscale, l.grad_b, dp_out, w.outer_scratch,
                                 h->f16, s);
        });
      if (rc != -3) return rc;
      IIC_TRY(run_lora_down(h, dout, n_out, M, l.bt32, l.b16, l.r4, dp_out, s, true));
      return timed(h, kLoraDown, s, [&] {
        return launch_lora_outer(p_fwd, h->lora_pad, dout, n_out, M, 0, l.rank, h->grad_unscale, 0, l.grad_b, w.outer_scratch, h->f16, s);
      });
    };
    if (l_pr.rank) {
      IIC_TRY(lora_dp_db(l_pr, t.p2, w.g16, d, w.dp2));   // dP2 = dY . B2^T, dB2 = P2^T . dY
      bw_pr.rank = l_pr.rank; bw_pr.r4 = l_pr.r4; bw_pr.r_pad = l_pr.r_pad; bw_pr.bt = l_pr.a16;
    }
    // the pre-activation u = y2 W1^T + b1 + P1 B1: kept by the forward (train_fused; 2 bytes x M x 4d per layer is small
    // change on a 180 GB part) or recomputed here
    if (!h->train_fused)
      IIC_TRY(run_gemm(h, t.y2, d, b.w_fc, M, mlp, d, &l_fc, t.p1, kEpiBiasBf16, b.b_fc, nullptr, t.u, mlp, 1, s));
    if (l_pr.rank)
      IIC_TRY(timed(h, kLoraDown, s, [&] {
        return launch_lora_outer(w.dp2, h->lora_pad, t.u, mlp, M, act, l_pr.rank, l_pr.scaling * h->grad_unscale, 1, l_pr.grad_a, w.outer_scratch, h->f16, s);
      }));
    if (h->train_fused) {   // dh := du = (dY . W2 + dP2 . (s2 A2)^T) o act'(u) in the GEMM epilogue
      IIC_TRY(run_gemm(h, w.g16, d, b.w_proj_t, M, mlp, d, &bw_pr, w.dp2, kEpiActGradBf16, nullptr, reinterpret_cast<const float*>(t.u),
                       w.dh, mlp, act, s));
    } else {
      IIC_TRY(run_gemm(h, w.g16, d, b.w_proj_t, M, mlp, d, &bw_pr, w.dp2, kEpiBiasBf16, nullptr, nullptr, w.dh, mlp, 1, s));
      IIC_TRY(timed(h, kMisc, s, [&] { return launch_act_bwd(w.dh, t.u, (long long)M * mlp, act, h->f16, s); }));   // dh := du
    }
    // ---- c_fc:  u = y2 W1^T + b1 + P1 B1   (P1 = s1 y2 A1) ----
    LoraSlot bw_fc;
    if (l_fc.rank) {
      IIC_TRY(lora_dp_db(l_fc, t.p1, w.dh, mlp, w.dp1));   // dP1 = dU . B1^T, dB1 = P1^T . dU
      IIC_TRY(timed(h, kLoraDown, s, [&] {
        return launch_lora_outer(w.dp1, h->lora_pad, t.y2, d, M, 0, l_fc.rank, l_fc.scaling * h->grad_unscale, 1, l_fc.grad_a, w.outer_scratch, h->f16, s);
      }));
      bw_fc.rank = l_fc.rank; bw_fc.r4 = l_fc.r4; bw_fc.r_pad = l_fc.r_pad; bw_fc.bt = l_fc.a16;
    }
    IIC_TRY(run_gemm(h, w.dh, mlp, b.w_fc_t, M, d, mlp, &bw_fc, w.dp1, kEpiBiasBf16, nullptr, nullptr, w.dy, d, 1, s));
    IIC_TRY(timed(h, kLayerNorm, s, [&] { return launch_layernorm_bwd(w.dy, t.x_mid, b.ln2_g, w.dx, w.g16, M, d, eps, h->f16, s); }));
    if (li == 0) return 0;   // nothing below the first block's MLP carries a LoRA parameter
    // ---- attention block:  x_mid = x_in + attn(ln_1(x_in)) W_o^T + b_o ----
    IIC_TRY(run_gemm(h, w.g16, d, b.w_out_t, M, d, d, &none, nullptr, kEpiBiasBf16, nullptr, nullptr, w.da, d, 1, s));
    IIC_TRY(timed(h, kAttention, s, [&] { return run_attention_bwd(h, t.qkv, t.attn, w.da, t.lse, w.attn_d, w.dqkv, B, T, H, d / H, s); }));
    IIC_TRY(run_gemm(h, w.dqkv, 3 * d, b.w_qkv_t, M, d, 3 * d, &none, nullptr, kEpiBiasBf16, nullptr, nullptr, w.dy, d, 1, s));
    IIC_TRY(timed(h, kLayerNorm, s, [&] { return launch_layernorm_bwd(w.dy, t.x_in, b.ln1_g, w.dx, w.g16, M, d, eps, h->f16, s); }));
  }
  return 0;
}

}  // namespace

extern "C" {

const char* iic_version(void) { return "iic-b200 0.1 (sm_100a; tcgen05 + TMA)"; }

int iic_create(iic_handle** out, const iic_config* cfg) {
  if (out == nullptr || cfg == nullptr) { g_create_error = "iic_create: null argument"; return IIC_ERR_ARG; }
  *out = nullptr;
  if (cfg->image_size <= 0 || cfg->patch_size <= 0 || cfg->image_size % cfg->patch_size != 0 || cfg->width <= 0 ||
      cfg->width % 128 != 0 || cfg->heads <= 0 || cfg->width != cfg->heads * 64 || cfg->layers <= 0 ||
      cfg->mlp_dim % 256 != 0 || cfg->mlp_dim * 2 < cfg->width * 4 || cfg->embed_dim <= 0 || cfg->width % 256 != 0 ||
      (cfg->operand_dtype != IIC_DTYPE_BF16 && cfg->operand_dtype != IIC_DTYPE_F16 && cfg->operand_dtype != 0)) {
    g_create_error = "iic_create: unsupported architecture (need head_dim 64, width % 256 == 0, mlp % 256 == 0)";
    return IIC_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    g_create_error = "iic_create: no CUDA device (this library has no CPU path)";
    return IIC_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "iic_create: bad device ordinal"; return IIC_ERR_ARG; }
  cudaDeviceProp prop;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
    g_create_error = "iic_create: cudaSetDevice failed";
    return IIC_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "iic_create: device is not sm_100 (Blackwell B200); kernels are built for sm_100a only";
    return IIC_ERR_CUDA;
  }
  iic_handle* h = new iic_handle();
  h->cfg = *cfg;
  h->g = cfg->image_size / cfg->patch_size;
  h->T = cfg->seq_tokens > 0 ? cfg->seq_tokens : h->g * h->g + 1;
  h->patch_k = 3 * cfg->patch_size * cfg->patch_size;
  h->patch_kpad = (h->patch_k + 7) / 8 * 8;
  h->lora_pad = 16;
  h->num_sms = prop.multiProcessorCount;
  h->ctas = cfg->gemm_ctas == 1 ? 1 : 2;
  h->f16 = cfg->operand_dtype == IIC_DTYPE_F16 ? 1 : 0;
  if (const char* e = getenv("IIC_ATTN_IMPL")) h->attn_impl = atoi(e);
  if (const char* e = getenv("IIC_ATTN_BWD_IMPL")) h->attn_bwd_impl = atoi(e);
  if (const char* e = getenv("IIC_PDL_MAX_BATCH")) h->pdl_max_batch = atoi(e);
  if (const char* e = getenv("IIC_TRAIN_FUSED")) h->train_fused = atoi(e);
  if (const char* e = getenv("IIC_LORA_BWD_FUSED")) h->lora_bwd_fused = atoi(e);
  if (const char* e = getenv("IIC_GEMM_CTAS")) { if (atoi(e) == 1) h->ctas = 1; else if (atoi(e) == 2) h->ctas = 2; }
  h->blocks.resize(cfg->layers);
  h->pre = preprocess_plan_create();
  if (h->pre == nullptr) { delete h; g_create_error = "iic_create: preprocess plan allocation failed"; return IIC_ERR_CUDA; }
  *out = h;
  return IIC_OK;
}

void iic_destroy(iic_handle* h) {
  if (h == nullptr) return;
  if (h->d_group_off) cudaFree(h->d_group_off);
  if (h->d_group_split) cudaFree(h->d_group_split);
  preprocess_plan_destroy(h->pre);
  delete h;
}

const char* iic_last_error(const iic_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int iic_get_dims(const iic_handle* h, iic_dims* out) {
  if (!h || !out) return IIC_ERR_ARG;
  out->tokens = h->T; out->grid = h->g; out->patch_k = h->patch_k; out->patch_kpad = h->patch_kpad;
  out->lora_pad = h->lora_pad;
  return IIC_OK;
}

int iic_load_weight(iic_handle* h, const char* name, const void* dev_ptr, int dtype, int ndim, const int64_t* shape) {
  if (!h || !name || !dev_ptr || !shape) return fail(h, IIC_ERR_ARG, "iic_load_weight: null argument");
  const int d = h->cfg.width, mlp = h->cfg.mlp_dim, E = h->cfg.embed_dim;
  auto want = [&](int dt, int nd, int64_t s0, int64_t s1) -> bool {
    if (dtype != dt || ndim != nd || shape[0] != s0 || (nd == 2 && shape[1] != s1)) {
      char buf[256];
      snprintf(buf, sizeof buf, "iic_load_weight(%s): expected dtype %d shape [%lld%s%lld], got dtype %d ndim %d [%lld,..]",
               name, dt, (long long)s0, nd == 2 ? "," : "", nd == 2 ? (long long)s1 : 0ll, dtype, ndim,
               (long long)shape[0]);
      h->err = buf;
      return false;
    }
    if ((reinterpret_cast<uintptr_t>(dev_ptr) & 15) != 0) { h->err = std::string(name) + ": pointer must be 16-byte aligned"; return false; }
    return true;
  };
  const std::string n(name);
  const int wdt = h->f16 ? IIC_DTYPE_F16 : IIC_DTYPE_BF16;  // matmul weights share the activation operand format
#define IIC_SET(field, type, dt, nd, s0, s1)                              \
  do {                                                                    \
    if (!want(dt, nd, s0, s1)) return IIC_ERR_ARG;                        \
    field = static_cast<type>(dev_ptr);                                   \
    return IIC_OK;                                                        \
  } while (0)
  if (n == "conv1.weight") IIC_SET(h->conv_w, const void*, wdt, 2, d, h->patch_kpad);
  if (n == "class_embedding") IIC_SET(h->cls, const float*, IIC_DTYPE_F32, 1, d, 0);
  if (n == "positional_embedding") IIC_SET(h->pos, const float*, IIC_DTYPE_F32, 2, h->T, d);
  if (n == "ln_pre.weight") IIC_SET(h->lnpre_g, const float*, IIC_DTYPE_F32, 1, d, 0);
  if (n == "ln_pre.bias") IIC_SET(h->lnpre_b, const float*, IIC_DTYPE_F32, 1, d, 0);
  if (n == "ln_post.weight") IIC_SET(h->lnpost_g, const float*, IIC_DTYPE_F32, 1, d, 0);
  if (n == "ln_post.bias") IIC_SET(h->lnpost_b, const float*, IIC_DTYPE_F32, 1, d, 0);
  if (n == "proj") IIC_SET(h->proj, const float*, IIC_DTYPE_F32, 2, d, E);
  const std::string pre = "transformer.resblocks.";
  if (n.compare(0, pre.size(), pre) == 0) {
    const size_t dot = n.find('.', pre.size());
    if (dot != std::string::npos) {
      const int li = atoi(n.substr(pre.size(), dot - pre.size()).c_str());
      if (li >= 0 && li < int(h->blocks.size())) {
        Block& b = h->blocks[li];
        const std::string r = n.substr(dot + 1);
        if (r == "ln_1.weight") IIC_SET(b.ln1_g, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "ln_1.bias") IIC_SET(b.ln1_b, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "ln_2.weight") IIC_SET(b.ln2_g, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "ln_2.bias") IIC_SET(b.ln2_b, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "attn.in_proj_weight") IIC_SET(b.w_qkv, const void*, wdt, 2, 3 * d, d);
        if (r == "attn.in_proj_bias") IIC_SET(b.b_qkv, const float*, IIC_DTYPE_F32, 1, 3 * d, 0);
        if (r == "attn.out_proj.weight") IIC_SET(b.w_out, const void*, wdt, 2, d, d);
        if (r == "attn.out_proj.bias") IIC_SET(b.b_out, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "mlp.c_fc.weight") IIC_SET(b.w_fc, const void*, wdt, 2, mlp, d);
        if (r == "mlp.c_fc.bias") IIC_SET(b.b_fc, const float*, IIC_DTYPE_F32, 1, mlp, 0);
        if (r == "mlp.c_proj.weight") IIC_SET(b.w_proj, const void*, wdt, 2, d, mlp);
        if (r == "mlp.c_proj.bias") IIC_SET(b.b_proj, const float*, IIC_DTYPE_F32, 1, d, 0);
        if (r == "attn.in_proj_weight_t") IIC_SET(b.w_qkv_t, const void*, wdt, 2, d, 3 * d);
        if (r == "attn.out_proj.weight_t") IIC_SET(b.w_out_t, const void*, wdt, 2, d, d);
        if (r == "mlp.c_fc.weight_t") IIC_SET(b.w_fc_t, const void*, wdt, 2, d, mlp);
        if (r == "mlp.c_proj.weight_t") IIC_SET(b.w_proj_t, const void*, wdt, 2, mlp, d);
      }
    }
  }
#undef IIC_SET
  return fail(h, IIC_ERR_ARG, std::string("iic_load_weight: unknown weight name ") + name);
}

int iic_set_lora(iic_handle* h, int layer, int which, const float* a_scaled, const void* b_t, int rank) {
  if (!h) return IIC_ERR_ARG;
  if (layer < 0 || layer >= int(h->blocks.size()) || which < 0 || which > 3)
    return fail(h, IIC_ERR_ARG, "iic_set_lora: bad layer / projection id");
  LoraSlot& s = h->blocks[layer].lora[which];
  if (rank <= 0) { s = LoraSlot(); return IIC_OK; }
  if (rank > h->lora_pad) return fail(h, IIC_ERR_ARG, "iic_set_lora: rank exceeds iic_dims.lora_pad (16)");
  if (!a_scaled || !b_t || (reinterpret_cast<uintptr_t>(a_scaled) & 15) || (reinterpret_cast<uintptr_t>(b_t) & 15))
    return fail(h, IIC_ERR_ARG, "iic_set_lora: null or unaligned pointer");
  s.a = a_scaled;
  s.bt = b_t;
  s.rank = rank;
  s.r4 = (rank + 3) / 4 * 4;
  s.r_pad = (rank + 15) / 16 * 16;
  return IIC_OK;
}

int iic_set_lora_operands16(iic_handle* h, int layer, int which, const void* a_t16, const void* b16) {
  if (!h) return IIC_ERR_ARG;
  if (layer < 0 || layer >= int(h->blocks.size()) || which < 0 || which > 3)
    return fail(h, IIC_ERR_ARG, "iic_set_lora_operands16: bad layer / projection id");
  LoraSlot& s = h->blocks[layer].lora[which];
  if (s.rank <= 0) return fail(h, IIC_ERR_STATE, "iic_set_lora_operands16: call iic_set_lora for this slot first");
  if ((reinterpret_cast<uintptr_t>(a_t16) & 15) || (reinterpret_cast<uintptr_t>(b16) & 15))
    return fail(h, IIC_ERR_ARG, "iic_set_lora_operands16: unaligned pointer");
  s.at16 = a_t16;
  s.b16 = b16;
  return IIC_OK;
}

int iic_set_lora_source(iic_handle* h, int layer, int which, const float* lora_a, const float* lora_b, float scaling) {
  if (!h) return IIC_ERR_ARG;
  if (layer < 0 || layer >= int(h->blocks.size()) || which < 0 || which > 3)
    return fail(h, IIC_ERR_ARG, "iic_set_lora_source: bad layer / projection id");
  LoraSlot& s = h->blocks[layer].lora[which];
  if (s.rank <= 0) return fail(h, IIC_ERR_STATE, "iic_set_lora_source: call iic_set_lora for this slot first");
  s.src_a = lora_a;
  s.src_b = lora_b;
  s.scaling = scaling;
  return IIC_OK;
}

int iic_refresh_lora(iic_handle* h, void* stream) {
  if (!h) return IIC_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = h->cfg.width, mlp = h->cfg.mlp_dim;
  const int dims[4][2] = {{d, 3 * d}, {d, d}, {d, mlp}, {mlp, d}};   // (in, out) of in_proj, out_proj, c_fc, c_proj
  // up to 16 slots per launch: the operand buffers are caller-owned device memory registered through iic_set_lora / _train /
  // _operands16
  LoraRefreshBatch batch;
  batch.n = 0;
  batch.pad = h->lora_pad;
  auto flush = [&]() -> int {
    if (batch.n == 0) return 0;
    Scope sc(h->prof, kMisc, st);
    int rc = launch_lora_refresh_batch(batch, h->f16, st);
    batch.n = 0;
    return rc;
  };
  for (Block& b : h->blocks)
    for (int which = 0; which < 4; ++which) {
      LoraSlot& s = b.lora[which];
      if (s.rank <= 0 || s.src_a == nullptr || s.src_b == nullptr) continue;
      LoraRefreshSlot& q = batch.slot[batch.n++];
      q.A = s.src_a; q.B = s.src_b;
      q.a = const_cast<float*>(s.a); q.bt = const_cast<void*>(s.bt); q.a16 = const_cast<void*>(s.a16);
      q.bt32 = const_cast<float*>(s.bt32); q.at16 = const_cast<void*>(s.at16); q.b16 = const_cast<void*>(s.b16);
      q.in = dims[which][0]; q.out = dims[which][1]; q.rank = s.rank; q.r4 = s.r4; q.scaling = s.scaling;
      if (batch.n == 16) {
        int rc = flush();
        if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "iic_refresh_lora: launch failed");
      }
    }
  {
    int rc = flush();
    if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "iic_refresh_lora: launch failed");
  }
  return IIC_OK;
}

int iic_set_labels(iic_handle* h, const float* text, int num_labels, const int* group_offsets, const int* group_split,
                   int num_groups, int topk, float logit_scale) {
  if (!h) return IIC_ERR_ARG;
  if (!text || num_labels <= 0 || !group_offsets || num_groups <= 0 || topk <= 0 || topk > 8)
    return fail(h, IIC_ERR_ARG, "iic_set_labels: bad argument (need L > 0, G > 0, 1 <= topk <= 8)");
  if (group_offsets[0] != 0 || group_offsets[num_groups] != num_labels)
    return fail(h, IIC_ERR_ARG, "iic_set_labels: group_offsets must start at 0 and end at num_labels");
  for (int g = 0; g < num_groups; ++g)
    if (group_offsets[g + 1] <= group_offsets[g]) return fail(h, IIC_ERR_ARG, "iic_set_labels: empty or unordered group");
  if (cudaSetDevice(h->cfg.device) != cudaSuccess) return fail(h, IIC_ERR_CUDA, "cudaSetDevice failed");
  if (h->d_group_off) cudaFree(h->d_group_off);
  if (h->d_group_split) cudaFree(h->d_group_split);
  h->d_group_off = h->d_group_split = nullptr;
  std::vector<int> split(num_groups, 0);
  if (group_split) split.assign(group_split, group_split + num_groups);
  if (cudaMalloc(&h->d_group_off, sizeof(int) * (num_groups + 1)) != cudaSuccess ||
      cudaMalloc(&h->d_group_split, sizeof(int) * num_groups) != cudaSuccess ||
      cudaMemcpy(h->d_group_off, group_offsets, sizeof(int) * (num_groups + 1), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->d_group_split, split.data(), sizeof(int) * num_groups, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(h, IIC_ERR_CUDA, "iic_set_labels: device allocation/copy failed");
  h->has_split = group_split != nullptr;
  h->text = text; h->L = num_labels; h->G = num_groups; h->topk = topk; h->logit_scale = logit_scale;
  return IIC_OK;
}

int iic_preprocess(iic_handle* h, const uint8_t* const* imgs, const int* hw, int B, void* out, int out_layout,
                   void* stream) {
  if (!h || !imgs || !hw || !out || B < 0 || out_layout < 0 || out_layout > 2)
    return fail(h, IIC_ERR_ARG, "iic_preprocess: bad argument");
  const char* e = nullptr;
  Scope sc(h->prof, kPreprocess, static_cast<cudaStream_t>(stream));
  h->prof.launches[kPreprocess]++;  // two kernels (horizontal + vertical pass)
  int rc = launch_preprocess(h->pre, imgs, hw, B, h->cfg.image_size, h->cfg.patch_size, h->patch_kpad, out, out_layout,
                             h->f16, static_cast<cudaStream_t>(stream), &e);
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, e ? e : "preprocess failed");
  return IIC_OK;
}

int iic_preprocess_same_size(iic_handle* h, const uint8_t* imgs, int B, void* out, int out_layout, void* stream) {
  if (!h || !imgs || !out || B < 0 || out_layout < 0 || out_layout > 2)
    return fail(h, IIC_ERR_ARG, "iic_preprocess_same_size: bad argument");
  PdlScope pdl(h, B);
  Scope sc(h->prof, kPreprocess, static_cast<cudaStream_t>(stream));
  int rc = launch_preprocess_fast(h->pre, imgs, B, h->cfg.image_size, h->cfg.patch_size, h->patch_kpad, out, out_layout,
                                  h->f16, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "preprocess_same_size launch failed");
  return IIC_OK;
}

int iic_patchify(iic_handle* h, const void* chw, int dtype, int B, void* patches_out, void* stream) {
  if (!h || !chw || !patches_out || B < 0) return fail(h, IIC_ERR_ARG, "iic_patchify: bad argument");
  if (h->patch_kpad != h->patch_k) {
    // padded tail columns must be zero for the GEMM
    if (cudaMemsetAsync(patches_out, 0, size_t(B) * h->g * h->g * h->patch_kpad * 2, static_cast<cudaStream_t>(stream)) !=
        cudaSuccess)
      return fail(h, IIC_ERR_CUDA, "iic_patchify: memset failed");
  }
  Scope sc(h->prof, kMisc, static_cast<cudaStream_t>(stream));
  int rc = launch_chw_to_patches(chw, dtype, patches_out, B, h->cfg.image_size, h->cfg.patch_size, h->patch_kpad,
                                 h->f16, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "patchify launch failed");
  return IIC_OK;
}

size_t iic_workspace_bytes(const iic_handle* h, int B) {
  if (!h || B <= 0) return 0;
  return carve(h, B, nullptr).total + 1024;
}

static int check_ws(iic_handle* h, int B, void* workspace, size_t bytes, Workspace* w) {
  if (B <= 0 || !workspace) return fail(h, IIC_ERR_ARG, "workspace / batch: bad argument");
  if (!check_ready(h)) return fail(h, IIC_ERR_STATE, "not all weights have been loaded (iic_load_weight)");
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  *w = carve(h, B, base);
  if (size_t(base - static_cast<uint8_t*>(workspace)) + w->total > bytes)
    return fail(h, IIC_ERR_ARG, "workspace too small: see iic_workspace_bytes");
  return 0;
}

int iic_encode(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* emb_out,
               void* stream) {
  if (!h || !patches || !emb_out) return fail(h, IIC_ERR_ARG, "iic_encode: null argument");
  PdlScope pdl(h, B);
  Workspace w;
  int rc = check_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = run_encoder(h, patches, B, w, s);
  if (rc) return rc;
  return run_head(h, w.x, (long long)h->T * h->cfg.width, nullptr, B, emb_out, nullptr, s, w.head_small);
}

int iic_encode_sequence(iic_handle* h, const float* x_in, const int32_t* row_index, int B, void* workspace,
                        size_t workspace_bytes, float* emb_out, void* stream) {
  if (!h || !x_in || !row_index || !emb_out) return fail(h, IIC_ERR_ARG, "iic_encode_sequence: null argument");
  if (h->cfg.seq_tokens <= 0) return fail(h, IIC_ERR_STATE, "iic_encode_sequence: handle was not created with iic_config.seq_tokens");
  Workspace w;
  int rc = check_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int d = h->cfg.width, T = h->T;
  h->err.clear();
  if (cudaMemcpyAsync(w.x, x_in, size_t(B) * T * d * sizeof(float), cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return fail(h, IIC_ERR_CUDA, "iic_encode_sequence: copy failed");
  rc = run_blocks(h, B, w, s);
  if (rc) return rc;
  // row row_index[b] of every sequence (the EOT token) -> ln_final -> @ text_projection
  float* rows = reinterpret_cast<float*>(w.xln);   // [B, d] f32 scratch (xln is free after the last block)
  {
    Scope sc(h->prof, kMisc, s);
    if (launch_gather_rows(w.x, row_index, T, d, B, rows, s) != 0) return fail(h, IIC_ERR_CUDA, "row gather launch failed");
  }
  return run_head(h, rows, (long long)d, nullptr, B, emb_out, nullptr, s, w.head_small);
}

int iic_head(iic_handle* h, const float* emb, int B, const iic_head_out* out, void* stream) {
  if (!h || !emb || !out || B <= 0) return fail(h, IIC_ERR_ARG, "iic_head: bad argument");
  return run_head(h, nullptr, 0, emb, B, nullptr, out, static_cast<cudaStream_t>(stream));
}

int iic_classify(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* emb_out,
                 const iic_head_out* out, void* stream) {
  if (!h || !patches || !out) return fail(h, IIC_ERR_ARG, "iic_classify: null argument");
  PdlScope pdl(h, B);
  Workspace w;
  int rc = check_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = run_encoder(h, patches, B, w, s);
  if (rc) return rc;
  return run_head(h, w.x, (long long)h->T * h->cfg.width, nullptr, B, emb_out, out, s, w.head_small);
}

int iic_profile(iic_handle* h, int enable) {
  if (!h) return IIC_ERR_ARG;
  h->prof.on = enable != 0;
  h->prof.spans.clear();
  h->prof.used = 0;
  for (long long& l : h->prof.launches) l = 0;
  return IIC_OK;
}

int iic_profile_read(iic_handle* h, double* ms_by_class, long long* launches_by_class, int n) {
  if (!h || n < kNumClasses) return fail(h, IIC_ERR_ARG, "iic_profile_read: need room for 12 classes");
  for (int i = 0; i < kNumClasses; ++i) {
    if (ms_by_class) ms_by_class[i] = 0.0;
    if (launches_by_class) launches_by_class[i] = h->prof.launches[i];
  }
  if (!h->prof.spans.empty()) {
    if (cudaEventSynchronize(h->prof.spans.back().b) != cudaSuccess)
      return fail(h, IIC_ERR_CUDA, "iic_profile_read: event synchronize failed (a kernel faulted?)");
    for (const Profiler::Span& sp : h->prof.spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && ms_by_class) ms_by_class[sp.cls] += ms;
    }
  }
  h->prof.spans.clear();
  h->prof.used = 0;
  for (long long& l : h->prof.launches) l = 0;
  return IIC_OK;
}


// ---- training ----
int iic_set_lora_train(iic_handle* h, int layer, int which, const void* a16, const float* bt32, float scaling,
                       float* grad_a, float* grad_b) {
  if (!h) return IIC_ERR_ARG;
  if (layer < 0 || layer >= int(h->blocks.size()) || which < 0 || which > 3)
    return fail(h, IIC_ERR_ARG, "iic_set_lora_train: bad layer / projection id");
  LoraSlot& s = h->blocks[layer].lora[which];
  if (s.rank <= 0) return fail(h, IIC_ERR_STATE, "iic_set_lora_train: call iic_set_lora for this slot first");
  if (!a16 || !bt32 || !grad_a || !grad_b) return fail(h, IIC_ERR_ARG, "iic_set_lora_train: null pointer");
  s.a16 = a16; s.bt32 = bt32; s.scaling = scaling; s.grad_a = grad_a; s.grad_b = grad_b;
  return IIC_OK;
}

int iic_train_set_loss_scale(iic_handle* h, float loss_scale) {
  if (!h || !(loss_scale > 0.f)) return fail(h, IIC_ERR_ARG, "iic_train_set_loss_scale: scale must be positive");
  h->grad_unscale = 1.0f / loss_scale;
  return IIC_OK;
}

size_t iic_train_workspace_bytes(const iic_handle* h, int B) {
  if (!h || B <= 0) return 0;
  return carve_train(h, B, nullptr).total + 1024;
}

static int check_train_ws(iic_handle* h, int B, void* workspace, size_t bytes, TrainWorkspace* w) {
  if (B <= 0 || !workspace) return fail(h, IIC_ERR_ARG, "workspace / batch: bad argument");
  if (!check_ready(h)) return fail(h, IIC_ERR_STATE, "not all weights have been loaded (iic_load_weight)");
  if (!check_train_ready(h)) return fail(h, IIC_ERR_STATE, "training needs the transposed weights (*_t) and iic_set_lora_train for every LoRA slot");
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 1024));
  *w = carve_train(h, B, base);
  if (size_t(base - static_cast<uint8_t*>(workspace)) + w->total > bytes)
    return fail(h, IIC_ERR_ARG, "training workspace too small: see iic_train_workspace_bytes");
  return 0;
}

int iic_train_forward(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* x_cls_out,
                      void* stream) {
  if (!h || !patches || !x_cls_out) return fail(h, IIC_ERR_ARG, "iic_train_forward: null argument");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  if (h->cfg.seq_tokens > 0) return fail(h, IIC_ERR_STATE, "iic_train_forward: sequence handle, use iic_train_forward_sequence");
  return run_train_forward(h, patches, nullptr, nullptr, B, w, x_cls_out, static_cast<cudaStream_t>(stream));
}

int iic_train_forward_sequence(iic_handle* h, const float* x_tokens, const int32_t* row_index, int B, void* workspace,
                               size_t workspace_bytes, float* x_rows_out, void* stream) {
  if (!h || !x_tokens || !row_index || !x_rows_out) return fail(h, IIC_ERR_ARG, "iic_train_forward_sequence: null argument");
  if (h->cfg.seq_tokens <= 0) return fail(h, IIC_ERR_STATE, "iic_train_forward_sequence: handle was not created with iic_config.seq_tokens");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  return run_train_forward(h, nullptr, x_tokens, row_index, B, w, x_rows_out, static_cast<cudaStream_t>(stream));
}

int iic_train_backward_begin_sequence(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_rows,
                                      const int32_t* row_index, void* stream) {
  if (!h || !dx_rows || !row_index) return fail(h, IIC_ERR_ARG, "iic_train_backward_begin_sequence: null argument");
  if (h->cfg.seq_tokens <= 0) return fail(h, IIC_ERR_STATE, "iic_train_backward_begin_sequence: not a sequence handle");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  return run_train_backward_begin(h, B, w, dx_rows, row_index, static_cast<cudaStream_t>(stream));
}

int iic_train_backward(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_cls, void* stream) {
  if (!h || !dx_cls) return fail(h, IIC_ERR_ARG, "iic_train_backward: null argument");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  rc = run_train_backward_begin(h, B, w, dx_cls, nullptr, static_cast<cudaStream_t>(stream));
  for (int li = int(h->blocks.size()) - 1; li >= 0 && rc == 0; --li)
    rc = run_train_backward_layer(h, B, w, li, static_cast<cudaStream_t>(stream));
  return rc;
}

int iic_train_backward_begin(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_cls, void* stream) {
  if (!h || !dx_cls) return fail(h, IIC_ERR_ARG, "iic_train_backward_begin: null argument");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  return run_train_backward_begin(h, B, w, dx_cls, nullptr, static_cast<cudaStream_t>(stream));
}

int iic_train_backward_layer(iic_handle* h, int B, void* workspace, size_t workspace_bytes, int layer, void* stream) {
  if (!h || layer < 0 || layer >= int(h->blocks.size())) return fail(h, IIC_ERR_ARG, "iic_train_backward_layer: bad argument");
  TrainWorkspace w;
  int rc = check_train_ws(h, B, workspace, workspace_bytes, &w);
  if (rc) return rc;
  return run_train_backward_layer(h, B, w, layer, static_cast<cudaStream_t>(stream));
}

int iic_op_attention_bwd(iic_handle* h, const void* qkv, void* out, const void* d_out, void* dqkv, float* lse_scratch,
                         int B, int T, int heads, void* stream) {
  if (!h || !qkv || !out || !d_out || !dqkv || !lse_scratch) return fail(h, IIC_ERR_ARG, "iic_op_attention_bwd: null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // recompute the forward into `out`'s twin to obtain the log-sum-exp (op-level test helper)
  int rc = run_attention(h, qkv, out, lse_scratch, B, T, heads, 64, 0, s);
  if (rc == 0) rc = run_attention_bwd(h, qkv, out, d_out, lse_scratch, lse_scratch + size_t(B) * heads * T, dqkv, B, T, heads, 64, s);
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "attention backward: unsupported shape (T <= 432) or launch failure");
  return IIC_OK;
}

int iic_op_layernorm_bwd(iic_handle* h, const void* dy, const float* x, const float* gamma, float* dx, void* dx16, int rows,
                         int D, void* stream) {
  if (!h || !dy || !x || !gamma || !dx) return fail(h, IIC_ERR_ARG, "iic_op_layernorm_bwd: null argument");
  int rc = launch_layernorm_bwd(dy, x, gamma, dx, dx16, rows, D, 1e-5f, h->f16, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "layernorm backward failed");
  return IIC_OK;
}

int iic_op_act_bwd(iic_handle* h, void* dh, const void* u, long long n, int act, void* stream) {
  if (!h || !dh || !u) return fail(h, IIC_ERR_ARG, "iic_op_act_bwd: null argument");
  int rc = launch_act_bwd(dh, u, n, act, h->f16, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "activation backward failed");
  return IIC_OK;
}

size_t iic_op_lora_scratch_bytes(int N, int M) { return (N > 0 && M > 0) ? lora_outer_scratch_bytes(N, M) : 0; }

int iic_op_lora_bwd(iic_handle* h, const void* P, int p_ld, const void* Y, int N, int M, const void* Bm, int rank, float scale,
                    float* out_db, void* out_dp16, void* scratch, size_t scratch_bytes, void* stream) {
  if (!h || !P || !Y || !Bm || !out_db || !out_dp16 || !scratch) return fail(h, IIC_ERR_ARG, "iic_op_lora_bwd: null argument");
  if (scratch_bytes < lora_outer_scratch_bytes(N, M)) return fail(h, IIC_ERR_ARG, "iic_op_lora_bwd: scratch too small (iic_op_lora_scratch_bytes)");
  int rc = launch_lora_bwd(P, p_ld, Y, N, M, Bm, rank, scale, out_db, out_dp16, static_cast<float*>(scratch), h->f16,
                           static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -2 ? IIC_ERR_CUDA : IIC_ERR_ARG, rc == -3 ? "lora_bwd: needs N % 256 == 0 and p_ld == 16" : "lora_bwd failed");
  return IIC_OK;
}

int iic_op_lora_outer(iic_handle* h, const void* P, int p_ld, const void* Y, int N, int M, int act, int rank, float scale,
                      int transpose, float* out, void* scratch, size_t scratch_bytes, void* stream) {
  if (!h || !P || !Y || !out || !scratch) return fail(h, IIC_ERR_ARG, "iic_op_lora_outer: null argument");
  if (scratch_bytes < lora_outer_scratch_bytes(N, M)) return fail(h, IIC_ERR_ARG, "iic_op_lora_outer: scratch too small (iic_op_lora_scratch_bytes)");
  int rc = launch_lora_outer(P, p_ld, Y, N, M, act, rank, scale, transpose, out, static_cast<float*>(scratch), h->f16,
                             static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "lora_outer failed");
  return IIC_OK;
}

// ---- single operators ----
int iic_op_gemm(iic_handle* h, const void* a, int lda, const void* w, int ldw, int M, int N, int K, const void* lora_p,
                const void* lora_bt, int r_pad, int lora_ld, int epilogue, const float* bias, const float* residual,
                void* out, int ldc, int group, int ctas, const float* down_a, float* down_part, void* stream) {
  if (!h || !a || !w || !out) return fail(h, IIC_ERR_ARG, "iic_op_gemm: null argument");
  GemmProblem g;
  g.down_a = down_a;
  g.down_part = down_part;
  if (down_part != nullptr) g.tile_n = 256;   // the caller sized down_part for the wide tile (include/iic.h)
  g.a = a; g.lda = lda;
  g.w = w; g.ldw = ldw;
  g.M = M; g.N = N; g.K = K;
  g.lora_p = lora_p;
  g.lora_bt = lora_bt;
  g.f16 = h->f16;
  g.r_pad = r_pad; g.lora_ld = lora_ld;
  g.epilogue = epilogue; g.bias = bias; g.residual = residual; g.out = out; g.ldc = ldc; g.group = group;
  const char* e = nullptr;
  int rc = launch_gemm(g, ctas == 1 ? 1 : (ctas == 2 ? 2 : h->ctas), h->num_sms, static_cast<cudaStream_t>(stream), &e);
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, e ? e : "gemm failed");
  return IIC_OK;
}

int iic_op_gemm_act_dual(iic_handle* h, const void* a, int lda, const void* w, int ldw, int M, int N, int K, const void* lora_p,
                         const void* lora_bt, int r_pad, int lora_ld, const float* bias, void* out_act, void* out_pre, int act,
                         int ctas, void* stream) {
  if (!h || !a || !w || !out_act || !out_pre) return fail(h, IIC_ERR_ARG, "iic_op_gemm_act_dual: null argument");
  GemmProblem g;
  g.a = a; g.lda = lda;
  g.w = w; g.ldw = ldw;
  g.M = M; g.N = N; g.K = K;
  g.lora_p = lora_p;
  g.lora_bt = lora_bt;
  g.f16 = h->f16;
  g.r_pad = r_pad; g.lora_ld = lora_ld;
  g.epilogue = kEpiBiasActDualBf16; g.bias = bias; g.residual = nullptr; g.out = out_act; g.out2 = out_pre; g.ldc = N;
  g.group = act == IIC_ACT_GELU_ERF ? 2 : 1;
  const char* e = nullptr;
  int rc = launch_gemm(g, ctas == 1 ? 1 : (ctas == 2 ? 2 : h->ctas), h->num_sms, static_cast<cudaStream_t>(stream), &e);
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, e ? e : "gemm failed");
  return IIC_OK;
}

int iic_op_layernorm(iic_handle* h, const float* x, const float* gamma, const float* beta, void* out_bf16,
                     float* out_f32, int rows, int D, const float* lora_a_scaled, int r4, void* p_out, int p_ld,
                     void* stream) {
  if (!h || !x || !gamma || !beta) return fail(h, IIC_ERR_ARG, "iic_op_layernorm: null argument");
  int rc = launch_layernorm(x, D, gamma, beta, out_bf16, out_f32, D, rows, D, 1e-5f, lora_a_scaled, r4, p_out, p_ld,
                            h->f16, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "layernorm: unsupported width or launch failure");
  return IIC_OK;
}

int iic_op_lora_down(iic_handle* h, const void* x_bf16, int K, int rows, const float* lora_a_scaled, int r4, void* p_out,
                     int p_ld, void* stream) {
  if (!h || !x_bf16 || !lora_a_scaled || !p_out) return fail(h, IIC_ERR_ARG, "iic_op_lora_down: null argument");
  int rc = launch_lora_down_bf16(x_bf16, K, rows, lora_a_scaled, r4, p_out, p_ld, h->f16,
                                 static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "lora_down launch failed");
  return IIC_OK;
}

int iic_op_attention(iic_handle* h, const void* qkv_bf16, void* out_bf16, int B, int T, int heads, int impl, void* stream) {
  if (!h || !qkv_bf16 || !out_bf16) return fail(h, IIC_ERR_ARG, "iic_op_attention: null argument");
  int rc = run_attention(h, qkv_bf16, out_bf16, nullptr, B, T, heads, 64, impl, static_cast<cudaStream_t>(stream));
  if (rc != 0) return fail(h, rc == -1 ? IIC_ERR_ARG : IIC_ERR_CUDA, "attention launch failed");
  return IIC_OK;
}

}  // extern "C"
