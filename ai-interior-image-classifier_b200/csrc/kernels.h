// Internal launcher interface between the C-ABI engine (iic_api.cu) and the kernel translation units.
// Every launcher returns 0 on success, -1 for an unsupported shape/argument, -2 for a CUDA launch error.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace iic {

// One-time per-DEVICE opt-in guard.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the current device only, so
// a process-wide "done" flag would leave a second engine on another GPU of the same process without the opt-in.  One bit per
// device ordinal; two threads racing on the first call both set the (idempotent) attribute, which is harmless.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0ull};
  static int device() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 63;
  }
  bool need() const { return ((done.load(std::memory_order_acquire) >> device()) & 1ull) == 0ull; }
  void mark() { done.fetch_or(1ull << device(), std::memory_order_release); }
};

// Programmatic dependent launch for the small-batch (latency) path: the ~100 dependent launches of a batch-1 pass overlap
// each kernel's launch latency and prologue with its predecessor's tail.  Set per call by the engine (iic_api.cu) on the calling
// thread; kernels on that path start with ptx::pdl_launch_dependents() and call ptx::pdl_wait() before their first global read.
extern thread_local int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- rowwise.cu ----
int launch_layernorm(const float* x, long long x_row_stride, const float* gamma, const float* beta,
                     void* out_bf16, float* out_f32, long long out_row_stride, int rows, int D, float eps,
                     const float* lora_a, int r4, void* p_out, int p_ld, int f16, cudaStream_t stream);
int launch_lora_down_bf16(const void* x, int K, int rows, const float* lora_a, int r4, void* p_out,
                          int p_ld, int f16, cudaStream_t stream);
// sums the per-column-tile partials written by the c_fc GEMM epilogue (GemmProblem::down_*) into the 16-bit P matrix
int launch_lora_reduce(const float* part, int n_tiles, int rows, void* p_out, int p_ld, int f16, cudaStream_t stream);
// out[b, :] = x[b * T + row_index[b], :]   (f32, D % 4 == 0)
int launch_gather_rows(const float* x, const int32_t* row_index, int T, int D, int B, float* out, cudaStream_t stream);
int launch_scatter_rows(const float* rows, const int32_t* row_index, int T, int D, int B, float* x, cudaStream_t stream);
int launch_fill_cls(float* x_pre, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream);
// dtype: 0 = f32, 1 = bf16, 2 = f16
int launch_chw_to_patches(const void* img, int dtype, void* patches, int B, int R, int P, int k_pad, int f16,
                          cudaStream_t stream);

// ---- attention.cu ----
// lse (nullable): f32 [B*H, T] log2-domain log-sum-exp of the scaled scores, saved for the backward pass
int launch_attention(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16,
                     cudaStream_t stream);
// tcgen05 / TMEM kernel (attention_sm100.cu): head_dim 64, K/V of one head resident in smem (T <= ~760);
// returns -3 if the shape is outside its envelope
int launch_attention_sm100(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16, bool causal,
                           int num_sms, cudaStream_t stream);
// whole-row tcgen05 kernel for short, unmasked sequences (attention_row_sm100.cu): T <= 208, head_dim 64; the score row of a
// query is complete in TMEM before the softmax reads it (exact maximum, no online rescale).  Returns -3 outside its envelope.
int launch_attention_row_sm100(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16, int num_sms,
                               cudaStream_t stream);
// dqkv[M, 3d] (16-bit) from d_out[M, d], the saved qkv / out / lse.  T <= 432 (everything of one head lives in smem).
int launch_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int B, int T,
                         int H, int head_dim, int f16, cudaStream_t stream);

// tcgen05 / TMEM variant for T <= 592 (attention_bwd_sm100.cu); dsum_scratch f32 [B*H*T]; returns -3 outside its envelope
int launch_attention_bwd_sm100(const void* qkv, const void* out, const void* d_out, const float* lse, float* dsum_scratch,
                               void* dqkv, int B, int T, int H, int head_dim, int f16, int causal, int num_sms,
                               cudaStream_t stream);

// D[bh, q] = rowsum(dO o O), f32 [B*H*T] (attention_bwd_sm100.cu): the pre-kernel of the tcgen05 backward kernels
int launch_attention_bwd_dsum(const void* d_out, const void* out, float* dsum, int B, int T, int H, int f16, cudaStream_t stream);
// one pass per key tile (attention_bwd_fused_sm100.cu), T <= 256; dsum already computed; -3 outside its envelope
int launch_attention_bwd_fused_sm100(const void* qkv, const void* d_out, const float* lse, const float* dsum, void* dqkv, int B, int T,
                                     int H, int head_dim, int f16, int causal, int num_sms, cudaStream_t stream);

// ---- train_ops.cu ----
int launch_layernorm_bwd(const void* dy, const float* x, const float* gamma, float* dx, void* dx16, int rows, int D,
                         float eps, int f16, cudaStream_t stream);
int launch_cast16(const float* in, void* out, long long n, int f16, cudaStream_t stream);
// act: 0 identity, 1 QuickGELU, 2 GELU(erf)
int launch_act_bwd(void* dh, const void* u, long long n, int act, int f16, cudaStream_t stream);
size_t lora_outer_scratch_bytes(int N, int M);
// One pass over Y [M, N] (16-bit) for the two LoRA gradients that read it: out_db f32 [rank, N] = scale * P^T . Y (P 16-bit
// [M, 16]) and out_dp16 16-bit [M, 16] = Y . Bm^T (Bm 16-bit [16, N], rows >= rank zero).  Returns -3 outside its envelope
// (N % 256, p_ld == 16); scratch as lora_outer_scratch_bytes(N, M).
int launch_lora_bwd(const void* P, int p_ld, const void* Y, int N, int M, const void* Bm, int rank, float scale, float* out_db,
                    void* out_dp16, float* scratch, int f16, cudaStream_t stream);
// out = scale * P[:, :rank]^T . act(Y)  ->  [rank, N] (transpose = 0: dB) or [N, rank] (transpose = 1: dA); deterministic
int launch_lora_outer(const void* P, int p_ld, const void* Y, int N, int M, int act, int rank, float scale, int transpose,
                      float* out, float* scratch, int f16, cudaStream_t stream);

// derived operands of one LoRA slot from its fp32 parameters A [in, rank], B [rank, out] (see train_ops.cu); null outputs skipped
int launch_lora_refresh(const float* A, const float* B, int in, int out, int rank, int r4, int pad, float scaling, float* a,
                        void* bt, void* a16, float* bt32, void* at16, void* b16, int f16, cudaStream_t stream);
// the same for up to 16 slots in ONE launch (the per-step refresh of a 12-block model is 24 slots: two launches, not 24)
struct LoraRefreshSlot {
  const float* A;   // f32 [in, rank]
  const float* B;   // f32 [rank, out]
  float* a;         // destinations as in launch_lora_refresh (null: skipped)
  void* bt;
  void* a16;
  float* bt32;
  void* at16;
  void* b16;
  int in, out, rank, r4;
  float scaling;
};
struct LoraRefreshBatch {
  LoraRefreshSlot slot[16];
  int n, pad;
};
int launch_lora_refresh_batch(const LoraRefreshBatch& batch, int f16, cudaStream_t stream);

// ---- head.cu ----
int launch_head(const float* x, long long x_img_stride, const float* ln_g, const float* ln_b, float eps,
                const float* proj, int W, int E, const float* text, int L, const int* group_off,
                const int* group_split, int G, int topk, float logit_scale, int B, float* emb_out, float* logits_out,
                float* probs_out, float* topk_val, int* topk_idx, float* split_sum, const float* emb_in,
                float* small_scratch, cudaStream_t stream);
// device scratch (floats: B*E + 2*B*L) that switches launch_head to its small-batch latency path (B <= 16); 0 otherwise
size_t head_small_scratch_bytes(int B, int E, int L);

// ---- preprocess.cu ----
struct PreprocessPlan;  // opaque: host-side coefficient tables + device scratch, owned by the engine handle
PreprocessPlan* preprocess_plan_create();
void preprocess_plan_destroy(PreprocessPlan*);
// imgs: HOST array of B DEVICE pointers to uint8 HWC RGB images; hw: HOST array [B][2] = (height, width).
// out_mode 0: 16-bit patch matrix [B*g*g, k_pad];  out_mode 1: f32 CHW [B,3,R,R];  out_mode 2: 16-bit CHW.
// f16: 16-bit outputs are fp16 instead of bf16.
int launch_preprocess(PreprocessPlan* plan, const uint8_t* const* imgs, const int* hw, int B, int R, int P, int k_pad,
                      void* out, int out_mode, int f16, cudaStream_t stream, const char** err);
// same-size fast path: one contiguous uint8 [B,R,R,3] device buffer, no resampling.
int launch_preprocess_fast(PreprocessPlan* plan, const uint8_t* imgs, int B, int R, int P, int k_pad, void* out,
                           int out_mode, int f16, cudaStream_t stream);

// ---- gemm_sm100.cu ----
struct GemmProblem {
  const void* a;   // [M, lda]   (lda = row pitch in elements, multiple of 8)
  int lda;
  const void* w;   // [N, ldw]
  int ldw;
  int M, N, K;
  const void* lora_p;  // [M, r_pad] (row pitch r_pad) or null
  const void* lora_bt; // [N, r_pad]
  int r_pad;                    // multiple of 16, <= 64: columns the MMA consumes
  int lora_ld;                  // row pitch (elements) of lora_p / lora_bt; 0 -> r_pad.  Columns [r_pad, 64) of the
                                // TMA box fall outside the tensor and are zero-filled (never consumed anyway).
  int epilogue;                 // GemmEpilogue
  const float* bias;
  const float* residual;        // residual [M, ldc] or pos table [G+1, N]
  void* out;
  int ldc;
  int group;
  int f16;  // 16-bit operand format: 0 = bf16, 1 = fp16
  // optional, activation epilogues only: fuse the NEXT projection's LoRA down-projection (rank <= 4) into this epilogue.
  // down_a f32 [N, 4] = scaling * lora_A of the consumer; down_part f32 [gemm_down_parts(...)][M][4] receives per-tile
  // partials (one per column tile and epilogue warp group).
  const float* down_a = nullptr;
  float* down_part = nullptr;
  void* out2 = nullptr;  // kEpiBiasActDualBf16: 16-bit [M, ldc] pre-activation output (out receives the activation)
  int tile_n = 0;        // 0: gemm_tile_n() decides; 256: force the wide tile (callers that sized down_part for it)
};
// ctas: 1 or 2 (tcgen05 cta_group).  num_sms: SM count of the device.
int launch_gemm(const GemmProblem& p, int ctas, int num_sms, cudaStream_t stream, const char** err);
// Output-tile width the launcher picks: 256 columns per CTA pair, except that problems whose 256-wide tiles would leave
// most SMs idle (the single-image path: M = 197 rows) are cut into 128- or 64-column tiles (inference epilogues, ctas == 2).
int gemm_tile_n(int M, int N, int epilogue, int ctas, int num_sms);
// number of partial slots the fused LoRA down-projection of an activation epilogue writes for this problem
int gemm_down_parts(int M, int N, int epilogue, int ctas, int num_sms);
size_t gemm_smem_bytes(int ctas);

}  // namespace iic
