/* iic.h -- C ABI of the B200-native LoRA-ViT image encoder + label-scoring head.
 *
 * The reference (M1A5TO/AI-interior-image-classifier) has no FFI of its own: its seam is the Python duck type
 * returned by `clip.load(...)` and used at
 *     /root/reference/main.py:152,241          model, preprocess = clip.load("ViT-B/16", device)
 *     /root/reference/main.py:201,438,489      preprocess(PIL.Image)            -> iic_preprocess*
 *     /root/reference/main.py:204,444,503      model.encode_image(batch)        -> iic_patchify + iic_encode
 *     /root/reference/main.py:205-211,445-459,504-509   L2-norm, 100*cos, softmax, topk -> iic_head / iic_classify
 *     /root/reference/main.py:30-31,42-43      LoRALinear / LoRALayer forward   -> iic_set_lora (fused in the GEMM)
 * Every entry point below is what a ctypes / cffi binding on the reference side would bind; INTEGRATION.md shows
 * that binding.  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - All data pointers are DEVICE pointers unless a parameter says "host".
 *   - The caller owns every input / output / workspace buffer and keeps weight buffers alive for the life of
 *     the handle (weights are borrowed, not copied).  The library owns only a few KB of descriptors (label group offsets,
 *     TMA maps are built per call on the host) and ONE documented exception: the grow-only resampling scratch of the
 *     general-size iic_preprocess (intermediate rows + coefficient tables, sized by the largest batch seen; the
 *     same-size fast path and every other entry point allocate nothing and never synchronise the stream).
 *   - All work is enqueued on the caller's stream (cudaStream_t passed as void*); calls are asynchronous.
 *   - Return value: 0 = IIC_OK, negative = error; iic_last_error(h) describes the last failure.
 *   - A handle is not re-entrant: serialise calls per handle (one handle per GPU per process for data parallel).
 *   - There is no CPU path.  Without a CUDA device every compute entry point fails.
 */
#ifndef IIC_H_
#define IIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIC_OK 0
#define IIC_ERR_ARG (-1)    /* bad argument / unsupported shape */
#define IIC_ERR_CUDA (-2)   /* CUDA runtime / launch failure    */
#define IIC_ERR_STATE (-3)  /* weights / labels missing         */

#define IIC_DTYPE_F32 0
#define IIC_DTYPE_BF16 1
#define IIC_DTYPE_F16 2

#define IIC_ACT_QUICK_GELU 0 /* OpenAI CLIP: x * sigmoid(1.702 x) */
#define IIC_ACT_GELU_ERF 1

/* which projection of a residual block a LoRA pair attaches to */
#define IIC_LORA_IN_PROJ 0  /* not reachable from the reference (F3): generic slot */
#define IIC_LORA_OUT_PROJ 1 /* reference wraps it but nn.MultiheadAttention never calls it (F4): off by default */
#define IIC_LORA_C_FC 2
#define IIC_LORA_C_PROJ 3

/* preprocess output layouts */
#define IIC_OUT_PATCHES_BF16 0 /* [B*g*g, k_pad] patch matrix, column = c*P*P + ky*P + kx (what iic_encode eats) */
#define IIC_OUT_CHW_F32 1      /* [B,3,R,R] float32: exactly what the reference's preprocess returns              */
#define IIC_OUT_CHW_BF16 2

typedef struct iic_handle iic_handle;

typedef struct iic_config {
  int image_size; /* R: 224 (ViT-B/16) or 336 (ViT-L/14@336px) */
  int patch_size; /* P: 16 or 14                               */
  int width;      /* 768 / 1024                                */
  int layers;     /* 12 / 24                                   */
  int heads;      /* width / 64                                */
  int mlp_dim;    /* 4 * width                                 */
  int embed_dim;  /* 512 / 768                                 */
  int activation; /* IIC_ACT_*                                 */
  int device;     /* CUDA ordinal                              */
  int gemm_ctas;  /* 0 = default, 1 = single-SM tiles, 2 = CTA pairs (tcgen05 cta_group::2) */
  int operand_dtype; /* 16-bit format of activations and matmul weights: IIC_DTYPE_BF16 (default, also for 0) or
                        IIC_DTYPE_F16 (OpenAI CLIP's own GPU dtype; same tensor-core rate, 3 more mantissa bits).
                        Accumulation, LayerNorm, softmax, the residual stream and the head are fp32 in both. */
  int seq_tokens; /* 0: vision tower (T = (R/P)^2 + 1).  > 0: a SEQUENCE handle for CLIP's text tower - T = seq_tokens (77),
                     no patch embedding / class token / ln_pre; see iic_encode_sequence.  image_size / patch_size are
                     then only validated, not used. */
  int causal;     /* 1: causal attention mask (CLIP text tower), 0: full attention */
} iic_config;

typedef struct iic_dims {
  int tokens;     /* T = g*g + 1                   */
  int grid;       /* g = R / P                     */
  int patch_k;    /* 3*P*P                         */
  int patch_kpad; /* patch_k rounded up to 8       */
  int lora_pad;   /* row pitch of LoRA operands    */
} iic_dims;

int iic_create(iic_handle** out, const iic_config* cfg);
void iic_destroy(iic_handle* h);
const char* iic_last_error(const iic_handle* h); /* h may be NULL: last create error */
int iic_get_dims(const iic_handle* h, iic_dims* out);
const char* iic_version(void);

/* ---- weights (borrowed device pointers; names are OpenAI-CLIP `visual.` state-dict names without the prefix) --
 *   ("bf16" below means the handle's operand_dtype: bf16 or fp16)
 *   conv1.weight                      bf16 [width, patch_kpad]   (conv weight flattened (c,ky,kx), zero padded)
 *   class_embedding                   f32  [width]
 *   positional_embedding              f32  [T, width]
 *   ln_pre.weight|bias, ln_post.weight|bias                      f32 [width]
 *   proj                              f32  [width, embed_dim]
 *   transformer.resblocks.{i}.ln_1.weight|bias, .ln_2.weight|bias           f32 [width]
 *   transformer.resblocks.{i}.attn.in_proj_weight   bf16 [3*width, width];  .attn.in_proj_bias   f32 [3*width]
 *   transformer.resblocks.{i}.attn.out_proj.weight  bf16 [width, width];    .attn.out_proj.bias  f32 [width]
 *   transformer.resblocks.{i}.mlp.c_fc.weight       bf16 [mlp, width];      .mlp.c_fc.bias       f32 [mlp]
 *   transformer.resblocks.{i}.mlp.c_proj.weight     bf16 [width, mlp];      .mlp.c_proj.bias     f32 [width]
 */
int iic_load_weight(iic_handle* h, const char* name, const void* dev_ptr, int dtype, int ndim, const int64_t* shape);

/* LoRA pair for one projection (reference: LoRALayer, /root/reference/main.py:19-31):
 *   a_scaled f32 [in, r4]   = lora_A * (alpha/rank), columns zero padded to r4 = round_up(rank, 4)
 *   b_t      bf16 [out, ld] = lora_B^T, columns zero padded to ld = iic_dims.lora_pad (16 for rank <= 16, ...)
 * Passing rank == 0 clears the slot.  A cleared slot costs nothing. */
int iic_set_lora(iic_handle* h, int layer, int which, const float* a_scaled, const void* b_t, int rank);

/* Optional 16-bit operands of one LoRA slot (after iic_set_lora), needed for ranks > 4: the down-projections
 * P = s (x . A) (forward) and dP = dY . B^T (training backward) then run on the tcgen05 GEMM with N = lora_pad instead of
 * the fp32-A row kernels.  a_t16: 16-bit [lora_pad, in] = (scaling * lora_A)^T, b16: 16-bit [lora_pad, out] = lora_B
 * (rows >= rank zero; b16 may be NULL for inference).  Borrowed device pointers.
 * Replaces: `x @ self.lora_A` of LoRALayer.forward (/root/reference/main.py:31, train_lora.py:28) for rank-16 adapters. */
int iic_set_lora_operands16(iic_handle* h, int layer, int which, const void* a_t16, const void* b16);

/* Training loops change lora_A / lora_B every step.  iic_set_lora_source registers the fp32 parameters of a slot
 * (lora_a f32 [in, rank], lora_b f32 [rank, out] as the reference stores them, main.py:26-27; scaling = alpha / rank);
 * iic_refresh_lora then rebuilds, with one small kernel per registered slot on `stream`, every derived operand buffer that
 * was handed to iic_set_lora / iic_set_lora_train / iic_set_lora_operands16 (those buffers must be writable).  Replaces the
 * implicit "parameters are read at every forward" of LoRALayer.forward (main.py:30-31) after optimizer.step() (train_lora.py:252). */
int iic_set_lora_source(iic_handle* h, int layer, int which, const float* lora_a, const float* lora_b, float scaling);
int iic_refresh_lora(iic_handle* h, void* stream);

/* Label text embeddings the head scores against (reference: text_features_cache, main.py:296-311 and
 * detector text_features, main.py:179-182): text f32 [L, embed_dim], rows L2-normalised by the caller exactly as
 * the reference does.  group_offsets (host, G+1 ints) partitions the L labels into softmax groups;
 * group_split (host, G ints or NULL): for each group the number of leading labels whose probabilities are summed
 * into split_sum (detector: 11).  topk <= 8.  logit_scale = 100.0 in the reference. */
int iic_set_labels(iic_handle* h, const float* text, int num_labels, const int* group_offsets, const int* group_split,
                   int num_groups, int topk, float logit_scale);

/* ---- preprocessing ------------------------------------------------------------------------------------------ */
/* General path: B images of arbitrary size.  imgs = HOST array of B device pointers to uint8 HWC RGB;
 * hw = HOST array [B][2] = (height, width).  PIL-compatible antialiased bicubic + centre crop + normalise. */
int iic_preprocess(iic_handle* h, const uint8_t* const* imgs, const int* hw, int B, void* out, int out_layout,
                   void* stream);
/* Fast path: one contiguous uint8 [B, R, R, 3] buffer already at the model resolution (no resampling). */
int iic_preprocess_same_size(iic_handle* h, const uint8_t* imgs, int B, void* out, int out_layout, void* stream);
/* ---- image ingest: baseline-JPEG decode on the device ----------------------------------------------
 * Replaces `Image.open(path).convert("RGB")` (/root/reference/main.py:330-334 load_image, called from main.py:165 and
 * main.py:412) for local JPEG files: the decoded uint8 HWC pixels are written where iic_preprocess reads them and are
 * bit-identical to Pillow / libjpeg-turbo (islow IDCT, fancy upsampling, fixed-point YCbCr -> RGB).  Envelope: sequential
 * Huffman JPEG (SOF0 / SOF1, 8 bit), grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0, restart intervals; files outside it are
 * reported by the plan and stay on the caller's host path (as PNG files and URLs do).
 *
 * A PLAN is a host-only object (no CUDA call): it parses the headers of n files laid back to back in `blob` (file i =
 * bytes [offsets[i], offsets[i+1])) and lays out the device scratch.  The caller then provides
 *   dev_blob   the same bytes on the device, readable for 8 bytes past offsets[n]
 *   out_rgb    HOST array of n DEVICE pointers, out_rgb[i] -> uint8 [height_i][width_i][3]; NULL skips file i (required for
 *              files whose status is not IIC_JPEG_OK)
 *   staging    iic_jpeg_plan_staging_bytes() of PINNED host memory (the descriptors are built there and copied to the device
 *              on the stream; it must stay untouched until that copy has run)
 *   scratch    iic_jpeg_plan_scratch_bytes() of device memory, 256-byte aligned (descriptors, Huffman tables, coefficient
 *              blocks, sample planes)
 * and iic_jpeg_decode enqueues one copy and three kernels on `stream`.  Nothing is allocated or synchronised inside. */
#define IIC_JPEG_OK 0
#define IIC_JPEG_UNSUPPORTED 1 /* a valid JPEG outside the envelope (progressive, CMYK, arithmetic, ...) */
#define IIC_JPEG_CORRUPT 2     /* not a JPEG / broken header                                              */
typedef struct iic_jpeg_plan iic_jpeg_plan;
int iic_jpeg_plan_create(const uint8_t* blob, const int64_t* offsets, int n, iic_jpeg_plan** out);
void iic_jpeg_plan_destroy(iic_jpeg_plan* plan);
int iic_jpeg_plan_info(const iic_jpeg_plan* plan, int i, int* width, int* height, int* status);
int iic_jpeg_plan_infos(const iic_jpeg_plan* plan, int* whs); /* host int [n][3] = (width, height, status) of every file */
const char* iic_jpeg_plan_reason(const iic_jpeg_plan* plan, int i); /* why file i is outside the envelope ("" if it is not) */
size_t iic_jpeg_plan_staging_bytes(const iic_jpeg_plan* plan);
size_t iic_jpeg_plan_scratch_bytes(const iic_jpeg_plan* plan);
int iic_jpeg_decode(const iic_jpeg_plan* plan, const uint8_t* dev_blob, uint8_t* const* out_rgb, void* staging, void* scratch,
                    void* stream);

/* [B,3,R,R] float tensor (what reference code feeds encode_image) -> patch matrix. */
int iic_patchify(iic_handle* h, const void* chw, int dtype, int B, void* patches_out, void* stream);

/* ---- encoder + head ----------------------------------------------------------------------------------------- */
size_t iic_workspace_bytes(const iic_handle* h, int B);
/* patches bf16 [B*g*g, patch_kpad] -> emb f32 [B, embed_dim] (un-normalised, == model.encode_image output) */
int iic_encode(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* emb_out,
               void* stream);

/* Text tower (reference: model.encode_text, /root/reference/main.py:181, 308; train_lora.py:237; main_API.py:161) on
 * a handle created with seq_tokens > 0.  x_in f32 [B*T, width] = token_embedding[tokens] + positional_embedding (the
 * caller's gather + add), row_index int32 [B] (device) = tokens.argmax(-1), the EOT position.  Runs the residual blocks
 * (LoRA slots as for the vision tower, causal mask per iic_config.causal), then ln_final (loaded as `ln_post.*`) of row
 * row_index[b] of every sequence, then @ text_projection (loaded as `proj`): emb f32 [B, embed_dim], un-normalised. */
int iic_encode_sequence(iic_handle* h, const float* x_in, const int32_t* row_index, int B, void* workspace,
                        size_t workspace_bytes, float* emb_out, void* stream);

typedef struct iic_head_out {
  float* logits;    /* [B, L]      logit_scale * cos            (nullable) */
  float* probs;     /* [B, L]      per-group softmax            (nullable) */
  float* topk_val;  /* [B, G, k]   descending                   (required) */
  int32_t* topk_idx;/* [B, G, k]   index inside the group, -1 pad (required) */
  float* split_sum; /* [B, G]      sum of the first group_split[g] probabilities (nullable) */
} iic_head_out;

/* emb f32 [B, embed_dim] (un-normalised) -> scores */
int iic_head(iic_handle* h, const float* emb, int B, const iic_head_out* out, void* stream);
/* fused: patches -> encoder -> head; emb_out nullable */
int iic_classify(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* emb_out,
                 const iic_head_out* out, void* stream);

/* ---- training step (reference: train_lora.py:231-252; LoRA on the vision MLPs) ------------------
 * Only LoRA parameters receive gradients; frozen weights only carry dX.  Needs, in addition to the inference state:
 *   - transposed copies of the four projection weights of every block, loaded with iic_load_weight under the names
 *     `...attn.in_proj_weight_t` [width, 3*width], `...attn.out_proj.weight_t` [width, width],
 *     `...mlp.c_fc.weight_t` [width, mlp], `...mlp.c_proj.weight_t` [mlp, width]   (operand dtype);
 *   - per LoRA slot (after iic_set_lora): a16 = scaling*lora_A in the operand dtype [in, lora_pad] (zero padded),
 *     bt32 = lora_B^T f32 [out, r4], scaling = alpha/rank, and where the gradients go:
 *     grad_a f32 [in, rank] (d loss / d lora_A), grad_b f32 [rank, out] (d loss / d lora_B) - the reference's layouts.
 * iic_train_forward keeps the activations the backward needs inside the (larger) training workspace and returns the
 * class-token rows of the final residual stream, x_cls f32 [B, width] (the input of ln_post): the head and the loss
 * (ln_post, proj, L2-norm, InfoNCE - O(B*width) work) stay with the caller, who hands back dx_cls = d loss / d x_cls.
 * iic_train_backward overwrites every registered grad_a / grad_b.  LoRA rank <= 16 on mlp.c_fc / mlp.c_proj. */
int iic_set_lora_train(iic_handle* h, int layer, int which, const void* a16, const float* bt32, float scaling,
                       float* grad_a, float* grad_b);
/* Loss scaling for the 16-bit gradient operands (needed with fp16: activation gradients underflow otherwise): the caller
 * multiplies dx_cls by `loss_scale` (a power of two), the engine divides the LoRA gradients by it.  Default 1. */
int iic_train_set_loss_scale(iic_handle* h, float loss_scale);
size_t iic_train_workspace_bytes(const iic_handle* h, int B);
int iic_train_forward(iic_handle* h, const void* patches, int B, void* workspace, size_t workspace_bytes, float* x_cls_out,
                      void* stream);
int iic_train_backward(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_cls, void* stream);
/* The same backward, one residual block at a time (layer = layers-1 ... 0 after _begin): block `layer`'s LoRA gradients
 * are final when the call's work completes, so the caller can all-reduce them while the blocks below are still running. */
int iic_train_backward_begin(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_cls, void* stream);
int iic_train_backward_layer(iic_handle* h, int B, void* workspace, size_t workspace_bytes, int layer, void* stream);
/* Text tower (handles created with iic_config.seq_tokens, causal = 1): the training step train_lora.py:236-252 actually runs -
 * `lora_wrapper.clip_model.encode_text(texts)` through the LoRA on the text MLPs (train_lora.py:62-100), loss.backward().
 * x_tokens f32 [B*T, width] = token_embedding(tokens) + positional_embedding; row_index [B] = tokens.argmax(-1) (EOT);
 * x_rows_out f32 [B, width] = those rows of the final residual stream (input of ln_final).  The backward starts from
 * dx_rows [B, width] scattered to the same rows; the blocks are then walked with iic_train_backward_layer. */
int iic_train_forward_sequence(iic_handle* h, const float* x_tokens, const int32_t* row_index, int B, void* workspace,
                               size_t workspace_bytes, float* x_rows_out, void* stream);
int iic_train_backward_begin_sequence(iic_handle* h, int B, void* workspace, size_t workspace_bytes, const float* dx_rows,
                                      const int32_t* row_index, void* stream);

/* ---- measurement ----------------------------------------------------------------------------------------------
 * Kernel classes: 0 tcgen05 GEMM (all), 1 LayerNorm, 2 attention, 3 LoRA down-projection, 4 head, 5 preprocess, 6 misc,
 * 7-11 the GEMM launches again by shape (qkv, out_proj, c_fc, c_proj, other): n >= 12.
 * iic_profile(h, 1) starts counting launches and bracketing every launch with CUDA events on the caller's stream;
 * iic_profile_read sums the elapsed milliseconds and launch counts per class since the last read and resets them
 * (it synchronises on the last recorded event).  iic_profile(h, 0) keeps counting launches but records no events. */
int iic_profile(iic_handle* h, int enable);
int iic_profile_read(iic_handle* h, double* ms_by_class, long long* launches_by_class, int n);

/* ---- single operators (exported for parity tests and profiling; same kernels the encoder runs) -------------- */
/* D = epilogue(A[M,K] . W[N,K]^T (+ P[M,r] . Bt[N,r]^T));  epilogue: 0 bias->bf16, 1 bias+QuickGELU->bf16,
 * 2 bias+residual->f32, 3 pos-emb scatter->f32, 4 bias+GELU(erf)->bf16, 8 (training backward of mlp.c_proj + activation)
 * acc * act'(u) -> 16-bit with u = `residual` reinterpreted as the 16-bit pre-activation [M, ldc] and group = 1 QuickGELU /
 * 2 erf GELU.  lora_p/lora_bt nullable.
 * down_a f32 [N,4] / down_part f32 [2*ceil(N/256)][M][4] (nullable, activation epilogues): the consumer's LoRA
 * down-projection fused into this epilogue as per-column-tile partials. */
int iic_op_gemm(iic_handle* h, const void* a, int lda, const void* w, int ldw, int M, int N, int K, const void* lora_p,
                const void* lora_bt, int r_pad, int lora_ld, int epilogue, const float* bias, const float* residual,
                void* out, int ldc, int group, int ctas, const float* down_a, float* down_part, void* stream);
/* Training forward of mlp.c_fc (train_lora.py:233 through LoRALinear.forward, train_lora.py:43-44): out_act 16-bit [M,N] =
 * act(A . W^T (+ LoRA) + bias) and out_pre 16-bit [M,N] = the pre-activation the backward needs, from one pass.
 * act: IIC_ACT_QUICK_GELU / IIC_ACT_GELU_ERF. */
int iic_op_gemm_act_dual(iic_handle* h, const void* a, int lda, const void* w, int ldw, int M, int N, int K, const void* lora_p,
                         const void* lora_bt, int r_pad, int lora_ld, const float* bias, void* out_act, void* out_pre, int act,
                         int ctas, void* stream);
int iic_op_layernorm(iic_handle* h, const float* x, const float* gamma, const float* beta, void* out_bf16,
                     float* out_f32, int rows, int D, const float* lora_a_scaled, int r4, void* p_out, int p_ld,
                     void* stream);
int iic_op_lora_down(iic_handle* h, const void* x_bf16, int K, int rows, const float* lora_a_scaled, int r4, void* p_out,
                     int p_ld, void* stream);
/* impl: 0 = what the encoder uses (tcgen05/TMEM kernel while K/V of one head fit in smem, T <= ~760; else mma.sync),
 * 1 = mma.sync kernel, 2 = tcgen05 kernel */
int iic_op_attention(iic_handle* h, const void* qkv_bf16, void* out_bf16, int B, int T, int heads, int impl, void* stream);
/* backward operators (same kernels iic_train_backward runs).  iic_op_attention_bwd recomputes the forward into `out`
 * (to obtain the log-sum-exp) and then writes dqkv [M, 3*d].  lse_scratch f32 [2*B*heads*T]: the log-sum-exp followed by
 * the backward's D = rowsum(dO o O) scratch. */
int iic_op_attention_bwd(iic_handle* h, const void* qkv, void* out, const void* d_out, void* dqkv, float* lse_scratch,
                         int B, int T, int heads, void* stream);
int iic_op_layernorm_bwd(iic_handle* h, const void* dy, const float* x, const float* gamma, float* dx, void* dx16, int rows,
                         int D, void* stream);
int iic_op_act_bwd(iic_handle* h, void* dh, const void* u, long long n, int act, void* stream);
/* Both gradients of a LoRA pair (main.py:30-31, 42-43) that read the output gradient Y 16-bit [M, N], in one pass:
 * out_db f32 [rank, N] = scale * P^T . Y with P 16-bit [M, p_ld = 16] the forward's s * x . A, and out_dp16 16-bit [M, 16] =
 * Y . Bm^T with Bm 16-bit [16, N] = lora_B (rows >= rank zero).  N % 256 == 0. */
int iic_op_lora_bwd(iic_handle* h, const void* P, int p_ld, const void* Y, int N, int M, const void* Bm, int rank, float scale,
                    float* out_db, void* out_dp16, void* scratch, size_t scratch_bytes, void* stream);
int iic_op_lora_outer(iic_handle* h, const void* P, int p_ld, const void* Y, int N, int M, int act, int rank, float scale,
                      int transpose, float* out, void* scratch, size_t scratch_bytes, void* stream);
/* Bytes of caller-owned device scratch the two LoRA gradient operators above need for an [M, N] gradient (deterministic
 * two-stage reduction: per-row-block partials).  The operators allocate nothing and do not synchronise the stream. */
size_t iic_op_lora_scratch_bytes(int N, int M);

#ifdef __cplusplus
}
#endif
#endif /* IIC_H_ */
