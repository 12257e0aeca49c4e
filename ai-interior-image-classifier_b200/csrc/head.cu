// Fused scoring head:  ln_post(CLS) -> @ visual.proj -> L2-normalise -> 100 * cos vs label text embeddings ->
// per-group softmax -> per-group top-k (+ optional leading-span probability sum for the interior detector).
//
// Reference semantics, all fp32:
//   encode_image tail (OpenAI CLIP VisionTransformer.forward): x = ln_post(x[:, 0, :]); x = x @ proj
//   /root/reference/main.py:205,445,504   f = f / f.norm(dim=-1, keepdim=True)          (no epsilon)
//   /root/reference/main.py:208,456,506   sims = (100.0 * f @ T.T).softmax(dim=-1)      (per label group)
//   /root/reference/main.py:211,457,507   vals, inds = sims[0].topk(min(5, |group|))
//   /root/reference/main.py:216-217       sum of the first 11 detector probabilities vs the remaining 29
//
// kImgs images per CTA so the [width, E] projection and the [L, E] label matrix are read from L2 once per kImgs.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "kernels.h"
#include "ptx_sm100.cuh"

namespace iic {

namespace {
constexpr int kThreads = 256;
constexpr int kImgs = 4;
constexpr int kMaxTopk = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
}  // namespace

// smem layout (floats): xln[kImgs][W] | emb[kImgs][E] | logits[kImgs][L] | red[kImgs][8]
__global__ void __launch_bounds__(kThreads)
head_kernel(const float* __restrict__ x, long long x_img_stride, const float* __restrict__ ln_g,
            const float* __restrict__ ln_b, float eps, const float* __restrict__ proj, int W, int E,
            const float* __restrict__ text, int L, const int* __restrict__ group_off, const int* __restrict__ group_split,
            int G, int topk, float logit_scale, int B, float* __restrict__ emb_out, float* __restrict__ logits_out,
            float* __restrict__ probs_out, float* __restrict__ topk_val, int* __restrict__ topk_idx,
            float* __restrict__ split_sum, const float* __restrict__ emb_in) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float sm[];
  float* s_x = sm;
  float* s_e = s_x + kImgs * W;
  float* s_l = s_e + kImgs * E;
  float* s_r = s_l + kImgs * L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int img0 = blockIdx.x * kImgs;
  const int n_img = min(kImgs, B - img0);

  if (emb_in != nullptr) {
    // embeddings supplied by the caller (iic_head): skip ln_post + projection
    for (int i = tid; i < kImgs * E; i += kThreads) {
      const int im = i / E;
      s_e[i] = im < n_img ? emb_in[size_t(img0) * E + i] : 0.f;
    }
  }
  // ---- 1. ln_post on the class-token rows: one warp per image ----
  if (emb_in == nullptr && warp < kImgs) {
    if (warp < n_img) {
      const float* xr = x + size_t(img0 + warp) * x_img_stride;
      float s = 0.f;
      for (int k = lane; k < W; k += 32) s += xr[k];
      const float mean = warp_sum(s) / W;
      float q = 0.f;
      for (int k = lane; k < W; k += 32) { const float dlt = xr[k] - mean; q += dlt * dlt; }
      const float rstd = rsqrtf(warp_sum(q) / W + eps);
      for (int k = lane; k < W; k += 32) s_x[warp * W + k] = (xr[k] - mean) * rstd * ln_g[k] + ln_b[k];
    } else {
      for (int k = lane; k < W; k += 32) s_x[warp * W + k] = 0.f;
    }
  }
  __syncthreads();

  // ---- 2. projection: emb[i][e] = sum_k xln[i][k] * proj[k][e]   (proj row-major [W, E], coalesced over e) ----
  // Summation order (shared with the small-batch path below, so that an image's result does not depend on the batch it travels
  // in): K is cut into 16 chunks of ceil(W / 16) rows, each chunk is a sequential fmaf chain in k order, the 16 partial sums
  // are added left to right.
  for (int e = tid; e < E && emb_in == nullptr; e += kThreads) {
    float tot[kImgs];
#pragma unroll
    for (int i = 0; i < kImgs; ++i) tot[i] = 0.f;
    const int kw = (W + 15) / 16;
    for (int c = 0; c < 16; ++c) {
      const int k_lo = c * kw, k_hi = min(W, k_lo + kw);
      float acc[kImgs];
#pragma unroll
      for (int i = 0; i < kImgs; ++i) acc[i] = 0.f;
      // 16 projection rows in flight per thread: each CTA streams the whole [W, E] matrix from L2 exactly once, so the loop is
      // bound by load latency, not bandwidth
      for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
        float wv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) wv[j] = (k0 + j < k_hi) ? __ldg(proj + size_t(k0 + j) * E + e) : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (k0 + j < k_hi) {
#pragma unroll
            for (int i = 0; i < kImgs; ++i) acc[i] = fmaf(s_x[i * W + k0 + j], wv[j], acc[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kImgs; ++i) tot[i] += acc[i];
    }
#pragma unroll
    for (int i = 0; i < kImgs; ++i) {
      s_e[i * E + e] = tot[i];
      if (i < n_img && emb_out != nullptr) emb_out[size_t(img0 + i) * E + e] = tot[i];
    }
  }
  __syncthreads();

  // ---- 3. L2 norm (no epsilon, as the reference): one warp per image ----
  if (warp < kImgs) {
    float q = 0.f;
    for (int e = lane; e < E; e += 32) { const float v = s_e[warp * E + e]; q += v * v; }
    q = warp_sum(q);
    if (lane == 0) s_r[warp] = rsqrtf(q);
  }
  __syncthreads();
  for (int i = tid; i < kImgs * E; i += kThreads) s_e[i] *= s_r[i / E];
  __syncthreads();

  // ---- 4. logits[i][l] = logit_scale * <f_i, text_l>: one warp per label, all kImgs images at once ----
  for (int l = warp; l < L; l += kThreads / 32) {
    const float* tr = text + size_t(l) * E;
    float acc[kImgs];
#pragma unroll
    for (int i = 0; i < kImgs; ++i) acc[i] = 0.f;
    for (int e0 = lane; e0 < E; e0 += 256) {   // eight label-row loads in flight per lane (same summation order)
      float tv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tv[j] = (e0 + 32 * j < E) ? __ldg(tr + e0 + 32 * j) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + 32 * j;
        if (e < E) {
#pragma unroll
          for (int i = 0; i < kImgs; ++i) acc[i] = fmaf(s_e[i * E + e], tv[j], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kImgs; ++i) {
      const float v = warp_sum(acc[i]) * logit_scale;
      if (lane == 0) {
        s_l[i * L + l] = v;
        if (i < n_img && logits_out != nullptr) logits_out[size_t(img0 + i) * L + l] = v;
      }
    }
  }
  __syncthreads();

  // ---- 5. per (image, group): softmax, top-k, leading-span sum: one warp per (image, group) pair ----
  for (int pair = warp; pair < n_img * G; pair += kThreads / 32) {
    const int i = pair / G, gidx = pair - i * G;
    const int lo = group_off[gidx], hi = group_off[gidx + 1];
    float* lg = s_l + i * L;
    float mx = -CUDART_INF_F;
    for (int l = lo + lane; l < hi; l += 32) mx = fmaxf(mx, lg[l]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int l = lo + lane; l < hi; l += 32) sum += expf(lg[l] - mx);
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    const int split = group_split != nullptr ? group_split[gidx] : 0;
    float ssum = 0.f;
    for (int l = lo + lane; l < hi; l += 32) {
      const float p = expf(lg[l] - mx) * inv;
      lg[l] = p;  // logits already written out; reuse the slot for the probability
      if (probs_out != nullptr) probs_out[size_t(img0 + i) * L + l] = p;
      if (l - lo < split) ssum += p;
    }
    ssum = warp_sum(ssum);
    if (lane == 0 && split_sum != nullptr) split_sum[size_t(img0 + i) * G + gidx] = ssum;
    __syncwarp();
    // iterative arg-max, ties -> lowest index (what a stable descending sort gives)
    const int kk = min(topk, hi - lo);
    for (int r = 0; r < topk; ++r) {
      float bv = -1.f;
      int bi = 0x7fffffff;
      if (r < kk) {
        for (int l = lo + lane; l < hi; l += 32) {
          const float p = lg[l];
          if (p > bv) { bv = p; bi = l; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
      }
      if (lane == 0) {
        const size_t oidx = (size_t(img0 + i) * G + gidx) * topk + r;
        if (r < kk) {
          topk_val[oidx] = bv;
          topk_idx[oidx] = bi - lo;  // index inside the group, like torch.topk on the group's slice
          lg[bi] = -2.f;             // remove from further rounds
        } else {
          topk_val[oidx] = 0.f;
          topk_idx[oidx] = -1;
        }
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Small batches (B <= kSmallB: the single-image calls of main.py:191-222, 472-510): the work of head_kernel spread over
// many CTAs in three short launches.  One CTA of 256 threads walking [W, E] and [L, E] for <= 4 images is ~0.29 ms of pure load
// latency - a quarter of the batch-1 latency of the whole path.  Same arithmetic in the same summation ORDER per output as
// head_kernel (16 K-chunks for the projection, the lane / shuffle-tree order for norms and logits), so a small batch gives
// the SAME BITS as the same images inside a large one (tests/test_parity_gpu.py::test_batch_invariance_and_determinism).
//   head_small_proj:   ln_post(CLS) per CTA (cheap, recomputed), then 32 output columns per CTA: lane = column, the 16 warps
//                      split K (all of a thread's projection rows in flight at once), cross-warp sum through smem
//   head_small_logits: row norm per CTA (recomputed), then one label per warp, lanes split E
//   head_small_groups: softmax / top-k / leading-span sum, one warp per (image, group)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kSmallB = 16;
constexpr int kSmallThreads = 512;

__global__ void __launch_bounds__(kSmallThreads)
head_small_proj_kernel(const float* __restrict__ x, long long x_img_stride, const float* __restrict__ ln_g,
                       const float* __restrict__ ln_b, float eps, const float* __restrict__ proj, int W, int E, int B,
                       float* __restrict__ emb_out) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float sm[];
  float* s_x = sm;                       // [B][W] normalised class-token rows
  float* s_p = s_x + size_t(B) * W;      // [16 warps][B][32] partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < B; i += kSmallThreads / 32) {
    const float* xr = x + size_t(i) * x_img_stride;
    float s = 0.f;
    for (int k = lane; k < W; k += 32) s += xr[k];
    const float mean = warp_sum(s) / W;
    float q = 0.f;
    for (int k = lane; k < W; k += 32) { const float dlt = xr[k] - mean; q += dlt * dlt; }
    const float rstd = rsqrtf(warp_sum(q) / W + eps);
    for (int k = lane; k < W; k += 32) s_x[i * W + k] = (xr[k] - mean) * rstd * ln_g[k] + ln_b[k];
  }
  __syncthreads();
  const int e = blockIdx.x * 32 + lane;
  const int kw = (W + 15) / 16;          // projection rows per warp
  const int k0 = warp * kw, k1 = min(W, k0 + kw);
  float acc[kSmallB];
#pragma unroll
  for (int i = 0; i < kSmallB; ++i) acc[i] = 0.f;
  for (int kb = k0; kb < k1; kb += 16) {
    float wv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wv[j] = (kb + j < k1 && e < E) ? __ldg(proj + size_t(kb + j) * E + e) : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (kb + j < k1) {
#pragma unroll
        for (int i = 0; i < kSmallB; ++i)
          if (i < B) acc[i] = fmaf(s_x[i * W + kb + j], wv[j], acc[i]);
      }
  }
#pragma unroll
  for (int i = 0; i < kSmallB; ++i)
    if (i < B) s_p[(warp * B + i) * 32 + lane] = acc[i];
  __syncthreads();
  for (int o = tid; o < B * 32; o += kSmallThreads) {
    const int i = o >> 5, l = o & 31;
    float v = 0.f;
    for (int w = 0; w < 16; ++w) v += s_p[(w * B + i) * 32 + l];   // fixed order: bit-reproducible
    if (blockIdx.x * 32 + l < E) emb_out[size_t(i) * E + blockIdx.x * 32 + l] = v;
  }
}

__global__ void __launch_bounds__(kSmallThreads)
head_small_logits_kernel(const float* __restrict__ emb, int E, const float* __restrict__ text, int L, float logit_scale, int B,
                         float* __restrict__ logits_out) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  extern __shared__ float sm[];
  float* s_e = sm;                       // [B][E] L2-normalised embeddings
  float* s_r = s_e + size_t(B) * E;      // [B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < B * E; i += kSmallThreads) s_e[i] = emb[i];
  __syncthreads();
  for (int i = warp; i < B; i += kSmallThreads / 32) {
    float q = 0.f;
    for (int k = lane; k < E; k += 32) { const float v = s_e[i * E + k]; q += v * v; }
    q = warp_sum(q);
    if (lane == 0) s_r[i] = rsqrtf(q);   // no epsilon, as the reference (main.py:205)
  }
  __syncthreads();
  const int l = blockIdx.x * (kSmallThreads / 32) + warp;
  if (l >= L) return;
  const float* tr = text + size_t(l) * E;
  float acc[kSmallB];
#pragma unroll
  for (int i = 0; i < kSmallB; ++i) acc[i] = 0.f;
  for (int e0 = lane; e0 < E; e0 += 32 * 16) {
    float tv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tv[j] = (e0 + 32 * j < E) ? __ldg(tr + e0 + 32 * j) : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int e = e0 + 32 * j;
      if (e < E) {
#pragma unroll
        for (int i = 0; i < kSmallB; ++i)
          if (i < B) acc[i] = fmaf(s_e[i * E + e] * s_r[i], tv[j], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kSmallB; ++i)
    if (i < B) {
      const float v = warp_sum(acc[i]) * logit_scale;
      if (lane == 0) logits_out[size_t(i) * L + l] = v;
    }
}

__global__ void __launch_bounds__(256)
head_small_groups_kernel(const float* __restrict__ logits, int L, const int* __restrict__ group_off,
                         const int* __restrict__ group_split, int G, int topk, int B, float* __restrict__ probs_out,
                         float* __restrict__ topk_val, int* __restrict__ topk_idx, float* __restrict__ split_sum,
                         float* __restrict__ scratch) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x * 8 + warp;
  if (pair >= B * G) return;
  const int i = pair / G, gidx = pair - i * G;
  const int lo = group_off[gidx], hi = group_off[gidx + 1];
  const float* lg = logits + size_t(i) * L;
  float* pr = scratch + size_t(i) * L;      // probabilities (working copy: entries are knocked out by the top-k rounds)
  float mx = -CUDART_INF_F;
  for (int l = lo + lane; l < hi; l += 32) mx = fmaxf(mx, lg[l]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int l = lo + lane; l < hi; l += 32) sum += expf(lg[l] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  const int split = group_split != nullptr ? group_split[gidx] : 0;
  float ssum = 0.f;
  for (int l = lo + lane; l < hi; l += 32) {
    const float p = expf(lg[l] - mx) * inv;
    pr[l] = p;
    if (probs_out != nullptr) probs_out[size_t(i) * L + l] = p;
    if (l - lo < split) ssum += p;
  }
  ssum = warp_sum(ssum);
  if (lane == 0 && split_sum != nullptr) split_sum[size_t(i) * G + gidx] = ssum;
  __syncwarp();
  const int kk = min(topk, hi - lo);
  for (int r = 0; r < topk; ++r) {
    float bv = -1.f;
    int bi = 0x7fffffff;
    if (r < kk) {
      for (int l = lo + lane; l < hi; l += 32) {
        const float p = pr[l];
        if (p > bv) { bv = p; bi = l; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
    }
    if (lane == 0) {
      const size_t oidx = (size_t(i) * G + gidx) * topk + r;
      if (r < kk) {
        topk_val[oidx] = bv;
        topk_idx[oidx] = bi - lo;
        pr[bi] = -2.f;
      } else {
        topk_val[oidx] = 0.f;
        topk_idx[oidx] = -1;
      }
    }
    __syncwarp();
  }
}
}  // namespace

size_t head_small_scratch_bytes(int B, int E, int L) {
  return (B > 0 && B <= kSmallB) ? sizeof(float) * (size_t(B) * E + 2 * size_t(B) * L) : 0;
}

int launch_head(const float* x, long long x_img_stride, const float* ln_g, const float* ln_b, float eps,
                const float* proj, int W, int E, const float* text, int L, const int* group_off,
                const int* group_split, int G, int topk, float logit_scale, int B, float* emb_out, float* logits_out,
                float* probs_out, float* topk_val, int* topk_idx, float* split_sum, const float* emb_in,
                float* small_scratch, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (topk > kMaxTopk || topk < 0) return -1;
  if (B <= kSmallB && small_scratch != nullptr && W <= 2048) {
    // ---- latency path: three short, wide launches ----
    float* emb = emb_out != nullptr ? emb_out : small_scratch;                       // [B, E]
    float* logits = logits_out != nullptr ? logits_out : small_scratch + size_t(B) * E;   // [B, L]
    float* pwork = small_scratch + size_t(B) * E + size_t(B) * L;                    // [B, L]
    static PerDeviceOnce attr_small;
    if (attr_small.need()) {
      if (cudaFuncSetAttribute(head_small_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(head_small_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return -2;
      attr_small.mark();
    }
    if (emb_in == nullptr) {
      const size_t smem = sizeof(float) * (size_t(B) * W + size_t(16) * B * 32);
      launch_k(head_small_proj_kernel, dim3((E + 31) / 32), dim3(kSmallThreads), smem, stream, x, x_img_stride, ln_g, ln_b, eps, proj, W, E, B, emb);
    } else {
      emb = const_cast<float*>(emb_in);
    }
    if (text == nullptr || L <= 0) return cudaGetLastError() == cudaSuccess ? 0 : -2;
    const size_t smem2 = sizeof(float) * (size_t(B) * E + 64);
    launch_k(head_small_logits_kernel, dim3((L + kSmallThreads / 32 - 1) / (kSmallThreads / 32)), dim3(kSmallThreads), smem2, stream,
             static_cast<const float*>(emb), E, text, L, logit_scale, B, logits);
    launch_k(head_small_groups_kernel, dim3((B * G + 7) / 8), dim3(256), 0, stream, static_cast<const float*>(logits), L, group_off,
             group_split, G, topk, B, probs_out, topk_val, topk_idx, split_sum, pwork);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  const size_t smem = sizeof(float) * (size_t(kImgs) * (W + E + L) + 64);
  if (smem > 200 * 1024) return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const int blocks = (B + kImgs - 1) / kImgs;
  head_kernel<<<blocks, kThreads, smem, stream>>>(x, x_img_stride, ln_g, ln_b, eps, proj, W, E, text, L, group_off,
                                                  group_split, G, topk, logit_scale, B, emb_out, logits_out, probs_out,
                                                  topk_val, topk_idx, split_sum, emb_in);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
