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
  for (int e = tid; e < E && emb_in == nullptr; e += kThreads) {
    float acc[kImgs];
#pragma unroll
    for (int i = 0; i < kImgs; ++i) acc[i] = 0.f;
    // 16 projection rows in flight per thread: each CTA streams the whole [W, E] matrix from L2 exactly once, so the loop is
    // bound by load latency, not bandwidth (4 in flight: 0.5 ms per 1024 images; same summation order, same bits)
    for (int k0 = 0; k0 < W; k0 += 16) {
      float wv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) wv[j] = (k0 + j < W) ? __ldg(proj + size_t(k0 + j) * E + e) : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (k0 + j < W) {
#pragma unroll
          for (int i = 0; i < kImgs; ++i) acc[i] = fmaf(s_x[i * W + k0 + j], wv[j], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kImgs; ++i) {
      s_e[i * E + e] = acc[i];
      if (i < n_img && emb_out != nullptr) emb_out[size_t(img0 + i) * E + e] = acc[i];
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

int launch_head(const float* x, long long x_img_stride, const float* ln_g, const float* ln_b, float eps,
                const float* proj, int W, int E, const float* text, int L, const int* group_off,
                const int* group_split, int G, int topk, float logit_scale, int B, float* emb_out, float* logits_out,
                float* probs_out, float* topk_val, int* topk_idx, float* split_sum, const float* emb_in,
                cudaStream_t stream) {
  if (B <= 0) return 0;
  if (topk > kMaxTopk || topk < 0) return -1;
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
