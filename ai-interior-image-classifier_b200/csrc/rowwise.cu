// Row-wise, HBM-bound kernels of the encoder: LayerNorm (+ fused LoRA down-projection), the stand-alone LoRA
// down-projection for bf16 inputs, class-token fill, CHW -> patch-matrix re-index.
//
// Reference semantics (OpenAI CLIP model.py as called from /root/reference/main.py:204,444,503):
//   LayerNorm subclass computes in fp32 with eps = 1e-5 and casts back; here the residual stream is kept in
//   fp32 end-to-end and the normalised rows are emitted as bf16 GEMM operands.
//   LoRALayer.forward (/root/reference/main.py:30-31): (x @ A @ B) * scaling  -> P = x @ (scaling*A) is produced
//   here (fp32 math, bf16 store) and B is applied inside the GEMM tile (gemm_sm100.cuh).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_types.cuh"
#include "kernels.h"

namespace iic {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row.  kVec = D / 128 float4 chunks per lane (D = 768 -> 6, 1024 -> 8).
// out_bf16 / out_f32 may each be null.  If lora_a != null also emits P[row, 0..r_pad) = xln . lora_a
// (lora_a is [D, r4] fp32 with the LoRA scaling already folded in, r4 = rank rounded up to 4; P row pitch = p_ld).
template <int kVec, bool kF16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, long long x_row_stride, const float* __restrict__ gamma,
                 const float* __restrict__ beta, uint16_t* __restrict__ out_bf16, float* __restrict__ out_f32,
                 long long out_row_stride, int rows, float eps, const float* __restrict__ lora_a, int r4,
                 uint16_t* __restrict__ p_out, int p_ld) {
  constexpr int D = kVec * 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + size_t(warp) * x_row_stride);
  float4 v[kVec];
#pragma unroll
  for (int j = 0; j < kVec; ++j) v[j] = xr[lane + 32 * j];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kVec; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int j = 0; j < kVec; ++j) {
    const float4 g = __ldg(g4 + lane + 32 * j), b = __ldg(b4 + lane + 32 * j);
    v[j].x = (v[j].x - mean) * rstd * g.x + b.x;
    v[j].y = (v[j].y - mean) * rstd * g.y + b.y;
    v[j].z = (v[j].z - mean) * rstd * g.z + b.z;
    v[j].w = (v[j].w - mean) * rstd * g.w + b.w;
  }
  if (out_f32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(out_f32 + size_t(warp) * out_row_stride);
#pragma unroll
    for (int j = 0; j < kVec; ++j) o[lane + 32 * j] = v[j];
  }
  if (out_bf16 != nullptr) {
    uint2* o = reinterpret_cast<uint2*>(out_bf16 + size_t(warp) * out_row_stride);
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      uint2 pk;
      pk.x = Act<kF16>::pack(v[j].x, v[j].y);
      pk.y = Act<kF16>::pack(v[j].z, v[j].w);
      o[lane + 32 * j] = pk;
    }
  }
  if (lora_a != nullptr) {
    // P[row, c] = sum_k xln[k] * A[k, c];  4 columns at a time, A rows read as float4 (L1/L2 resident)
    for (int c0 = 0; c0 < r4; c0 += 4) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < kVec; ++j) {
        const int k = 4 * (lane + 32 * j);
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k) * r4 + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k + 1) * r4 + c0));
        const float4 w2 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k + 2) * r4 + c0));
        const float4 w3 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k + 3) * r4 + c0));
        a0 += v[j].x * w0.x + v[j].y * w1.x + v[j].z * w2.x + v[j].w * w3.x;
        a1 += v[j].x * w0.y + v[j].y * w1.y + v[j].z * w2.y + v[j].w * w3.y;
        a2 += v[j].x * w0.z + v[j].y * w1.z + v[j].z * w2.z + v[j].w * w3.z;
        a3 += v[j].x * w0.w + v[j].y * w1.w + v[j].z * w2.w + v[j].w * w3.w;
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
      if (lane == 0) {
        uint2 pk;
        pk.x = Act<kF16>::pack(a0, a1);
        pk.y = Act<kF16>::pack(a2, a3);
        *reinterpret_cast<uint2*>(p_out + size_t(warp) * p_ld + c0) = pk;
      }
    }
  }
}

// P[row, 0..r4) = X[row, :] . A   for bf16 X [rows, K] (K % 256 == 0).  One warp per row, fp32 accumulate.
template <bool kF16>
__global__ void __launch_bounds__(256)
lora_down_bf16_kernel(const uint16_t* __restrict__ x, int K, int rows, const float* __restrict__ lora_a, int r4,
                      uint16_t* __restrict__ p_out, int p_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + size_t(warp) * K);
  for (int c0 = 0; c0 < r4; c0 += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int ch = lane; ch < K / 8; ch += 32) {
      const uint4 raw = xr[ch];
      const uint32_t* h = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 xv = Act<kF16>::unpack(h[e]);
        const int k = ch * 8 + 2 * e;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k) * r4 + c0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(k + 1) * r4 + c0));
        a0 += xv.x * w0.x + xv.y * w1.x;
        a1 += xv.x * w0.y + xv.y * w1.y;
        a2 += xv.x * w0.z + xv.y * w1.z;
        a3 += xv.x * w0.w + xv.y * w1.w;
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    if (lane == 0) {
      uint2 pk;
      pk.x = Act<kF16>::pack(a0, a1);
      pk.y = Act<kF16>::pack(a2, a3);
      *reinterpret_cast<uint2*>(p_out + size_t(warp) * p_ld + c0) = pk;
    }
  }
}

// x_pre[b*T + 0, :] = class_embedding + positional_embedding[0]
__global__ void fill_cls_kernel(float* __restrict__ x_pre, const float* __restrict__ cls, const float* __restrict__ pos,
                                int B, int T, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x_pre[size_t(b) * T * D + d] = cls[d] + pos[d];
}

// float/bf16 CHW image batch [B,3,R,R] -> bf16 patch matrix [B*g*g, k_pad], column = c*P*P + ky*P + kx
// (the im2col of a stride==kernel conv is a pure re-index).  One thread per (b, c, y, patch-x) handles P pixels.
template <typename T, bool kF16>
__global__ void __launch_bounds__(256)
chw_to_patches_kernel(const T* __restrict__ img, typename Act<kF16>::T* __restrict__ patches, int B, int R, int P,
                      int k_pad) {
  const int g = R / P;
  const long long total = (long long)B * 3 * R * g;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int px = int(i % g);
  long long t = i / g;
  const int y = int(t % R); t /= R;
  const int c = int(t % 3);
  const int b = int(t / 3);
  const int py = y / P, ky = y - py * P;
  const T* src = img + ((size_t(b) * 3 + c) * R + y) * R + px * P;
  typename Act<kF16>::T* dst = patches + (size_t(b) * g * g + py * g + px) * k_pad + c * P * P + ky * P;
  for (int kx = 0; kx < P; ++kx) dst[kx] = Act<kF16>::from_float(float(src[kx]));
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
template <bool kF16>
static int layernorm_dispatch(const float* x, long long x_row_stride, const float* gamma, const float* beta, uint16_t* ob,
                              float* of, long long ors, int rows, int D, float eps, const float* la, int r4, uint16_t* po,
                              int pld, cudaStream_t stream) {
  const int threads = 256;
  const int blocks = (rows + (threads / 32) - 1) / (threads / 32);
  switch (D) {
    case 512: layernorm_kernel<4, kF16><<<blocks, threads, 0, stream>>>(x, x_row_stride, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld); break;
    case 768: layernorm_kernel<6, kF16><<<blocks, threads, 0, stream>>>(x, x_row_stride, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld); break;
    case 1024: layernorm_kernel<8, kF16><<<blocks, threads, 0, stream>>>(x, x_row_stride, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld); break;
    case 1280: layernorm_kernel<10, kF16><<<blocks, threads, 0, stream>>>(x, x_row_stride, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld); break;
    default: return -1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_layernorm(const float* x, long long x_row_stride, const float* gamma, const float* beta, void* out_bf16,
                     float* out_f32, long long out_row_stride, int rows, int D, float eps, const float* lora_a, int r4,
                     void* p_out, int p_ld, int f16, cudaStream_t stream) {
  if (rows <= 0) return 0;
  uint16_t* ob = static_cast<uint16_t*>(out_bf16);
  uint16_t* po = static_cast<uint16_t*>(p_out);
  return f16 ? layernorm_dispatch<true>(x, x_row_stride, gamma, beta, ob, out_f32, out_row_stride, rows, D, eps, lora_a,
                                        r4, po, p_ld, stream)
             : layernorm_dispatch<false>(x, x_row_stride, gamma, beta, ob, out_f32, out_row_stride, rows, D, eps, lora_a,
                                         r4, po, p_ld, stream);
}

int launch_lora_down_bf16(const void* x, int K, int rows, const float* lora_a, int r4, void* p_out, int p_ld, int f16,
                          cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (K % 8 != 0) return -1;
  const int threads = 256;
  const int blocks = (rows + 7) / 8;
  if (f16)
    lora_down_bf16_kernel<true><<<blocks, threads, 0, stream>>>(static_cast<const uint16_t*>(x), K, rows, lora_a, r4,
                                                                static_cast<uint16_t*>(p_out), p_ld);
  else
    lora_down_bf16_kernel<false><<<blocks, threads, 0, stream>>>(static_cast<const uint16_t*>(x), K, rows, lora_a, r4,
                                                                 static_cast<uint16_t*>(p_out), p_ld);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_fill_cls(float* x_pre, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream) {
  const int n = B * D;
  if (n <= 0) return 0;
  fill_cls_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x_pre, cls, pos, B, T, D);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <bool kF16>
static int patches_dispatch(const void* img, int dtype, void* patches, int B, int R, int P, int k_pad, int blocks,
                            cudaStream_t stream) {
  using OT = typename Act<kF16>::T;
  if (dtype == 0)
    chw_to_patches_kernel<float, kF16><<<blocks, 256, 0, stream>>>(static_cast<const float*>(img), static_cast<OT*>(patches), B, R, P, k_pad);
  else if (dtype == 1)
    chw_to_patches_kernel<__nv_bfloat16, kF16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(img), static_cast<OT*>(patches), B, R, P, k_pad);
  else if (dtype == 2)
    chw_to_patches_kernel<__half, kF16><<<blocks, 256, 0, stream>>>(static_cast<const __half*>(img), static_cast<OT*>(patches), B, R, P, k_pad);
  else
    return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_chw_to_patches(const void* img, int dtype, void* patches, int B, int R, int P, int k_pad, int f16,
                          cudaStream_t stream) {
  const long long total = (long long)B * 3 * R * (R / P);
  if (total <= 0) return 0;
  const int blocks = int((total + 255) / 256);
  return f16 ? patches_dispatch<true>(img, dtype, patches, B, R, P, k_pad, blocks, stream)
             : patches_dispatch<false>(img, dtype, patches, B, R, P, k_pad, blocks, stream);
}

}  // namespace iic
