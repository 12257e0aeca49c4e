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
#include "ptx_sm100.cuh"

namespace iic {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LayerNorm rows (fp32 in, fp32 stats) -> 16-bit GEMM operand and/or fp32, optionally fused with the LoRA
// down-projection P[row, :] = xln[row, :] . (scaling * lora_A) for the projection that consumes the normalised row.
//   kVec  = D / 128 float4 chunks per lane (D = 768 -> 6, 1024 -> 8)
//   kMode = 0: no LoRA;  1: r4 == 4 with the lane's slice of A held in registers (96 floats at D = 768), amortised
//           over every row the warp processes;  2: any r4 (multiple of 4), A re-read through L1 per row.
// Each warp walks rows warp_id, warp_id + n_warps, ... and prefetches the next row while it works on the current one.
// P rows are written full width (p_ld columns, zeros beyond r4): the GEMM's LoRA k-step reads 16 columns.
template <int kVec, bool kF16, int kMode>
__global__ void __launch_bounds__(256, kMode == 1 ? 1 : 2)
layernorm_kernel(const float* __restrict__ x, long long x_row_stride, const float* __restrict__ gamma,
                 const float* __restrict__ beta, uint16_t* __restrict__ out_bf16, float* __restrict__ out_f32,
                 long long out_row_stride, int rows, float eps, const float* __restrict__ lora_a, int r4,
                 uint16_t* __restrict__ p_out, int p_ld) {
  constexpr int D = kVec * 128;
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);

  float4 areg[kMode == 1 ? kVec * 4 : 1];
  if constexpr (kMode == 1) {
#pragma unroll
    for (int j = 0; j < kVec; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        areg[j * 4 + q] = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(4 * (lane + 32 * j) + q) * 4));
  }

  float4 nxt[kVec];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + size_t(row) * x_row_stride);
#pragma unroll
    for (int j = 0; j < kVec; ++j) nxt[j] = xr[lane + 32 * j];
  }
  for (; row < rows; row += n_warps) {
    float4 v[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) v[j] = nxt[j];
    if (row + n_warps < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + size_t(row + n_warps) * x_row_stride);
#pragma unroll
      for (int j = 0; j < kVec; ++j) nxt[j] = xr[lane + 32 * j];
    }
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
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const float4 g = __ldg(g4 + lane + 32 * j), b = __ldg(b4 + lane + 32 * j);
      v[j].x = (v[j].x - mean) * rstd * g.x + b.x;
      v[j].y = (v[j].y - mean) * rstd * g.y + b.y;
      v[j].z = (v[j].z - mean) * rstd * g.z + b.z;
      v[j].w = (v[j].w - mean) * rstd * g.w + b.w;
    }
    if (out_f32 != nullptr) {
      float4* o = reinterpret_cast<float4*>(out_f32 + size_t(row) * out_row_stride);
#pragma unroll
      for (int j = 0; j < kVec; ++j) o[lane + 32 * j] = v[j];
    }
    if (out_bf16 != nullptr) {
      uint2* o = reinterpret_cast<uint2*>(out_bf16 + size_t(row) * out_row_stride);
#pragma unroll
      for (int j = 0; j < kVec; ++j) {
        uint2 pk;
        pk.x = Act<kF16>::pack(v[j].x, v[j].y);
        pk.y = Act<kF16>::pack(v[j].z, v[j].w);
        o[lane + 32 * j] = pk;
      }
    }
    if constexpr (kMode == 1) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < kVec; ++j) {
        const float4 w0 = areg[j * 4], w1 = areg[j * 4 + 1], w2 = areg[j * 4 + 2], w3 = areg[j * 4 + 3];
        a0 += v[j].x * w0.x + v[j].y * w1.x + v[j].z * w2.x + v[j].w * w3.x;
        a1 += v[j].x * w0.y + v[j].y * w1.y + v[j].z * w2.y + v[j].w * w3.y;
        a2 += v[j].x * w0.z + v[j].y * w1.z + v[j].z * w2.z + v[j].w * w3.z;
        a3 += v[j].x * w0.w + v[j].y * w1.w + v[j].z * w2.w + v[j].w * w3.w;
      }
      a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
      // lanes 0 .. p_ld/4-1 each write one 4-column group: group 0 carries the values, the rest are zeros
      if (lane * 4 < p_ld) {
        uint2 pk = make_uint2(0u, 0u);
        if (lane == 0) { pk.x = Act<kF16>::pack(a0, a1); pk.y = Act<kF16>::pack(a2, a3); }
        *reinterpret_cast<uint2*>(p_out + size_t(row) * p_ld + lane * 4) = pk;
      }
    } else if constexpr (kMode == 2) {
      for (int c0 = 0; c0 < p_ld; c0 += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (c0 < r4) {
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
        }
        if (lane == 0) {
          uint2 pk;
          pk.x = Act<kF16>::pack(a0, a1);
          pk.y = Act<kF16>::pack(a2, a3);
          *reinterpret_cast<uint2*>(p_out + size_t(row) * p_ld + c0) = pk;
        }
      }
    }
  }
}

// P[row, 0..p_ld) = X[row, :] . A (zeros beyond r4) for 16-bit X [rows, K] (K % 8 == 0).  Stand-alone fallback used
// where the down-projection cannot ride in a producer kernel (attn.out_proj slot, ranks other than 4): one warp per
// kRows rows so every A row fetched from L1 is used kRows times; fp32 accumulate.
template <bool kF16, int kRows>
__global__ void __launch_bounds__(256)
lora_down_bf16_kernel(const uint16_t* __restrict__ x, int K, int rows, const float* __restrict__ lora_a, int r4,
                      uint16_t* __restrict__ p_out, int p_ld) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * kRows;
  if (row0 >= rows) return;
  for (int c0 = 0; c0 < p_ld; c0 += 4) {
    float acc[kRows][4];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    if (c0 < r4) {
      for (int ch = lane; ch < K / 8; ch += 32) {
        float4 w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = __ldg(reinterpret_cast<const float4*>(lora_a + size_t(ch * 8 + e) * r4 + c0));
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (row0 + r < rows) {
            const uint4 raw = *reinterpret_cast<const uint4*>(x + size_t(row0 + r) * K + ch * 8);
            const uint32_t* h = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 xv = Act<kF16>::unpack(h[e]);
              acc[r][0] += xv.x * w[2 * e].x + xv.y * w[2 * e + 1].x;
              acc[r][1] += xv.x * w[2 * e].y + xv.y * w[2 * e + 1].y;
              acc[r][2] += xv.x * w[2 * e].z + xv.y * w[2 * e + 1].z;
              acc[r][3] += xv.x * w[2 * e].w + xv.y * w[2 * e + 1].w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float a0 = warp_sum(acc[r][0]), a1 = warp_sum(acc[r][1]), a2 = warp_sum(acc[r][2]), a3 = warp_sum(acc[r][3]);
      if (lane == 0 && row0 + r < rows) {
        uint2 pk;
        pk.x = Act<kF16>::pack(a0, a1);
        pk.y = Act<kF16>::pack(a2, a3);
        *reinterpret_cast<uint2*>(p_out + size_t(row0 + r) * p_ld + c0) = pk;
      }
    }
  }
}

// Finishes the down-projection the c_fc GEMM epilogue started: part[n_tile][row][4] fp32 partial dot products
// (one per 256-column tile of the hidden activation) -> P[row, 0..p_ld) 16-bit, zeros beyond column 4.
// Deterministic (fixed summation order), unlike an atomic accumulation.
template <bool kF16>
__global__ void __launch_bounds__(256)
lora_reduce_kernel(const float* __restrict__ part, int n_tiles, int rows, uint16_t* __restrict__ p_out, int p_ld) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < n_tiles; ++t) {
    const float4 v = *reinterpret_cast<const float4*>(part + (size_t(t) * rows + row) * 4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  uint2 pk;
  pk.x = Act<kF16>::pack(s.x, s.y);
  pk.y = Act<kF16>::pack(s.z, s.w);
  uint2* o = reinterpret_cast<uint2*>(p_out + size_t(row) * p_ld);
  o[0] = pk;
  for (int c = 1; c < p_ld / 4; ++c) o[c] = make_uint2(0u, 0u);
}

// x_pre[b*T + 0, :] = class_embedding + positional_embedding[0]
__global__ void fill_cls_kernel(float* __restrict__ x_pre, const float* __restrict__ cls, const float* __restrict__ pos,
                                int B, int T, int D) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
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
template <int kVec, bool kF16>
static int layernorm_launch(const float* x, long long xs, const float* gamma, const float* beta, uint16_t* ob, float* of,
                            long long ors, int rows, float eps, const float* la, int r4, uint16_t* po, int pld,
                            cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int threads = 256, wpb = threads / 32;
  const int need = (rows + wpb - 1) / wpb;
  if (la == nullptr) {
    const int blocks = need < sms * 8 ? need : sms * 8;   // <= 8 resident CTAs/SM worth of warps, grid-stride beyond
    launch_k(layernorm_kernel<kVec, kF16, 0>, dim3(blocks), dim3(threads), 0, stream, x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld);
  } else if (r4 == 4) {
    const int blocks = need < sms ? need : sms;           // A slice lives in registers (1 CTA/SM): few, long-lived warps
    launch_k(layernorm_kernel<kVec, kF16, 1>, dim3(blocks), dim3(threads), 0, stream, x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld);
  } else {
    const int blocks = need < sms * 8 ? need : sms * 8;
    launch_k(layernorm_kernel<kVec, kF16, 2>, dim3(blocks), dim3(threads), 0, stream, x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

template <bool kF16>
static int layernorm_dispatch(const float* x, long long xs, const float* gamma, const float* beta, uint16_t* ob,
                              float* of, long long ors, int rows, int D, float eps, const float* la, int r4, uint16_t* po,
                              int pld, cudaStream_t stream) {
  switch (D) {
    case 512: return layernorm_launch<4, kF16>(x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld, stream);
    case 768: return layernorm_launch<6, kF16>(x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld, stream);
    case 1024: return layernorm_launch<8, kF16>(x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld, stream);
    case 1280: return layernorm_launch<10, kF16>(x, xs, gamma, beta, ob, of, ors, rows, eps, la, r4, po, pld, stream);
    default: return -1;
  }
}

int launch_layernorm(const float* x, long long x_row_stride, const float* gamma, const float* beta, void* out_bf16,
                     float* out_f32, long long out_row_stride, int rows, int D, float eps, const float* lora_a, int r4,
                     void* p_out, int p_ld, int f16, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (lora_a != nullptr && (r4 % 4 != 0 || p_ld % 4 != 0 || r4 > p_ld || p_ld > 128)) return -1;
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
  if (K % 8 != 0 || r4 % 4 != 0 || p_ld % 4 != 0 || r4 > p_ld) return -1;
  constexpr int kRows = 4;
  const int threads = 256;
  const int warps = (rows + kRows - 1) / kRows;
  const int blocks = (warps + 7) / 8;
  if (f16)
    lora_down_bf16_kernel<true, kRows><<<blocks, threads, 0, stream>>>(static_cast<const uint16_t*>(x), K, rows, lora_a, r4,
                                                                       static_cast<uint16_t*>(p_out), p_ld);
  else
    lora_down_bf16_kernel<false, kRows><<<blocks, threads, 0, stream>>>(static_cast<const uint16_t*>(x), K, rows, lora_a,
                                                                        r4, static_cast<uint16_t*>(p_out), p_ld);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_lora_reduce(const float* part, int n_tiles, int rows, void* p_out, int p_ld, int f16, cudaStream_t stream) {
  if (rows <= 0) return 0;
  if (p_ld % 4 != 0 || p_ld < 4) return -1;
  const int blocks = (rows + 255) / 256;
  if (f16)
    launch_k(lora_reduce_kernel<true>, dim3(blocks), dim3(256), 0, stream, part, n_tiles, rows, static_cast<uint16_t*>(p_out), p_ld);
  else
    launch_k(lora_reduce_kernel<false>, dim3(blocks), dim3(256), 0, stream, part, n_tiles, rows, static_cast<uint16_t*>(p_out), p_ld);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// text tower: the EOT row of every sequence, out[b, :] = x[b * T + row_index[b], :]
__global__ void gather_rows_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_index, int T, int D4, int B,
                                   float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D4) return;
  const int b = i / D4, c = i - b * D4;
  int r = row_index[b];
  r = r < 0 ? 0 : (r >= T ? T - 1 : r);
  out[i] = reinterpret_cast<const float4*>(x)[(size_t(b) * T + r) * D4 + c];
}

int launch_gather_rows(const float* x, const int32_t* row_index, int T, int D, int B, float* out, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (D % 4 != 0) return -1;
  const int n = B * (D / 4);
  gather_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x, row_index, T, D / 4, B, reinterpret_cast<float4*>(out));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// training of the text tower: dx[b * T + row_index[b], :] = d_rows[b, :] (dx zeroed by the caller)
__global__ void scatter_rows_kernel(const float4* __restrict__ rows, const int32_t* __restrict__ row_index, int T, int D4, int B,
                                    float4* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D4) return;
  const int b = i / D4, c = i - b * D4;
  int r = row_index[b];
  r = r < 0 ? 0 : (r >= T ? T - 1 : r);
  x[(size_t(b) * T + r) * D4 + c] = rows[i];
}

int launch_scatter_rows(const float* rows, const int32_t* row_index, int T, int D, int B, float* x, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (D % 4 != 0) return -1;
  const int n = B * (D / 4);
  scatter_rows_kernel<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const float4*>(rows), row_index, T, D / 4, B,
                                                            reinterpret_cast<float4*>(x));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_fill_cls(float* x_pre, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream) {
  const int n = B * D;
  if (n <= 0) return 0;
  launch_k(fill_cls_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, x_pre, cls, pos, B, T, D);
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
