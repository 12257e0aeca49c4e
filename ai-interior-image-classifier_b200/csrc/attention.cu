// Fused softmax attention for the short ViT sequence (T = 197 ... 577 tokens, head_dim 64, no mask, no dropout).
//
// Replaces the scaled-dot-product core of nn.MultiheadAttention inside CLIP's ResidualAttentionBlock
// (called through model.encode_image at /root/reference/main.py:204, 444, 503):
//     softmax(Q K^T / sqrt(64)) V          per (image, head)
// Input is the packed in_proj output qkv[M, 3*d] (q | k | v, heads contiguous, 64 each), output is [M, d] bf16
// in the layout attn.out_proj's GEMM consumes.
//
// One CTA per (image, head): the whole K and V of that head live in shared memory (XOR-swizzled 128-byte rows,
// conflict-free ldmatrix), 4 warps each own 16-query tiles, keys are streamed in chunks of 64 with an online
// (running max / running sum) softmax kept in registers; row max/sum use warp shuffles inside the 4-lane quads.
// Tensor-core path: mma.sync m16n8k16 bf16 -> fp32 (attention is ~4% of the encoder FLOPs; the tcgen05 budget is
// spent on the projection GEMMs).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_types.cuh"
#include "kernels.h"

namespace iic {

namespace {

constexpr int kHd = 64;      // head dim
constexpr int kChunk = 64;   // keys per online-softmax step
constexpr int kWarps = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <bool kF16>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kF16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
// byte offset of (row, 16-byte chunk) inside a [rows][64] bf16 tile with 128-byte rows, XOR swizzled
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kWarps * 32, 4)
attention_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, float* __restrict__ lse, int T, int TP,
                 int H, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int d = H * kHd;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const uint32_t ks = smem_u32(smem);
  const uint32_t vs = ks + uint32_t(TP) * 128u;
  const uint16_t* base = qkv + size_t(b) * T * (3 * d) + h * kHd;

  // ---- stage K and V of this (image, head) ----
  for (int i = threadIdx.x; i < TP * 8; i += kWarps * 32) {
    const int row = i >> 3, ch = i & 7;
    if (row < T) {
      const uint16_t* src = base + size_t(row) * (3 * d) + ch * 8;
      cp_async16(ks + swz(row, ch), src + d);
      cp_async16(vs + swz(row, ch), src + 2 * d);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(smem + swz(row, ch)) = z;
      *reinterpret_cast<uint4*>(smem + size_t(TP) * 128 + swz(row, ch)) = z;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int n_qtiles = TP / 16;
  const int n_chunks = (TP + kChunk - 1) / kChunk;

  for (int qt = warp; qt < n_qtiles; qt += kWarps) {
    const int q0 = qt * 16;
    // ---- Q fragments straight from global (each element is read exactly once) ----
    uint32_t qf[4][4];
    {
      const int r0 = q0 + g, r1 = q0 + g + 8;
      const uint16_t* p0 = base + size_t(r0) * (3 * d);
      const uint16_t* p1 = base + size_t(r1) * (3 * d);
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const int c = kt * 16 + 2 * t;
        qf[kt][0] = r0 < T ? *reinterpret_cast<const uint32_t*>(p0 + c) : 0u;
        qf[kt][1] = r1 < T ? *reinterpret_cast<const uint32_t*>(p1 + c) : 0u;
        qf[kt][2] = r0 < T ? *reinterpret_cast<const uint32_t*>(p0 + c + 8) : 0u;
        qf[kt][3] = r1 < T ? *reinterpret_cast<const uint32_t*>(p1 + c + 8) : 0u;
      }
    }
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int c = 0; c < n_chunks; ++c) {
      const int key0 = c * kChunk;
      const int n_nt = min(8, (TP - key0) >> 3);  // 8-key n-tiles present in this chunk (even number)
      float s[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;

      // ---- S = Q K^T ----
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < n_nt) {
          const int krow = key0 + nt * 8 + (lane & 7);
#pragma unroll
          for (int kp = 0; kp < 2; ++kp) {  // two pairs of 16-wide k-tiles
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(ks + swz(krow, kp * 4 + (lane >> 3)), b0, b1, b2, b3);
            mma_16816<kF16>(s[nt], qf[kp * 2], b0, b1);
            mma_16816<kF16>(s[nt], qf[kp * 2 + 1], b2, b3);
          }
        }
      }
      // ---- scale, mask the padded keys, online softmax ----
      float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = key0 + nt * 8 + 2 * t;
        const bool v0 = key < T, v1 = key + 1 < T;
        s[nt][0] = v0 ? s[nt][0] * scale_log2e : -INFINITY;
        s[nt][1] = v1 ? s[nt][1] * scale_log2e : -INFINITY;
        s[nt][2] = v0 ? s[nt][2] * scale_log2e : -INFINITY;
        s[nt][3] = v1 ? s[nt][3] * scale_log2e : -INFINITY;
        cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
        cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);  // finite: key 0 is always valid in chunk 0
      const float al0 = exp2f(m0 - mn0), al1 = exp2f(m1 - mn1);
      m0 = mn0; m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mn0);
        s[nt][1] = exp2f(s[nt][1] - mn0);
        s[nt][2] = exp2f(s[nt][2] - mn1);
        s[nt][3] = exp2f(s[nt][3] - mn1);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * al0 + rs0;
      l1 = l1 * al1 + rs1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i][0] *= al0; o[i][1] *= al0; o[i][2] *= al1; o[i][3] *= al1;
      }
      // ---- O += P V ----
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        if (2 * kt < n_nt) {
          uint32_t pf[4];
          pf[0] = Act<kF16>::pack(s[2 * kt][0], s[2 * kt][1]);
          pf[1] = Act<kF16>::pack(s[2 * kt][2], s[2 * kt][3]);
          pf[2] = Act<kF16>::pack(s[2 * kt + 1][0], s[2 * kt + 1][1]);
          pf[3] = Act<kF16>::pack(s[2 * kt + 1][2], s[2 * kt + 1][3]);
          const int vrow = key0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
          for (int np = 0; np < 4; ++np) {  // pairs of 8-wide head-dim n-tiles
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(vs + swz(vrow, np * 2 + (lane >> 4)), b0, b1, b2, b3);
            mma_16816<kF16>(o[np * 2], pf, b0, b1);
            mma_16816<kF16>(o[np * 2 + 1], pf, b2, b3);
          }
        }
      }
    }
    // ---- normalise and store ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    if (lse != nullptr && t == 0) {
      // training: log2-domain log-sum-exp of the scaled scores, so the backward recomputes P = exp2(s*c - lse)
      if (r0 < T) lse[size_t(bh) * T + r0] = m0 + log2f(l0);
      if (r1 < T) lse[size_t(bh) * T + r1] = m1 + log2f(l1);
    }
    uint16_t* ob = out + size_t(b) * T * d + h * kHd;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + size_t(r0) * d + c) = Act<kF16>::pack(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + size_t(r1) * d + c) = Act<kF16>::pack(o[nt][2] * i1, o[nt][3] * i1);
    }
  }
}

int launch_attention(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16,
                     cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd) return -1;
  const int TP = (T + 15) / 16 * 16;
  const size_t smem = size_t(TP) * 128 * 2;
  if (smem > 227 * 1024) return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(head_dim));
  const uint16_t* q = static_cast<const uint16_t*>(qkv);
  uint16_t* o = static_cast<uint16_t*>(out);
  if (f16)
    attention_kernel<true><<<B * H, kWarps * 32, smem, stream>>>(q, o, lse, T, TP, H, scale_log2e);
  else
    attention_kernel<false><<<B * H, kWarps * 32, smem, stream>>>(q, o, lse, T, TP, H, scale_log2e);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
