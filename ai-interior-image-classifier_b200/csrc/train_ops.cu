// Backward-pass kernels of the LoRA fine-tuning step (everything that is not a GEMM or attention):
//   LayerNorm backward, QuickGELU / GELU backward, the skinny LoRA gradient reductions dB = P^T dY and
//   dA = s * act(X)^T dP, fp32 -> 16-bit casts.
//
// Reference semantics: torch.autograd through OpenAI CLIP's ResidualAttentionBlock with /root/reference/main.py's
// LoRALinear on mlp.c_fc / mlp.c_proj (main.py:30-31, 42-43) and the training loop of train_lora.py:231-252
// (only parameters whose name contains 'lora' receive gradients; the frozen weights only propagate dX).
// Activation gradients travel in the 16-bit operand format (they are GEMM operands), every reduction is fp32,
// the residual-stream gradient is fp32.  All reductions are deterministic (fixed partial order, no float atomics).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "act_types.cuh"
#include "kernels.h"

namespace iic {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float act_fwd(float u, int act) {
  if (act == 1) {   // QuickGELU, the forward epilogue's one-MUFU form: u sigmoid(1.702 u) = 0.5 u + 0.5 u tanh(0.851 u).  The
    // exp + divide form made the dA reduction over gelu(u) issue-bound (ncu: 60 % issue active, tensor pipe 5 %).
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * u));
    const float hu = 0.5f * u;
    return fmaf(hu, t, hu);
  }
  if (act == 2) return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));     // GELU (erf)
  return u;
}
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == 1) {
    const float s = 1.0f / (1.0f + __expf(-1.702f * u));
    return s * (1.0f + 1.702f * u * (1.0f - s));
  }
  if (act == 2) {
    const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
    return cdf + u * 0.3989422804014327f * __expf(-0.5f * u * u);
  }
  return 1.0f;
}

// dx[row,:] += LN'(dy[row,:]; x[row,:], gamma);  optional 16-bit copy of the updated dx (next GEMM operand).
// One warp per row, grid-stride.  kVec = D / 128.
template <int kVec, bool kF16>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const uint16_t* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                     float* __restrict__ dx, uint16_t* __restrict__ dx16, int rows, float eps) {
  constexpr int D = kVec * 128;
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += n_warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + size_t(row) * D);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + size_t(row) * D);
    float4 xv[kVec], gv[kVec], acc[kVec];
    uint2 draw[kVec];
    float4* o = reinterpret_cast<float4*>(dx + size_t(row) * D);
    // all three input streams of the row are requested before the first reduction: one round of memory latency per row
#pragma unroll
    for (int j = 0; j < kVec; ++j) xv[j] = xr[lane + 32 * j];
#pragma unroll
    for (int j = 0; j < kVec; ++j) draw[j] = dr[lane + 32 * j];
#pragma unroll
    for (int j = 0; j < kVec; ++j) acc[j] = o[lane + 32 * j];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) s += (xv[j].x + xv[j].y) + (xv[j].z + xv[j].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      xv[j].x -= mean; xv[j].y -= mean; xv[j].z -= mean; xv[j].w -= mean;
      q += (xv[j].x * xv[j].x + xv[j].y * xv[j].y) + (xv[j].z * xv[j].z + xv[j].w * xv[j].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      const uint2 raw = draw[j];
      const float2 d01 = Act<kF16>::unpack(raw.x), d23 = Act<kF16>::unpack(raw.y);
      const float4 g = __ldg(g4 + lane + 32 * j);
      xv[j].x *= rstd; xv[j].y *= rstd; xv[j].z *= rstd; xv[j].w *= rstd;   // xhat
      gv[j] = make_float4(d01.x * g.x, d01.y * g.y, d23.x * g.z, d23.y * g.w);
      c1 += (gv[j].x + gv[j].y) + (gv[j].z + gv[j].w);
      c2 += (gv[j].x * xv[j].x + gv[j].y * xv[j].y) + (gv[j].z * xv[j].z + gv[j].w * xv[j].w);
    }
    c1 = warp_sum(c1) * (1.0f / D);
    c2 = warp_sum(c2) * (1.0f / D);
    uint2* o16 = dx16 ? reinterpret_cast<uint2*>(dx16 + size_t(row) * D) : nullptr;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      float4 r = acc[j];
      r.x += rstd * (gv[j].x - c1 - xv[j].x * c2);
      r.y += rstd * (gv[j].y - c1 - xv[j].y * c2);
      r.z += rstd * (gv[j].z - c1 - xv[j].z * c2);
      r.w += rstd * (gv[j].w - c1 - xv[j].w * c2);
      o[lane + 32 * j] = r;
      if (o16) {
        uint2 pk;
        pk.x = Act<kF16>::pack(r.x, r.y);
        pk.y = Act<kF16>::pack(r.z, r.w);
        o16[lane + 32 * j] = pk;
      }
    }
  }
}

template <bool kF16>
__global__ void __launch_bounds__(256)
cast16_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  uint2 pk;
  pk.x = Act<kF16>::pack(v.x, v.y);
  pk.y = Act<kF16>::pack(v.z, v.w);
  reinterpret_cast<uint2*>(out)[i] = pk;
}

// dh <- dh * act'(u), elementwise, 8 elements per thread
template <bool kF16>
__global__ void __launch_bounds__(256)
act_bwd_kernel(uint16_t* __restrict__ dh, const uint16_t* __restrict__ u, long long n8, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 a = reinterpret_cast<uint4*>(dh)[i];
  const uint4 b = reinterpret_cast<const uint4*>(u)[i];
  uint32_t* pa = reinterpret_cast<uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 g = Act<kF16>::unpack(pa[e]), uu = Act<kF16>::unpack(pb[e]);
    pa[e] = Act<kF16>::pack(g.x * act_grad(uu.x, act), g.y * act_grad(uu.y, act));
  }
  reinterpret_cast<uint4*>(dh)[i] = a;
}

// part[split][c][n] = sum over the split's rows of P[m, c] * f(Y[m, n]),  c < 4, f = act_fwd (act = 0: identity).
// P 16-bit [M, p_ld] (already offset to the 4-column group), Y 16-bit [M, N] (N % 8 == 0).
// Thread = 8 adjacent columns of Y (one 16-byte load per row, 8 rows in flight); CTA = 128 threads = 1024 columns x
// one split of `rows_per_split` rows.  Many short splits keep >100k threads busy: the kernel is a pure stream over Y.
template <bool kF16>
__global__ void __launch_bounds__(128)
lora_outer_kernel(const uint16_t* __restrict__ P, int p_ld, const uint16_t* __restrict__ Y, int N, int M, int rows_per_split,
                  int act, float* __restrict__ part) {
  const int n = blockIdx.x * 1024 + threadIdx.x * 8;
  const int split = blockIdx.y;
  const int m0 = split * rows_per_split;
  const int m1 = min(M, m0 + rows_per_split);
  if (n >= N) return;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const uint2 praw = *reinterpret_cast<const uint2*>(P + size_t(m) * p_ld);   // warp-uniform: broadcast
    const float2 p01 = Act<kF16>::unpack(praw.x), p23 = Act<kF16>::unpack(praw.y);
    const uint4 yraw = *reinterpret_cast<const uint4*>(Y + size_t(m) * N + n);
    const uint32_t* yw = reinterpret_cast<const uint32_t*>(&yraw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 y = Act<kF16>::unpack(yw[e]);
      if (act != 0) { y.x = act_fwd(y.x, act); y.y = act_fwd(y.y, act); }
      acc[2 * e][0] = fmaf(p01.x, y.x, acc[2 * e][0]); acc[2 * e][1] = fmaf(p01.y, y.x, acc[2 * e][1]);
      acc[2 * e][2] = fmaf(p23.x, y.x, acc[2 * e][2]); acc[2 * e][3] = fmaf(p23.y, y.x, acc[2 * e][3]);
      acc[2 * e + 1][0] = fmaf(p01.x, y.y, acc[2 * e + 1][0]); acc[2 * e + 1][1] = fmaf(p01.y, y.y, acc[2 * e + 1][1]);
      acc[2 * e + 1][2] = fmaf(p23.x, y.y, acc[2 * e + 1][2]); acc[2 * e + 1][3] = fmaf(p23.y, y.y, acc[2 * e + 1][3]);
    }
  }
  float* o = part + (size_t(split) * 4) * N + n;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<float4*>(o + size_t(c) * N) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    *reinterpret_cast<float4*>(o + size_t(c) * N + 4) = make_float4(acc[4][c], acc[5][c], acc[6][c], acc[7][c]);
  }
}

// Ranks 5..16 in ONE pass over Y (the 4-column kernel above would stream Y once per 4 ranks): thread = 2 adjacent columns of Y
// and all 16 columns of P (zero padded beyond the rank), 32 accumulators; CTA = 128 threads = 256 columns x one split.
// part[split][c][n], c < 16.
template <bool kF16>
__global__ void __launch_bounds__(128)
lora_outer16_kernel(const uint16_t* __restrict__ P, int p_ld, const uint16_t* __restrict__ Y, int N, int M, int rows_per_split,
                    int act, float* __restrict__ part) {
  const int n = blockIdx.x * 256 + threadIdx.x * 2;
  const int split = blockIdx.y;
  const int m0 = split * rows_per_split;
  const int m1 = min(M, m0 + rows_per_split);
  if (n >= N) return;
  float a0[16], a1[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a0[c] = a1[c] = 0.f;
#pragma unroll 4
  for (int m = m0; m < m1; ++m) {
    const uint4 pa = *reinterpret_cast<const uint4*>(P + size_t(m) * p_ld);        // warp-uniform: broadcast
    const uint4 pb = *reinterpret_cast<const uint4*>(P + size_t(m) * p_ld + 8);
    float2 y = Act<kF16>::unpack(*reinterpret_cast<const uint32_t*>(Y + size_t(m) * N + n));
    if (act != 0) { y.x = act_fwd(y.x, act); y.y = act_fwd(y.y, act); }
    const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 pv = Act<kF16>::unpack(pw[e]);
      a0[2 * e] = fmaf(pv.x, y.x, a0[2 * e]);         a1[2 * e] = fmaf(pv.x, y.y, a1[2 * e]);
      a0[2 * e + 1] = fmaf(pv.y, y.x, a0[2 * e + 1]); a1[2 * e + 1] = fmaf(pv.y, y.y, a1[2 * e + 1]);
    }
  }
  float* o = part + (size_t(split) * 16) * N + n;
#pragma unroll
  for (int c = 0; c < 16; ++c) *reinterpret_cast<float2*>(o + size_t(c) * N) = make_float2(a0[c], a1[c]);
}

// ---- LoRA gradient reductions on the (legacy) tensor cores -----------------------------------------------------------------
// part[split][c][n] = sum over the split's rows m of P[m, c] * f(Y[m, n]),  c < 16 (P zero padded beyond the rank).
// A standalone kernel, so warp-level mma.sync does not fight a tcgen05 mainloop for the tensor pipe: the reduction is a
// [16 x rows] . [rows x 64] product per warp - 2 FLOP per byte of Y per rank - which the CUDA-core kernels above could not
// keep HBM-bound beyond rank 4.  CTA = 4 warps x 128 rows of one 64-column tile; per 16-row step a warp cp.asyncs its
// [16 x 64] slice of Y (XOR-swizzled 128-byte rows) and [16 x 16] slice of P into a private 2-stage buffer, loads
// A = P^T and B = Y with ldmatrix.trans and issues 8 m16n8k16 MMAs; the four warps' accumulators are added through smem.
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
template <bool kF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

constexpr int kOuterStages = 4;

template <bool kF16>
__global__ void __launch_bounds__(128)
lora_outer_mma_kernel(const uint16_t* __restrict__ P, int p_ld, const uint16_t* __restrict__ Y, int N, int M, int act,
                      int rank, int rows_per_warp, float* __restrict__ part) {
  // per warp: kOuterStages x (Y slice 16 x 128 B = 2 KB, P slice 16 x 32 B = 512 B) - three slices in flight per warp keep
  // the kernel on the HBM roofline; the CTA-wide reduction buffer red[4][16][65] aliases the stage buffers after the loop
  __shared__ __align__(128) uint8_t sbuf[4][kOuterStages][2560];
  static_assert(sizeof(float) * 4 * 16 * 65 <= 4 * kOuterStages * 2560, "reduction buffer must fit in the stage buffers");
  float (*red)[16][65] = reinterpret_cast<float (*)[16][65]>(&sbuf[0][0][0]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 64;
  const int m_begin = (blockIdx.y * 4 + warp) * rows_per_warp;
  const int m_end = min(M, m_begin + rows_per_warp);
  const uint32_t sb = static_cast<uint32_t>(__cvta_generic_to_shared(&sbuf[warp][0][0]));
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  auto issue = [&](int step, int stage) {
    const int m0 = m_begin + step * 16;
    const uint32_t ys = sb + uint32_t(stage) * 2560u, ps = ys + 2048u;
    // Y slice: 16 rows x 8 chunks of 16 B; lane -> (row = lane / 2 (+8 on the second round), 4 chunks)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * 32 + lane;            // 0..127
      const int row = idx >> 3, ch = idx & 7;
      const int m = m0 + row;
      const bool ok = m < m_end && n0 + ch * 8 < N;
      cp16(ys + uint32_t(row * 128 + ((ch ^ (row & 7)) << 4)), Y + size_t(ok ? m : 0) * N + n0 + ch * 8, ok);
    }
    {   // P slice: 16 rows x 2 chunks
      const int row = lane >> 1, ch = lane & 1;
      const int m = m0 + row;
      const bool ok = m < m_end;
      cp16(ps + uint32_t(row * 32 + ch * 16), P + size_t(ok ? m : 0) * p_ld + ch * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int steps = m_begin < m_end ? (m_end - m_begin + 15) / 16 : 0;
  // one commit group per step slot (empty beyond the last step), so "all but the newest kOuterStages - 1 groups" is always
  // the right wait
#pragma unroll
  for (int st = 0; st < kOuterStages - 1; ++st) {
    if (st < steps) issue(st, st);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int st = 0; st < steps; ++st) {
    const int stage = st % kOuterStages;
    if (st + kOuterStages - 1 < steps) issue(st + kOuterStages - 1, (st + kOuterStages - 1) % kOuterStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(kOuterStages - 1) : "memory");
    __syncwarp();
    const uint32_t ys = sb + uint32_t(stage) * 2560u, ps = ys + 2048u;
    // A = P^T (16 ranks x 16 rows): four transposed 8x8 blocks of the [row][rank] slice
    uint32_t a[4];
    {
      const int mat = lane >> 3, r8 = lane & 7;
      const int row = (mat >> 1) * 8 + r8;        // matrices 0,1: rows 0-7; 2,3: rows 8-15
      const int chunk = mat & 1;                  // matrices 0,2: ranks 0-7; 1,3: ranks 8-15
      ldsm4t(ps + uint32_t(row * 32 + chunk * 16), a);
    }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {              // pairs of 8-column tiles
      uint32_t b[4];
      const int mat = lane >> 3, r8 = lane & 7;
      const int row = (mat & 1) * 8 + r8;         // matrices 0,2: rows 0-7; 1,3: rows 8-15
      const int ch = jp * 2 + (mat >> 1);         // matrices 0,1: tile 2 jp; 2,3: tile 2 jp + 1
      ldsm4t(ys + uint32_t(row * 128 + ((ch ^ (row & 7)) << 4)), b);
      if (act != 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 y = Act<kF16>::unpack(b[e]);
          b[e] = Act<kF16>::pack(act_fwd(y.x, act), act_fwd(y.y, act));
        }
      }
      mma16816<kF16>(acc[2 * jp], a, b[0], b[1]);
      mma16816<kF16>(acc[2 * jp + 1], a, b[2], b[3]);
    }
    __syncwarp();   // every lane is done with this stage before it is refilled two steps later
  }
  // accumulator layout of m16n8: lane (g, t) holds ranks g, g + 8 and columns 2t, 2t + 1 of tile j
  const int g = lane >> 2, t = lane & 3;
  __syncthreads();   // every warp is done with its stage buffers before they become the reduction buffer
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[warp][g][8 * j + 2 * t] = acc[j][0];
    red[warp][g][8 * j + 2 * t + 1] = acc[j][1];
    red[warp][g + 8][8 * j + 2 * t] = acc[j][2];
    red[warp][g + 8][8 * j + 2 * t + 1] = acc[j][3];
  }
  __syncthreads();
  const int used = (rank + 3) & ~3;   // the reduction reads rows c < rank only
  for (int i = threadIdx.x; i < used * 64; i += 128) {
    const int c = i >> 6, n = i & 63;
    if (n0 + n < N)
      part[(size_t(blockIdx.y) * 16 + c) * N + n0 + n] = (red[0][c][n] + red[1][c][n]) + (red[2][c][n] + red[3][c][n]);
  }
}

// ---- both LoRA gradients that read the same activation gradient, in one pass ---------------------------------------------
// For Y = dOut [M, N] of a LoRALinear (main.py:42-43):  dB = P^T . Y  (P = s x A, [M, 16])  and  dP = Y . B^T  (B [16, N]).
// CTA = R rows x C columns, 4 warps; warp w owns C/4 columns (kTW 64-column tiles) for all R rows: its dB accumulators
// [16 x C/4] live in registers for the whole CTA; the [16 rows x 16] dP blocks of the four warps of a 16-row step are summed
// through shared memory in a fixed order (one barrier per step: deterministic, unlike atomics).
// Partials: part_db[row block][16][N], part_dp[column group][M][16].
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
constexpr int kBwdStages = 3;

template <bool kF16, int kTW>
__global__ void __launch_bounds__(128)
lora_bwd_kernel(const uint16_t* __restrict__ P, int p_ld, const uint16_t* __restrict__ Y, int N, int M,
                const uint16_t* __restrict__ Bm, int rows_per_cta, int rank, float* __restrict__ part_db,
                float* __restrict__ part_dp) {
  constexpr int kWarpCols = kTW * 64, kCtaCols = 4 * kWarpCols;
  constexpr int kYBytes = 16 * kTW * 128, kStageBytes = kYBytes + 512;
  extern __shared__ __align__(128) uint8_t dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* bm_gen = dsm + 4 * kBwdStages * kStageBytes;              // [16 ranks][kCtaCols] 16-bit, 16-byte chunks XOR-swizzled
  float* dp_s = reinterpret_cast<float*>(bm_gen + 16 * kCtaCols * 2);   // [rows_per_cta][16]
  float* dpw = dp_s + rows_per_cta * 16;                                // [2 step parities][4 warps][16 x 16]
  const uint32_t sb = static_cast<uint32_t>(__cvta_generic_to_shared(dsm)) + uint32_t(warp * kBwdStages * kStageBytes);
  const uint32_t bm_s = static_cast<uint32_t>(__cvta_generic_to_shared(bm_gen));
  const int c_cta = blockIdx.x * kCtaCols, c_warp = c_cta + warp * kWarpCols;
  const int m_begin = blockIdx.y * rows_per_cta;
  const int m_end = min(M, m_begin + rows_per_cta);
  const int steps = (m_end - m_begin + 15) / 16;   // the same for the four warps (they meet at a barrier every step)

  auto issue = [&](int step, int stage) {
    const int m0 = m_begin + step * 16;
    const uint32_t ys = sb + uint32_t(stage * kStageBytes), ps = ys + kYBytes;
#pragma unroll
    for (int q = 0; q < kTW * 4; ++q) {
      const int idx = q * 32 + lane;                    // 16 rows x (kTW * 8) chunks
      const int row = idx / (kTW * 8), chf = idx % (kTW * 8);
      const int t = chf >> 3, ch = chf & 7;
      const int m = m0 + row;
      const bool ok = m < m_end;
      cp16(ys + uint32_t(t * 2048 + row * 128 + ((ch ^ (row & 7)) << 4)), Y + size_t(ok ? m : 0) * N + c_warp + chf * 8, ok);
    }
    {
      const int row = lane >> 1, ch = lane & 1;
      const int m = m0 + row;
      const bool ok = m < m_end;
      cp16(ps + uint32_t(row * 32 + ch * 16), P + size_t(ok ? m : 0) * p_ld + ch * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int st = 0; st < kBwdStages - 1; ++st) {
    if (st < steps) issue(st, st);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // B slice of the CTA's columns and the zeroed dP buffer (plain stores; visible after the barrier below)
  for (int i = threadIdx.x; i < 16 * (kCtaCols / 8); i += 128) {
    const int r = i / (kCtaCols / 8), ch = i % (kCtaCols / 8);
    const uint4 v = *reinterpret_cast<const uint4*>(Bm + size_t(r) * N + c_cta + ch * 8);
    *reinterpret_cast<uint4*>(bm_gen + r * (kCtaCols * 2) + ((ch ^ (r & 7)) << 4)) = v;
  }
  __syncthreads();

  float acc[kTW * 8][4];
#pragma unroll
  for (int j = 0; j < kTW * 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  const int mat = lane >> 3, r8 = lane & 7;
  const int g = lane >> 2, t4 = lane & 3;
  for (int st = 0; st < steps; ++st) {
    const int stage = st % kBwdStages;
    if (st + kBwdStages - 1 < steps) issue(st + kBwdStages - 1, (st + kBwdStages - 1) % kBwdStages);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(kBwdStages - 1) : "memory");
    __syncwarp();
    const uint32_t ys = sb + uint32_t(stage * kStageBytes), ps = ys + kYBytes;
    uint32_t a[4];   // P^T: 16 ranks x 16 rows
    ldsm4t(ps + uint32_t(((mat >> 1) * 8 + r8) * 32 + (mat & 1) * 16), a);
    float dp[2][4], dq[2][4];   // two accumulation chains (even / odd k-steps): half the dependent-MMA latency per step
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < kTW; ++t) {
      const uint32_t yt = ys + uint32_t(t * 2048);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {   // dB: B operand = Y (k = rows), pairs of 8-column tiles
        uint32_t b[4];
        const int row = (mat & 1) * 8 + r8, ch = jp * 2 + (mat >> 1);
        ldsm4t(yt + uint32_t(row * 128 + ((ch ^ (row & 7)) << 4)), b);
        mma16816<kF16>(acc[t * 8 + 2 * jp], a, b[0], b[1]);
        mma16816<kF16>(acc[t * 8 + 2 * jp + 1], a, b[2], b[3]);
      }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {   // dP: A operand = Y (k = 16 columns per step), B operand = the B slice
        uint32_t ya[4], bb[4];
        {
          const int row = (mat & 1) * 8 + r8, ch = ks * 2 + (mat >> 1);
          ldsm4(yt + uint32_t(row * 128 + ((ch ^ (row & 7)) << 4)), ya);
        }
        {
          const int rk = (mat >> 1) * 8 + r8, ch = (warp * kWarpCols + t * 64 + ks * 16) / 8 + (mat & 1);
          ldsm4(bm_s + uint32_t(rk * (kCtaCols * 2) + ((ch ^ (rk & 7)) << 4)), bb);
        }
        if (ks & 1) {
          mma16816<kF16>(dq[0], ya, bb[0], bb[1]);
          if (rank > 8) mma16816<kF16>(dq[1], ya, bb[2], bb[3]);
        } else {
          mma16816<kF16>(dp[0], ya, bb[0], bb[1]);
          if (rank > 8) mma16816<kF16>(dp[1], ya, bb[2], bb[3]);   // ranks 8..15 are zero rows of B otherwise
        }
      }
    }
    {
      float* d0 = dpw + ((st & 1) * 4 + warp) * 256 + g * 16 + 2 * t4;
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        *reinterpret_cast<float2*>(d0 + n * 8) = make_float2(dp[n][0] + dq[n][0], dp[n][1] + dq[n][1]);
        *reinterpret_cast<float2*>(d0 + 128 + n * 8) = make_float2(dp[n][2] + dq[n][2], dp[n][3] + dq[n][3]);
      }
    }
    __syncthreads();   // also: every lane of every warp is done with this step's stage buffers
    {
      const float* w0 = dpw + (st & 1) * 4 * 256;   // the other parity is being written by the next step: no second barrier
#pragma unroll
      for (int e = threadIdx.x; e < 256; e += 128)
        dp_s[st * 256 + e] = ((w0[e] + w0[256 + e]) + w0[512 + e]) + w0[768 + e];
    }
  }
  // dB partial of this row block: rank g / g + 8, columns 2 t4, 2 t4 + 1 of each 8-column tile
#pragma unroll
  for (int j = 0; j < kTW * 8; ++j) {
    const int col = c_warp + j * 8 + 2 * t4;
    if (g < rank) *reinterpret_cast<float2*>(part_db + (size_t(blockIdx.y) * 16 + g) * N + col) = make_float2(acc[j][0], acc[j][1]);
    if (g + 8 < rank)
      *reinterpret_cast<float2*>(part_db + (size_t(blockIdx.y) * 16 + g + 8) * N + col) = make_float2(acc[j][2], acc[j][3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (m_end - m_begin) * 4; i += 128)
    reinterpret_cast<float4*>(part_dp + (size_t(blockIdx.x) * M + m_begin) * 16)[i] = reinterpret_cast<const float4*>(dp_s)[i];
}

// dP 16-bit [M, 16] = sum over column groups of part_dp[group][M][16]
template <bool kF16>
__global__ void __launch_bounds__(256)
lora_dp_reduce_kernel(const float* __restrict__ part, int groups, int M, uint16_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (row, 4-rank group)
  if (i >= M * 4) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int gq = 0; gq < groups; ++gq) {
    const float4 v = reinterpret_cast<const float4*>(part)[size_t(gq) * M * 4 + i];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  reinterpret_cast<uint2*>(out)[i] = make_uint2(Act<kF16>::pack(s.x, s.y), Act<kF16>::pack(s.z, s.w));
}

// out = scale * sum_split part[split][c][n], c < rc (4 or 16 columns per pass; only the c0 + c < rank rows are touched)
// written either as [c0+c][n] (dB: (r, out)) or [n][c0+c] (dA: (in, r)).  CTA = 32 consecutive n x 8 split lanes: lane j
// adds splits j, j + 8, ... (independent loads, four in flight), the eight partial sums are added in lane order - a fixed
// summation tree, so the result is deterministic; ~2 memory round trips instead of one per split.
__global__ void __launch_bounds__(256)
lora_outer_reduce_kernel(const float* __restrict__ part, int splits, int N, int rank, int c0, int rc, float scale, int transpose,
                         float* __restrict__ out) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int n_chunks = (N + 31) / 32;
  const int c = blockIdx.x / n_chunks, n = (blockIdx.x - c * n_chunks) * 32 + lane;
  float s = 0.f;
  if (n < N) {
    const float* p = part + size_t(c) * N + n;
    const size_t stride = size_t(rc) * N;
    for (int sp = j; sp < splits; sp += 32) {
      const float v0 = p[size_t(sp) * stride];
      const float v1 = sp + 8 < splits ? p[size_t(sp + 8) * stride] : 0.f;
      const float v2 = sp + 16 < splits ? p[size_t(sp + 16) * stride] : 0.f;
      const float v3 = sp + 24 < splits ? p[size_t(sp + 24) * stride] : 0.f;
      s = (((s + v0) + v1) + v2) + v3;
    }
  }
  red[j][lane] = s;
  __syncthreads();
  if (j != 0 || n >= N) return;
  float t = red[0][lane];
#pragma unroll
  for (int k = 1; k < 8; ++k) t += red[k][lane];
  t *= scale;
  if (transpose) out[size_t(n) * rank + c0 + c] = t;
  else out[size_t(c0 + c) * N + n] = t;
}
static void launch_outer_reduce(const float* part, int splits, int N, int rank, int c0, int rc, float scale, int transpose, float* out,
                                cudaStream_t stream) {
  const int used = rank - c0 < rc ? rank - c0 : rc;
  if (used <= 0) return;
  lora_outer_reduce_kernel<<<used * ((N + 31) / 32), 256, 0, stream>>>(part, splits, N, rank, c0, rc, scale, transpose, out);
}

}  // namespace

int launch_layernorm_bwd(const void* dy, const float* x, const float* gamma, float* dx, void* dx16, int rows, int D,
                         float eps, int f16, cudaStream_t stream) {
  if (rows <= 0) return 0;
  const int threads = 256;
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const uint16_t* d = static_cast<const uint16_t*>(dy);
  uint16_t* o = static_cast<uint16_t*>(dx16);
#define IIC_LNB(KV)                                                                                          \
  if (f16) layernorm_bwd_kernel<KV, true><<<blocks, threads, 0, stream>>>(d, x, gamma, dx, o, rows, eps);    \
  else layernorm_bwd_kernel<KV, false><<<blocks, threads, 0, stream>>>(d, x, gamma, dx, o, rows, eps);
  switch (D) {
    case 512: IIC_LNB(4) break;
    case 768: IIC_LNB(6) break;
    case 1024: IIC_LNB(8) break;
    case 1280: IIC_LNB(10) break;
    default: return -1;
  }
#undef IIC_LNB
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_cast16(const float* in, void* out, long long n, int f16, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 4 != 0) return -1;
  const long long n4 = n / 4;
  const unsigned blocks = unsigned((n4 + 255) / 256);
  if (f16) cast16_kernel<true><<<blocks, 256, 0, stream>>>(in, static_cast<uint16_t*>(out), n4);
  else cast16_kernel<false><<<blocks, 256, 0, stream>>>(in, static_cast<uint16_t*>(out), n4);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_act_bwd(void* dh, const void* u, long long n, int act, int f16, cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n % 8 != 0) return -1;
  const long long n8 = n / 8;
  const unsigned blocks = unsigned((n8 + 255) / 256);
  if (f16) act_bwd_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<uint16_t*>(dh), static_cast<const uint16_t*>(u), n8, act);
  else act_bwd_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<uint16_t*>(dh), static_cast<const uint16_t*>(u), n8, act);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// SMs of the current device (grid shaping only; 148 when there is no device to ask)
static int sm_count() {
  static int cached[16] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0 ? n : 148;
  }
  return cached[dev];
}
// rows per CTA (a multiple of `step`, <= max_rows) that minimise  waves x (rows + fixed cost of a CTA, ~48 rows' worth:
// pipeline fill, B slice, partial write-out)  for `col_groups` CTAs per row block - these kernels stream Y once, so a nearly
// empty last wave (300 CTAs on 296 slots: what 512-row CTAs gave at M = 25216) costs a whole CTA time.
// smem_of(rows) = dynamic + static shared memory of one CTA.  Ties go to the larger CTA (fewer partials).
template <class SmemOf>
static int rows_for_whole_waves(int M, int col_groups, int step, int min_rows, int max_rows, SmemOf smem_of) {
  const int sms = sm_count();
  long long best_cost = -1;
  int best = max_rows;
  for (int rows = min_rows; rows <= max_rows; rows += step) {
    long long per_sm = (228 * 1024) / (long long)(smem_of(rows) + 1024);
    if (per_sm < 1) continue;
    if (per_sm > 16) per_sm = 16;
    const long long ctas = (long long)col_groups * ((M + rows - 1) / rows), slots = per_sm * sms;
    const long long cost = ((ctas + slots - 1) / slots) * (rows + 48);
    if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = rows; }
  }
  return best;
}
static size_t lora_bwd_smem(int tw, int rows) {
  return size_t(4 * kBwdStages * (16 * tw * 128 + 512)) + size_t(16) * (256 * tw) * 2 + size_t(rows) * 64 + 2 * 4 * 256 * 4;
}
// fused dB + dP pass (lora_bwd_kernel): 512-column CTAs when N % 512 == 0 and there are enough of them, else 256 columns;
// the row count fills whole waves (M = 25216: 528 x 512 -> 288 CTAs on 296 slots; 128 x 256 -> 591 on 592)
static void lora_bwd_geometry(int N, int M, int* tw, int* rows, int* cta_cols) {
  const bool big = N % 512 == 0 && (long long)(N / 512) * ((M + 511) / 512) >= 2 * 148;
  *tw = big ? 2 : 1;
  *cta_cols = big ? 512 : 256;
  const int t = *tw;
  *rows = rows_for_whole_waves(M, N / *cta_cols, 16, 64, big ? 528 : 512, [t](int r) { return lora_bwd_smem(t, r); });
}
// lora_outer_mma_kernel: rows per warp (x 4 warps per CTA of 64 columns; 40 KB of static shared memory)
static int lora_outer_rows_per_warp(int N, int M) {
  return rows_for_whole_waves(M, (N + 63) / 64, 64, 256, 2048, [](int) { return size_t(4 * kOuterStages * 2560); }) / 4;
}

int launch_lora_bwd(const void* P, int p_ld, const void* Y, int N, int M, const void* Bm, int rank, float scale, float* out_db,
                    void* out_dp16, float* scratch, int f16, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (rank < 1 || rank > 16 || p_ld != 16 || N % 256 != 0 || (reinterpret_cast<uintptr_t>(P) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(Bm) & 15) != 0)
    return -3;   // not this kernel's envelope: the caller runs the separate dP GEMM + dB reduction
  int tw, rows, cta_cols;
  lora_bwd_geometry(N, M, &tw, &rows, &cta_cols);
  const int row_blocks = (M + rows - 1) / rows, col_groups = N / cta_cols;
  float* part_db = scratch;
  float* part_dp = scratch + size_t(row_blocks) * 16 * N;
  const size_t smem = lora_bwd_smem(tw, rows);
  const uint16_t* p = static_cast<const uint16_t*>(P);
  const uint16_t* y = static_cast<const uint16_t*>(Y);
  const uint16_t* bm = static_cast<const uint16_t*>(Bm);
  const dim3 grid = dim3(static_cast<unsigned>(col_groups), static_cast<unsigned>(row_blocks), 1u);
#define IIC_LB(F16, TW)                                                                                                   \
  {                                                                                                                       \
    auto kern = lora_bwd_kernel<F16, TW>;                                                                                 \
    static PerDeviceOnce attr;                                                                                            \
    if (attr.need()) {                                                                                                    \
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024) != cudaSuccess) return -2;  \
      attr.mark();                                                                                                        \
    }                                                                                                                     \
    kern<<<grid, 128, smem, stream>>>(p, p_ld, y, N, M, bm, rows, rank, part_db, part_dp);                                    \
  }
  if (f16) { if (tw == 2) IIC_LB(true, 2) else IIC_LB(true, 1) }
  else { if (tw == 2) IIC_LB(false, 2) else IIC_LB(false, 1) }
#undef IIC_LB
  launch_outer_reduce(part_db, row_blocks, N, rank, 0, 16, scale, 0, out_db, stream);
  if (f16) lora_dp_reduce_kernel<true><<<(M * 4 + 255) / 256, 256, 0, stream>>>(part_dp, col_groups, M, static_cast<uint16_t*>(out_dp16));
  else lora_dp_reduce_kernel<false><<<(M * 4 + 255) / 256, 256, 0, stream>>>(part_dp, col_groups, M, static_cast<uint16_t*>(out_dp16));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int lora_outer_splits(int M) { return (M + 63) / 64; }   // 64 rows per split
size_t lora_outer_scratch_bytes(int N, int M) {   // the largest partial buffer of the three reduction kernels
  const int rpc = 4 * lora_outer_rows_per_warp(N, M);
  const size_t rows = size_t(lora_outer_splits(M)) * 4, rows_mma = size_t((M + rpc - 1) / rpc) * 16, rows16 = size_t((M + 255) / 256) * 16;
  size_t bytes = (rows > rows_mma ? (rows > rows16 ? rows : rows16) : (rows_mma > rows16 ? rows_mma : rows16)) * N * sizeof(float);
  if (N % 256 == 0) {   // lora_bwd_kernel: dB partials per row block + dP partials per column group
    int tw, r, c;
    lora_bwd_geometry(N, M, &tw, &r, &c);
    const size_t fused = (size_t((M + r - 1) / r) * 16 * N + size_t(N / c) * M * 16) * sizeof(float);
    if (fused > bytes) bytes = fused;
  }
  return bytes;
}

int launch_lora_outer(const void* P, int p_ld, const void* Y, int N, int M, int act, int rank, float scale, int transpose,
                      float* out, float* scratch, int f16, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (rank < 1 || (rank + 3) / 4 * 4 > p_ld || N % 8 != 0) return -1;
  const uint16_t* p = static_cast<const uint16_t*>(P);
  const uint16_t* y = static_cast<const uint16_t*>(Y);
  static const int impl = [] { const char* e = getenv("IIC_LORA_OUTER_IMPL"); return e ? atoi(e) : 0; }();   // 1: CUDA-core kernels
  if (impl == 0 && rank <= 16 && p_ld >= 16 && (reinterpret_cast<uintptr_t>(P) & 15) == 0 && (p_ld % 8) == 0) {
    // tensor-core reduction: one pass over Y for every rank; one partial per 512-row CTA
    const int rpw = lora_outer_rows_per_warp(N, M), splits = (M + 4 * rpw - 1) / (4 * rpw);
    dim3 grid(unsigned((N + 63) / 64), unsigned(splits));
    if (f16) lora_outer_mma_kernel<true><<<grid, 128, 0, stream>>>(p, p_ld, y, N, M, act, rank, rpw, scratch);
    else lora_outer_mma_kernel<false><<<grid, 128, 0, stream>>>(p, p_ld, y, N, M, act, rank, rpw, scratch);
    launch_outer_reduce(scratch, splits, N, rank, 0, 16, scale, transpose, out, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  if (rank > 4 && rank <= 16 && p_ld >= 16) {
    // one pass over Y for all ranks: 256-row splits keep the partial buffer at M/16 * N floats (same as the rank-4 path)
    const int rps = 256, splits = (M + rps - 1) / rps;
    dim3 grid(unsigned((N + 255) / 256), unsigned(splits));
    if (f16) lora_outer16_kernel<true><<<grid, 128, 0, stream>>>(p, p_ld, y, N, M, rps, act, scratch);
    else lora_outer16_kernel<false><<<grid, 128, 0, stream>>>(p, p_ld, y, N, M, rps, act, scratch);
    launch_outer_reduce(scratch, splits, N, rank, 0, 16, scale, transpose, out, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
  }
  const int splits = lora_outer_splits(M);
  const int rps = 64;
  dim3 grid(unsigned((N + 1023) / 1024), unsigned(splits));
  for (int c0 = 0; c0 < rank; c0 += 4) {   // 4 LoRA columns per pass over Y (rank 4: one pass)
    if (f16) lora_outer_kernel<true><<<grid, 128, 0, stream>>>(p + c0, p_ld, y, N, M, rps, act, scratch);
    else lora_outer_kernel<false><<<grid, 128, 0, stream>>>(p + c0, p_ld, y, N, M, rps, act, scratch);
    launch_outer_reduce(scratch, splits, N, rank, c0, 4, scale, transpose, out, stream);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}


// ---- per-step refresh of the derived LoRA operands of one slot from the fp32 parameters (after the optimizer step) -------
// A f32 [in, rank], B f32 [rank, out]  ->  a f32 [in, r4] = s*A;  bt 16-bit [out, pad] = B^T;  a16 16-bit [in, pad] = s*A;
// bt32 f32 [out, r4] = B^T;  at16 16-bit [pad, in] = (s*A)^T;  b16 16-bit [pad, out] = B.   Null destinations are skipped;
// padding columns / rows are written as zeros.  Same roundings as the host-side preparation (round to nearest).
template <bool kF16>
__global__ void lora_refresh_kernel(const float* __restrict__ A, const float* __restrict__ B, int in, int out, int rank, int r4,
                                    int pad, float s, float* __restrict__ a, typename Act<kF16>::T* __restrict__ bt,
                                    typename Act<kF16>::T* __restrict__ a16, float* __restrict__ bt32,
                                    typename Act<kF16>::T* __restrict__ at16, typename Act<kF16>::T* __restrict__ b16) {
  using T = typename Act<kF16>::T;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < in) {
    for (int c = 0; c < pad; ++c) {
      const float v = c < rank ? A[size_t(i) * rank + c] * s : 0.f;
      if (c < r4 && a != nullptr) a[size_t(i) * r4 + c] = v;
      if (a16 != nullptr) a16[size_t(i) * pad + c] = Act<kF16>::from_float(v);
      if (at16 != nullptr) at16[size_t(c) * in + i] = Act<kF16>::from_float(v);
    }
  }
  if (i < out) {
    for (int c = 0; c < pad; ++c) {
      const float v = c < rank ? B[size_t(c) * out + i] : 0.f;
      if (bt != nullptr) bt[size_t(i) * pad + c] = Act<kF16>::from_float(v);
      if (c < r4 && bt32 != nullptr) bt32[size_t(i) * r4 + c] = v;
      if (b16 != nullptr) b16[size_t(c) * out + i] = Act<kF16>::from_float(v);
    }
  }
  (void)sizeof(T);
}

template <bool kF16>
__global__ void lora_refresh_batch_kernel(const __grid_constant__ LoraRefreshBatch b) {
  using T = typename Act<kF16>::T;
  const LoraRefreshSlot& q = b.slot[blockIdx.y];
  const int in = q.in, out = q.out, rank = q.rank, r4 = q.r4, pad = b.pad;
  T* bt = static_cast<T*>(q.bt);
  T* a16 = static_cast<T*>(q.a16);
  T* at16 = static_cast<T*>(q.at16);
  T* b16 = static_cast<T*>(q.b16);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < in) {
    for (int c = 0; c < pad; ++c) {
      const float v = c < rank ? q.A[size_t(i) * rank + c] * q.scaling : 0.f;
      if (c < r4 && q.a != nullptr) q.a[size_t(i) * r4 + c] = v;
      if (a16 != nullptr) a16[size_t(i) * pad + c] = Act<kF16>::from_float(v);
      if (at16 != nullptr) at16[size_t(c) * in + i] = Act<kF16>::from_float(v);
    }
  }
  if (i < out) {
    for (int c = 0; c < pad; ++c) {
      const float v = c < rank ? q.B[size_t(c) * out + i] : 0.f;
      if (bt != nullptr) bt[size_t(i) * pad + c] = Act<kF16>::from_float(v);
      if (c < r4 && q.bt32 != nullptr) q.bt32[size_t(i) * r4 + c] = v;
      if (b16 != nullptr) b16[size_t(c) * out + i] = Act<kF16>::from_float(v);
    }
  }
}

int launch_lora_refresh_batch(const LoraRefreshBatch& batch, int f16, cudaStream_t stream) {
  if (batch.n <= 0) return 0;
  if (batch.n > 16) return -1;
  int n_max = 0;
  for (int k = 0; k < batch.n; ++k) {
    const LoraRefreshSlot& q = batch.slot[k];
    if (q.A == nullptr || q.B == nullptr || q.rank <= 0 || q.rank > batch.pad || q.in <= 0 || q.out <= 0) return -1;
    n_max = n_max > q.in ? n_max : q.in;
    n_max = n_max > q.out ? n_max : q.out;
  }
  const dim3 grid = dim3(static_cast<unsigned>((n_max + 127) / 128), static_cast<unsigned>(batch.n), 1u);
  if (f16) lora_refresh_batch_kernel<true><<<grid, 128, 0, stream>>>(batch);
  else lora_refresh_batch_kernel<false><<<grid, 128, 0, stream>>>(batch);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_lora_refresh(const float* A, const float* B, int in, int out, int rank, int r4, int pad, float scaling, float* a,
                        void* bt, void* a16, float* bt32, void* at16, void* b16, int f16, cudaStream_t stream) {
  if (A == nullptr || B == nullptr || rank <= 0 || rank > pad || in <= 0 || out <= 0) return -1;
  const int n = in > out ? in : out;
  if (f16)
    lora_refresh_kernel<true><<<(n + 127) / 128, 128, 0, stream>>>(A, B, in, out, rank, r4, pad, scaling, a, static_cast<__half*>(bt),
                                                                 static_cast<__half*>(a16), bt32, static_cast<__half*>(at16),
                                                                 static_cast<__half*>(b16));
  else
    lora_refresh_kernel<false><<<(n + 127) / 128, 128, 0, stream>>>(A, B, in, out, rank, r4, pad, scaling, a,
                                                                  static_cast<__nv_bfloat16*>(bt), static_cast<__nv_bfloat16*>(a16),
                                                                  bt32, static_cast<__nv_bfloat16*>(at16),
                                                                  static_cast<__nv_bfloat16*>(b16));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
