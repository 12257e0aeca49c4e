// Attention backward for the short ViT sequence: (dO, saved Q/K/V, O, log-sum-exp) -> dQ, dK, dV, packed like qkv.
//
// Reference semantics: torch.autograd through nn.MultiheadAttention's softmax(Q K^T / sqrt(64)) V inside CLIP's
// ResidualAttentionBlock; needed only to carry dX through the frozen attention blocks down to the LoRA-adapted MLPs
// (train_lora.py:249 `loss.backward()` restricted to parameters named '*lora*').
//
// One CTA per (image, head); Q, K, V, dO of that head live in shared memory (XOR-swizzled 128-byte rows), P is
// recomputed from the saved log-sum-exp (no T x T matrix is ever stored).  Two passes over the same tiles, both with
// mma.sync m16n8k16 and everything else in registers:
//   pass Q (query-major):  S = Q K^T, dP = dO V^T, dS = P o (dP - D) / 8,  dQ = dS K
//   pass K (key-major):    S^T = K Q^T, dP^T = V dO^T,  dV = P^T dO,  dK = dS^T Q
// with D[q] = sum_d dO[q,d] O[q,d].  Recomputing S and dP in the second pass costs 2 extra tile products but keeps
// every accumulator private to one warp: no atomics, no cross-warp reduction, deterministic.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_types.cuh"
#include "kernels.h"

namespace iic {

namespace {

constexpr int kHd = 64;
constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool kF16>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (kF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// A-operand fragments (16 rows x 64 dims) of the tile starting at `row0` of a swizzled [rows][64] smem matrix
__device__ __forceinline__ void load_a_frags(uint32_t base, int row0, int lane, uint32_t (&f)[4][4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) ldsm4(base + swz(r, kt * 2 + (lane >> 4)), f[kt][0], f[kt][1], f[kt][2], f[kt][3]);
}

// acc[nt] (16 x 8 per n-tile) = A(16 x 64, fragments) . M[rows n0 + nt*8 ..][:]^T   for nt < n_nt   (M row-major [row][64])
template <bool kF16>
__device__ __forceinline__ void tile_abt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t mbase, int n0, int n_nt,
                                         int lane) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    if (nt < n_nt) {
      const int r = n0 + nt * 8 + (lane & 7);
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        ldsm4(mbase + swz(r, kp * 4 + (lane >> 3)), b0, b1, b2, b3);
        mma<kF16>(acc[nt], a[kp * 2], b0, b1);
        mma<kF16>(acc[nt], a[kp * 2 + 1], b2, b3);
      }
    }
  }
}

// out(16 x 64) += W(16 x (8*n_nt), accumulator layout, packed on the fly) . M[rows k0 ..][:]      (M row-major [row][64])
template <bool kF16>
__device__ __forceinline__ void tile_ab(float (&out)[8][4], const float (&w)[8][4], uint32_t mbase, int k0, int n_nt,
                                        int lane) {
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    if (2 * kt < n_nt) {
      uint32_t pf[4];
      pf[0] = Act<kF16>::pack(w[2 * kt][0], w[2 * kt][1]);
      pf[1] = Act<kF16>::pack(w[2 * kt][2], w[2 * kt][3]);
      pf[2] = Act<kF16>::pack(w[2 * kt + 1][0], w[2 * kt + 1][1]);
      pf[3] = Act<kF16>::pack(w[2 * kt + 1][2], w[2 * kt + 1][3]);
      const int r = k0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm4t(mbase + swz(r, np * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma<kF16>(out[np * 2], pf, b0, b1);
        mma<kF16>(out[np * 2 + 1], pf, b2, b3);
      }
    }
  }
}

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kWarps * 32, 1)
attention_bwd_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ o, const uint16_t* __restrict__ d_o,
                     const float* __restrict__ lse, uint16_t* __restrict__ dqkv, int T, int TP, int H, float scale,
                     float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int d = H * kHd;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_bytes = uint32_t(TP) * 128u;
  const uint32_t qs = smem_u32(smem), ks = qs + tile_bytes, vs = ks + tile_bytes, dos = vs + tile_bytes;
  float* s_lse = reinterpret_cast<float*>(smem + 4 * size_t(tile_bytes));
  float* s_d = s_lse + TP;

  const uint16_t* qbase = qkv + size_t(b) * T * (3 * d) + h * kHd;
  const uint16_t* obase = o + size_t(b) * T * d + h * kHd;
  const uint16_t* dobase = d_o + size_t(b) * T * d + h * kHd;

  // ---- stage Q, K, V, dO; D = rowsum(dO o O); lse ----
  for (int i = threadIdx.x; i < TP * 8; i += kWarps * 32) {
    const int row = i >> 3, ch = i & 7;
    if (row < T) {
      const uint16_t* src = qbase + size_t(row) * (3 * d) + ch * 8;
      cp_async16(qs + swz(row, ch), src);
      cp_async16(ks + swz(row, ch), src + d);
      cp_async16(vs + swz(row, ch), src + 2 * d);
      cp_async16(dos + swz(row, ch), dobase + size_t(row) * d + ch * 8);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) *reinterpret_cast<uint4*>(smem + m * size_t(tile_bytes) + swz(row, ch)) = z;
    }
    // D: 8 lanes per row, 8 elements each
    float part = 0.f;
    if (row < T) {
      const uint4 a = *reinterpret_cast<const uint4*>(dobase + size_t(row) * d + ch * 8);
      const uint4 c = *reinterpret_cast<const uint4*>(obase + size_t(row) * d + ch * 8);
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* pc = reinterpret_cast<const uint32_t*>(&c);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = Act<kF16>::unpack(pa[e]), y = Act<kF16>::unpack(pc[e]);
        part += x.x * y.x + x.y * y.y;
      }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (ch == 0) {
      s_d[row] = part;
      s_lse[row] = row < T ? lse[size_t(bh) * T + row] : INFINITY;   // padded queries: P = exp2(-inf) = 0
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int n_tiles16 = TP / 16;
  const int n_chunks = (TP + 63) / 64;

  // ================= pass Q: dQ =================
  for (int qt = warp; qt < n_tiles16; qt += kWarps) {
    const int q0 = qt * 16;
    uint32_t qf[4][4], dof[4][4];
    load_a_frags(qs, q0, lane, qf);
    load_a_frags(dos, q0, lane, dof);
    const float l0 = s_lse[q0 + g], l1 = s_lse[q0 + g + 8], d0 = s_d[q0 + g], d1 = s_d[q0 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      const int key0 = c * 64;
      const int n_nt = min(8, (TP - key0) >> 3);
      float s[8][4], dp[8][4];
      tile_abt<kF16>(s, qf, ks, key0, n_nt, lane);
      tile_abt<kF16>(dp, dof, vs, key0, n_nt, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int key = key0 + nt * 8 + 2 * t;
        const bool v0 = key < T, v1 = key + 1 < T;
        const float p0 = v0 ? exp2f(s[nt][0] * scale_log2e - l0) : 0.f;
        const float p1 = v1 ? exp2f(s[nt][1] * scale_log2e - l0) : 0.f;
        const float p2 = v0 ? exp2f(s[nt][2] * scale_log2e - l1) : 0.f;
        const float p3 = v1 ? exp2f(s[nt][3] * scale_log2e - l1) : 0.f;
        s[nt][0] = p0 * (dp[nt][0] - d0) * scale;
        s[nt][1] = p1 * (dp[nt][1] - d0) * scale;
        s[nt][2] = p2 * (dp[nt][2] - d1) * scale;
        s[nt][3] = p3 * (dp[nt][3] - d1) * scale;
      }
      tile_ab<kF16>(dq, s, ks, key0, n_nt, lane);
    }
    const int r0 = q0 + g, r1 = q0 + g + 8;
    uint16_t* ob = dqkv + size_t(b) * T * (3 * d) + h * kHd;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (r0 < T) *reinterpret_cast<uint32_t*>(ob + size_t(r0) * (3 * d) + col) = Act<kF16>::pack(dq[nt][0], dq[nt][1]);
      if (r1 < T) *reinterpret_cast<uint32_t*>(ob + size_t(r1) * (3 * d) + col) = Act<kF16>::pack(dq[nt][2], dq[nt][3]);
    }
  }

  // ================= pass K: dK, dV =================
  for (int kt16 = warp; kt16 < n_tiles16; kt16 += kWarps) {
    const int k0 = kt16 * 16;
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(ks, k0, lane, kf);
    load_a_frags(vs, k0, lane, vf);
    const bool kv0 = k0 + g < T, kv1 = k0 + g + 8 < T;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    for (int c = 0; c < n_chunks; ++c) {
      const int qq0 = c * 64;
      const int n_nt = min(8, (TP - qq0) >> 3);
      float st[8][4], dpt[8][4];
      tile_abt<kF16>(st, kf, qs, qq0, n_nt, lane);     // S^T  [16 keys x 64 queries]
      tile_abt<kF16>(dpt, vf, dos, qq0, n_nt, lane);   // dP^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int qc = qq0 + nt * 8 + 2 * t;           // query columns qc, qc+1 (padded queries carry lse = +inf)
        const float la = nt < n_nt ? s_lse[qc] : INFINITY, lb = nt < n_nt ? s_lse[qc + 1] : INFINITY;
        const float da = nt < n_nt ? s_d[qc] : 0.f, db = nt < n_nt ? s_d[qc + 1] : 0.f;
        const float p0 = kv0 ? exp2f(st[nt][0] * scale_log2e - la) : 0.f;
        const float p1 = kv0 ? exp2f(st[nt][1] * scale_log2e - lb) : 0.f;
        const float p2 = kv1 ? exp2f(st[nt][2] * scale_log2e - la) : 0.f;
        const float p3 = kv1 ? exp2f(st[nt][3] * scale_log2e - lb) : 0.f;
        st[nt][0] = p0; st[nt][1] = p1; st[nt][2] = p2; st[nt][3] = p3;
        dpt[nt][0] = p0 * (dpt[nt][0] - da) * scale;
        dpt[nt][1] = p1 * (dpt[nt][1] - db) * scale;
        dpt[nt][2] = p2 * (dpt[nt][2] - da) * scale;
        dpt[nt][3] = p3 * (dpt[nt][3] - db) * scale;
      }
      tile_ab<kF16>(dv, st, dos, qq0, n_nt, lane);     // dV += P^T dO
      tile_ab<kF16>(dk, dpt, qs, qq0, n_nt, lane);     // dK += dS^T Q
    }
    const int r0 = k0 + g, r1 = k0 + g + 8;
    uint16_t* ob = dqkv + size_t(b) * T * (3 * d) + h * kHd;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (r0 < T) {
        *reinterpret_cast<uint32_t*>(ob + size_t(r0) * (3 * d) + d + col) = Act<kF16>::pack(dk[nt][0], dk[nt][1]);
        *reinterpret_cast<uint32_t*>(ob + size_t(r0) * (3 * d) + 2 * d + col) = Act<kF16>::pack(dv[nt][0], dv[nt][1]);
      }
      if (r1 < T) {
        *reinterpret_cast<uint32_t*>(ob + size_t(r1) * (3 * d) + d + col) = Act<kF16>::pack(dk[nt][2], dk[nt][3]);
        *reinterpret_cast<uint32_t*>(ob + size_t(r1) * (3 * d) + 2 * d + col) = Act<kF16>::pack(dv[nt][2], dv[nt][3]);
      }
    }
  }
}

int launch_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* dqkv, int B, int T,
                         int H, int head_dim, int f16, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd) return -1;
  const int TP = (T + 15) / 16 * 16;
  const size_t smem = size_t(TP) * 128 * 4 + size_t(TP) * 8;
  if (smem > 227 * 1024) return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const float scale = 1.0f / sqrtf(float(head_dim));
  const float scale_log2e = 1.4426950408889634f * scale;
  const uint16_t* q = static_cast<const uint16_t*>(qkv);
  const uint16_t* o = static_cast<const uint16_t*>(out);
  const uint16_t* g = static_cast<const uint16_t*>(d_out);
  uint16_t* dq = static_cast<uint16_t*>(dqkv);
  if (f16)
    attention_bwd_kernel<true><<<B * H, kWarps * 32, smem, stream>>>(q, o, g, lse, dq, T, TP, H, scale, scale_log2e);
  else
    attention_bwd_kernel<false><<<B * H, kWarps * 32, smem, stream>>>(q, o, g, lse, dq, T, TP, H, scale, scale_log2e);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
