// tcgen05 attention for the short ViT sequence (T <= 208 keys, head_dim 64, no mask): the WHOLE score row lives in TMEM.
//
//   softmax(Q K^T / sqrt(64)) V   per (image, head)        [reference: nn.MultiheadAttention inside CLIP's
//   ResidualAttentionBlock, reached through model.encode_image at /root/reference/main.py:204, 444, 503]
//
// One persistent CTA per SM walks (image, head) items; K and V of two items are resident in 128B-swizzled smem.  Queries form
// 128-row tiles; the two tiles of a 197-token item go to the two TMEM REGIONS (256 columns each), each served by its own
// softmax GROUP of 8 warps and its own MMA issuer warp:
//
//   S = Q K^T              ONE tcgen05.mma chain (4 k-steps, N = all keys rounded up to 16: 208 at T = 197) into region
//                          columns [0, N): the score row of a query is complete in TMEM before the softmax touches it.
//   softmax                exact row maximum first (pass 1), then p = exp2(s*c - m*c) (pass 2): no online rescale, no
//                          speculative exponentials, no redo path.  TWO threads per query row (column halves A / B, in
//                          different warps of the same TMEM lane quadrant): 16 softmax warps = 4 per scheduler.  Every
//                          tcgen05.wait::ld costs ~100+ cycles, so a half reads its scores in three loads per tile: its
//                          lower <= 4 units (pass 1), its top <= 3 units (pass 1; they stay in registers for pass 2), the
//                          lower units again (pass 2).  The halves meet through smem + a 64-thread named barrier (row
//                          maximum, then row sum).  P (16 bit) is written IN PLACE: P(u) into the first 8 columns of unit u's
//                          own S columns, except unit 12, whose columns are O's: into the second half of unit 11's.
//   O = P V                tcgen05.mma with A = P from TMEM (one k-step per 16 keys), B = V in its natural [key][dim]
//                          layout (MN-major), accumulator in region columns [192, 256) - S columns that are dead by then.
//                          Each half releases its P in two chunks (top units, lower units); half B's first chunk holds unit
//                          12, so the first MMA (which overwrites O's columns) waits for it.
//   epilogue               O row * 1 / sum -> 16 bit -> 64B-swizzled smem slab -> ONE TMA store per warp (32 rows x 32 dims;
//                          rows >= T are clipped by the [B][T][d] tensor map); optional log2-domain LSE for the backward pass.
//
//   warp 0  TMA producer (K, V per item; Q tile per group)     warp 1 / 2  MMA issuer of group 0 / 1     warp 2 also owns
//   the TMEM allocation     warps 4-19  softmax: (warp - 4) & 3 = lane quadrant, bit 2 = group, bit 3 = column half.
//
// Measured at B = 1024, T = 197, H = 12 (12288 items, 83 per SM): 0.357 ms against 0.423 ms for the block-wise kernel
// (attention_sm100.cu).  Floors of the launch: HBM 1.24 GB = 0.19 ms; MUFU 0.12-0.16 ms; tensor pipe ~0.16 ms (S 4 x ~130 clk
// + P.V 13 x 58-77 clk per tile: consecutive MMAs into one accumulator serialise).  Ablations on the same box
// (tools/build_variant.sh): TMA loads + barriers only 0.140 ms (the HBM read floor), + both MMA chains 0.204, exponentials
// +0.045, pass 1 + exchange +0.04.  What bounds it is the serial chain of a tile (S -> pass 1 -> pass 2 -> P.V -> epilogue,
// ~8000 clk per item with ~100-250 clk per hop) and issue slots, not a pipe: running a share of the exponentials as a
// polynomial on the FMA pipe (the block-wise kernel's trick) makes it SLOWER here (0.383 / 0.398 / 0.408 ms for 2 / 3 / 4 of 8
// pairs).  Tried and dropped, each measured: one thread per row x two passes of x16 double-buffered loads (0.387); all 16
// warps on one tile with four threads per row, scores read once and kept in registers, the previous tile's epilogue deferred
// into the next tile's body (0.411 - every phase runs in lock-step and nothing fills the gaps); two O accumulators for
// alternating units, densely packed P, two-stage release (0.48-0.56); the MMA issuer on the scheduler of the mostly idle lane
// quadrant 3 (no change); a staggered start of the two groups (no change); mbarrier.try_wait with a suspend hint (no change,
// kept: fewer polling instructions).  Longer sequences (ViT-L/14 @ 336: 577 keys) and the causal text tower stay on
// attention_sm100.cu (online softmax over 80-key blocks).  All waits are bounded (trap, never hang).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "act_types.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace iic {

namespace {

constexpr int kHd = 64;
constexpr int kThreads = 640;            // 4 control warps + 16 softmax warps
constexpr int kQTileBytes = 128 * 128;   // 128 query rows x 64 dims x 2 B
constexpr int kRegionCols = 256;         // TMEM columns per group: S [0, 16U), P in place, O [192, 256)
constexpr int kOCol = 192;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kXchgBytes = 2 * 2 * 2 * 2 * 128 * 4;   // {max, sum} x tile parity x group x half x 128 rows, f32
constexpr int kSlabBytes = 16 * 2048;    // per softmax warp: 32 rows x 64 B of output, staged for its TMA store
constexpr int kBarBytes = 512;

struct RowAttnParams {
  uint16_t* out;
  float* lse;
  int items, T, H;
  int nq;            // 128-row query tiles per item (1 or 2)
  int U;             // 16-key units (keys rounded up to 16): S has 16 U columns
  int kv_bytes;      // bytes of one K (or V) buffer = 16 U rows x 128 B
  float scale_log2e;
#ifdef IIC_ATTN_TRACE
  long long* trace;  // debug builds only: [warp][tile][8] clock64 stamps of CTA 0
#endif
};

#ifdef IIC_ATTN_TRACE
#define TRACE(slot) do { if (blockIdx.x == 0 && lane == 0 && trace_tile < 32) prm.trace[(warp * 32 + trace_tile) * 8 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

__host__ __device__ constexpr uint32_t idesc_row(uint32_t m, uint32_t n, bool f16, bool b_mn) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t id, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(id), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait_row() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#ifndef IIC_ATTN_ROW_POLY_MASK
#define IIC_ATTN_ROW_POLY_MASK 0x00   // bit e: pair e of the 8 pairs of a 16-column unit runs its exponentials as a polynomial on the FMA
                                      // pipe.  Measured at B = 1024, T = 197: 0x00 0.357 ms, 0x24 (2 of 8) 0.383, 0x94 (3 of 8) 0.398,
                                      // 0x55 (4 of 8) 0.408 - this kernel is bound by issue slots and latency, not by the MUFU pipe
#endif

__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// exp2 of two values on the FMA pipe: x = n + f with n = rint(x) from the magic-number add, 2^f by a degree-3 polynomial on
// [-0.5, 0.5] (max relative error 7.5e-5, below the 16-bit rounding of P), 2^n by adding n to the exponent field.
__device__ __forceinline__ void ex2_poly_pair(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  upk2(x2, x0, x1);
  const uint64_t xc = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t t2 = fadd2(xc, pk2(12582912.f, 12582912.f));
  const uint64_t n2 = fadd2(t2, pk2(-12582912.f, -12582912.f));
  const uint64_t f2 = ffma2(n2, pk2(-1.f, -1.f), xc);
  uint64_t p2 = ffma2(pk2(0.0551716685f, 0.0551716685f), f2, pk2(0.2426111251f, 0.2426111251f));
  p2 = ffma2(p2, f2, pk2(0.6932609677f, 0.6932609677f));
  p2 = ffma2(p2, f2, pk2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  upk2(t2, t0, t1);
  upk2(p2, p0, p1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tma_store_3d(const void* desc, uint32_t src_smem, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// maximum over N consecutive fp32 columns held in v (two independent chains)
template <int N>
__device__ __forceinline__ float max_regs(const uint32_t* v, float m) {
  float a = m, b = -INFINITY;
#pragma unroll
  for (int e = 0; e < N; e += 4) {
    a = max3f(a, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
    b = max3f(b, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
  }
  return fmaxf(a, b);
}

// one 16-key unit of pass 2: p = exp2(s*c - mc) for the 16 scores in v -> 8 packed 16-bit pairs in pk, row sum into acc
template <bool kF16>
__device__ __forceinline__ void exp_unit(const uint32_t* v, uint32_t* pk, uint64_t c2, uint64_t nm2, uint64_t& acc0, uint64_t& acc1) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint64_t xx = ffma2(pk2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), c2, nm2);
    float e0, e1;
    if ((IIC_ATTN_ROW_POLY_MASK >> e) & 1) {
      ex2_poly_pair(xx, e0, e1);
    } else {
      float x0, x1;
      upk2(xx, x0, x1);
      e0 = ex2f(x0);
      e1 = ex2f(x1);
    }
    if (e & 1) acc1 = fadd2(acc1, pk2(e0, e1)); else acc0 = fadd2(acc0, pk2(e0, e1));
    pk[e] = Act<kF16>::pack(e0, e1);
  }
}

__device__ __forceinline__ void mask_unit(uint32_t* v, int nvalid) {
#pragma unroll
  for (int e = 0; e < 16; ++e)
    if (e >= nvalid) v[e] = 0xff800000u;   // -inf: never the maximum, exp2 -> 0 (padding keys carry no weight)
}

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_row_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                           const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ RowAttnParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t q_s = base;                                    // Q tile of group g at g * 16 KB
  const uint32_t slab_s = base + 2 * kQTileBytes;               // output slab of softmax warp w at w * 2 KB (32 rows x 64 B)
  const uint32_t kv_s = slab_s + kSlabBytes;                    // stage s: K at kv_s + s*2*kv_bytes, V right behind
  const uint32_t xchg_off = uint32_t(2 * kQTileBytes + kSlabBytes) + 4u * uint32_t(prm.kv_bytes);
  float* xchg = reinterpret_cast<float*>(gen + xchg_off);       // [kind][parity][group][half][128]
  const uint32_t bar = base + xchg_off + kXchgBytes;
  auto k_full = [&](int s) { return bar + 8u * s; };
  auto v_full = [&](int s) { return bar + 16 + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 32 + 8u * s; };
  auto q_full = [&](int g) { return bar + 48 + 8u * g; };
  auto q_empty = [&](int g) { return bar + 64 + 8u * g; };
  auto s_full = [&](int g) { return bar + 80 + 8u * g; };
  auto o_full = [&](int g) { return bar + 96 + 8u * g; };
  auto o_free = [&](int g) { return bar + 112 + 8u * g; };
  // P of a pair of 16-key units is complete (all four lane quadrants): chunk c of column half hf of group g
  auto p_done = [&](int g, int hf, int c) { return bar + 128 + 8u * uint32_t((g * 2 + hf) * 4 + c); };
  const uint32_t tmem_slot = bar + 256;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + xchg_off + kXchgBytes + 256);

  ptx::pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = prm.T, H = prm.H, nq = prm.nq, U = prm.U, items = prm.items;
  const int d = H * kHd;
  const int ua = (U + 1) >> 1, ub = U - ua;        // 16-key units of column half A (units [0, ua)) / B (units [ua, U))

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_kv);
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(k_full(s), 1);
      ptx::mbar_init(v_full(s), 1);
      ptx::mbar_init(kv_empty(s), uint32_t(nq));   // one tcgen05.commit per group that reads the stage
      ptx::mbar_init(q_full(s), 1);
      ptx::mbar_init(q_empty(s), 1);               // tcgen05.commit
      ptx::mbar_init(s_full(s), 1);                // tcgen05.commit
      ptx::mbar_init(o_full(s), 1);                // tcgen05.commit
      ptx::mbar_init(o_free(s), 8);                // one elected lane per softmax warp of the group
      for (int hf = 0; hf < 2; ++hf)
        for (int c = 0; c < 4; ++c) ptx::mbar_init(p_done(s, hf, c), 4);   // the four quadrant warps of that half
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::pdl_wait();   // set-up done: from here on the previous kernel's output (qkv) is read

  // tile t of the CTA's n-th item belongs to group t ^ (n & 1) (two tiles) or n & 1 (one tile): the group that gets the short
  // second tile of a 197-token sequence alternates from item to item
  auto group_of = [&](int n, int t) { return (t ^ n) & 1; };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ======================= TMA producer: K, V of every item (2 stages), Q tile per group =======================
      if (ptx::elect_one()) {
        uint32_t cnt0 = 0, cnt1 = 0;   // Q tiles handed to group 0 / 1 so far (scalars: a dynamically indexed array would live in local memory)
        int n = 0;
        const uint32_t kv_tx = uint32_t(prm.kv_bytes);
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
          const int b = item / H, h = item - b * H;
          const int row0 = b * T;
          const int s = n & 1;
          const uint32_t par = uint32_t(n >> 1) & 1u;
          const uint32_t ks = kv_s + uint32_t(s) * 2u * uint32_t(prm.kv_bytes), vs = ks + uint32_t(prm.kv_bytes);
          ptx::mbar_wait(kv_empty(s), par ^ 1u);
          ptx::mbar_arrive_expect_tx(k_full(s), kv_tx);
          ptx::tma_load_2d(&tm_kv, k_full(s), ks, d + h * kHd, row0, ptx::kEvictFirst);
          for (int t = 0; t < nq; ++t) {
            const int g = group_of(n, t);
            const uint32_t cg = g ? cnt1 : cnt0;
            ptx::mbar_wait(q_empty(g), (cg & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(q_full(g), kQTileBytes);
            ptx::tma_load_2d(&tm_q, q_full(g), q_s + uint32_t(g) * kQTileBytes, h * kHd, row0 + t * 128, ptx::kEvictFirst);
            if (g) ++cnt1; else ++cnt0;
          }
          ptx::mbar_arrive_expect_tx(v_full(s), kv_tx);
          ptx::tma_load_2d(&tm_kv, v_full(s), vs, 2 * d + h * kHd, row0, ptx::kEvictFirst);
        }
      }
      __syncwarp();
    } else if (warp == 1 || warp == 2) {
      // ======================= MMA issuer of group g: S, then P.V chunk by chunk as the softmax releases them ===========
      const int g = warp - 1;
      const uint32_t region = tmem_base + uint32_t(g * kRegionCols);
      const uint32_t id_s = idesc_row(128, uint32_t(16 * U), kF16, false);
      const uint32_t id_o = idesc_row(128, kHd, kF16, true);
      // unit 12 (T > 192) shares its S columns with the O accumulator (columns [192, 256)): half B walks its units from the top,
      // and the first P.V MMA (which overwrites O) waits for half B's first chunk, the one that consumes unit 12
      const int gate = ub == 0 ? 0 : 1;
      uint32_t c = 0;   // tiles of this group so far
      int n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        if (nq == 1 && group_of(n, 0) != g) continue;
        const int stage = n & 1;
        const uint32_t kv_par = uint32_t(n >> 1) & 1u, par = c & 1u;
        const uint32_t ks = kv_s + uint32_t(stage) * 2u * uint32_t(prm.kv_bytes), vs = ks + uint32_t(prm.kv_bytes);
        const uint32_t trace_tile = c; (void)trace_tile;
        TRACE(0);
        // ---- S = Q K^T (whole row) ----
        ptx::mbar_wait(q_full(g), par);
        ptx::mbar_wait(k_full(stage), kv_par);
        TRACE(1);
        if (c > 0) ptx::mbar_wait(o_free(g), par ^ 1u);   // the previous tile's O (and P) have left the region
        TRACE(2);
        ptx::tcgen05_fence_after();
        if (ptx::elect_one()) {
          const uint64_t dq = ptx::make_kmajor_sw128_desc(q_s + uint32_t(g) * kQTileBytes);
          const uint64_t dk = ptx::make_kmajor_sw128_desc(ks);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_f16<1>(region, dq + uint64_t(2 * kk), dk + uint64_t(2 * kk), id_s, kk != 0 ? 1u : 0u);
          ptx::umma_commit<1>(s_full(g));
          ptx::umma_commit<1>(q_empty(g));
        }
        __syncwarp();
        TRACE(3);
        // ---- O = P V: one k-step per 16-key unit, P(u) sits in the first 8 columns of the unit's own S columns ----
        // V rows are keys: 16 keys per k-step = 2048 bytes further down the [key][64 dims] tile (MN-major operand, 128B
        // swizzle: 8-key groups 1024 bytes apart (SBO), one 64-wide atom along N)
        uint64_t dv = uint64_t((vs >> 4) & 0x3FFFu);
        dv |= uint64_t(1) << 16;
        dv |= uint64_t(1024 >> 4) << 32;
        dv |= uint64_t(1) << 46;
        dv |= uint64_t(2) << 61;
        ptx::mbar_wait(v_full(stage), kv_par);
        uint32_t acc = 0;
        // chunk 0 of a half = its kept set (the top <= 3 units, walked downwards), chunk 1 = the rest (the <= 4 units below,
        // upwards): the softmax warps' order.  One k-step per unit, P(u) where the softmax warps put it.
        auto issue_chunk = [&](int hf, int ch) {
          const int lo_h = hf ? ua : 0, hi_h = hf ? U : ua;
          const int keep_h = hi_h - lo_h < 3 ? hi_h - lo_h : 3;
          const int count = ch ? hi_h - lo_h - keep_h : keep_h;
          ptx::mbar_wait(p_done(g, hf, ch), par);
          ptx::tcgen05_fence_after();
          if (ptx::elect_one()) {
            for (int k = 0; k < count; ++k) {
              const int u = ch ? lo_h + k : hi_h - 1 - k;
              const uint32_t pa = region + uint32_t(u == 12 ? 16 * 11 + 8 : 16 * u);
              umma_ts(region + kOCol, pa, dv + uint64_t(u * 128), id_o, acc);
              acc = 1u;
            }
          }
          __syncwarp();
          acc = 1u;
        };
        const int nch_a = ua > 3 ? 2 : (ua > 0 ? 1 : 0), nch_b = ub > 3 ? 2 : (ub > 0 ? 1 : 0);
        int ia = 0, ib = 0;
        for (; ib < gate; ++ib) issue_chunk(1, ib);
        TRACE(4);
        while (ia < nch_a || ib < nch_b) {
          if (ia < nch_a) { issue_chunk(0, ia); ++ia; }
          if (ib < nch_b) { issue_chunk(1, ib); ++ib; }
        }
        TRACE(5);
        if (ptx::elect_one()) {
          ptx::umma_commit<1>(o_full(g));
          ptx::umma_commit<1>(kv_empty(stage));   // this group's reads of the item's K / V are complete with these MMAs
        }
        __syncwarp();
        TRACE(6);
        ++c;
      }
    }
  } else {
    // ======================= softmax / output: two threads (column halves) per query row =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int w = warp - 4;
    const int quad = w & 3, g = (w >> 2) & 1, half = w >> 3;
    const int r = quad * 32 + lane;                 // row inside the 128-row query tile
    const uint32_t region = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(g * kRegionCols);
    const int my_units = half ? ub : ua;
    // the last unit of the row holds padding keys when T is not a multiple of 16: it belongs to half B (half A when U == 1)
    const bool owns_tail = (16 * U != T) && (half == 1 || ub == 0);
    const int tail_valid = T - 16 * (U - 1);                  // valid columns of the last unit (1..16)
    const int lo = half ? ua : 0, hi = half ? U : ua;                              // this half's units [lo, hi)
    const int n_keep = my_units < 3 ? my_units : 3, n_rest = my_units - n_keep;   // see pass 1
    const int n_chunks = n_rest > 0 ? 2 : (my_units > 0 ? 1 : 0);                  // releases to the P.V issuer per tile
    const float c = prm.scale_log2e;
    const uint32_t bar_id = 1u + uint32_t(g * 4 + quad);      // named barrier of the two warps (halves) sharing these rows
    const uint32_t slab = slab_s + uint32_t(w) * 2048u;
    uint8_t* slab_ptr = gen + (2 * kQTileBytes) + w * 2048;
    uint32_t cnt = 0;                                         // tiles of this group so far
    int n = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const int t = nq == 2 ? ((g ^ n) & 1) : 0;
      if (nq == 1 && group_of(n, 0) != g) continue;
      const int b = item / H, h = item - b * H;
      const uint32_t par = cnt & 1u;
      const bool live = t * 128 + quad * 32 < T;    // warps whose 32 query rows are all padding do no math
      const uint32_t trace_tile = cnt; (void)trace_tile;
      TRACE(0);
      ptx::mbar_wait(s_full(g), par);
      TRACE(1);
      ptx::tcgen05_fence_after();
      // Every tcgen05.wait::ld costs ~100+ cycles whatever it waits for, so the row is read in as few loads as possible: the
      // half's units [lo, hi) split into a KEPT set (its top <= 3 units: read last in pass 1 and still in registers when pass 2
      // starts) and a REST set (the <= 4 units below: read first in pass 1 and once more in pass 2) - three waits per tile for
      // the scores, never more than 64 score registers live.  Unit 12 (top of half B at T = 197) is therefore the first unit
      // half B turns into P: it shares its columns with O, see the issuer's gate.
      uint32_t kv[48];
      float m = -INFINITY;
      uint64_t acc0 = 0ull, acc1 = 0ull;
      if (live && my_units > 0) {
        // ---- pass 1: exact row maximum of this half ----
        if (n_rest > 0) {
          uint32_t rv[64];
#pragma unroll
          for (int j = 0; j < 4; ++j)   // unconditional loads (slots past n_rest re-read the last unit): the array stays in registers
            ld16(region + uint32_t(16 * (lo + (j < n_rest ? j : n_rest - 1))), rv + 16 * j);
          ptx::tmem_ld_wait();
          TRACE(2);
#pragma unroll
          for (int j = 0; j < 4; ++j) m = max_regs<16>(rv + 16 * j, m);
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)     // slots past n_keep re-read the lowest kept unit
          ld16(region + uint32_t(16 * (hi - 1 - (i < n_keep ? i : n_keep - 1))), kv + 16 * i);
        ptx::tmem_ld_wait();
        TRACE(3);
        if (owns_tail) mask_unit(kv, tail_valid);       // the padded unit U - 1 is kept slot 0 of the half that owns it
        if (owns_tail && n_keep < 2) mask_unit(kv + 16, tail_valid);   // ... and so are the slots that repeat it
        if (owns_tail && n_keep < 2) mask_unit(kv + 32, tail_valid);
#pragma unroll
        for (int i = 0; i < 3; ++i) m = max_regs<16>(kv + 16 * i, m);
      }
      xchg[((0 * 2 + int(par)) * 2 + g) * 256 + half * 128 + r] = m;
      TRACE(4);
      named_barrier(bar_id, 64);
      TRACE(5);
      m = fmaxf(m, xchg[((0 * 2 + int(par)) * 2 + g) * 256 + (half ^ 1) * 128 + r]);
      // ---- pass 2: exponentials, P in place, two releases per half (kept set, rest set) to the P.V issuer ----
      if (live && my_units > 0) {
        // P(u) goes into the first 8 columns of the unit's own S columns; unit 12 (whose columns are O's) into the second half
        // of unit 11's columns, which this thread holds in registers by then
        const float mc = m * c;
        const uint64_t c2 = pk2(c, c), nm2 = pk2(-mc, -mc);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (i < n_keep) {
            const int u = hi - 1 - i;
            exp_unit<kF16>(kv + 16 * i, pk, c2, nm2, acc0, acc1);
            st8(region + uint32_t(u == 12 ? 16 * 11 + 8 : 16 * u), pk);
          }
        if (n_rest > 0) {
          uint32_t rv[64];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ld16(region + uint32_t(16 * (lo + (j < n_rest ? j : n_rest - 1))), rv + 16 * j);
          // release the kept set while the rest set is in flight
          tmem_st_wait_row();
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(p_done(g, half, 0));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < n_rest) {
              exp_unit<kF16>(rv + 16 * j, pk, c2, nm2, acc0, acc1);
              st8(region + uint32_t(16 * (lo + j)), pk);
            }
          tmem_st_wait_row();
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(p_done(g, half, 1));
        } else {
          tmem_st_wait_row();
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(p_done(g, half, 0));
        }
      } else {
        // padding rows only (or a half without units): nothing to compute, but the issuer counts four quadrant warps per chunk
        __syncwarp();
        if (lane == 0)
          for (int ch = 0; ch < n_chunks; ++ch) ptx::mbar_arrive(p_done(g, half, ch));
      }
      float a0, a1, a2, a3;
      upk2(acc0, a0, a1);
      upk2(acc1, a2, a3);
      const float my_sum = (a0 + a1) + (a2 + a3);
      xchg[((1 * 2 + int(par)) * 2 + g) * 256 + half * 128 + r] = my_sum;
      TRACE(6);
      // the previous tile's output slab must have been read by its TMA store before it is rewritten (issued a tile ago)
      if (lane == 0) ptx::tma_store_wait_read<0>();
      // make the row sum visible to the partner half before either of us can pass o_full (both halves arrive on the barrier)
      named_barrier(bar_id, 64);
      // ---- epilogue: this half's 32 of the 64 output dims -> swizzled smem slab -> one TMA store per warp ----
      ptx::mbar_wait(o_full(g), par);
      ptx::tcgen05_fence_after();
      if (live) {
        uint32_t o[32];
        ld32(region + uint32_t(kOCol + 32 * half), o);
        ptx::tmem_ld_wait();
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_free(g));
        const float sum = my_sum + xchg[((1 * 2 + int(par)) * 2 + g) * 256 + (half ^ 1) * 128 + r];
        const float inv = 1.0f / sum;
        // 64B-swizzled slab (CU_TENSOR_MAP_SWIZZLE_64B): 16-byte chunk j of row `lane` lands at chunk j ^ ((lane >> 1) & 3)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            wv[e] = Act<kF16>::pack(__uint_as_float(o[8 * jj + 2 * e]) * inv, __uint_as_float(o[8 * jj + 2 * e + 1]) * inv);
          *reinterpret_cast<uint4*>(slab_ptr + lane * 64 + ((jj ^ ((lane >> 1) & 3)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        // rows >= T of the tile fall outside dimension 1 of the [B][T][d] tensor map: the store clips them
        if (lane == 0) {
          tma_store_3d(&tm_out, slab, h * kHd + half * 32, t * 128 + quad * 32, b);
          ptx::tma_store_commit();
        }
        const int q = t * 128 + r;
        if (prm.lse != nullptr && half == 0 && q < T) prm.lse[(size_t(b) * H + h) * T + q] = m * c + log2f(sum);
      } else {
        ptx::tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_free(g));
      }
      TRACE(7);
      ++cnt;
    }
    if (lane == 0) ptx::tma_store_wait<0>();   // all output stores have landed before the CTA exits
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

namespace {
PFN_cuTensorMapEncodeTiled_v12000 encode_fn_row() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}
bool make_map_row(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f16) {
  auto fn = encode_fn_row();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr,
            box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// output [B][T][d] as a 3-D map: a warp stores 32 rows x 32 dims (64 B, 64B-swizzled slab); rows >= T of a tile are clipped
bool make_map_out(CUtensorMap* m, void* ptr, int B, int T, int d, bool f16) {
  auto fn = encode_fn_row();
  if (!fn) return false;
  cuuint64_t gdim[3] = {cuuint64_t(d), cuuint64_t(T), cuuint64_t(B)};
  cuuint64_t gstr[2] = {cuuint64_t(d) * 2, cuuint64_t(T) * cuuint64_t(d) * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// returns -3 when the shape is outside this kernel's envelope (T <= 208, head_dim 64): the caller falls back to the
// block-wise kernel (attention_sm100.cu)
int launch_attention_row_sm100(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16, int num_sms,
                               cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1 || T > 208) return -3;   // 13 units: at most one unit overlaps the O columns
  RowAttnParams p;
  p.U = (T + 15) / 16;
  p.kv_bytes = p.U * 16 * 128;
  p.nq = (T + 127) / 128;
  p.items = B * H;
  p.T = T;
  p.H = H;
  p.out = static_cast<uint16_t*>(out);
  p.lse = lse;
  p.scale_log2e = 1.4426950408889634f / sqrtf(float(head_dim));
  const int smem = 2 * kQTileBytes + kSlabBytes + 4 * p.kv_bytes + kXchgBytes + kBarBytes + 1024 /*alignment slack*/;
  if (smem > kMaxSmem) return -3;
  const int d = H * kHd;
  CUtensorMap tq, tkv, tout;
  const uint64_t rows = uint64_t(B) * T;
  if (!make_map_row(&tq, qkv, rows, uint64_t(3 * d), 128, f16 != 0) ||
      !make_map_row(&tkv, qkv, rows, uint64_t(3 * d), uint32_t(16 * p.U), f16 != 0) ||
      !make_map_out(&tout, out, B, T, d, f16 != 0))
    return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_row_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_row_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
#ifdef IIC_ATTN_TRACE
  static long long* d_trace = nullptr;
  if (!d_trace) cudaMalloc(&d_trace, 20 * 32 * 8 * sizeof(long long));
  cudaMemsetAsync(d_trace, 0, 20 * 32 * 8 * sizeof(long long), stream);
  p.trace = d_trace;
#endif
  if (f16) launch_k(attention_row_sm100_kernel<true>, dim3(grid), dim3(kThreads), size_t(smem), stream, tq, tkv, tout, p);
  else launch_k(attention_row_sm100_kernel<false>, dim3(grid), dim3(kThreads), size_t(smem), stream, tq, tkv, tout, p);
#ifdef IIC_ATTN_TRACE
  if (const char* path = getenv("IIC_ATTN_TRACE_OUT")) {
    static long long h_trace[20 * 32 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h_trace, d_trace, sizeof(h_trace), cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(path, "w")) {
      for (int w = 0; w < 20; ++w)
        for (int t = 0; t < 32; ++t) {
          fprintf(f, "%d %d", w, t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %lld", h_trace[(w * 32 + t) * 8 + k]);
          fprintf(f, "\n");
        }
      fclose(f);
    }
  }
#endif
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
