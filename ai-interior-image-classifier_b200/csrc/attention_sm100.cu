// tcgen05 attention, head_dim 64, any ViT sequence length whose K/V of one head fit one SM's shared memory (T <= ~640):
//
//   softmax(Q K^T / sqrt(64)) V   per (image, head)        [reference: nn.MultiheadAttention inside CLIP's
//   ResidualAttentionBlock, reached through model.encode_image at /root/reference/main.py:204, 444, 503]
//
// One persistent CTA per SM walks (image, head) items; K and V of the item are resident in 128B-swizzled smem (two items
// in flight when they fit).  The queries form 128-row tiles; two softmax groups (4 warps each) take alternate tiles.
// The keys are cut into nb blocks of <= 80 (a multiple of 16 each), and per (tile, block) "op":
//
//   S = Q K_j^T            tcgen05.mma, fp32 accumulator in one of the group's TWO 80-column S buffers in TMEM, so the
//                          tensor core always runs one or two ops AHEAD of the softmax (also across tiles and items).
//   softmax                ONE THREAD PER QUERY ROW: tcgen05.ld 32x32b hands a thread its own row of the block (<= 80
//                          registers), so max and sum are thread-local: no shuffles, S is read from TMEM exactly once.
//                          Online softmax with a LAZY running maximum: m only moves when a block exceeds it by more than
//                          2^8 (then O and the row sum are rescaled, a rare path taken warp-uniformly); otherwise
//                          p = exp2(s*c - m) may reach 2^8, which is harmless in fp32 / bf16 / fp16.  Packed FFMA2 /
//                          FADD2, FMNMX3; P -> 16 bit -> written back IN PLACE over S (tcgen05.st): P never touches smem.
//   O (+)= P_j V_j         tcgen05.mma with A = P from TMEM, B = V in its natural [key][dim] layout (MN-major operand),
//                          accumulator in columns [192, 256) of the group's region.
//   epilogue               O row * 1/sum -> 16 bit -> global; optional log2-domain LSE for the backward pass.
//
//   warp 0  TMA producer K/V      warp 1 / 2  MMA issuers of group 0 / 1 (warp 2 also owns the TMEM allocation)
//   warp 3  TMA producer Q tiles (2-deep ring per group)           warps 4-7 / 8-11  softmax groups 0 / 1
//
// The softmax is bound by the MUFU pipe (16 ex2/clk/SM, tools/ubench_sm100.cu); everything else in the inner loop is
// ~2 issue slots per element.  tcgen05.mma instructions of one CTA execute in issue order, which is what orders
// "PV_k reads P_k" before "S_{k+2} overwrites it".  All waits are bounded (trap, never hang).
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
constexpr int kThreads = 384;
constexpr int kQTileBytes = 128 * 128;   // 128 query rows x 64 dims x 2 B
constexpr int kRegionCols = 256;         // TMEM columns per softmax group: S buffers [0,80) [80,160), O [192,256)
constexpr int kOCol = 192;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kBarBytes = 512;
constexpr float kLazyLog2 = 8.0f;        // the running maximum is left alone until a block exceeds it by 2^8

struct AttnParams {
  uint16_t* out;
  float* lse;
  int items, T, H;
  int nq;            // 128-row query tiles per item
  int nb;            // key blocks per item
  int bq, brem;      // block j holds bq + (j < brem) units of 16 keys
  int TP;            // keys rounded up to 16
  int kv_stages;     // items whose K/V are resident at once (1 or 2)
  int kv_bytes;      // bytes of one K (or V) buffer (TMA boxes may overshoot TP rows)
  int kv_box, kv_loads;
  float scale_log2e;
  int causal;                 // 1: key j is visible to query i only if j <= i (CLIP text tower)
};

// instruction descriptor: D f32, A/B 16-bit, A K-major (or TMEM), B K-major (b_mn = false) or MN-major (b_mn = true)
__host__ __device__ constexpr uint32_t idesc(uint32_t m, uint32_t n, bool f16, bool b_mn) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t id, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(id), "r"(acc)
      : "memory");
}

// 16 consecutive fp32 columns of this thread's TMEM lane <-> v[o .. o+16) (o is a constant after unrolling)
__device__ __forceinline__ void tmem_ld16_at(uint32_t taddr, uint32_t* v, int o) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]),
        "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]),
        "=r"(v[o + 14]), "=r"(v[o + 15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_at(uint32_t taddr, const uint32_t* v, int o) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]),
        "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]),
        "r"(v[o + 14]), "r"(v[o + 15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8_at(uint32_t taddr, const uint32_t* v, int o) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),
                 "r"(v[o + 6]), "r"(v[o + 7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// pairs of a 16-column unit whose exponentials run on the FMA pipe (polynomial) instead of the MUFU pipe: bit e of the mask
#ifndef IIC_ATTN_POLY_MASK
#define IIC_ATTN_POLY_MASK 0x24   // pairs 2 and 5 of 8: a quarter of the exponentials (measured best: 0.47 -> 0.42 ms at T = 197;
                                  // 1/8 and 3/8 are both slower than 2/8 - the MUFU pipe is not the only limiter)
#endif

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2f(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// exp2 of two values on the FMA pipe (the MUFU pipe, 16 ex2/clk/SM, is what bounds the softmax): x = n + f with n = rint(x)
// from the magic-number add, 2^f by a degree-3 polynomial on [-0.5, 0.5] (max relative error 7.5e-5 = 2^-13.7, below the
// 16-bit rounding of P), 2^n by adding n to the exponent field.  Valid for x in [-125, 100].
__device__ __forceinline__ void ex2_poly2(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  const uint64_t xc = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t t2 = add2(xc, pack2(12582912.f, 12582912.f));             // 1.5 * 2^23: the integer part lands in the low bits
  const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
  const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), xc);
  uint64_t p2 = fma2(pack2(0.0551716685f, 0.0551716685f), f2, pack2(0.2426111251f, 0.2426111251f));
  p2 = fma2(p2, f2, pack2(0.6932609677f, 0.6932609677f));
  p2 = fma2(p2, f2, pack2(0.9999280572f, 0.9999280572f));
  float t0, t1, p0, p1;
  unpack2(t2, t0, t1);
  unpack2(p2, p0, p1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

struct SoftmaxRow {
  float m;               // running maximum, already scaled: max(s) * c (lazy: only moves in steps of > 2^8)
  uint64_t acc0, acc1;   // packed partial row sums
};

// Rare path of softmax_block: some row's maximum rises by more than 2^8 in this block.  The speculative P of the block
// (computed with the old maximum) is dropped: the row sum and O are rescaled, S is read again from TMEM (it has not been
// overwritten yet) and the block is redone with the new maximum.  Out of line: never on the fast path.
template <bool kF16>
__device__ __noinline__ void softmax_block_redo(uint32_t sb, uint32_t o_addr, SoftmaxRow& st, float mblk, float c, int units,
                                                int nvalid, uint32_t pv_bar, uint32_t pv_par) {
  const bool mine = mblk > st.m + kLazyLog2;
  const float f = mine ? exp2f(st.m - mblk) : 1.0f;
  if (mine) st.m = mblk;
  const uint64_t f2 = pack2(f, f);
  uint64_t acc0 = mul2f(st.acc0, f2), acc1 = mul2f(st.acc1, f2);
  ptx::mbar_wait(pv_bar, pv_par);   // P.V of the previous op must have landed in O before O is touched
  ptx::tcgen05_fence_after();
  uint32_t o[16];
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    tmem_ld16_at(o_addr + uint32_t(16 * cc), o, 0);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
    tmem_st16_at(o_addr + uint32_t(16 * cc), o, 0);
  }
  const float m = st.m;
#pragma unroll 1
  for (int i = 0; i < units; ++i) {
    tmem_ld16_at(sb + uint32_t(16 * i), o, 0);
    ptx::tmem_ld_wait();
    const int lim = nvalid - 16 * i;   // visible keys of this unit (>= 16: all)
    uint32_t pk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float e0 = ex2(fmaf(__uint_as_float(o[2 * e]), c, -m)), e1 = ex2(fmaf(__uint_as_float(o[2 * e + 1]), c, -m));
      e0 = 2 * e < lim ? e0 : 0.f;
      e1 = 2 * e + 1 < lim ? e1 : 0.f;
      if (e & 1) acc1 = add2(acc1, pack2(e0, e1)); else acc0 = add2(acc0, pack2(e0, e1));
      pk[e] = Act<kF16>::pack(e0, e1);
    }
    tmem_st8_at(sb + uint32_t(8 * i), pk, 0);   // columns [8i, 8i+8) lie behind the S columns still to be read
  }
  st.acc0 = acc0;
  st.acc1 = acc1;
  tmem_st_wait();
}

// One key block (U units of 16 keys) of one query row: S (fp32, TMEM columns [sb, sb + 16U)) -> P (16 bit, in place,
// columns [sb, sb + 8U)), online softmax state in st.  Straight-line code for the compile-time U.  Only the block's last
// unit can hold padding keys (tail < 16 real keys): they are set to -inf once, right after the load.
// First block of a tile: maximum, then exponentials.  Later blocks: the exponentials are computed SPECULATIVELY with the
// running maximum while the block maximum is reduced alongside (independent instruction streams, no max -> exp
// serialisation); only if some row needs a new maximum (rare) the block is redone.  All 32 lanes call this together.
template <bool kF16, int U>
__device__ __forceinline__ void softmax_block(uint32_t sb, uint32_t o_addr, SoftmaxRow& st, float c, bool first, int nvalid,
                                              bool causal, uint32_t pv_bar, uint32_t pv_par) {
  // nvalid: leading columns of the block this row may see (padding keys beyond T; with a causal mask also keys after the
  // query).  Without a causal mask it is warp-uniform and only ever cuts into the last unit.
  uint32_t v[16 * U];
#pragma unroll
  for (int i = 0; i < U; ++i) tmem_ld16_at(sb + uint32_t(16 * i), v, 16 * i);
  ptx::tmem_ld_wait();
  if (!causal) {
    if (nvalid < 16 * U) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (16 * (U - 1) + e >= nvalid) v[16 * (U - 1) + e] = 0xff800000u;   // -inf: exp2 -> 0, never the maximum
    }
  } else if (__any_sync(0xffffffffu, nvalid < 16 * U)) {
#pragma unroll
    for (int e = 0; e < 16 * U; ++e)
      if (e >= nvalid) v[e] = 0xff800000u;
  }
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  auto unit_max = [&](int i) {
#pragma unroll
    for (int e = 0; e < 16; e += 2)
      mx[(e >> 1) & 3] = max3(mx[(e >> 1) & 3], __uint_as_float(v[16 * i + e]), __uint_as_float(v[16 * i + e + 1]));
  };
  if (first) {
#pragma unroll
    for (int i = 0; i < U; ++i) unit_max(i);
    st.m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c;
  }
  const float m = st.m;
  const uint64_t c2 = pack2(c, c), nm2 = pack2(-m, -m);
  uint64_t acc0 = st.acc0, acc1 = st.acc1;
  uint32_t pk[8 * U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    if (!first) unit_max(i);   // independent of the exponentials: the two instruction streams interleave
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint64_t xx = fma2(pack2(__uint_as_float(v[16 * i + 2 * e]), __uint_as_float(v[16 * i + 2 * e + 1])), c2, nm2);
      float e0, e1;
      if ((IIC_ATTN_POLY_MASK >> e) & 1) {
        ex2_poly2(xx, e0, e1);
      } else {
        float x0, x1;
        unpack2(xx, x0, x1);
        e0 = ex2(x0);
        e1 = ex2(x1);
      }
      if (e & 1) acc1 = add2(acc1, pack2(e0, e1)); else acc0 = add2(acc0, pack2(e0, e1));
      pk[8 * i + e] = Act<kF16>::pack(e0, e1);
    }
  }
  if (!first) {
    const float mblk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * c;
    if (__any_sync(0xffffffffu, mblk > m + kLazyLog2)) {
      softmax_block_redo<kF16>(sb, o_addr, st, mblk, c, U, nvalid, pv_bar, pv_par);
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < U; ++i) tmem_st8_at(sb + uint32_t(8 * i), pk, 8 * i);
  st.acc0 = acc0;
  st.acc1 = acc1;
  tmem_st_wait();
}

}  // namespace

// kMaxUnits: key-block size limit in units of 16 keys (5 -> 80-column S buffers, 6 -> 96)
template <bool kF16, int kMaxUnits>
__global__ void __launch_bounds__(kThreads, 1)
attention_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                       const __grid_constant__ AttnParams prm) {
  constexpr int kSBufCols = 16 * kMaxUnits;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t q_s = base;                                    // Q ring: group g, slot s at (2g + s) * 16 KB
  const uint32_t kv_s = base + 4 * kQTileBytes;                 // stage s: K at kv_s + s*2*kv_bytes, V right behind
  const uint32_t bar = kv_s + uint32_t(prm.kv_stages) * 2u * uint32_t(prm.kv_bytes);
  const uint32_t bar_off = bar - base;
  auto k_full = [&](int s) { return bar + 8u * s; };
  auto v_full = [&](int s) { return bar + 16 + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 32 + 8u * s; };
  auto o_full = [&](int g) { return bar + 48 + 8u * g; };
  auto o_free = [&](int g) { return bar + 64 + 8u * g; };
  auto q_full = [&](int g, int s) { return bar + 80 + 8u * (2 * g + s); };
  auto q_empty = [&](int g, int s) { return bar + 112 + 8u * (2 * g + s); };
  auto s_full = [&](int g, int s) { return bar + 144 + 8u * (2 * g + s); };
  auto sm_done = [&](int g, int s) { return bar + 176 + 8u * (2 * g + s); };
  auto pv_done = [&](int g, int s) { return bar + 208 + 8u * (2 * g + s); };
  const uint32_t tmem_slot = bar + 240;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + bar_off + 240);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = prm.T, H = prm.H, nq = prm.nq, nb = prm.nb, items = prm.items;
  const int d = H * kHd;

  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tm_kv);
  if (warp == 3 && lane == 0) ptx::prefetch_tensormap(&tm_q);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(k_full(s), 1); ptx::mbar_init(v_full(s), 1);
      ptx::mbar_init(kv_empty(s), nq > 1 ? 2 : 1);   // one tcgen05.commit per MMA issuer that reads the stage (single-tile
                                                      // sequences: the item belongs to one group)
      ptx::mbar_init(o_full(s), 1);      // tcgen05.commit
      ptx::mbar_init(o_free(s), 128);    // every thread of the group
      for (int b = 0; b < 2; ++b) {
        ptx::mbar_init(q_full(s, b), 1); ptx::mbar_init(q_empty(s, b), 1);
        ptx::mbar_init(s_full(s, b), 1);      // tcgen05.commit
        ptx::mbar_init(sm_done(s, b), 128);   // every thread of the group
        ptx::mbar_init(pv_done(s, b), 1);     // tcgen05.commit
      }
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_items = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  // Tile t of the CTA's n-th item belongs to softmax group (t & 1) ^ (n & nq & 1): with an odd tile count the group that gets
  // the extra tile alternates from item to item (ViT-L/14: 5 tiles; text tower: 1 tile, so both groups work at all).
  auto first_tile = [&](int g, int n) { return g ^ (n & nq & 1); };
  auto blk_units = [&](int j) { return prm.bq + (j < prm.brem ? 1 : 0); };
  auto blk_start = [&](int j) { return j * prm.bq + min(j, prm.brem); };   // in units of 16 keys

  // register budget: the producer / issuer warp group gives registers to the two softmax warp groups
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ======================= K/V producer =======================
    if (ptx::elect_one()) {
      int n = 0;
      const uint32_t bytes = uint32_t(prm.kv_loads * prm.kv_box) * 128u;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const int b = item / H, h = item - b * H;
        const int row0 = b * T;
        const int s = n % prm.kv_stages;
        const uint32_t par = uint32_t(n / prm.kv_stages) & 1u;
        const uint32_t ks = kv_s + uint32_t(s) * 2u * uint32_t(prm.kv_bytes), vs = ks + uint32_t(prm.kv_bytes);
        ptx::mbar_wait(kv_empty(s), par ^ 1u);
        ptx::mbar_arrive_expect_tx(k_full(s), bytes);
        for (int l = 0; l < prm.kv_loads; ++l)
          ptx::tma_load_2d(&tm_kv, k_full(s), ks + uint32_t(l * prm.kv_box) * 128u, d + h * kHd, row0 + l * prm.kv_box,
                           ptx::kEvictFirst);
        ptx::mbar_arrive_expect_tx(v_full(s), bytes);
        for (int l = 0; l < prm.kv_loads; ++l)
          ptx::tma_load_2d(&tm_kv, v_full(s), vs + uint32_t(l * prm.kv_box) * 128u, 2 * d + h * kHd, row0 + l * prm.kv_box,
                           ptx::kEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ======================= Q producer: tile t of every item goes to group t & 1 (2-deep ring per group) ===========
    if (ptx::elect_one()) {
      uint32_t cnt[2] = {0, 0};
      int n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const int b = item / H, h = item - b * H;
        for (int t = 0; t < nq; ++t) {
          const int g = (t & 1) ^ (n & nq & 1), slot = int(cnt[g] & 1u);
          ptx::mbar_wait(q_empty(g, slot), ((cnt[g] >> 1) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(q_full(g, slot), kQTileBytes);
          ptx::tma_load_2d(&tm_q, q_full(g, slot), q_s + uint32_t(2 * g + slot) * kQTileBytes, h * kHd, b * T + t * 128,
                           ptx::kEvictFirst);
          ++cnt[g];
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ======================= MMA issuers: warp 1 serves softmax group 0, warp 2 group 1 =======================
    // The whole warp runs the loop (so that every operand stays in uniform registers); one elected lane issues.
    // Two cursors walk the group's (item, tile, block) sequence: s_* = next S to issue, p_* = next P.V to issue.
    // S op k goes to S buffer k & 1 and may be issued once P.V of op k-2 has been issued (in-order tensor pipe), so
    // the tensor core runs up to two ops ahead of the softmax, across tiles and items.
    const int g = warp - 1;
    uint32_t total_ops = 0;
    for (int n = 0; n < n_items; ++n) {
      const int t0 = first_tile(g, n);
      if (t0 < nq) total_ops += uint32_t((nq - t0 + 1) / 2) * uint32_t(nb);
    }
    if (total_ops > 0) {
      const uint32_t region = tmem_base + uint32_t(g * kRegionCols);
      const uint32_t idesc_o = idesc(128, kHd, kF16, true);
      // cursor = (item, tile, block); items in which this group has no tile (single-tile sequences) are skipped
      auto next_item = [&](int& n, int& t) {
        do { ++n; t = first_tile(g, n); } while (n < n_items && t >= nq);
      };
      int s_n = -1, s_t = 0, s_j = 0, p_n = -1, p_t = 0, p_j = 0;
      next_item(s_n, s_t);
      next_item(p_n, p_t);
      uint32_t s_k = 0, p_k = 0, s_tiles = 0, p_tiles = 0;
      long long t_last = clock64();
      while (p_k < total_ops) {
        bool progress = false;
        // ---- O (+)= P V of the oldest pending op ----
        if (p_k < s_k) {
          const int buf = int(p_k & 1u), stage = p_n % prm.kv_stages;
          bool ready = ptx::mbar_test_wait(sm_done(g, buf), (p_k >> 1) & 1u) &&
                       ptx::mbar_test_wait(v_full(stage), uint32_t(p_n / prm.kv_stages) & 1u);
          if (p_j == 0 && p_tiles > 0) ready = ready && ptx::mbar_test_wait(o_free(g), (p_tiles - 1u) & 1u);
          if (__all_sync(0xffffffffu, ready)) {
            ptx::tcgen05_fence_after();
            const uint32_t vs = kv_s + uint32_t(stage) * 2u * uint32_t(prm.kv_bytes) + uint32_t(prm.kv_bytes);
            const int u = blk_units(p_j), u0 = blk_start(p_j);
            // V rows are keys: 16 keys per K-step = 2048 bytes further down the [key][64 dims] tile (MN-major operand,
            // 128B swizzle: 8-key groups 1024 bytes apart (SBO), one 64-wide atom along N (LBO unused))
            uint64_t dv = uint64_t(((vs + uint32_t(u0) * 2048u) >> 4) & 0x3FFFu);
            dv |= uint64_t(1) << 16;
            dv |= uint64_t(1024 >> 4) << 32;
            dv |= uint64_t(1) << 46;
            dv |= uint64_t(2) << 61;
            const uint32_t pa = region + uint32_t(buf * kSBufCols);
            const bool last_blk = p_j == nb - 1, last_tile = p_t + 2 >= nq;
            if (ptx::elect_one()) {
              for (int i = 0; i < u; ++i)
                umma_f16_ts(region + kOCol, pa + uint32_t(8 * i), dv + uint64_t(i * 128), idesc_o, (p_j | i) != 0 ? 1u : 0u);
              ptx::umma_commit<1>(pv_done(g, buf));
              if (last_blk) {
                ptx::umma_commit<1>(o_full(g));
                if (last_tile) ptx::umma_commit<1>(kv_empty(stage));   // this group's share of the item's K/V reads
              }
            }
            __syncwarp();
            ++p_k;
            if (last_blk) {
              ++p_tiles;
              p_j = 0;
              p_t += 2;
              if (p_t >= nq) next_item(p_n, p_t);
            } else {
              ++p_j;
            }
            progress = true;
          }
        }
        // ---- S = Q K_j^T of the next op ----
        if (s_k < total_ops && s_k < p_k + 2u) {
          const int buf = int(s_k & 1u), stage = s_n % prm.kv_stages, slot = int(s_tiles & 1u);
          bool ready = true;
          if (s_j == 0)
            ready = ptx::mbar_test_wait(q_full(g, slot), (s_tiles >> 1) & 1u) &&
                    ptx::mbar_test_wait(k_full(stage), uint32_t(s_n / prm.kv_stages) & 1u);
          if (__all_sync(0xffffffffu, ready)) {
            ptx::tcgen05_fence_after();
            const uint32_t ks = kv_s + uint32_t(stage) * 2u * uint32_t(prm.kv_bytes);
            const int u = blk_units(s_j), u0 = blk_start(s_j);
            const uint64_t dq = ptx::make_kmajor_sw128_desc(q_s + uint32_t(2 * g + slot) * kQTileBytes);
            const uint64_t dk = ptx::make_kmajor_sw128_desc(ks + uint32_t(u0) * 2048u);
            const uint32_t id = idesc(128, uint32_t(16 * u), kF16, false);
            const bool last_blk = s_j == nb - 1;
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16<1>(region + uint32_t(buf * kSBufCols), dq + uint64_t(2 * kk), dk + uint64_t(2 * kk), id, kk != 0 ? 1u : 0u);
              ptx::umma_commit<1>(s_full(g, buf));
              if (last_blk) ptx::umma_commit<1>(q_empty(g, slot));   // last product that reads this Q tile
            }
            __syncwarp();
            ++s_k;
            if (last_blk) {
              ++s_tiles;
              s_j = 0;
              s_t += 2;
              if (s_t >= nq) next_item(s_n, s_t);
            } else {
              ++s_j;
            }
            progress = true;
          }
        }
        if (progress) {
          t_last = clock64();
        } else if ((clock64() - t_last) > IIC_MBAR_TIMEOUT_CYCLES) {
          // (a __nanosleep back-off of the idle issuer was measured: no effect on the kernel or on the step)
          if (lane == 0)
            printf("iic: attention MMA issuer stalled (block %d group %d: S ops %u, P.V ops %u of %u)\n", int(blockIdx.x), g, s_k, p_k, total_ops);
          __trap();
        }
      }
    }
  }
  } else {
    // ======================= softmax / output: one thread per query row =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                 // row inside the 128-row query tile
    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(g * kRegionCols);
    const float c = prm.scale_log2e;
    uint32_t k = 0, o_cnt = 0;   // ops / tiles processed by this group so far
    // The epilogue of a tile (wait for its last P.V, read O, normalise, store) is deferred until the FIRST block of the
    // group's next tile has been processed: the tensor-core round trip of the last P.V hides under those exponentials.
    struct Pending {
      bool valid, live;
      int b, h, t;
      float m;
      uint64_t acc0, acc1;
    } pend;
    pend.valid = false;
    auto epilogue = [&](const Pending& pd) {
      ptx::mbar_wait(o_full(g), o_cnt & 1u);
      ++o_cnt;
      ptx::tcgen05_fence_after();
      uint32_t v[64];
      if (pd.live) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld16_at(taddr + uint32_t(kOCol + 16 * cc), v, 16 * cc);
        ptx::tmem_ld_wait();
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(o_free(g));
      const int q = pd.t * 128 + r;
      if (q < T) {
        float s0, s1, s2, s3;
        unpack2(pd.acc0, s0, s1);
        unpack2(pd.acc1, s2, s3);
        const float sum = (s0 + s1) + (s2 + s3);
        const float inv = 1.0f / sum;
        const uint64_t inv2 = pack2(inv, inv);
        uint16_t* orow = prm.out + (size_t(pd.b) * T + q) * d + pd.h * kHd;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x0, x1;
            unpack2(mul2f(pack2(__uint_as_float(v[8 * jj + 2 * e]), __uint_as_float(v[8 * jj + 2 * e + 1])), inv2), x0, x1);
            w[e] = Act<kF16>::pack(x0, x1);
          }
          *reinterpret_cast<uint4*>(orow + 8 * jj) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (prm.lse != nullptr) prm.lse[(size_t(pd.b) * H + pd.h) * T + q] = pd.m + log2f(sum);
      }
    };
    {
      int n_loc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n_loc) {
        const int b = item / H, h = item - b * H;
        for (int t = first_tile(g, n_loc); t < nq; t += 2) {
          const bool warp_live = t * 128 + quad * 32 < T;   // warps whose 32 query rows are all padding do no math
          SoftmaxRow st;
          st.m = 0.f;
          st.acc0 = 0ull;
          st.acc1 = 0ull;
          for (int j = 0; j < nb; ++j, ++k) {
            const int buf = int(k & 1u);
            const uint32_t sb = taddr + uint32_t(buf * kSBufCols);
            ptx::mbar_wait(s_full(g, buf), (k >> 1) & 1u);
            ptx::tcgen05_fence_after();
            if (warp_live) {
              const int u = blk_units(j);
              // columns of the block this row may see: real keys (< T) and, under a causal mask, keys up to the query itself
              const int k0 = 16 * blk_start(j);
              const int nvalid = (prm.causal ? min(T, t * 128 + r + 1) : T) - k0;
              const bool causal = prm.causal != 0;
              const uint32_t pv_bar = pv_done(g, buf ^ 1), pv_par = ((k - 1u) >> 1) & 1u;
              const uint32_t o_addr = taddr + uint32_t(kOCol);
              switch (u) {
                case 1: softmax_block<kF16, 1>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par); break;
                case 2: softmax_block<kF16, 2>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par); break;
                case 3: softmax_block<kF16, 3>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par); break;
                case 4: softmax_block<kF16, 4>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par); break;
                case 5: softmax_block<kF16, 5>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par); break;
                default:
                  if constexpr (kMaxUnits >= 6) softmax_block<kF16, 6>(sb, o_addr, st, c, j == 0, nvalid, causal, pv_bar, pv_par);
                  break;
              }
            }
            ptx::tcgen05_fence_before();
            ptx::mbar_arrive(sm_done(g, buf));
            if (j == 0 && pend.valid) {
              epilogue(pend);
              pend.valid = false;
            }
          }
          pend.valid = true; pend.live = warp_live; pend.b = b; pend.h = h; pend.t = t;
          pend.m = st.m; pend.acc0 = st.acc0; pend.acc1 = st.acc1;
        }
      }
      if (pend.valid) epilogue(pend);
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

namespace {
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
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
bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f16) {
  auto fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  return fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr,
            box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// returns -3 when the shape is outside this kernel's envelope (caller falls back to the mma.sync kernel)
int launch_attention_sm100(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16, bool causal,
                           int num_sms, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1) return -3;
  AttnParams p;
  p.TP = (T + 15) / 16 * 16;
  const int units = p.TP / 16;
  int max_units = 5;   // 80-key blocks: S row + packed P stay in registers without spills (96 spills in the hot loop)
  if (const char* e = getenv("IIC_ATTN_MAXU")) max_units = atoi(e) == 6 ? 6 : 5;
  p.nb = (units + max_units - 1) / max_units;
  p.bq = units / p.nb;
  p.brem = units % p.nb;
  // K/V arrive in TMA boxes of 16*dd rows; pick the box (<= 256 rows) that overshoots TP the least
  int best_d = 1, best_rows = 1 << 30;
  for (int dd = 16; dd >= 4; --dd) {
    const int rows = (units + dd - 1) / dd * dd;
    if (rows < best_rows) { best_rows = rows; best_d = dd; }
  }
  if (units <= 16) { best_d = units; best_rows = units; }
  p.kv_box = 16 * best_d;
  p.kv_loads = best_rows / best_d;
  p.kv_bytes = best_rows * 16 * 128;
  const int fixed = 4 * kQTileBytes + kBarBytes + 1024 /*alignment slack*/;
  if (fixed + 2 * p.kv_bytes > kMaxSmem) return -3;
  p.kv_stages = fixed + 4 * p.kv_bytes <= kMaxSmem ? 2 : 1;
  const int smem = fixed + p.kv_stages * 2 * p.kv_bytes;
  p.nq = (T + 127) / 128;
  p.items = B * H;
  p.T = T;
  p.H = H;
  p.out = static_cast<uint16_t*>(out);
  p.lse = lse;
  p.scale_log2e = 1.4426950408889634f / sqrtf(float(head_dim));
  p.causal = causal ? 1 : 0;
  const int d = H * kHd;
  CUtensorMap tq, tkv;
  const uint64_t rows = uint64_t(B) * T;
  if (!make_map(&tq, qkv, rows, uint64_t(3 * d), 128, f16 != 0) || !make_map(&tkv, qkv, rows, uint64_t(3 * d), uint32_t(p.kv_box), f16 != 0))
    return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_sm100_kernel<false, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_sm100_kernel<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_sm100_kernel<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_sm100_kernel<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
  if (max_units == 6) {
    if (f16) attention_sm100_kernel<true, 6><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
    else attention_sm100_kernel<false, 6><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
  } else {
    if (f16) attention_sm100_kernel<true, 5><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
    else attention_sm100_kernel<false, 5><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
