// tcgen05 attention, head_dim 64, any ViT sequence length that fits one SM's shared memory (T <= ~760):
//
//   softmax(Q K^T / sqrt(64)) V   per (image, head)        [reference: nn.MultiheadAttention inside CLIP's
//   ResidualAttentionBlock, reached through model.encode_image at /root/reference/main.py:204, 444, 503]
//
// One persistent CTA per SM walks (image, head) items; K and V of the item are resident in 128B-swizzled smem (two items
// in flight when they fit).  The queries form 128-row tiles; two softmax groups (4 warps each) take alternate tiles, so
// one group's tensor-core round trip hides under the other group's exponentials.  Per tile, with the keys cut into
// nb blocks of Nb <= 192 (a single block of <= 256 when the whole row fits: ViT-B/16 @ 224, T = 197):
//
//   pass A (nb > 1 only)   S_j = Q K_j^T -> TMEM, row maximum only.  The scores are simply computed twice (the tensor
//                          pipe is idle most of the time) so that pass B needs no running maximum and no rescaling
//                          of O: the numerics are those of a plain two-pass softmax for every T.
//   pass B                 S_j = Q K_j^T -> TMEM (fp32, columns [0, Nj) of the group's 256-column region);
//                          ONE THREAD PER QUERY ROW: tcgen05.ld 32x32b hands a thread its own row, so max and sum are
//                          thread-local (no shuffles); p = exp2(s*c - m*c) with packed FFMA2/FADD2, converted to 16 bit
//                          and written back IN PLACE over S (tcgen05.st, columns [0, Nj/2)): P never touches smem;
//                          O += P_j V_j with A = P from TMEM, B = V in its natural [key][dim] layout (MN-major operand),
//                          accumulator in columns [192, 256) of the region.
//   epilogue               O row * 1/sum -> 16 bit -> global; optional log2-domain LSE for the backward pass.
//
//   warp 0  TMA producer K/V      warp 1  MMA issuer (one thread, event loop over both groups)
//   warp 2  TMEM allocator        warp 3  TMA producer Q tiles      warps 4-7 / 8-11  softmax groups 0 / 1
//
// The softmax is bound by the MUFU pipe (16 ex2/clk/SM, tools/ubench_sm100.cu); everything else in the inner loop is
// ~2 issue slots per element.  TMEM reads are not a limit (>= 700 B/clk/SM measured), which is why S is read twice
// rather than kept in registers.  tcgen05.mma instructions of one CTA execute in issue order, which is what orders
// "PV_j reads P_j" before "S_{j+1} overwrites it".  All waits are bounded (trap, never hang).
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
constexpr int kRegionCols = 256;         // TMEM columns per softmax group
constexpr int kOCol = 192;               // O accumulator: columns [192, 256) of the region
constexpr int kMaxSmem = 227 * 1024;
constexpr int kBarBytes = 256;

struct AttnParams {
  uint16_t* out;
  float* lse;
  int items, T, H;
  int nq;         // 128-row query tiles per item
  int nb, Nb;     // key blocks per item, rows per block (multiple of 16)
  int TP;         // keys rounded up to 16
  int kv_stages;  // items whose K/V are resident at once (1 or 2)
  int kv_bytes;   // bytes of one K (or V) buffer = nb * Nb * 128
  float scale_log2e;
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

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
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

// running maximum over the first `valid` of the 32 columns in v (valid >= 32: no masking)
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32], int valid, float m) {
  if (valid >= 32) {
    float a = m, b = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      a = max3(a, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      b = max3(b, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
    }
    return fmaxf(a, b);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) m = fmaxf(m, i < valid ? __uint_as_float(v[i]) : -INFINITY);
  return m;
}

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                       const __grid_constant__ AttnParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t q_s = base;                                    // two Q tiles (one per group)
  const uint32_t kv_s = base + 2 * kQTileBytes;                 // stage s: K at kv_s + s*2*kv_bytes, V right behind
  const uint32_t bar = kv_s + uint32_t(prm.kv_stages) * 2u * uint32_t(prm.kv_bytes);
  const uint32_t bar_off = bar - base;
  auto k_full = [&](int s) { return bar + 8u * s; };
  auto v_full = [&](int s) { return bar + 16 + 8u * s; };
  auto kv_empty = [&](int s) { return bar + 32 + 8u * s; };
  auto q_full = [&](int g) { return bar + 48 + 8u * g; };
  auto q_empty = [&](int g) { return bar + 64 + 8u * g; };
  auto s_full = [&](int g) { return bar + 80 + 8u * g; };
  auto sm_done = [&](int g) { return bar + 96 + 8u * g; };
  auto o_full = [&](int g) { return bar + 112 + 8u * g; };
  auto o_free = [&](int g) { return bar + 128 + 8u * g; };
  const uint32_t tmem_slot = bar + 144;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + bar_off + 144);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = prm.T, H = prm.H, nq = prm.nq, nb = prm.nb, Nb = prm.Nb, TP = prm.TP, items = prm.items;
  const int d = H * kHd;

  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tm_kv);
  if (warp == 3 && lane == 0) ptx::prefetch_tensormap(&tm_q);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(k_full(s), 1); ptx::mbar_init(v_full(s), 1); ptx::mbar_init(kv_empty(s), 1);
      ptx::mbar_init(q_full(s), 1); ptx::mbar_init(q_empty(s), 1);
      ptx::mbar_init(s_full(s), 1);      // tcgen05.commit
      ptx::mbar_init(sm_done(s), 128);   // every thread of the group
      ptx::mbar_init(o_full(s), 1);      // tcgen05.commit
      ptx::mbar_init(o_free(s), 128);    // every thread of the group
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_items = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int first_pass = nb > 1 ? 0 : 1;

  if (warp == 0) {
    // ======================= K/V producer =======================
    if (ptx::elect_one()) {
      int n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const int b = item / H, h = item - b * H;
        const int row0 = b * T;
        const int s = n % prm.kv_stages;
        const uint32_t par = uint32_t(n / prm.kv_stages) & 1u;
        const uint32_t ks = kv_s + uint32_t(s) * 2u * uint32_t(prm.kv_bytes), vs = ks + uint32_t(prm.kv_bytes);
        ptx::mbar_wait(kv_empty(s), par ^ 1u);
        ptx::mbar_arrive_expect_tx(k_full(s), uint32_t(prm.kv_bytes));
        for (int j = 0; j < nb; ++j)
          ptx::tma_load_2d(&tm_kv, k_full(s), ks + uint32_t(j * Nb) * 128u, d + h * kHd, row0 + j * Nb, ptx::kEvictFirst);
        ptx::mbar_arrive_expect_tx(v_full(s), uint32_t(prm.kv_bytes));
        for (int j = 0; j < nb; ++j)
          ptx::tma_load_2d(&tm_kv, v_full(s), vs + uint32_t(j * Nb) * 128u, 2 * d + h * kHd, row0 + j * Nb, ptx::kEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ======================= Q producer: tile t of every item goes to group t & 1 =======================
    if (ptx::elect_one()) {
      uint32_t cnt[2] = {0, 0};
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        for (int t = 0; t < nq; ++t) {
          const int g = t & 1;
          ptx::mbar_wait(q_empty(g), (cnt[g] & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(q_full(g), kQTileBytes);
          ptx::tma_load_2d(&tm_q, q_full(g), q_s + g * kQTileBytes, h * kHd, b * T + t * 128, ptx::kEvictFirst);
          ++cnt[g];
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer: one thread, a small event loop over both groups =======================
    if (ptx::elect_one()) {
      const uint32_t idesc_o = idesc(128, kHd, kF16, true);
      struct GState {
        int n, t, pass, j, st;
        uint32_t q_cnt, smd_cnt, o_cnt;
        bool active;
      } gs[2];
      for (int g = 0; g < 2; ++g) gs[g] = {0, g, first_pass, 0, 0, 0u, 0u, 0u, g < nq && n_items > 0};
      int tiles_done[2] = {0, 0};
      long long t_last = clock64();
      while (gs[0].active || gs[1].active) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          GState& s = gs[g];
          if (!s.active) continue;
          const int stage = s.n % prm.kv_stages;
          const uint32_t kv_par = uint32_t(s.n / prm.kv_stages) & 1u;
          const uint32_t ks = kv_s + uint32_t(stage) * 2u * uint32_t(prm.kv_bytes), vs = ks + uint32_t(prm.kv_bytes);
          const uint32_t region = tmem_base + uint32_t(g * kRegionCols);
          const int Nj = min(Nb, TP - s.j * Nb);
          if (s.st == 0) {
            // ---- S_j = Q K_j^T ----
            if (s.j == 0 && s.pass == first_pass) {
              if (!ptx::mbar_test_wait(q_full(g), s.q_cnt & 1u)) continue;
              if (!ptx::mbar_test_wait(k_full(stage), kv_par)) continue;
              // a single-block S wider than 192 columns overlaps the O accumulator of the previous tile
              if (nb == 1 && TP > kOCol && s.o_cnt > 0 && !ptx::mbar_test_wait(o_free(g), (s.o_cnt - 1u) & 1u)) continue;
            }
            ptx::tcgen05_fence_after();
            const uint64_t dq = ptx::make_kmajor_sw128_desc(q_s + g * kQTileBytes);
            const uint64_t dk = ptx::make_kmajor_sw128_desc(ks + uint32_t(s.j * Nb) * 128u);
            const uint32_t id = idesc(128, uint32_t(Nj), kF16, false);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(region, dq + uint64_t(2 * k), dk + uint64_t(2 * k), id, k != 0 ? 1u : 0u);
            ptx::umma_commit<1>(s_full(g));
            if (s.pass == 1 && s.j == nb - 1) {   // last product that reads this Q tile
              ptx::umma_commit<1>(q_empty(g));
              ++s.q_cnt;
            }
            s.st = 1;
            progress = true;
          } else {
            if (!ptx::mbar_test_wait(sm_done(g), s.smd_cnt & 1u)) continue;
            if (s.pass == 0) {
              // the group has taken the block's row maxima: next block (or start pass B)
              ++s.smd_cnt;
              if (++s.j == nb) { s.pass = 1; s.j = 0; }
              s.st = 0;
              progress = true;
              continue;
            }
            // ---- O (+)= P_j V_j ----
            if (!ptx::mbar_test_wait(v_full(stage), kv_par)) continue;
            if (s.j == 0 && s.o_cnt > 0 && !ptx::mbar_test_wait(o_free(g), (s.o_cnt - 1u) & 1u)) continue;
            ++s.smd_cnt;
            ptx::tcgen05_fence_after();
            const int n_kstep = Nj >> 4;
            for (int k = 0; k < n_kstep; ++k) {
              // V rows are keys: 16 keys per K-step = 2048 bytes further down the [key][64 dims] tile (MN-major operand,
              // 128B swizzle: 8-key groups 1024 bytes apart (SBO), one 64-wide atom along N (LBO unused))
              uint64_t dv = uint64_t(((vs + uint32_t(s.j * Nb) * 128u + uint32_t(k) * 2048u) >> 4) & 0x3FFFu);
              dv |= uint64_t(1) << 16;
              dv |= uint64_t(1024 >> 4) << 32;
              dv |= uint64_t(1) << 46;
              dv |= uint64_t(2) << 61;
              umma_f16_ts(region + kOCol, region + uint32_t(8 * k), dv, idesc_o, (s.j | k) != 0 ? 1u : 0u);
            }
            if (s.j == nb - 1) {
              ptx::umma_commit<1>(o_full(g));
              ++s.o_cnt;
              if (++tiles_done[stage] == nq) {   // every product that reads this item's K/V has been issued
                ptx::umma_commit<1>(kv_empty(stage));
                tiles_done[stage] = 0;
              }
            }
            s.st = 0;
            if (++s.j == nb) {
              s.j = 0;
              s.pass = first_pass;
              s.t += 2;
              if (s.t >= nq) {
                s.t = g;
                if (++s.n >= n_items) s.active = false;
              }
            }
            progress = true;
          }
        }
        if (progress) {
          t_last = clock64();
        } else if ((clock64() - t_last) > IIC_MBAR_TIMEOUT_CYCLES) {
          printf("iic: attention MMA scheduler stalled (block %d)\n", int(blockIdx.x));
          __trap();
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================= softmax / output: one thread per query row =======================
    const int g = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                 // row inside the 128-row query tile
    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(g * kRegionCols);
    const float c = prm.scale_log2e;
    const uint64_t c2 = pack2(c, c);
    uint32_t s_cnt = 0, o_cnt = 0;
    if (g < nq) {
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        for (int t = g; t < nq; t += 2) {
          const bool warp_live = t * 128 + quad * 32 < T;   // warps whose 32 query rows are all padding do no math
          float m = -INFINITY;
          uint32_t va[32], vb[32];
          // ---------------- pass A: row maximum over all key blocks ----------------
          if (nb > 1) {
            for (int j = 0; j < nb; ++j) {
              ptx::mbar_wait(s_full(g), s_cnt & 1u);
              ++s_cnt;
              ptx::tcgen05_fence_after();
              if (warp_live) {
                const int Nj = min(Nb, TP - j * Nb), valid = min(Nj, T - j * Nb);
                const int n_ch = (Nj + 31) >> 5;
                ptx::tmem_ld_32x32b_x32(taddr, va);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int ch = 0; ch < 6; ch += 2) {
                  if (ch < n_ch) {
                    if (ch + 1 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 1) * 32), vb);
                    m = chunk_max(va, valid - ch * 32, m);
                    ptx::tmem_ld_wait();
                  }
                  if (ch + 1 < n_ch) {
                    if (ch + 2 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 2) * 32), va);
                    m = chunk_max(vb, valid - (ch + 1) * 32, m);
                    ptx::tmem_ld_wait();
                  }
                }
              }
              ptx::tcgen05_fence_before();
              ptx::mbar_arrive(sm_done(g));
            }
          }
          // ---------------- pass B: P = exp2(S*c - m*c) in place, row sum ----------------
          uint64_t acc0 = 0ull, acc1 = 0ull;   // packed partial row sums (0.0f bit patterns)
          float mc = m * c;
          for (int j = 0; j < nb; ++j) {
            ptx::mbar_wait(s_full(g), s_cnt & 1u);
            ++s_cnt;
            ptx::tcgen05_fence_after();
            if (warp_live) {
              const int Nj = min(Nb, TP - j * Nb), valid = min(Nj, T - j * Nb);
              const int n_ch = (Nj + 31) >> 5;
              if (nb == 1) {
                ptx::tmem_ld_32x32b_x32(taddr, va);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int ch = 0; ch < 8; ch += 2) {
                  if (ch < n_ch) {
                    if (ch + 1 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 1) * 32), vb);
                    m = chunk_max(va, valid - ch * 32, m);
                    ptx::tmem_ld_wait();
                  }
                  if (ch + 1 < n_ch) {
                    if (ch + 2 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 2) * 32), va);
                    m = chunk_max(vb, valid - (ch + 1) * 32, m);
                    ptx::tmem_ld_wait();
                  }
                }
                mc = m * c;
              }
              const uint64_t nmc2 = pack2(-mc, -mc);
              auto emit = [&](const uint32_t (&v)[32], int ch) {
                uint32_t pk[16];
                const int vl = valid - ch * 32;
                if (vl >= 32) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    float x0, x1;
                    unpack2(fma2(pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), c2, nmc2), x0, x1);
                    const float e0 = ex2(x0), e1 = ex2(x1);
                    if (i & 1) acc1 = add2(acc1, pack2(e0, e1)); else acc0 = add2(acc0, pack2(e0, e1));
                    pk[i] = Act<kF16>::pack(e0, e1);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    float e0 = ex2(fmaf(__uint_as_float(v[2 * i]), c, -mc)), e1 = ex2(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
                    e0 = 2 * i < vl ? e0 : 0.f;
                    e1 = 2 * i + 1 < vl ? e1 : 0.f;
                    acc0 = add2(acc0, pack2(e0, e1));
                    pk[i] = Act<kF16>::pack(e0, e1);
                  }
                }
                tmem_st16(taddr + uint32_t(ch * 16), pk);   // always behind the S columns still to be read
              };
              ptx::tmem_ld_32x32b_x32(taddr, va);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int ch = 0; ch < 8; ch += 2) {
                if (ch < n_ch) {
                  if (ch + 1 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 1) * 32), vb);
                  emit(va, ch);
                  ptx::tmem_ld_wait();
                }
                if (ch + 1 < n_ch) {
                  if (ch + 2 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((ch + 2) * 32), va);
                  emit(vb, ch + 1);
                  ptx::tmem_ld_wait();
                }
              }
              tmem_st_wait();
            }
            ptx::tcgen05_fence_before();
            ptx::mbar_arrive(sm_done(g));
          }
          // ---------------- O row ----------------
          ptx::mbar_wait(o_full(g), o_cnt & 1u);
          ++o_cnt;
          ptx::tcgen05_fence_after();
          const int q = t * 128 + r;
          if (warp_live) {
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(kOCol), va);
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(kOCol + 32), vb);
            ptx::tmem_ld_wait();
          }
          ptx::tcgen05_fence_before();
          ptx::mbar_arrive(o_free(g));
          if (q < T) {
            float s0, s1, s2, s3;
            unpack2(acc0, s0, s1);
            unpack2(acc1, s2, s3);
            const float sum = (s0 + s1) + (s2 + s3);
            const float inv = 1.0f / sum;
            const uint64_t inv2 = pack2(inv, inv);
            uint16_t* orow = prm.out + (size_t(b) * T + q) * d + h * kHd;
            auto store32 = [&](const uint32_t (&v)[32], int off) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float x0, x1;
                  unpack2(mul2f(pack2(__uint_as_float(v[8 * jj + 2 * e]), __uint_as_float(v[8 * jj + 2 * e + 1])), inv2), x0, x1);
                  w[e] = Act<kF16>::pack(x0, x1);
                }
                *reinterpret_cast<uint4*>(orow + off + 8 * jj) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            };
            store32(va, 0);
            store32(vb, 32);
            if (prm.lse != nullptr) prm.lse[(size_t(b) * H + h) * T + q] = mc + log2f(sum);
          }
        }
      }
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
int launch_attention_sm100(const void* qkv, void* out, float* lse, int B, int T, int H, int head_dim, int f16, int num_sms,
                           cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1) return -3;
  AttnParams p;
  p.TP = (T + 15) / 16 * 16;
  if (p.TP <= 256) {
    p.nb = 1;
    p.Nb = p.TP;
  } else {
    p.nb = (p.TP + kOCol - 1) / kOCol;
    p.Nb = ((p.TP + p.nb - 1) / p.nb + 15) / 16 * 16;
    p.nb = (p.TP + p.Nb - 1) / p.Nb;
  }
  p.kv_bytes = p.nb * p.Nb * 128;
  const int fixed = 2 * kQTileBytes + kBarBytes + 1024 /*alignment slack*/;
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
  const int d = H * kHd;
  CUtensorMap tq, tkv;
  const uint64_t rows = uint64_t(B) * T;
  if (!make_map(&tq, qkv, rows, uint64_t(3 * d), 128, f16 != 0) || !make_map(&tkv, qkv, rows, uint64_t(3 * d), uint32_t(p.Nb), f16 != 0))
    return -1;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attention_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done = true;
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
  if (f16)
    attention_sm100_kernel<true><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
  else
    attention_sm100_kernel<false><<<grid, kThreads, smem, stream>>>(tq, tkv, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
