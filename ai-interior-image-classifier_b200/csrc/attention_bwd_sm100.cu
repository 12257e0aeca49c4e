// tcgen05 attention backward, head_dim 64, T <= 592 (ViT-B/16 @ 224: T = 197; ViT-L/14 @ 336: T = 577):
//   (dO, saved Q/K/V, O, log-sum-exp) -> dQ, dK, dV, packed like qkv.
//
// Reference semantics: torch.autograd through nn.MultiheadAttention's softmax(Q K^T / sqrt(64)) V inside CLIP's
// ResidualAttentionBlock; needed only to carry dX through the frozen attention blocks down to the LoRA-adapted MLPs
// (/root/reference/train_lora.py:249 `loss.backward()` restricted to parameters named '*lora*').
//
// One persistent CTA per SM walks (image, head) items; every item is two passes, all products on tcgen05:
//
//   pass Q, query tile t:  S = Q_t K^T, dP = dO_t V^T  (fp32, TMEM)  ->  dS = P o (dP - D) / 8  ->  dQ_t = dS K
//   pass K, key tile u:    S^T = K_u Q^T, dP^T = V_u dO^T            ->  P^T, dS^T              ->  dV_u = P^T dO, dK_u = dS^T Q
//
// with P = exp2(S c - lse) recomputed from the saved log-sum-exp and D[q] = sum_d dO[q,d] O[q,d] from a small pre-kernel.
// Shared memory per pass: the two operands that span the whole sequence (X, Y = K, V in pass Q; Q, dO in pass K; TMA,
// 128B swizzle; the next pass is prefetched when two stages fit) and a 2-deep ring of 128-row tiles (Q_t, dO_t / K_u, V_u).
// X and Y serve both as K-major operands (S = Q_t K^T ...) and as MN-major operands (dQ = dS K ...): same bytes, different
// descriptors.  Sequences longer than 256 are cut into column blocks of <= 192 whose second products accumulate in TMEM.
// The backward softmax is purely elementwise (no row reductions), so the 8 compute warps split the COLUMNS of a block:
// warps 4-7 take the first half, warps 8-11 the second half of every row (a thread still owns one TMEM lane = one row).
// P / dS go back to TMEM as 16-bit A operands IN PLACE, each group packing into the start of ITS OWN fp32 columns (a K-step of
// the second product takes its A operand from any column, so the two halves need not be contiguous): a group only ever
// overwrites columns it has already read, and the groups never wait for each other.  The 64-column accumulators of the second
// products sit in the top columns [192, 256) of the two 256-column regions.
//
//   warp 0  TMA producer X/Y    warp 3  TMA producer tiles    warp 1  MMA issuer (whole warp, one elected lane issues)
//   warp 2  TMEM allocator      warps 4-11  compute
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "act_types.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace iic {

namespace {

constexpr int kHd = 64;
constexpr int kThreads = 384;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kBarBytes = 256;
constexpr int kTileBytes = 128 * 128;   // one 128-row tile
constexpr int kVecFloats = 640;         // lse / D of one item (T <= 592 rounded up)

struct BwdParams {
  const float* lse;   // [B*H, T] log2-domain log-sum-exp of the scaled scores
  const float* dsum;  // [B*H, T] D = rowsum(dO o O)
  uint16_t* dqkv;     // [B*T, 3d]
  int items, T, H, TP;
  int nblk, bq, brem;   // column blocks: block j holds bq + (j < brem) units of 16 columns
  int stages;           // (item, pass) units whose X / Y are resident at once (1 or 2)
  int xy_bytes;         // bytes of one X (or Y) buffer (TMA boxes may overshoot TP rows)
  int xy_box, xy_loads;
  float scale, scale_log2e;
  int causal;   // text tower: P[q, k] = 0 for k > q (the saved log-sum-exp already covers the visible keys only)
};

__host__ __device__ constexpr uint32_t idesc(uint32_t m, uint32_t n, bool f16, bool b_mn) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}
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
// MN-major operand over a [row][64 dims] 128B-swizzled tile: rows are the reduction index (16 per K-step = 2048 bytes),
// 8-row groups 1024 bytes apart (SBO), one 64-wide atom along N (LBO unused)
__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr) {
  uint64_t d = uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t(1) << 16;
  d |= uint64_t(1024 >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_qkv_full, const __grid_constant__ CUtensorMap tm_qkv_tile,
                           const __grid_constant__ CUtensorMap tm_do_full, const __grid_constant__ CUtensorMap tm_do_tile,
                           const __grid_constant__ BwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int T = prm.T, H = prm.H, items = prm.items;
  const int d = H * kHd;
  // smem: tile ring (2 slots x (A0, A1)), X / Y stages, per-item vectors, barriers
  auto ta_s = [&](int slot, int which) { return base + uint32_t(2 * slot + which) * kTileBytes; };
  const uint32_t xy_base = base + 4u * kTileBytes;
  auto x_s = [&](int s) { return xy_base + uint32_t(2 * s) * uint32_t(prm.xy_bytes); };
  auto y_s = [&](int s) { return xy_base + uint32_t(2 * s + 1) * uint32_t(prm.xy_bytes); };
  const uint32_t vec_off = 4u * kTileBytes + uint32_t(prm.stages) * 2u * uint32_t(prm.xy_bytes);
  float* s_lse = reinterpret_cast<float*>(gen + vec_off);
  float* s_d = s_lse + kVecFloats;
  const uint32_t bar_off = vec_off + 2u * kVecFloats * 4u;
  const uint32_t bar = base + bar_off;
  auto xy_full = [&](int s) { return bar + 8u * s; };
  auto xy_empty = [&](int s) { return bar + 16 + 8u * s; };
  auto tile_full = [&](int s) { return bar + 32 + 8u * s; };
  auto tile_empty = [&](int s) { return bar + 48 + 8u * s; };
  auto sd_full = [&](int g) { return bar + 64 + 8u * g; };
  auto ds_ready = [&](int g) { return bar + 80 + 8u * g; };
  const uint32_t acc_full = bar + 96, acc_free = bar + 104;
  const uint32_t tmem_slot = bar + 112;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + bar_off + 112u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv_full);
    ptx::prefetch_tensormap(&tm_do_full);
  }
  if (warp == 3 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv_tile);
    ptx::prefetch_tensormap(&tm_do_tile);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(xy_full(s), 1); ptx::mbar_init(xy_empty(s), 1);
      ptx::mbar_init(tile_full(s), 1); ptx::mbar_init(tile_empty(s), 1);
      ptx::mbar_init(sd_full(s), 1);       // tcgen05.commit
      ptx::mbar_init(ds_ready(s), 128);    // every thread of the column group
    }
    ptx::mbar_init(acc_full, 1);     // tcgen05.commit
    ptx::mbar_init(acc_free, 256);   // every compute thread
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int n_items = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int n_tiles = (T + 127) / 128;          // 128-row tiles per pass
  const int nblk = prm.nblk;
  auto blk_units = [&](int j) { return prm.bq + (j < prm.brem ? 1 : 0); };
  auto blk_start = [&](int j) { return j * prm.bq + min(j, prm.brem); };   // in units of 16 columns
  // TMEM columns: S / S^T block at [0, 16 u), dP / dP^T block at [256, 256 + 16 u).  Column group 0 owns the block's units
  // [0, ua), group 1 units [ua, u); the packed 16-bit unit v of a group sits at the group's first column + 8 * (v - first).
  // 64-column accumulators in the top columns of the regions: [192, 256) and [448, 512).
  constexpr uint32_t kColS = 0, kColDp = 256, kAcc0 = 192, kAcc1 = 448;
  auto packed_col = [&](int v, int ua) { return uint32_t(v < ua ? 8 * v : 16 * ua + 8 * (v - ua)); };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ======================= TMA producer: X / Y of every (item, pass) =======================
    if (ptx::elect_one()) {
      const uint32_t bytes = uint32_t(prm.xy_loads * prm.xy_box) * 128u;
      int m = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        for (int pass = 0; pass < 2; ++pass, ++m) {
          const int s = m % prm.stages;
          ptx::mbar_wait(xy_empty(s), (uint32_t(m / prm.stages) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(xy_full(s), 2u * bytes);
          for (int l = 0; l < prm.xy_loads; ++l) {
            const uint32_t off = uint32_t(l * prm.xy_box) * 128u;
            const int row = b * T + l * prm.xy_box;
            if (pass == 0) {   // X = K, Y = V
              ptx::tma_load_2d(&tm_qkv_full, xy_full(s), x_s(s) + off, d + h * kHd, row, ptx::kEvictFirst);
              ptx::tma_load_2d(&tm_qkv_full, xy_full(s), y_s(s) + off, 2 * d + h * kHd, row, ptx::kEvictFirst);
            } else {           // X = Q, Y = dO
              ptx::tma_load_2d(&tm_qkv_full, xy_full(s), x_s(s) + off, h * kHd, row, ptx::kEvictFirst);
              ptx::tma_load_2d(&tm_do_full, xy_full(s), y_s(s) + off, h * kHd, row, ptx::kEvictFirst);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ======================= TMA producer: 128-row tiles (A operands of the first products) =======================
    if (ptx::elect_one()) {
      uint32_t tc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        for (int pass = 0; pass < 2; ++pass) {
          for (int t = 0; t < n_tiles; ++t, ++tc) {
            const int slot = int(tc & 1u);
            ptx::mbar_wait(tile_empty(slot), ((tc >> 1) & 1u) ^ 1u);
            ptx::mbar_arrive_expect_tx(tile_full(slot), 2u * kTileBytes);
            const int row = b * T + t * 128;
            if (pass == 0) {   // Q_t, dO_t
              ptx::tma_load_2d(&tm_qkv_tile, tile_full(slot), ta_s(slot, 0), h * kHd, row, ptx::kEvictFirst);
              ptx::tma_load_2d(&tm_do_tile, tile_full(slot), ta_s(slot, 1), h * kHd, row, ptx::kEvictFirst);
            } else {           // K_u, V_u
              ptx::tma_load_2d(&tm_qkv_tile, tile_full(slot), ta_s(slot, 0), d + h * kHd, row, ptx::kEvictFirst);
              ptx::tma_load_2d(&tm_qkv_tile, tile_full(slot), ta_s(slot, 1), 2 * d + h * kHd, row, ptx::kEvictFirst);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t id_acc = idesc(128, kHd, kF16, true);
    uint32_t tile_cnt = 0, blk_cnt = 0;   // phases of acc_full / acc_free / tile ring, and of sd_full / ds_ready
    int m = 0;
    for (int n = 0; n < n_items; ++n) {
      for (int pass = 0; pass < 2; ++pass, ++m) {
        const int s = m % prm.stages;
        ptx::mbar_wait(xy_full(s), uint32_t(m / prm.stages) & 1u);
        for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
          const int slot = int(tile_cnt & 1u);
          ptx::mbar_wait(tile_full(slot), (tile_cnt >> 1) & 1u);
          // a single block wider than 192 columns overlaps the accumulators of the previous tile: they must have been read
          // out before S / dP are written; with several blocks (<= 192 columns) only the first accumulating product must wait
          bool acc_waited = tile_cnt == 0;
          if (!acc_waited && nblk == 1) { ptx::mbar_wait(acc_free, (tile_cnt - 1u) & 1u); acc_waited = true; }
          const uint64_t da0 = ptx::make_kmajor_sw128_desc(ta_s(slot, 0)), da1 = ptx::make_kmajor_sw128_desc(ta_s(slot, 1));
          for (int j = 0; j < nblk; ++j, ++blk_cnt) {
            const int u = blk_units(j), u0 = blk_start(j), ua = (u + 1) / 2;
            ptx::tcgen05_fence_after();
            // first products over the block's columns = rows [16 u0, 16 (u0 + u)) of X / Y, as two halves with their own barriers
            const uint64_t db0 = ptx::make_kmajor_sw128_desc(x_s(s) + uint32_t(u0) * 2048u);
            const uint64_t db1 = ptx::make_kmajor_sw128_desc(y_s(s) + uint32_t(u0) * 2048u);
            const uint32_t id_h0 = idesc(128, uint32_t(16 * ua), kF16, false), id_h1 = idesc(128, uint32_t(16 * (u - ua)), kF16, false);
            const uint64_t boff = uint64_t((uint32_t(ua) * 2048u) >> 4);
            if (ptx::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16<1>(tmem_base + kColS, da0 + uint64_t(2 * kk), db0 + uint64_t(2 * kk), id_h0, kk != 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16<1>(tmem_base + kColDp, da1 + uint64_t(2 * kk), db1 + uint64_t(2 * kk), id_h0, kk != 0 ? 1u : 0u);
              ptx::umma_commit<1>(sd_full(0));
              if (u > ua) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16<1>(tmem_base + kColS + uint32_t(16 * ua), da0 + uint64_t(2 * kk), db0 + boff + uint64_t(2 * kk), id_h1, kk != 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  ptx::umma_f16<1>(tmem_base + kColDp + uint32_t(16 * ua), da1 + uint64_t(2 * kk), db1 + boff + uint64_t(2 * kk), id_h1, kk != 0 ? 1u : 0u);
              }
              ptx::umma_commit<1>(sd_full(1));
              if (j == nblk - 1) ptx::umma_commit<1>(tile_empty(slot));   // last product that reads this tile pair
            }
            __syncwarp();
            // both halves must be done before the second products: they read the packed operands of both
            ptx::mbar_wait(ds_ready(0), blk_cnt & 1u);
            ptx::mbar_wait(ds_ready(1), blk_cnt & 1u);
            if (!acc_waited) { ptx::mbar_wait(acc_free, (tile_cnt - 1u) & 1u); acc_waited = true; }
            ptx::tcgen05_fence_after();
            // second products: reduction over the block's columns = rows [16 u0, ...) of X / Y as MN-major operands
            const uint64_t dx = mn_desc(x_s(s) + uint32_t(u0) * 2048u), dy = mn_desc(y_s(s) + uint32_t(u0) * 2048u);
            if (ptx::elect_one()) {
              if (pass == 0) {       // dQ_t += dS K
                for (int ks = 0; ks < u; ++ks)
                  umma_ts(tmem_base + kAcc0, tmem_base + kColDp + packed_col(ks, ua), dx + uint64_t(ks * 128), id_acc, (j | ks) != 0 ? 1u : 0u);
              } else {               // dV_u += P^T dO,  dK_u += dS^T Q
                for (int ks = 0; ks < u; ++ks)
                  umma_ts(tmem_base + kAcc0, tmem_base + kColS + packed_col(ks, ua), dy + uint64_t(ks * 128), id_acc, (j | ks) != 0 ? 1u : 0u);
                for (int ks = 0; ks < u; ++ks)
                  umma_ts(tmem_base + kAcc1, tmem_base + kColDp + packed_col(ks, ua), dx + uint64_t(ks * 128), id_acc, (j | ks) != 0 ? 1u : 0u);
              }
              if (j == nblk - 1) {
                ptx::umma_commit<1>(acc_full);
                if (t == n_tiles - 1) ptx::umma_commit<1>(xy_empty(s));   // last product that reads this pass's X / Y
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  } else {
    // ======================= compute: thread = one TMEM lane (row), warp group = one half of the columns =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    const float c = prm.scale_log2e, scale = prm.scale;
    const int ctid = threadIdx.x - 128;                    // 0..255
    uint32_t tile_cnt = 0, blk_cnt = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / H, h = item - b * H;
      // ---- per-item vectors: lse (log2 domain) and D of every query; padded queries get lse = +inf (P = 0), D = 0 ----
      asm volatile("bar.sync 1, 256;" ::: "memory");       // previous item's readers are done with the vectors
      for (int i = ctid; i < kVecFloats; i += 256) {
        const bool ok = i < T;
        s_lse[i] = ok ? prm.lse[(size_t(b) * H + h) * T + i] : INFINITY;
        s_d[i] = ok ? prm.dsum[(size_t(b) * H + h) * T + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
          const int row = t * 128 + r;                     // query (pass Q) or key (pass K) this thread owns
          const bool row_ok = row < T;
          const bool warp_live = t * 128 + quad * 32 < T;  // warps whose 32 rows are all padding skip the math (rows are independent)
          const float lse_r = (pass == 0 && row < kVecFloats) ? s_lse[row] : 0.f;
          const float d_r = (pass == 0 && row < kVecFloats) ? s_d[row] : 0.f;
          for (int j = 0; j < nblk; ++j, ++blk_cnt) {
            const int u = blk_units(j), u0 = blk_start(j), ua = (u + 1) / 2;
            const int v_lo = grp == 0 ? 0 : ua, v_hi = grp == 0 ? ua : u;
            ptx::mbar_wait(sd_full(grp), blk_cnt & 1u);
            ptx::tcgen05_fence_after();
#pragma unroll
            for (int vv = 0; vv < 8; ++vv) {
              const int v = v_lo + vv;                     // unit inside the block
              if (v < v_hi && warp_live) {
                uint32_t sv[16], dv[16];
                ld16(lane_addr + kColS + uint32_t(16 * v), sv);
                ld16(lane_addr + kColDp + uint32_t(16 * v), dv);
                ptx::tmem_ld_wait();
                uint32_t pp[8], pds[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const int c0 = 16 * (u0 + v) + 2 * e;    // column inside the whole sequence
                  float l0, l1, dd0, dd1;
                  if (pass == 0) {
                    l0 = l1 = lse_r; dd0 = dd1 = d_r;
                  } else {
                    const float2 lv = *reinterpret_cast<const float2*>(&s_lse[c0]);   // warp-uniform address: broadcast
                    const float2 dvv = *reinterpret_cast<const float2*>(&s_d[c0]);
                    l0 = lv.x; l1 = lv.y; dd0 = dvv.x; dd1 = dvv.y;
                  }
                  float p0 = ex2(fmaf(__uint_as_float(sv[2 * e]), c, -l0));
                  float p1 = ex2(fmaf(__uint_as_float(sv[2 * e + 1]), c, -l1));
                  if (pass == 0) {                         // key columns beyond T hold another image's keys
                    p0 = c0 < T ? p0 : 0.f;
                    p1 = c0 + 1 < T ? p1 : 0.f;
                    if (prm.causal) {                      // row = query, column = key: keys after the query are masked
                      p0 = c0 <= row ? p0 : 0.f;
                      p1 = c0 + 1 <= row ? p1 : 0.f;
                    }
                  } else if (prm.causal) {                 // row = key, column = query: queries before the key never saw it
                    p0 = c0 >= row ? p0 : 0.f;
                    p1 = c0 + 1 >= row ? p1 : 0.f;
                  }
                  const float ds0 = p0 * (__uint_as_float(dv[2 * e]) - dd0) * scale;
                  const float ds1 = p1 * (__uint_as_float(dv[2 * e + 1]) - dd1) * scale;
                  pp[e] = Act<kF16>::pack(p0, p1);
                  pds[e] = Act<kF16>::pack(ds0, ds1);
                }
                // in place: the packed unit lands in fp32 columns this group has already read (see packed_col)
                if (pass == 1) st8(lane_addr + kColS + packed_col(v, ua), pp);
                st8(lane_addr + kColDp + packed_col(v, ua), pds);
              }
            }
            st_wait();
            ptx::tcgen05_fence_before();
            ptx::mbar_arrive(ds_ready(grp));
          }
          // ---- accumulators -> global ----
          ptx::mbar_wait(acc_full, tile_cnt & 1u);
          ptx::tcgen05_fence_after();
          uint32_t a[64];
          // pass Q: dQ (acc0), columns split between the groups.  pass K: group 0 stores dV (acc0), group 1 dK (acc1).
          const int ncol = pass == 0 ? 32 : 64;
          const uint32_t acc_addr = lane_addr + (pass == 0 ? kAcc0 + uint32_t(32 * grp) : (grp == 0 ? kAcc0 : kAcc1));
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            if (16 * cc < ncol && warp_live) {
              uint32_t tmp[16];
              ld16(acc_addr + uint32_t(16 * cc), tmp);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) a[16 * cc + e] = tmp[e];
            }
          }
          ptx::tcgen05_fence_before();
          ptx::mbar_arrive(acc_free);
          if (row_ok) {
            const int col0 = pass == 0 ? h * kHd + 32 * grp : (grp == 0 ? 2 * d : d) + h * kHd;
            uint16_t* dst = prm.dqkv + (size_t(b) * T + row) * (3 * d) + col0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              if (8 * jj < ncol) {
                uint4 w;
                w.x = Act<kF16>::pack(__uint_as_float(a[8 * jj]), __uint_as_float(a[8 * jj + 1]));
                w.y = Act<kF16>::pack(__uint_as_float(a[8 * jj + 2]), __uint_as_float(a[8 * jj + 3]));
                w.z = Act<kF16>::pack(__uint_as_float(a[8 * jj + 4]), __uint_as_float(a[8 * jj + 5]));
                w.w = Act<kF16>::pack(__uint_as_float(a[8 * jj + 6]), __uint_as_float(a[8 * jj + 7]));
                *reinterpret_cast<uint4*>(dst + 8 * jj) = w;
              }
            }
          }
        }
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 512);
}

// D[bh, q] = sum_d dO[q, h*64 + d] * O[q, h*64 + d]   (8 lanes per (row, head), 8 elements each)
template <bool kF16>
__global__ void attention_bwd_dsum_kernel(const uint16_t* __restrict__ d_o, const uint16_t* __restrict__ o, float* __restrict__ dsum,
                                          int B, int T, int H) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * H * 8;
  const long long rh = i >> 3;
  const int ch = int(i & 7);
  float part = 0.f;
  int b = 0, q = 0, h = 0;
  if (i < total) {
    const long long row = rh / H;
    h = int(rh - row * H);
    b = int(row / T);
    q = int(row - (long long)b * T);
    const size_t off = size_t(row) * (size_t(H) * kHd) + size_t(h) * kHd + size_t(ch) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(d_o + off);
    const uint4 cc = *reinterpret_cast<const uint4*>(o + off);
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* pc = reinterpret_cast<const uint32_t*>(&cc);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = Act<kF16>::unpack(pa[e]), y = Act<kF16>::unpack(pc[e]);
      part += x.x * y.x + x.y * y.y;
    }
  }
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  part += __shfl_xor_sync(0xffffffffu, part, 4);
  if (i < total && ch == 0) dsum[(size_t(b) * H + h) * T + q] = part;
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

int launch_attention_bwd_dsum(const void* d_out, const void* out, float* dsum, int B, int T, int H, int f16, cudaStream_t stream) {
  if (B <= 0) return 0;
  const long long total = (long long)B * T * H * 8;
  const int blocks = int((total + 255) / 256);
  if (f16)
    attention_bwd_dsum_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(d_out), static_cast<const uint16_t*>(out), dsum, B, T, H);
  else
    attention_bwd_dsum_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(d_out), static_cast<const uint16_t*>(out), dsum, B, T, H);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// returns -3 when the shape is outside this kernel's envelope (caller falls back to the mma.sync kernel)
int launch_attention_bwd_sm100(const void* qkv, const void* out, const void* d_out, const float* lse, float* dsum_scratch,
                               void* dqkv, int B, int T, int H, int head_dim, int f16, int causal, int num_sms,
                               cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1 || T > kVecFloats - 48 || dsum_scratch == nullptr) return -3;
  BwdParams p;
  p.TP = (T + 15) / 16 * 16;
  const int units = p.TP / 16;
  // one column block while the sequence fits a 256-column TMEM region, else blocks of <= 192 columns (12 units)
  p.nblk = units <= 16 ? 1 : (units + 11) / 12;
  p.bq = units / p.nblk;
  p.brem = units % p.nblk;
  // X / Y arrive in TMA boxes of 16*dd rows; pick the box (<= 256 rows) that overshoots TP the least
  int best_d = 1, best_rows = 1 << 30;
  for (int dd = 16; dd >= 4; --dd) {
    const int rows = (units + dd - 1) / dd * dd;
    if (rows < best_rows) { best_rows = rows; best_d = dd; }
  }
  if (units <= 16) { best_d = units; best_rows = units; }
  p.xy_box = 16 * best_d;
  p.xy_loads = best_rows / best_d;
  p.xy_bytes = best_rows * 16 * 128;
  const int fixed = 4 * kTileBytes + 2 * kVecFloats * 4 + kBarBytes + 1024 /*alignment slack*/;
  if (fixed + 2 * p.xy_bytes > kMaxSmem) return -3;
  p.stages = fixed + 4 * p.xy_bytes <= kMaxSmem ? 2 : 1;
  const int smem = fixed + p.stages * 2 * p.xy_bytes;
  p.items = B * H;
  p.causal = causal;
  p.T = T;
  p.H = H;
  p.lse = lse;
  p.dsum = dsum_scratch;
  p.dqkv = static_cast<uint16_t*>(dqkv);
  p.scale = 1.0f / sqrtf(float(head_dim));
  p.scale_log2e = 1.4426950408889634f * p.scale;
  const int d = H * kHd;
  if (int rc = launch_attention_bwd_dsum(d_out, out, dsum_scratch, B, T, H, f16, stream)) return rc;
  CUtensorMap tqf, tqt, tdf, tdt;
  const uint64_t rows = uint64_t(B) * T;
  const bool h16 = f16 != 0;
  if (!make_map(&tqf, qkv, rows, uint64_t(3 * d), uint32_t(p.xy_box), h16) || !make_map(&tqt, qkv, rows, uint64_t(3 * d), 128, h16) ||
      !make_map(&tdf, d_out, rows, uint64_t(d), uint32_t(p.xy_box), h16) || !make_map(&tdt, d_out, rows, uint64_t(d), 128, h16))
    return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_bwd_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
  if (f16)
    attention_bwd_sm100_kernel<true><<<grid, kThreads, smem, stream>>>(tqf, tqt, tdf, tdt, p);
  else
    attention_bwd_sm100_kernel<false><<<grid, kThreads, smem, stream>>>(tqf, tqt, tdf, tdt, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
