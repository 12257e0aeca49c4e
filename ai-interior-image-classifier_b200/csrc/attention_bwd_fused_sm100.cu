// tcgen05 attention backward in ONE pass per key tile, head_dim 64, T <= 256 (ViT-B/16 @ 224: T = 197; the text tower: 77):
//   (dO, saved Q/K/V, O, log-sum-exp) -> dQ, dK, dV, packed like qkv.
//
// Reference semantics: torch.autograd through nn.MultiheadAttention's softmax(Q K^T / sqrt(64)) V inside CLIP's
// ResidualAttentionBlock (/root/reference/train_lora.py:249 `loss.backward()` restricted to parameters named '*lora*'):
// the backward of the frozen attention only carries dX down to the LoRA-adapted MLPs.
//
// attention_bwd_sm100.cu walks every (image, head) item twice - a query pass for dQ and a key pass for dK / dV - and recomputes
// the scores and their exponentials in both.  Here only the KEY pass exists.  For key tile u (128 keys = TMEM lanes) and query
// block j (<= 128 queries = TMEM columns):
//
//   S^T = K_u Q_j^T,  dP^T = V_u dO_j^T   (fp32, TMEM)   ->   P^T = exp2(S^T c - lse[q]),  dS^T = P^T o (dP^T - D[q]) / 8
//   dV_u += P^T dO_j,  dK_u += dS^T Q_j                   A operands: P^T / dS^T as 16 bit IN PLACE in TMEM (as before)
//   dQ_j += dS K_u                                        A operand: the same dS^T values, which the compute warps ALSO store
//                                                         (16 bit, [key][query], 128B swizzle) in a shared-memory tile: read
//                                                         with an MN-major descriptor that tile IS dS (m = query, k = key)
//
// so every exponential and every first product is computed once.  dQ of the (at most two) query blocks accumulates over the
// key tiles in its own TMEM columns and is read out once per item.  TMEM: S^T [0, 128)  dP^T [128, 256)  dV [256, 320)
// dK [320, 384)  dQ_0 [384, 448)  dQ_1 [448, 512).  Keys beyond the sequence hold another image's rows: their dS is forced to
// zero in the shared-memory tile (they are the REDUCTION index of dQ; as rows of dV / dK they are simply not stored).
//
//   warp 0  TMA producer Q / dO (whole sequence)   warp 3  TMA producer K_u / V_u tiles   warp 1  MMA issuer
//   warp 2  TMEM allocator      warps 4-11  compute (thread = one key; warps 4-7 / 8-11 = first / second half of a block's queries)
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
constexpr int kVecFloats = 256;         // lse / D of one item (T <= 256)
constexpr int kStageBytes = 2 * kTileBytes;   // dS^T of one (key tile, query block): two 64-query atoms of [128 keys][128 B]

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
// the same over TWO 64-wide atoms along MN that lie `atom_bytes` apart (LBO): M = 128 rows of an MN-major A operand
__device__ __forceinline__ uint64_t mn_desc_atoms(uint32_t smem_addr, uint32_t atom_bytes) {
  uint64_t d = uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t((atom_bytes >> 4) & 0x3FFFu) << 16;
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
attention_bwd_fused_kernel(const __grid_constant__ CUtensorMap tm_qkv_full, const __grid_constant__ CUtensorMap tm_qkv_tile,
                           const __grid_constant__ CUtensorMap tm_do_full, const __grid_constant__ CUtensorMap tm_do_tile,
                           const __grid_constant__ BwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int T = prm.T, H = prm.H, items = prm.items;
  const int d = H * kHd;
  // smem: tile ring (2 slots x (A0, A1)), X / Y stages, per-item vectors, barriers
  auto ta_s = [&](int slot, int which) { return base + uint32_t(2 * slot + which) * kTileBytes; };
  const uint32_t stage_s = base + 4u * kTileBytes;          // dS^T tile (MN-major A operand of the dQ product)
  uint8_t* stage_gen = gen + 4u * kTileBytes;
  const uint32_t xy_base = stage_s + uint32_t(kStageBytes);
  auto x_s = [&](int s) { return xy_base + uint32_t(2 * s) * uint32_t(prm.xy_bytes); };
  auto y_s = [&](int s) { return xy_base + uint32_t(2 * s + 1) * uint32_t(prm.xy_bytes); };
  const uint32_t vec_off = 4u * kTileBytes + uint32_t(kStageBytes) + uint32_t(prm.stages) * 2u * uint32_t(prm.xy_bytes);
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
  const uint32_t dq_full = bar + 112, dq_free = bar + 120;
  const uint32_t tmem_slot = bar + 128;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + bar_off + 128u);

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
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dq_free, 256);
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
  // TMEM columns: S^T block at [0, 16 u), dP^T block at [128, 128 + 16 u) (u <= 8 units of 16 queries).  Column group 0 owns the
  // block's units [0, ua), group 1 units [ua, u); the packed 16-bit unit v of a group sits at the group's first column + 8 * (v - first).
  constexpr uint32_t kColS = 0, kColDp = 128, kAcc0 = 256, kAcc1 = 320, kAccQ = 384;
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
        {
          const int s = m % prm.stages;
          ptx::mbar_wait(xy_empty(s), (uint32_t(m / prm.stages) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(xy_full(s), 2u * bytes);
          for (int l = 0; l < prm.xy_loads; ++l) {   // X = Q, Y = dO
            const uint32_t off = uint32_t(l * prm.xy_box) * 128u;
            const int row = b * T + l * prm.xy_box;
            ptx::tma_load_2d(&tm_qkv_full, xy_full(s), x_s(s) + off, h * kHd, row, ptx::kEvictFirst);
            ptx::tma_load_2d(&tm_do_full, xy_full(s), y_s(s) + off, h * kHd, row, ptx::kEvictFirst);
          }
          ++m;
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
        for (int t = 0; t < n_tiles; ++t, ++tc) {   // K_u, V_u
          const int slot = int(tc & 1u);
          ptx::mbar_wait(tile_empty(slot), ((tc >> 1) & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(tile_full(slot), 2u * kTileBytes);
          const int row = b * T + t * 128;
          ptx::tma_load_2d(&tm_qkv_tile, tile_full(slot), ta_s(slot, 0), d + h * kHd, row, ptx::kEvictFirst);
          ptx::tma_load_2d(&tm_qkv_tile, tile_full(slot), ta_s(slot, 1), 2 * d + h * kHd, row, ptx::kEvictFirst);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t id_acc = idesc(128, kHd, kF16, true);                     // A from TMEM, B MN-major
    const uint32_t id_dq = idesc(128, kHd, kF16, true) | (1u << 15);         // A (the shared-memory dS^T tile) MN-major as well
    uint32_t tile_cnt = 0, blk_cnt = 0;   // phases of acc_full / acc_free / tile ring, and of sd_full / ds_ready
    for (int n = 0; n < n_items; ++n) {
      const int s = n % prm.stages;
      ptx::mbar_wait(xy_full(s), uint32_t(n / prm.stages) & 1u);
      for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
        const int slot = int(tile_cnt & 1u);
        ptx::mbar_wait(tile_full(slot), (tile_cnt >> 1) & 1u);
        bool acc_waited = tile_cnt == 0;
        const uint64_t da0 = ptx::make_kmajor_sw128_desc(ta_s(slot, 0)), da1 = ptx::make_kmajor_sw128_desc(ta_s(slot, 1));
        const uint64_t dk_mn = mn_desc(ta_s(slot, 0));                        // K_u as [key][dim]: B operand of dQ += dS K_u
        for (int j = 0; j < nblk; ++j, ++blk_cnt) {
          const int u = blk_units(j), u0 = blk_start(j), ua = (u + 1) / 2;
          ptx::tcgen05_fence_after();
          // first products over the block's queries = rows [16 u0, 16 (u0 + u)) of Q / dO, as two halves with their own barriers
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
          }
          __syncwarp();
          // both halves must be done before the second products: they read the packed operands (and the dS^T tile) of both
          ptx::mbar_wait(ds_ready(0), blk_cnt & 1u);
          ptx::mbar_wait(ds_ready(1), blk_cnt & 1u);
          if (!acc_waited) { ptx::mbar_wait(acc_free, (tile_cnt - 1u) & 1u); acc_waited = true; }
          if (n > 0 && t == 0 && j == 0) ptx::mbar_wait(dq_free, uint32_t(n - 1) & 1u);   // the previous item's dQ has been read out
          ptx::tcgen05_fence_after();
          // second products.  dV_u += P^T dO_j, dK_u += dS^T Q_j: reduction over the block's queries = rows [16 u0, ...) of dO / Q as
          // MN-major operands.  dQ_j += dS K_u: reduction over the tile's 128 keys, A = the dS^T tile read MN-major (m = query:
          // two 64-query atoms 16 KB apart), B = K_u read MN-major.
          const uint64_t dx = mn_desc(x_s(s) + uint32_t(u0) * 2048u), dy = mn_desc(y_s(s) + uint32_t(u0) * 2048u);
          const uint64_t dsm = mn_desc_atoms(stage_s, kTileBytes);
          if (ptx::elect_one()) {
            for (int ks = 0; ks < u; ++ks)
              umma_ts(tmem_base + kAcc0, tmem_base + kColS + packed_col(ks, ua), dy + uint64_t(ks * 128), id_acc, (j | ks) != 0 ? 1u : 0u);
            for (int ks = 0; ks < u; ++ks)
              umma_ts(tmem_base + kAcc1, tmem_base + kColDp + packed_col(ks, ua), dx + uint64_t(ks * 128), id_acc, (j | ks) != 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::umma_f16<1>(tmem_base + kAccQ + uint32_t(64 * j), dsm + uint64_t(ks * 128), dk_mn + uint64_t(ks * 128), id_dq,
                               (t | ks) != 0 ? 1u : 0u);
            if (j == nblk - 1) {
              ptx::umma_commit<1>(tile_empty(slot));   // the dQ product was the last reader of K_u
              ptx::umma_commit<1>(acc_full);
              if (t == n_tiles - 1) {
                ptx::umma_commit<1>(dq_full);
                ptx::umma_commit<1>(xy_empty(s));       // last product that reads this item's Q / dO
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  } else {
    // ======================= compute: thread = one TMEM lane (key), warp group = one half of the block's queries =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    const float c = prm.scale_log2e, scale = prm.scale;
    const int ctid = threadIdx.x - 128;                    // 0..255
    uint32_t tile_cnt = 0, blk_cnt = 0, item_cnt = 0;
    uint8_t* my_stage_row = stage_gen + r * 128;           // this key's 128-byte row in either atom of the dS^T tile
    const int sw = r & 7;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_cnt) {
      const int b = item / H, h = item - b * H;
      // ---- per-item vectors: lse (log2 domain) and D of every query; padded queries get lse = +inf (P = 0), D = 0 ----
      asm volatile("bar.sync 1, 256;" ::: "memory");       // previous item's readers are done with the vectors
      for (int i = ctid; i < kVecFloats; i += 256) {
        const bool ok = i < T;
        s_lse[i] = ok ? prm.lse[(size_t(b) * H + h) * T + i] : INFINITY;
        s_d[i] = ok ? prm.dsum[(size_t(b) * H + h) * T + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
        const int row = t * 128 + r;                       // the key this thread owns
        const bool row_ok = row < T;
        const bool warp_live = t * 128 + quad * 32 < T;    // warps whose 32 keys are all padding only zero their dS^T rows
        for (int j = 0; j < nblk; ++j, ++blk_cnt) {
          const int u = blk_units(j), u0 = blk_start(j), ua = (u + 1) / 2;
          const int v_lo = grp == 0 ? 0 : ua, v_hi = grp == 0 ? ua : u;
          ptx::mbar_wait(sd_full(grp), blk_cnt & 1u);
          ptx::tcgen05_fence_after();
#pragma unroll
          for (int vv = 0; vv < 4; ++vv) {
            const int v = v_lo + vv;                       // unit inside the block
            if (v < v_hi && !warp_live) {
              uint8_t* rowp = my_stage_row + (v >> 2) * kTileBytes;
              const int ch = (v & 3) * 2;
              *reinterpret_cast<uint4*>(rowp + (((ch) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            } else if (v < v_hi) {
              uint32_t sv[16], dv[16];
              ld16(lane_addr + kColS + uint32_t(16 * v), sv);
              ld16(lane_addr + kColDp + uint32_t(16 * v), dv);
              ptx::tmem_ld_wait();
              uint32_t pp[8], pds[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c0 = 16 * (u0 + v) + 2 * e;      // query index inside the whole sequence
                const float2 lv = *reinterpret_cast<const float2*>(&s_lse[c0]);   // warp-uniform address: broadcast
                const float2 dvv = *reinterpret_cast<const float2*>(&s_d[c0]);
                float p0 = ex2(fmaf(__uint_as_float(sv[2 * e]), c, -lv.x));
                float p1 = ex2(fmaf(__uint_as_float(sv[2 * e + 1]), c, -lv.y));
                if (prm.causal) {                          // row = key, column = query: queries before the key never saw it
                  p0 = c0 >= row ? p0 : 0.f;
                  p1 = c0 + 1 >= row ? p1 : 0.f;
                }
                // keys beyond the sequence are another image's rows: as the REDUCTION index of dQ their dS must vanish
                const float ds0 = row_ok ? p0 * (__uint_as_float(dv[2 * e]) - dvv.x) * scale : 0.f;
                const float ds1 = row_ok ? p1 * (__uint_as_float(dv[2 * e + 1]) - dvv.y) * scale : 0.f;
                pp[e] = Act<kF16>::pack(p0, p1);
                pds[e] = Act<kF16>::pack(ds0, ds1);
              }
              // in place: the packed unit lands in fp32 columns this group has already read (see packed_col)
              st8(lane_addr + kColS + packed_col(v, ua), pp);
              st8(lane_addr + kColDp + packed_col(v, ua), pds);
              // and dS^T into the shared-memory tile: unit v = queries [16 v, 16 v + 16) of the block = two 16-byte chunks of
              // this key's row in atom v / 4
              uint8_t* rowp = my_stage_row + (v >> 2) * kTileBytes;
              const int ch = (v & 3) * 2;
              *reinterpret_cast<uint4*>(rowp + (((ch) ^ sw) << 4)) = make_uint4(pds[0], pds[1], pds[2], pds[3]);
              *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ sw) << 4)) = make_uint4(pds[4], pds[5], pds[6], pds[7]);
            }
          }
          st_wait();
          ptx::fence_proxy_async_smem();                   // the dS^T tile is read by the tensor core (async proxy)
          ptx::tcgen05_fence_before();
          ptx::mbar_arrive(ds_ready(grp));
        }
        // ---- dV_u (acc0, group 0) and dK_u (acc1, group 1) -> global ----
        ptx::mbar_wait(acc_full, tile_cnt & 1u);
        ptx::tcgen05_fence_after();
        {
          const uint32_t acc_addr = lane_addr + (grp == 0 ? kAcc0 : kAcc1);
          uint16_t* dst = prm.dqkv + (size_t(b) * T + row) * (3 * d) + (grp == 0 ? 2 * d : d) + h * kHd;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t a[16];
            ld16(acc_addr + uint32_t(16 * cc), a);
            ptx::tmem_ld_wait();
            if (cc == 3) {
              ptx::tcgen05_fence_before();
              ptx::mbar_arrive(acc_free);
            }
            if (row_ok) {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                uint4 w;
                w.x = Act<kF16>::pack(__uint_as_float(a[8 * jj]), __uint_as_float(a[8 * jj + 1]));
                w.y = Act<kF16>::pack(__uint_as_float(a[8 * jj + 2]), __uint_as_float(a[8 * jj + 3]));
                w.z = Act<kF16>::pack(__uint_as_float(a[8 * jj + 4]), __uint_as_float(a[8 * jj + 5]));
                w.w = Act<kF16>::pack(__uint_as_float(a[8 * jj + 6]), __uint_as_float(a[8 * jj + 7]));
                *reinterpret_cast<uint4*>(dst + 16 * cc + 8 * jj) = w;
              }
            }
          }
        }
      }
      // ---- dQ of the item: block j sits in lanes [0, 16 u_j) of its accumulator; group g stores dims [32 g, 32 g + 32) ----
      ptx::mbar_wait(dq_full, item_cnt & 1u);
      ptx::tcgen05_fence_after();
      for (int j = 0; j < nblk; ++j) {
        const int q = 16 * blk_start(j) + r;               // the query in this thread's lane
        const bool q_ok = r < 16 * blk_units(j) && q < T;
        uint32_t a0[16], a1[16];
        ld16(lane_addr + kAccQ + uint32_t(64 * j + 32 * grp), a0);
        ld16(lane_addr + kAccQ + uint32_t(64 * j + 32 * grp + 16), a1);
        ptx::tmem_ld_wait();
        if (q_ok) {
          uint16_t* dst = prm.dqkv + (size_t(b) * T + q) * (3 * d) + h * kHd + 32 * grp;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t* a = jj < 2 ? a0 + 8 * jj : a1 + 8 * (jj - 2);
            uint4 w;
            w.x = Act<kF16>::pack(__uint_as_float(a[0]), __uint_as_float(a[1]));
            w.y = Act<kF16>::pack(__uint_as_float(a[2]), __uint_as_float(a[3]));
            w.z = Act<kF16>::pack(__uint_as_float(a[4]), __uint_as_float(a[5]));
            w.w = Act<kF16>::pack(__uint_as_float(a[6]), __uint_as_float(a[7]));
            *reinterpret_cast<uint4*>(dst + 8 * jj) = w;
          }
        }
      }
      ptx::tcgen05_fence_before();
      ptx::mbar_arrive(dq_free);
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

// one pass per key tile, T <= 256; dsum (D = rowsum(dO o O), launch_attention_bwd_dsum) must already be on the stream;
// returns -3 when the shape is outside this kernel's envelope
int launch_attention_bwd_fused_sm100(const void* qkv, const void* d_out, const float* lse, const float* dsum, void* dqkv, int B, int T,
                                     int H, int head_dim, int f16, int causal, int num_sms, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1 || T > 256 || dsum == nullptr) return -3;
  BwdParams p;
  p.TP = (T + 15) / 16 * 16;
  const int units = p.TP / 16;
  p.nblk = units <= 8 ? 1 : 2;             // query blocks of <= 8 units (128 TMEM columns)
  p.bq = units / p.nblk;
  p.brem = units % p.nblk;
  p.xy_box = 16 * units;                    // Q / dO of the whole sequence in one TMA box (<= 256 rows)
  p.xy_loads = 1;
  p.xy_bytes = units * 16 * 128;
  const int fixed = 4 * kTileBytes + kStageBytes + 2 * kVecFloats * 4 + kBarBytes + 1024 /*alignment slack*/;
  if (fixed + 2 * p.xy_bytes > kMaxSmem) return -3;
  p.stages = fixed + 4 * p.xy_bytes <= kMaxSmem ? 2 : 1;
  const int smem = fixed + p.stages * 2 * p.xy_bytes;
  p.items = B * H;
  p.causal = causal;
  p.T = T;
  p.H = H;
  p.lse = lse;
  p.dsum = dsum;
  p.dqkv = static_cast<uint16_t*>(dqkv);
  p.scale = 1.0f / sqrtf(float(head_dim));
  p.scale_log2e = 1.4426950408889634f * p.scale;
  const int d = H * kHd;
  CUtensorMap tqf, tqt, tdf, tdt;
  const uint64_t rows = uint64_t(B) * T;
  const bool h16 = f16 != 0;
  if (!make_map(&tqf, qkv, rows, uint64_t(3 * d), uint32_t(p.xy_box), h16) || !make_map(&tqt, qkv, rows, uint64_t(3 * d), 128, h16) ||
      !make_map(&tdf, d_out, rows, uint64_t(d), uint32_t(p.xy_box), h16) || !make_map(&tdt, d_out, rows, uint64_t(d), 128, h16))
    return -1;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(attention_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done.mark();
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
  if (f16)
    attention_bwd_fused_kernel<true><<<grid, kThreads, smem, stream>>>(tqf, tqt, tdf, tdt, p);
  else
    attention_bwd_fused_kernel<false><<<grid, kThreads, smem, stream>>>(tqf, tqt, tdf, tdt, p);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
