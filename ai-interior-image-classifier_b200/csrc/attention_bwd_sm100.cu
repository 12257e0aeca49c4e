// tcgen05 attention backward for sequences that fit one TMEM tile pair (T <= 256: ViT-B/16 @ 224, T = 197):
//   (dO, saved Q/K/V, O, log-sum-exp) -> dQ, dK, dV, packed like qkv.
//
// Reference semantics: torch.autograd through nn.MultiheadAttention's softmax(Q K^T / sqrt(64)) V inside CLIP's
// ResidualAttentionBlock; needed only to carry dX through the frozen attention blocks down to the LoRA-adapted MLPs
// (/root/reference/train_lora.py:249 `loss.backward()` restricted to parameters named '*lora*').
//
// One persistent CTA per SM walks (image, head) items; Q, K, V, dO of the item are TMA-loaded into 128B-swizzled smem (two
// items in flight when they fit) and serve both as K-major operands (S = Q K^T ...) and as MN-major operands
// (dQ = dS K ...) - same bytes, different descriptors.  Per item, four 128-row tiles, all products on tcgen05:
//
//   pass Q, query tile t:  S = Q_t K^T, dP = dO_t V^T  (fp32, TMEM)  ->  dS = P o (dP - D) / 8  ->  dQ_t = dS K
//   pass K, key tile u:    S^T = K_u Q^T, dP^T = V_u dO^T            ->  P^T, dS^T              ->  dV_u = P^T dO, dK_u = dS^T Q
//
// with P = exp2(S c - lse) recomputed from the saved log-sum-exp and D[q] = sum_d dO[q,d] O[q,d] from a small pre-kernel.
// The backward softmax is purely elementwise (no row reductions), so the 8 compute warps split the COLUMNS of a tile:
// warps 4-7 take the first half, warps 8-11 the second half of every row (a thread still owns one TMEM lane = one row).
// P / dS go back to TMEM as 16-bit A operands IN PLACE, each group packing into the start of ITS OWN fp32 columns (a K-step of
// the second product takes its A operand from any column, so the two halves need not be contiguous): a group only ever
// overwrites columns it has already read, and the groups never wait for each other.  The 64-column accumulators of the second
// products sit in the dead top columns [192, 256) of the two regions, so the whole pipeline needs 2 x 256 = 512 TMEM columns.
// The two halves of S / dP are separate products with their own barriers: the first group starts while the tensor core still
// works on the second half.
//
//   warp 0  TMA producer      warp 1  MMA issuer (whole warp, uniform operands, one elected lane issues)
//   warp 2  TMEM allocator    warps 4-11  compute
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

struct BwdParams {
  const float* lse;   // [B*H, T] log2-domain log-sum-exp of the scaled scores
  const float* dsum;  // [B*H, T] D = rowsum(dO o O)
  uint16_t* dqkv;     // [B*T, 3d]
  int items, T, H, TP;
  int stages;         // items resident at once (1 or 2)
  float scale, scale_log2e;
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

// kHU: compile-time bound on the 16-column units one warp group handles (7: TP <= 224, i.e. ViT-B/16; 8: TP <= 256)
template <bool kF16, int kHU>
__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                           const __grid_constant__ BwdParams prm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int T = prm.T, H = prm.H, TP = prm.TP, items = prm.items;
  const int d = H * kHd;
  const uint32_t tile_bytes = uint32_t(TP) * 128u;                      // one of Q, K, V, dO
  const uint32_t stage_bytes = 4u * tile_bytes;
  auto q_s = [&](int s) { return base + uint32_t(s) * stage_bytes; };
  auto k_s = [&](int s) { return q_s(s) + tile_bytes; };
  auto v_s = [&](int s) { return q_s(s) + 2u * tile_bytes; };
  auto do_s = [&](int s) { return q_s(s) + 3u * tile_bytes; };
  const uint32_t vec_off = uint32_t(prm.stages) * stage_bytes;          // lse / D of the current item: 2 x 256 floats
  float* s_lse = reinterpret_cast<float*>(gen + vec_off);
  float* s_d = s_lse + 256;
  const uint32_t bar = base + vec_off + 2048u;
  auto op_full = [&](int s) { return bar + 8u * s; };
  auto op_empty = [&](int s) { return bar + 16 + 8u * s; };
  auto sd_full = [&](int g) { return bar + 32 + 8u * g; };
  auto ds_ready = [&](int g) { return bar + 48 + 8u * g; };
  const uint32_t acc_full = bar + 64, acc_free = bar + 72;
  const uint32_t tmem_slot = bar + 80;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + vec_off + 2048u + 80u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv);
    ptx::prefetch_tensormap(&tm_do);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(op_full(s), 1); ptx::mbar_init(op_empty(s), 1); }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(sd_full(g), 1);      // tcgen05.commit
      ptx::mbar_init(ds_ready(g), 128);   // every thread of the group
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
  const int n_tiles = (T + 127) / 128;          // 128-row tiles per pass (1 or 2)
  // TMEM columns: S / S^T at [0, TP), dP / dP^T at [256, 256 + TP).  Column group 0 owns units [0, ua) of 16 columns,
  // group 1 units [ua, units); the packed 16-bit unit u of a group sits at the group's first column + 8 * (u - first unit).
  // 64-column accumulators in the top columns of the regions: [192, 256) and [448, 512).
  constexpr uint32_t kColS = 0, kColDp = 256, kAcc0 = 192, kAcc1 = 448;
  const int units = TP / 16;
  const int ua = (units + 1) / 2;
  auto packed_col = [&](int u) { return uint32_t(u < ua ? 8 * u : 16 * ua + 8 * (u - ua)); };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (warp == 0) {
    // ======================= TMA producer =======================
    if (ptx::elect_one()) {
      int n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const int b = item / H, h = item - b * H;
        const int s = n % prm.stages;
        const uint32_t par = uint32_t(n / prm.stages) & 1u;
        ptx::mbar_wait(op_empty(s), par ^ 1u);
        ptx::mbar_arrive_expect_tx(op_full(s), stage_bytes);
        ptx::tma_load_2d(&tm_qkv, op_full(s), q_s(s), h * kHd, b * T, ptx::kEvictFirst);
        ptx::tma_load_2d(&tm_qkv, op_full(s), k_s(s), d + h * kHd, b * T, ptx::kEvictFirst);
        ptx::tma_load_2d(&tm_qkv, op_full(s), v_s(s), 2 * d + h * kHd, b * T, ptx::kEvictFirst);
        ptx::tma_load_2d(&tm_do, op_full(s), do_s(s), h * kHd, b * T, ptx::kEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    const uint32_t id_acc = idesc(128, kHd, kF16, true);
    uint32_t tile_cnt = 0;   // tiles issued so far: phase of sd_full / ds_ready / acc_full / acc_free
    for (int n = 0; n < n_items; ++n) {
      const int s = n % prm.stages;
      ptx::mbar_wait(op_full(s), uint32_t(n / prm.stages) & 1u);
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
          // the accumulators of the previous tile live inside the S / dP regions: they must have been read out
          if (tile_cnt > 0) ptx::mbar_wait(acc_free, (tile_cnt - 1u) & 1u);
          ptx::tcgen05_fence_after();
          // pass Q: A = Q_t / dO_t, B = K / V.   pass K: A = K_u / V_u, B = Q / dO.   (all K-major)
          const uint32_t a0 = (pass == 0 ? q_s(s) : k_s(s)) + uint32_t(t) * 128u * 128u;
          const uint32_t a1 = (pass == 0 ? do_s(s) : v_s(s)) + uint32_t(t) * 128u * 128u;
          const uint64_t da0 = ptx::make_kmajor_sw128_desc(a0), da1 = ptx::make_kmajor_sw128_desc(a1);
          const uint64_t db0 = ptx::make_kmajor_sw128_desc(pass == 0 ? k_s(s) : q_s(s));
          const uint64_t db1 = ptx::make_kmajor_sw128_desc(pass == 0 ? v_s(s) : do_s(s));
          // the two column halves are separate products with their own barriers (N = 16 * ua and 16 * (units - ua))
          const uint32_t id_h0 = idesc(128, uint32_t(16 * ua), kF16, false), id_h1 = idesc(128, uint32_t(16 * (units - ua)), kF16, false);
          const uint64_t boff = uint64_t((uint32_t(ua) * 16u * 128u) >> 4);   // B rows [16 ua, TP): descriptor address offset
          if (ptx::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16<1>(tmem_base + kColS, da0 + uint64_t(2 * kk), db0 + uint64_t(2 * kk), id_h0, kk != 0 ? 1u : 0u);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::umma_f16<1>(tmem_base + kColDp, da1 + uint64_t(2 * kk), db1 + uint64_t(2 * kk), id_h0, kk != 0 ? 1u : 0u);
            ptx::umma_commit<1>(sd_full(0));
            if (units > ua) {
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
          const uint64_t dmn0 = mn_desc(pass == 0 ? k_s(s) : do_s(s));   // dQ = dS K   |  dV = P^T dO
          const uint64_t dmn1 = mn_desc(q_s(s));                          //             |  dK = dS^T Q
          // both halves must be done before the accumulators (which overlap the top fp32 columns of the second half) are written
          ptx::mbar_wait(ds_ready(0), tile_cnt & 1u);
          ptx::mbar_wait(ds_ready(1), tile_cnt & 1u);
          ptx::tcgen05_fence_after();
          if (ptx::elect_one()) {
            if (pass == 0) {
              for (int ks = 0; ks < units; ++ks)
                umma_ts(tmem_base + kAcc0, tmem_base + kColDp + packed_col(ks), dmn0 + uint64_t(ks * 128), id_acc, ks != 0 ? 1u : 0u);
            } else {
              for (int ks = 0; ks < units; ++ks)
                umma_ts(tmem_base + kAcc0, tmem_base + kColS + packed_col(ks), dmn0 + uint64_t(ks * 128), id_acc, ks != 0 ? 1u : 0u);
              for (int ks = 0; ks < units; ++ks)
                umma_ts(tmem_base + kAcc1, tmem_base + kColDp + packed_col(ks), dmn1 + uint64_t(ks * 128), id_acc, ks != 0 ? 1u : 0u);
            }
          }
          __syncwarp();
          if (ptx::elect_one()) {
            ptx::umma_commit<1>(acc_full);
            if (pass == 1 && t == n_tiles - 1) ptx::umma_commit<1>(op_empty(s));   // last product that reads the item's operands
          }
          __syncwarp();
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
    const int u_lo = grp == 0 ? 0 : ua, u_hi = grp == 0 ? ua : units;
    const float c = prm.scale_log2e, scale = prm.scale;
    const int ctid = threadIdx.x - 128;                    // 0..255
    uint32_t tile_cnt = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / H, h = item - b * H;
      // ---- per-item vectors: lse (log2 domain) and D of every query; padded queries get lse = +inf (P = 0), D = 0 ----
      asm volatile("bar.sync 1, 256;" ::: "memory");       // previous item's readers are done with the vectors
      if (ctid < 256) {
        const bool ok = ctid < T;
        s_lse[ctid] = ok ? prm.lse[(size_t(b) * H + h) * T + ctid] : INFINITY;
        s_d[ctid] = ok ? prm.dsum[(size_t(b) * H + h) * T + ctid] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < n_tiles; ++t, ++tile_cnt) {
          const int row = t * 128 + r;                     // query (pass Q) or key (pass K) this thread owns
          const bool row_ok = row < T;
          const bool warp_live = t * 128 + quad * 32 < T;  // warps whose 32 rows are all padding skip the math (rows are independent)
          const float lse_r = (pass == 0 && row < 256) ? s_lse[row] : 0.f;
          const float d_r = (pass == 0 && row < 256) ? s_d[row] : 0.f;
          ptx::mbar_wait(sd_full(grp), tile_cnt & 1u);
          ptx::tcgen05_fence_after();
#pragma unroll
          for (int uu = 0; uu < kHU; ++uu) {
            const int u = u_lo + uu;
            if (u < u_hi && warp_live) {
              uint32_t sv[16], dv[16];
              ld16(lane_addr + kColS + uint32_t(16 * u), sv);
              ld16(lane_addr + kColDp + uint32_t(16 * u), dv);
              ptx::tmem_ld_wait();
              uint32_t pp[8], pds[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c0 = 16 * u + 2 * e;
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
                if (pass == 0) {                           // key columns beyond T hold another image's keys
                  p0 = c0 < T ? p0 : 0.f;
                  p1 = c0 + 1 < T ? p1 : 0.f;
                }
                const float ds0 = p0 * (__uint_as_float(dv[2 * e]) - dd0) * scale;
                const float ds1 = p1 * (__uint_as_float(dv[2 * e + 1]) - dd1) * scale;
                pp[e] = Act<kF16>::pack(p0, p1);
                pds[e] = Act<kF16>::pack(ds0, ds1);
              }
              // in place: the packed unit lands in fp32 columns this group has already read (see packed_col)
              if (pass == 1) st8(lane_addr + kColS + packed_col(u), pp);
              st8(lane_addr + kColDp + packed_col(u), pds);
            }
          }
          st_wait();
          ptx::tcgen05_fence_before();
          ptx::mbar_arrive(ds_ready(grp));
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
            for (int j = 0; j < 8; ++j) {
              if (8 * j < ncol) {
                uint4 w;
                w.x = Act<kF16>::pack(__uint_as_float(a[8 * j]), __uint_as_float(a[8 * j + 1]));
                w.y = Act<kF16>::pack(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3]));
                w.z = Act<kF16>::pack(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5]));
                w.w = Act<kF16>::pack(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7]));
                *reinterpret_cast<uint4*>(dst + 8 * j) = w;
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

// returns -3 when the shape is outside this kernel's envelope (caller falls back to the mma.sync kernel)
int launch_attention_bwd_sm100(const void* qkv, const void* out, const void* d_out, const float* lse, float* dsum_scratch,
                               void* dqkv, int B, int T, int H, int head_dim, int f16, int num_sms, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T < 1 || T > 256 || dsum_scratch == nullptr) return -3;
  BwdParams p;
  p.TP = (T + 15) / 16 * 16;
  const int tile_bytes = p.TP * 128;
  const int fixed = 2048 + kBarBytes + 1024 /*alignment slack*/;
  // an A operand always covers 128 rows: the last tile of a TP-row buffer is read up to row n_tiles * 128 - 1, past the buffer
  // (those rows only produce output rows that are never stored); keep the over-read inside the allocation
  const int n_tiles = (T + 127) / 128;
  const int over = n_tiles * 128 * 128 - tile_bytes;
  if (fixed + 4 * tile_bytes + over > kMaxSmem) return -3;
  p.stages = fixed + 8 * tile_bytes + over <= kMaxSmem ? 2 : 1;
  const int smem = fixed + p.stages * 4 * tile_bytes + over;
  p.items = B * H;
  p.T = T;
  p.H = H;
  p.lse = lse;
  p.dsum = dsum_scratch;
  p.dqkv = static_cast<uint16_t*>(dqkv);
  p.scale = 1.0f / sqrtf(float(head_dim));
  p.scale_log2e = 1.4426950408889634f * p.scale;
  const int d = H * kHd;
  {
    const long long total = (long long)B * T * H * 8;
    const int blocks = int((total + 255) / 256);
    if (f16)
      attention_bwd_dsum_kernel<true><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(d_out), static_cast<const uint16_t*>(out), dsum_scratch, B, T, H);
    else
      attention_bwd_dsum_kernel<false><<<blocks, 256, 0, stream>>>(static_cast<const uint16_t*>(d_out), static_cast<const uint16_t*>(out), dsum_scratch, B, T, H);
    if (cudaGetLastError() != cudaSuccess) return -2;
  }
  CUtensorMap tq, tdo;
  const uint64_t rows = uint64_t(B) * T;
  if (!make_map(&tq, qkv, rows, uint64_t(3 * d), uint32_t(p.TP), f16 != 0) || !make_map(&tdo, d_out, rows, uint64_t(d), uint32_t(p.TP), f16 != 0))
    return -1;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attention_bwd_sm100_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_sm100_kernel<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_sm100_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(attention_bwd_sm100_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
      return -2;
    attr_done = true;
  }
  const int grid = p.items < num_sms ? p.items : num_sms;
  if (p.TP <= 224) {
    if (f16) attention_bwd_sm100_kernel<true, 7><<<grid, kThreads, smem, stream>>>(tq, tdo, p);
    else attention_bwd_sm100_kernel<false, 7><<<grid, kThreads, smem, stream>>>(tq, tdo, p);
  } else {
    if (f16) attention_bwd_sm100_kernel<true, 8><<<grid, kThreads, smem, stream>>>(tq, tdo, p);
    else attention_bwd_sm100_kernel<false, 8><<<grid, kThreads, smem, stream>>>(tq, tdo, p);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
