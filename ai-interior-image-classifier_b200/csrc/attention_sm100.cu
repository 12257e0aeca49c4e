// tcgen05 attention for sequences that fit one TMEM accumulator (T <= 256: ViT-B/16 @ 224 has T = 197).
//
//   softmax(Q K^T / sqrt(64)) V   per (image, head)        [reference: nn.MultiheadAttention inside CLIP's
//   ResidualAttentionBlock, reached through model.encode_image at /root/reference/main.py:204, 444, 503]
//
// One persistent CTA per SM walks (image, head) items.  Per item the 197 queries form two 128-row tiles handled
// concurrently; everything between the two tensor-core products stays on chip:
//
//   warp 0   producer   TMA: Q (256 rows) + K (TP rows) and V (TP rows) of the item into 128B-swizzled smem
//   warp 1   MMA        S_g = Q_g K^T      (tcgen05.mma, M=128, N=TP, fp32 accumulator in TMEM)            g = 0, 1
//                       O_g = P_g V        (A = P_g from smem, B = V in its natural [key][dim] layout = MN-major operand;
//                                           the accumulator re-uses the TMEM columns of S_g)
//   warp 2   TMEM allocator (512 columns: two 256-column regions)
//   warps 4-7 / 8-11   softmax group g: ONE THREAD PER QUERY ROW (tcgen05.ld 32x32b gives a thread its whole row, so
//                       row max and row sum are thread-local: no shuffles at all): pass 1 max, pass 2 p = exp2(s*c - m*c),
//                       row sum, P -> 16-bit -> swizzled smem (K-major UMMA operand); later O row * 1/sum -> global.
//
// mbarrier pipelines: qk_full/qk_empty, v_full/v_empty (TMA <-> MMA), s_full[g] / p_full[g] / o_full[g] / s_free[g]
// (MMA <-> softmax group g).  The loads of item i+1 start as soon as the MMAs that read Q/K (resp. V) of item i retire,
// so they overlap the softmax of item i.  All waits are bounded (trap, never hang).
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
constexpr int kQBytes = 256 * 128;        // two 128-row query tiles
constexpr int kKvBytesMax = 256 * 128;    // TP <= 256 rows of 128 bytes
constexpr int kPBlockBytes = 128 * 128;   // [128 rows x 64 keys] 16-bit, one 128B swizzle span per row
constexpr int kPBytes = 4 * kPBlockBytes; // up to 256 keys
constexpr int kSmemTotal = kQBytes + 2 * kKvBytesMax + 2 * kPBytes + 1024 /*barriers*/ + 1024 /*align*/;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// instruction descriptor: D f32, A/B 16-bit, A K-major, B K-major (b_mn = false) or MN-major (b_mn = true)
__host__ __device__ constexpr uint32_t idesc(uint32_t m, uint32_t n, bool f16, bool b_mn) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((b_mn ? 1u : 0u) << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}

}  // namespace

template <bool kF16>
__global__ void __launch_bounds__(kThreads, 1)
attention_sm100_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                       uint16_t* __restrict__ out, int items, int T, int TP, int H, float scale_log2e, uint32_t v_lbo_enc,
                       uint32_t v_sbo_enc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t q_s = base, k_s = q_s + kQBytes, v_s = k_s + kKvBytesMax, p_s = v_s + kKvBytesMax;
  const uint32_t bar = p_s + 2 * kPBytes;
  uint8_t* p_gen = gen + kQBytes + 2 * kKvBytesMax;
  // barrier slots
  const uint32_t qk_full = bar, qk_empty = bar + 8, v_full = bar + 16, v_empty = bar + 24;
  auto s_full = [&](int g) { return bar + 32 + 8u * g; };
  auto p_full = [&](int g) { return bar + 48 + 8u * g; };
  auto o_full = [&](int g) { return bar + 64 + 8u * g; };
  auto s_free = [&](int g) { return bar + 80 + 8u * g; };
  const uint32_t tmem_slot = bar + 96;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + kQBytes + 2 * kKvBytesMax + 2 * kPBytes + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = H * kHd;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_q);
    ptx::prefetch_tensormap(&tm_kv);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(qk_full, 1); ptx::mbar_init(qk_empty, 1); ptx::mbar_init(v_full, 1); ptx::mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(s_full(g), 1);    // tcgen05.commit
      ptx::mbar_init(p_full(g), 1);    // one thread of the group after the group barrier
      ptx::mbar_init(o_full(g), 1);    // tcgen05.commit
      ptx::mbar_init(s_free(g), 1);    // one thread of the group after the group barrier
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const uint32_t kv_bytes = uint32_t(TP) * 128u;
  const int n_pblk = (TP + 63) / 64;      // 64-key P blocks
  const int n_kstep = TP / 16;            // UMMA K-steps of the P.V product

  if (warp == 0) {
    // ======================= producer =======================
    if (ptx::elect_one()) {
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1u) {
        const int b = item / H, h = item - b * H;
        const int row0 = b * T;
        ptx::mbar_wait(qk_empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(qk_full, kQBytes + kv_bytes);
        ptx::tma_load_2d(&tm_q, qk_full, q_s, h * kHd, row0, ptx::kEvictFirst);
        ptx::tma_load_2d(&tm_kv, qk_full, k_s, d + h * kHd, row0, ptx::kEvictFirst);
        ptx::mbar_wait(v_empty, ph ^ 1u);
        ptx::mbar_arrive_expect_tx(v_full, kv_bytes);
        ptx::tma_load_2d(&tm_kv, v_full, v_s, 2 * d + h * kHd, row0, ptx::kEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    // One thread serves both query-tile groups as a small event loop, so that group g's chain
    //   S_g(i) -> softmax -> P_g V(i) -> O_g read -> S_g(i+1) ...
    // never waits behind the other group's products (the two chains only meet at the single-buffered Q/K and V tiles).
    if (ptx::elect_one()) {
      const uint32_t idesc_s = idesc(128, uint32_t(TP), kF16, false);
      const uint32_t idesc_o = idesc(128, kHd, kF16, true);
      const uint64_t dk = ptx::make_kmajor_sw128_desc(k_s);
      const int n_items = (items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
      int it[2] = {0, 0};        // item counter per group
      int stage[2] = {0, 0};     // 0: S_g(it) to issue, 1: P_g V(it) to issue
      int s_issued = 0, pv_issued = 0;   // products issued so far (2 per item): drive qk_empty / v_empty
      const long long t0 = clock64();
      while (it[0] < n_items || it[1] < n_items) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (it[g] >= n_items) continue;
          const uint32_t ph = uint32_t(it[g]) & 1u;
          if (stage[g] == 0) {
            // Q/K of this item landed, and group g has drained O_g of its previous item
            if (!ptx::mbar_test_wait(qk_full, ph) || !ptx::mbar_test_wait(s_free(g), ph ^ 1u)) continue;
            ptx::tcgen05_fence_after();
            const uint64_t dq = ptx::make_kmajor_sw128_desc(q_s + g * (128 * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16<1>(tmem_base + g * 256, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k != 0 ? 1u : 0u);
            ptx::umma_commit<1>(s_full(g));
            if ((++s_issued & 1) == 0) ptx::umma_commit<1>(qk_empty);   // both S products of the item issued: Q/K reusable
            stage[g] = 1;
            progress = true;
          } else {
            if (!ptx::mbar_test_wait(v_full, ph) || !ptx::mbar_test_wait(p_full(g), ph)) continue;
            ptx::tcgen05_fence_after();
            for (int j = 0; j < n_kstep; ++j) {
              const uint64_t dp = ptx::make_kmajor_sw128_desc(p_s + g * kPBytes + (j >> 2) * kPBlockBytes) + uint64_t(2 * (j & 3));
              // V rows are keys: 16 keys per K-step = 2048 bytes further down the [key][64 dims] tile (MN-major operand)
              uint64_t dv = uint64_t(((v_s + uint32_t(j) * 2048u) >> 4) & 0x3FFFu);
              dv |= uint64_t(v_lbo_enc) << 16;
              dv |= uint64_t(v_sbo_enc) << 32;
              dv |= uint64_t(1) << 46;
              dv |= uint64_t(2) << 61;
              ptx::umma_f16<1>(tmem_base + g * 256, dp, dv, idesc_o, j != 0 ? 1u : 0u);
            }
            ptx::umma_commit<1>(o_full(g));
            if ((++pv_issued & 1) == 0) ptx::umma_commit<1>(v_empty);
            stage[g] = 0;
            ++it[g];
            progress = true;
          }
        }
        if (!progress && (clock64() - t0) > 20 * IIC_MBAR_TIMEOUT_CYCLES) {
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
    const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(g * 256);
    uint8_t* my_p = p_gen + g * kPBytes + r * 128;  // this row inside each 64-key P block
    const int sw = r & 7;
    const int bar_id = 1 + g;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ph ^= 1u) {
      const int b = item / H, h = item - b * H;
      ptx::mbar_wait(s_full(g), ph);
      ptx::tcgen05_fence_after();
      // TMEM -> register bandwidth (64 B/clk per SM) is what bounds this kernel, so S is read exactly ONCE: an online
      // softmax over 32-column chunks whose running maximum is kept as an INTEGER power of two, m = ceil(max * c).  A P
      // chunk written under an older (smaller) m is later fixed up in shared memory by an exact multiplication with
      // 2^(m_old - m_final) - rare after the first chunks, and exact in bf16/fp16, so the result does not depend on the
      // chunking.  The next chunk's tcgen05.ld is always in flight while the current one is processed.
      const int n_ch = (TP + 31) >> 5;
      const bool warp_live = g * 128 + quad * 32 < T;   // warps whose 32 query rows are all padding do no math
      uint32_t va[32], vb[32];
      float m_run = -126.f, sum = 0.f;                  // exp2(s*c - m) stays a normal fp32 for any m >= -126
      float m_used[8];
      auto emit = [&](uint32_t (&v)[32], int c) {
        float p[32];
        if (!warp_live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) p[i] = 0.f;
        } else {
          const bool full = (c + 1) * 32 <= T;
          float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            c0 = fmaxf(c0, (full || c * 32 + i < T) ? __uint_as_float(v[i]) : -INFINITY);
            c1 = fmaxf(c1, (full || c * 32 + i + 1 < T) ? __uint_as_float(v[i + 1]) : -INFINITY);
            c2 = fmaxf(c2, (full || c * 32 + i + 2 < T) ? __uint_as_float(v[i + 2]) : -INFINITY);
            c3 = fmaxf(c3, (full || c * 32 + i + 3 < T) ? __uint_as_float(v[i + 3]) : -INFINITY);
          }
          const float mi = ceilf(fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * scale_log2e);
          if (mi > m_run) {
            sum *= exp2f(m_run - mi);   // exact power of two
            m_run = mi;
          }
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float e0, e1, e2, e3;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(__uint_as_float(v[i]), scale_log2e, -m_run)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(__uint_as_float(v[i + 1]), scale_log2e, -m_run)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fmaf(__uint_as_float(v[i + 2]), scale_log2e, -m_run)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fmaf(__uint_as_float(v[i + 3]), scale_log2e, -m_run)));
            if (!full) {
              e0 = c * 32 + i < T ? e0 : 0.f;     e1 = c * 32 + i + 1 < T ? e1 : 0.f;
              e2 = c * 32 + i + 2 < T ? e2 : 0.f; e3 = c * 32 + i + 3 < T ? e3 : 0.f;
            }
            s0 += e0; s1 += e1; s2 += e2; s3 += e3;
            p[i] = e0; p[i + 1] = e1; p[i + 2] = e2; p[i + 3] = e3;
          }
          sum += (s0 + s1) + (s2 + s3);
        }
        m_used[c] = m_run;
        uint8_t* blk = my_p + (c >> 1) * kPBlockBytes;
        const int j0 = (c & 1) * 4;   // first 16-byte chunk of this 32-key run inside the 64-key block
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = Act<kF16>::pack(p[8 * j], p[8 * j + 1]);
          q.y = Act<kF16>::pack(p[8 * j + 2], p[8 * j + 3]);
          q.z = Act<kF16>::pack(p[8 * j + 4], p[8 * j + 5]);
          q.w = Act<kF16>::pack(p[8 * j + 6], p[8 * j + 7]);
          *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ sw) << 4)) = q;
        }
      };
      ptx::tmem_ld_32x32b_x32(taddr, va);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        if (c < n_ch) {
          if (c + 1 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((c + 1) * 32), vb);
          emit(va, c);
          ptx::tmem_ld_wait();
          if (c + 1 < n_ch) {
            if (c + 2 < n_ch) ptx::tmem_ld_32x32b_x32(taddr + uint32_t((c + 2) * 32), va);
            emit(vb, c + 1);
            ptx::tmem_ld_wait();
          }
        }
      }
      // fix-up of the chunks written before the row maximum settled (this thread's own row only)
      if (warp_live) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < n_ch && m_used[c] != m_run) {
            const float f = exp2f(fmaxf(m_used[c] - m_run, -24.f));   // power of two, representable in bf16 and fp16
            const uint32_t f2 = Act<kF16>::pack(f, f);
            uint8_t* blk = my_p + (c >> 1) * kPBlockBytes;
            const int j0 = (c & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4* ptr = reinterpret_cast<uint4*>(blk + (((j0 + j) ^ sw) << 4));
              uint4 q = *ptr;
              q.x = Act<kF16>::mul2(q.x, f2); q.y = Act<kF16>::mul2(q.y, f2);
              q.z = Act<kF16>::mul2(q.z, f2); q.w = Act<kF16>::mul2(q.w, f2);
              *ptr = q;
            }
          }
        }
      }
      // P_g complete (and S_g fully read): make it visible to the tensor core, then hand over
      ptx::tcgen05_fence_before();
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if ((threadIdx.x & 127) == 0) ptx::mbar_arrive(p_full(g));
      // ---- O row ----
      ptx::mbar_wait(o_full(g), ph);
      ptx::tcgen05_fence_after();
      uint32_t o0[32], o1[32];
      ptx::tmem_ld_32x32b_x32(taddr, o0);
      ptx::tmem_ld_32x32b_x32(taddr + 32u, o1);
      ptx::tmem_ld_wait();
      ptx::tcgen05_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if ((threadIdx.x & 127) == 0) ptx::mbar_arrive(s_free(g));   // TMEM region g may be overwritten by the next S_g
      const int q = g * 128 + r;
      if (q < T) {
        const float inv = 1.0f / sum;
        uint16_t* orow = out + (size_t(b) * T + q) * d + h * kHd;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = Act<kF16>::pack(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
          w.y = Act<kF16>::pack(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
          w.z = Act<kF16>::pack(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
          w.w = Act<kF16>::pack(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 8 * j) = w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = Act<kF16>::pack(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
          w.y = Act<kF16>::pack(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
          w.z = Act<kF16>::pack(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
          w.w = Act<kF16>::pack(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + 32 + 8 * j) = w;
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
int launch_attention_sm100(const void* qkv, void* out, int B, int T, int H, int head_dim, int f16, int num_sms,
                           cudaStream_t stream) {
  if (B <= 0) return 0;
  if (head_dim != kHd || T > 256 || T < 16) return -3;
  const int TP = (T + 15) / 16 * 16;
  const int d = H * kHd;
  CUtensorMap tq, tkv;
  const uint64_t rows = uint64_t(B) * T;
  if (!make_map(&tq, qkv, rows, uint64_t(3 * d), 256, f16 != 0) || !make_map(&tkv, qkv, rows, uint64_t(3 * d), uint32_t(TP), f16 != 0))
    return -1;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(attention_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal) != cudaSuccess ||
        cudaFuncSetAttribute(attention_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal) != cudaSuccess)
      return -2;
    attr_done = true;
  }
  // MN-major SW128 operand descriptor for V [key][64 dims]: 8-key groups are 1024 bytes apart (SBO); one 64-wide atom in N (LBO unused)
  uint32_t lbo = 1, sbo = 1024 >> 4;
  if (const char* e = getenv("IIC_ATTN_VLBO")) lbo = uint32_t(atoi(e));
  if (const char* e = getenv("IIC_ATTN_VSBO")) sbo = uint32_t(atoi(e));
  const int items = B * H;
  const int grid = items < num_sms ? items : num_sms;
  const float scale_log2e = 1.4426950408889634f / sqrtf(float(head_dim));
  uint16_t* o = static_cast<uint16_t*>(out);
  if (f16)
    attention_sm100_kernel<true><<<grid, kThreads, kSmemTotal, stream>>>(tq, tkv, o, items, T, TP, H, scale_log2e, lbo, sbo);
  else
    attention_sm100_kernel<false><<<grid, kThreads, kSmemTotal, stream>>>(tq, tkv, o, items, T, TP, H, scale_log2e, lbo, sbo);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace iic
