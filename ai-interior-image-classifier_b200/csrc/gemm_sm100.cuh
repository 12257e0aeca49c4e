// Persistent, warp-specialised tcgen05 GEMM for the ViT encoder:   D[M,N] = epilogue( A[M,K] . W[N,K]^T  (+ P[M,r] . Bl[N,r]^T) )
//
//   A  : activations, bf16 (or fp16: template switch kF16), row-major (K contiguous)           -> "K-major" UMMA operand A
//   W  : nn.Linear weight [out,in] exactly as PyTorch stores it -> "K-major" UMMA operand B (no transpose needed)
//   P  : s * (x . lora_A), bf16 [M, r_pad] produced upstream;  Bl = lora_B^T, bf16 [N, r_pad].
//        The LoRA update is one extra (short) k-block accumulated into the SAME TMEM tile as the frozen W.x
//        product, i.e. reference main.py:42-43 `linear(x) + lora(x)` happens inside the accumulator.
//
// Reference call sites this replaces: nn.MultiheadAttention in_proj/out_proj, mlp.c_fc / mlp.c_proj
// (`LoRALinear.forward`, /root/reference/main.py:42-43, `LoRALayer.forward` main.py:30-31) and visual.conv1.
//
// Structure (one CTA per SM, 256 threads):
//   warp 0      TMA producer      (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier complete_tx)
//   warp 1      MMA issuer        (one elected thread; tcgen05.mma, fp32 accumulators in TMEM; leader CTA only)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue          (tcgen05.ld -> bias / QuickGELU / residual / pos-emb -> global)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer full/empty (MMA <-> epilogue), static
// persistent tile schedule (n fastest so the CTAs working at the same time share A rows in L2).
// kCtas == 2 pairs two SMs on one 256 x kBlockN tile (tcgen05 cta_group::2): each CTA loads its own 128 rows of
// A and half of the W tile, the leader issues the MMAs for both, commits are multicast to both CTAs.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "act_types.cuh"
#include "ptx_sm100.cuh"

namespace iic {

enum GemmEpilogue : int {
  kEpiBiasBf16 = 0,      // out bf16 = acc + bias                          (attn in_proj)
  kEpiBiasGeluBf16 = 1,  // out bf16 = quick_gelu(acc + bias)              (mlp.c_fc)
  kEpiBiasResF32 = 2,    // out f32  = acc + bias + residual (in place ok) (attn.out_proj, mlp.c_proj)
  kEpiPosF32 = 3,        // out f32[row + row/G + 1] = acc + pos[row%G + 1] (visual.conv1 patch embedding)
  kEpiGeluExactBf16 = 4, // out bf16 = gelu_erf(acc + bias)                (non-OpenAI checkpoints)
};

struct GemmArgs {
  int M;             // rows of A / D
  int N;             // rows of W / columns of D
  int num_k_blocks;  // ceil(K / 64) of the frozen product
  int lora_ksteps;   // 0: no LoRA block; else r_pad/16 (1..4) UMMA K-steps from the (P, Bl) tile pair
  const float* bias;      // [N] or nullptr
  const float* residual;  // kEpiBiasResF32: f32 [M, ldc];  kEpiPosF32: pos table f32 [G+1, N]
  void* out;              // bf16 or f32, leading dimension ldc
  int ldc;
  int group;  // kEpiPosF32: G = patches per image
};

constexpr int kBlockM = 128;  // rows per CTA (TMEM lanes)
constexpr int kBlockK = 64;   // one 128-byte swizzle span of bf16
constexpr int kUmmaK = 16;

template <int kCtas, int kBlockN>
struct GemmSmem {
  static constexpr int kLoadN = kBlockN / kCtas;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (200 * 1024) / kStageBytes;  // 4 x 48 KB (1 CTA) or 6 x 32 KB (2 CTA)
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = kAccStages * kBlockN;  // 512 when kBlockN = 256
  static constexpr int kBarBytes = 1024;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024 /*alignment slack*/;
  static_assert(kTmemCols == 32 || kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512,
                "TMEM allocation must be a power of two >= 32 columns");
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)  (OpenAI CLIP QuickGELU)
  const float e = exp2f(-1.702f * 1.4426950408889634f * x);
  return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int kCtas, int kBlockN, int kEpi, bool kF16>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_al, const __grid_constant__ CUtensorMap tm_bl,
                    const GemmArgs args) {
  using S = GemmSmem<kCtas, kBlockN>;
  constexpr int kStages = S::kStages;
  constexpr int kTileM = kBlockM * kCtas;

  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * S::kStageBytes;
  auto smem_a = [&](int s) { return smem_base + s * S::kStageBytes; };
  auto smem_b = [&](int s) { return smem_base + s * S::kStageBytes + S::kABytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + S::kAccStages + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 2 * S::kAccStages);
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kStages * S::kStageBytes + 8 * (2 * kStages + 2 * S::kAccStages));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtas == 1) ? 0u : ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    if (args.lora_ksteps > 0) {
      ptx::prefetch_tensormap(&tm_al);
      ptx::prefetch_tensormap(&tm_bl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), kCtas);  // one producer arrive per CTA of the pair (on the leader's barrier)
      ptx::mbar_init(empty_bar(s), 1);     // one tcgen05.commit
    }
    for (int a = 0; a < S::kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);             // one tcgen05.commit
      ptx::mbar_init(tempty_bar(a), kCtas * 128);  // every epilogue thread of the pair
    }
    ptx::fence_mbar_init();
  }
  if constexpr (kCtas > 1) ptx::cluster_sync();  // peer barriers exist before anybody signals them / allocs TMEM
  if (warp == 2) ptx::tmem_alloc<kCtas>(tmem_slot, S::kTmemCols);
  ptx::tcgen05_fence_before();
  if constexpr (kCtas > 1) ptx::cluster_sync(); else __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int m_tiles = (args.M + kTileM - 1) / kTileM;
  const int n_tiles = (args.N + kBlockN - 1) / kBlockN;
  const int total_tiles = m_tiles * n_tiles;
  const int cluster_id = blockIdx.x / kCtas;
  const int num_clusters = gridDim.x / kCtas;
  const int k_iters = args.num_k_blocks + (args.lora_ksteps > 0 ? 1 : 0);

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int m_idx = m_blk * kTileM + int(cta_rank) * kBlockM;
        const int n_idx = n_blk * kBlockN + int(cta_rank) * S::kLoadN;
        for (int kb = 0; kb < k_iters; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const bool is_lora = kb >= args.num_k_blocks;
          const CUtensorMap* ma = is_lora ? &tm_al : &tm_a;
          const CUtensorMap* mb = is_lora ? &tm_bl : &tm_b;
          const int k_idx = is_lora ? 0 : kb * kBlockK;
          if constexpr (kCtas == 1) {
            ptx::mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes);
            ptx::tma_load_2d(ma, full_bar(stage), smem_a(stage), k_idx, m_idx, ptx::kEvictNormal);
            ptx::tma_load_2d(mb, full_bar(stage), smem_b(stage), k_idx, n_idx, ptx::kEvictLast);
          } else {
            // both CTAs' bytes are accounted on the leader's barrier
            if (leader) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * S::kStageBytes);
            else ptx::mbar_arrive_cluster(full_bar(stage), 0);
            ptx::tma_load_2d_2sm(ma, full_bar(stage), smem_a(stage), k_idx, m_idx, ptx::kEvictNormal);
            ptx::tma_load_2d_2sm(mb, full_bar(stage), smem_b(stage), k_idx, n_idx, ptx::kEvictLast);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();  // reconverge before the (.aligned) teardown barriers
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA) =======================
    if (leader && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(kTileM, kBlockN, kF16);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * kBlockN);
        for (int kb = 0; kb < k_iters; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(smem_a(stage));
          const uint64_t db = ptx::make_kmajor_sw128_desc(smem_b(stage));
          const int ksteps = (kb >= args.num_k_blocks) ? args.lora_ksteps : (kBlockK / kUmmaK);
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k) {
            // advance 16 bf16 = 32 bytes inside the swizzle span: +2 in the (addr >> 4) field
            ptx::umma_f16<kCtas>(tmem_d, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit<kCtas>(empty_bar(stage));  // frees the smem slot (both CTAs) when these MMAs retire
          if (kb == k_iters - 1) ptx::umma_commit<kCtas>(tfull_bar(acc));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ======================= epilogue =======================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int it = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_blk * kTileM + int(cta_rank) * kBlockM + quad * 32 + lane;
      const int col0 = n_blk * kBlockN;
      const bool row_ok = row < args.M;

      long long out_row = row;
      const float* addend = nullptr;  // residual row or pos-emb row
      if constexpr (kEpi == kEpiBiasResF32) {
        addend = args.residual + size_t(row) * args.ldc;
      } else if constexpr (kEpi == kEpiPosF32) {
        const int img = row / args.group;
        out_row = row + img + 1;
        addend = args.residual + size_t(row - img * args.group + 1) * args.N;
      }
      if constexpr (kEpi == kEpiBiasResF32) {
        // The residual tile of the NEXT tile of this CTA: pull it into L2 while this tile's math runs so the
        // epilogue's dependent loads see L2 latency, not HBM latency.
        const int nt = tile + num_clusters;
        if (nt < total_tiles) {
          const int nm = nt / n_tiles, nn = nt - nm * n_tiles;
          const int nrow = nm * kTileM + int(cta_rank) * kBlockM + quad * 32 + lane;
          if (nrow < args.M) {
            const char* pr = reinterpret_cast<const char*>(args.residual + size_t(nrow) * args.ldc + nn * kBlockN);
#pragma unroll
            for (int i = 0; i < kBlockN * 4 / 128; ++i)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(pr + i * 128));
          }
        }
      }

      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * kBlockN);

#pragma unroll 1
      for (int c = 0; c < kBlockN / 32; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + uint32_t(c * 32), v);
        ptx::tmem_ld_wait();
        if (c == kBlockN / 32 - 1) {
          // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
          ptx::tcgen05_fence_before();
          if constexpr (kCtas == 1) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_cluster(tempty_bar(acc), 0);
        }
        const int col = col0 + c * 32;
        if (!row_ok || col >= args.N) continue;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if constexpr (kEpi != kEpiPosF32) {
          if (args.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + col + i));
              f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
          }
        }
        if constexpr (kEpi == kEpiBiasResF32 || kEpi == kEpiPosF32) {
          const float* ad = addend + (kEpi == kEpiBiasResF32 ? col : col);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 r = *reinterpret_cast<const float4*>(ad + i);
            f[i] += r.x; f[i + 1] += r.y; f[i + 2] += r.z; f[i + 3] += r.w;
          }
          float* o = reinterpret_cast<float*>(args.out) + size_t(out_row) * args.ldc + col;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        } else {
          if constexpr (kEpi == kEpiBiasGeluBf16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = quick_gelu(f[i]);
          } else if constexpr (kEpi == kEpiGeluExactBf16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
          }
          uint16_t* o = reinterpret_cast<uint16_t*>(args.out) + size_t(out_row) * args.ldc + col;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            pk.x = Act<kF16>::pack(f[i], f[i + 1]);
            pk.y = Act<kF16>::pack(f[i + 2], f[i + 3]);
            pk.z = Act<kF16>::pack(f[i + 4], f[i + 5]);
            pk.w = Act<kF16>::pack(f[i + 6], f[i + 7]);
            *reinterpret_cast<uint4*>(o + i) = pk;
          }
        }
      }
    }
  }

  // ======================= teardown =======================
  ptx::tcgen05_fence_before();
  if constexpr (kCtas > 1) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kCtas>(tmem_base, S::kTmemCols);
}

}  // namespace iic
