// Persistent, warp-specialised tcgen05 GEMM for the ViT encoder:   D[M,N] = epilogue( A[M,K] . W[N,K]^T  (+ P[M,r] . Bl[N,r]^T) )
//
//   A  : activations, bf16 (or fp16: template switch kF16), row-major (K contiguous) -> "K-major" UMMA operand A
//   W  : nn.Linear weight [out,in] exactly as PyTorch stores it -> "K-major" UMMA operand B (no transpose needed)
//   P  : s * (x . lora_A), 16-bit [M, r_pad] produced upstream;  Bl = lora_B^T, 16-bit [N, r_pad].
//        The LoRA update is one extra (short) k-block accumulated into the SAME TMEM tile as the frozen W.x
//        product, i.e. reference main.py:42-43 `linear(x) + lora(x)` happens inside the accumulator.
//
// Reference call sites this replaces: nn.MultiheadAttention in_proj/out_proj, mlp.c_fc / mlp.c_proj
// (`LoRALinear.forward`, /root/reference/main.py:42-43, `LoRALayer.forward` main.py:30-31) and visual.conv1.
//
// Structure (one CTA per SM, 384 threads):
//   warp 0      TMA producer of the A/W ring (cp.async.bulk.tensor, 128B-swizzled tiles, mbarrier complete_tx)
//   warp 1      MMA issuer        (one elected thread; tcgen05.mma, fp32 accumulators in TMEM; leader CTA only)
//   warp 2      TMEM allocator
//   warp 3      residual producer (fp32-residual epilogue only): TMA-loads 128x32 fp32 slabs of the residual stream
//               into the epilogue staging ring ahead of the epilogue
//   warps 4..11 epilogue, two groups of 4 warps (one warp per TMEM lane quadrant each): group g owns the slabs s with
//               s % 2 == g of every tile: tcgen05.ld -> bias / QuickGELU / residual -> swizzled smem slab -> TMA store.
//               Two warps per scheduler let one group's MUFU / TMEM latency hide under the other's ALU work.
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM double buffer full/empty (MMA <-> epilogue), staging ring
// (residual TMA load -> epilogue threads -> TMA store), static persistent tile schedule (n fastest so the CTAs working
// at the same time share A rows in L2).
// kCtas == 2 pairs two SMs on one 256 x kBlockN tile (tcgen05 cta_group::2): each CTA loads its own 128 rows of
// A and half of the W tile, the leader issues the MMAs for both, commits are multicast to both CTAs.
//
// Why the epilogue goes through shared memory + TMA: with one thread per accumulator row, direct global stores touch
// 32 different 128-byte lines per warp instruction (and the fp32 residual read does the same with a dependent load);
// ncu on the first version showed the tensor pipe only 21-54 % active with the tile time set by the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "act_types.cuh"
#include "ptx_sm100.cuh"

namespace iic {

enum GemmEpilogue : int {
  kEpiBiasBf16 = 0,      // out 16-bit = acc + bias                        (attn in_proj)
  kEpiBiasGeluBf16 = 1,  // out 16-bit = quick_gelu(acc + bias)            (mlp.c_fc)
  kEpiBiasResF32 = 2,    // out f32  = acc + bias + residual (in place ok) (attn.out_proj, mlp.c_proj)
  kEpiPosF32 = 3,        // out f32[row + row/G + 1] = acc + pos[row%G + 1] (visual.conv1 patch embedding)
  kEpiGeluExactBf16 = 4, // out 16-bit = gelu_erf(acc + bias)              (non-OpenAI checkpoints)
  kEpiBiasResF32DeepK = 5,  // kEpiBiasResF32 with a deeper operand ring and a shorter residual ring: chosen by the
                            // launcher for long reductions (K >= 2048: mlp.c_proj), where TMA look-ahead matters more
  // (6, 7: the LayerNorm fused behind the residual GEMM - built in three versions in round 1, parity-tested, measured, never a
  // win on the power-capped step, removed in round 2; history in DESIGN.md)
  // training: out 16-bit = acc * act'(u), u = the forward's pre-activation tile, TMA-loaded (16-bit, through tm_res) into the
  // staging slab ahead of the epilogue: dU = (dY . W2 + dP2 . A2^T) o act'(U) without a round trip of dH through HBM.
  // GemmArgs.group = activation (1 QuickGELU, 2 erf GELU).
  kEpiActGradBf16 = 8,
  // training forward of mlp.c_fc: out 16-bit = act(acc + bias) AND out2 16-bit = acc + bias (the pre-activation the
  // backward needs) from one pass over the accumulator.  GemmArgs.group = activation (1 QuickGELU, 2 erf GELU).
  kEpiBiasActDualBf16 = 9,
};

struct GemmArgs {
  int M;             // rows of A / D
  int N;             // rows of W / columns of D
  int num_k_blocks;  // ceil(K / 64) of the frozen product
  int lora_ksteps;   // 0: no LoRA block; else r_pad/16 (1..4) UMMA K-steps from the (P, Bl) tile pair
  const float* bias;      // [N] or nullptr
  const float* residual;  // kEpiPosF32: pos table f32 [G+1, N] (kEpiBiasResF32 reads the residual through tm_res)
  void* out;              // kEpiPosF32 only: f32 base pointer, leading dimension ldc (other modes store through tm_out)
  int ldc;
  int group;  // kEpiPosF32: G = patches per image
  // activation epilogues: per-tile partial of the consumer's LoRA down-projection, part[n_blk][row][0..3] (see kernels.h)
  const float* down_a;
  float* down_part;
};

constexpr int kBlockM = 128;  // rows per CTA (TMEM lanes)
constexpr int kBlockK = 64;   // one 128-byte swizzle span of bf16
constexpr int kUmmaK = 16;
constexpr int kSlabBytes = kBlockM * 128;  // epilogue staging slab: 128 rows x 128 bytes (64 x 16-bit or 32 x fp32)

template <int kCtas, int kBlockN, int kEpi>
struct GemmSmem {
  static constexpr bool kDeepK = kEpi == kEpiBiasResF32DeepK;
  static constexpr bool kResidual = kEpi == kEpiBiasResF32 || kEpi == kEpiBiasResF32DeepK;
  static constexpr bool kActGrad = kEpi == kEpiActGradBf16;
  static constexpr bool kDual = kEpi == kEpiBiasActDualBf16;
  static constexpr bool kSlabLoad = kResidual || kActGrad;   // warp 3 TMA-loads a tile of a second input into the staging slabs
  static constexpr bool kDirect = kEpi == kEpiPosF32;  // old direct-store epilogue, no staging ring
  static constexpr int kLoadN = kBlockN / kCtas;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kLoadN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // epilogue warp groups working on alternate slabs of a tile: 2 for the MUFU-heavy activation epilogues (one group's
  // MUFU / TMEM latency hides under the other's ALU work), 1 otherwise (measured: a second group only adds contention
  // for the bias-only and the residual epilogues)
  // (narrow tiles of the small-batch path: a group needs at least one staging slab of the tile)
  static constexpr int kSlabsPerTile = kBlockN / (kResidual ? 32 : 64);
  static constexpr int kGroups =
      ((kEpi == kEpiBiasGeluBf16 || kEpi == kEpiGeluExactBf16 || kActGrad || kDual) && kSlabsPerTile >= 2) || kDirect ? 2 : 1;
  static constexpr int kSlabs =   // staging ring depth
      kDirect ? 0 : (kResidual ? (kDeepK ? 3 : 4) : ((kActGrad || kDual) ? 4 : 2));
  static constexpr int kBufPerGroup = kDirect ? 1 : kSlabs / kGroups;
  static constexpr int kRingBudget =
      (kDeepK ? 208 : 192) * 1024 - kSlabs * kSlabBytes;
  static constexpr int kStages = kRingBudget / kStageBytes;  // 2 CTA: 6 / 5 / 4;  1 CTA: 4 / 3 / 2 ... see static_assert
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = kAccStages * kBlockN;  // 512 when kBlockN = 256
  static constexpr int kBarBytes = 1024;
  // consumer LoRA-A slice of the current tile [kBlockN][4] fp32 + the tile's bias slice [kBlockN] fp32 (activation epilogues)
  static constexpr int kDownBytes = kBlockN * 20;
  static constexpr int kTotal = kStages * kStageBytes + kSlabs * kSlabBytes + kBarBytes + kDownBytes +
                                1024 /*alignment slack*/;
  static_assert(kStages >= 2, "smem ring too shallow");
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
  static_assert(kTmemCols == 32 || kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512,
                "TMEM allocation must be a power of two >= 32 columns");
};

__device__ __forceinline__ float quick_gelu(float x) {
  // x * sigmoid(1.702 x)  (OpenAI CLIP QuickGELU) with sigmoid(z) = 0.5 + 0.5 tanh(z / 2):  h = 0.5x + 0.5x * tanh(0.851 x)
  // = FMUL, MUFU.TANH, FMUL, FFMA: ONE MUFU op per element instead of two (ex2 + rcp) - the MUFU pipe (16/clk/SM) is what
  // bounds the c_fc epilogue.  Measured on the B200 against double precision over [-12, 12] (tools/ubench_sm100.cu): max
  // absolute error 6.0e-6 (ex2+rcp form: 7.6e-7), i.e. far below the 16-bit rounding of the result.
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// (o0, o1) = (a0 + b0, a1 + b1) as one FADD2 (same rounding as two FADDs)
__device__ __forceinline__ void add2(float& o0, float& o1, float a0, float a1, float b0, float b1) {
  uint64_t a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(a));
}
// Two QuickGELUs on the packed fp32 pipe (FMUL2, FMUL2, FFMA2 + two MUFU.TANH): 2.5 issue slots per element instead of 4.
// mul/fma.rn.f32x2 round each lane exactly like the scalar forms, so the result is bit-identical to quick_gelu().
__device__ __forceinline__ void quick_gelu2(float& a, float& b) {
  uint64_t x, z, hx, t;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(0x3F59DB233F59DB23ull));   // 0.851f in both lanes
  float z0, z1, t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(z0), "=f"(z1) : "l"(z));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(t0), "f"(t1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(hx) : "l"(x), "l"(0x3F0000003F000000ull));  // 0.5f
  asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(x) : "l"(hx), "l"(t));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
}
// (a, b) *= QuickGELU'(u0), QuickGELU'(u1) on the packed pipe; lane arithmetic and rounding order identical to
//   s = fmaf(0.5, tanh(0.851 u), 0.5);  g = fmaf((1.702 u) s, 1 - s, s);  a *= g
__device__ __forceinline__ void quick_gelu_grad_mul2(float& a, float& b, float u0, float u1) {
  uint64_t u, z, t, sg, w, om, g, ab;
  asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(u0), "f"(u1));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(u), "l"(0x3F59DB233F59DB23ull));    // 0.851f
  float z0, z1, t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(z0), "=f"(z1) : "l"(z));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(z0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(z1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "f"(t0), "f"(t1));
  asm("fma.rn.f32x2 %0, %1, %2, %1;" : "=l"(sg) : "l"(0x3F0000003F000000ull), "l"(t));            // 0.5 t + 0.5
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w) : "l"(u), "l"(0x3FD9DB233FD9DB23ull));    // 1.702f u
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(w) : "l"(w), "l"(sg));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(om) : "l"(sg), "l"(0xBF800000BF800000ull), "l"(0x3F8000003F800000ull));   // 1 - s (exact as in the scalar form)
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(g) : "l"(w), "l"(om), "l"(sg));
  asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(a), "f"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ab) : "l"(ab), "l"(g));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(ab));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

constexpr int kGemmThreads = 384;
// named barrier of one epilogue group (ids 1, 2) / of both groups (id 3)
__device__ __forceinline__ void epi_bar_sync(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }
__device__ __forceinline__ void epi_bar_sync_all() { asm volatile("bar.sync 3, 256;" ::: "memory"); }

template <int kCtas, int kBlockN, int kEpi, bool kF16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ CUtensorMap tm_al, const __grid_constant__ CUtensorMap tm_bl,
                    const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
                    const __grid_constant__ CUtensorMap tm_out2, const GemmArgs args) {
  using S = GemmSmem<kCtas, kBlockN, kEpi>;
  constexpr int kStages = S::kStages;
  constexpr int kSlabs = S::kSlabs;
  constexpr int kTileM = kBlockM * kCtas;
  constexpr bool kResidual = S::kResidual;
  constexpr bool kDirect = S::kDirect;
  constexpr bool kActGrad = S::kActGrad;
  constexpr bool kDual = S::kDual;
  constexpr bool kAct = kEpi == kEpiBiasGeluBf16 || kEpi == kEpiGeluExactBf16 || kDual;
  constexpr int kSlabCols = kResidual ? 32 : 64;           // columns per staging slab
  constexpr int kSlabsPerTile = kBlockN / kSlabCols;       // 8 (fp32) or 4 (16-bit)

  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const uint32_t slab_base = smem_base + kStages * S::kStageBytes;
  const uint32_t bar_base = slab_base + kSlabs * kSlabBytes;
  auto smem_a = [&](int s) { return smem_base + s * S::kStageBytes; };
  auto smem_b = [&](int s) { return smem_base + s * S::kStageBytes + S::kABytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + S::kAccStages + a); };
  auto rfull_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 2 * S::kAccStages + b); };
  auto rempty_bar = [&](int b) { return bar_base + 8u * (2 * kStages + 2 * S::kAccStages + 4 + b); };
  constexpr int kTmemSlotOff = 8 * (2 * kStages + 2 * S::kAccStages + 12);
  const uint32_t tmem_slot = bar_base + kTmemSlotOff;
  uint8_t* bar_gen = smem_gen + kStages * S::kStageBytes + kSlabs * kSlabBytes;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(bar_gen + kTmemSlotOff);
  float4* down_s = reinterpret_cast<float4*>(bar_gen + S::kBarBytes);
  float* bias_s = reinterpret_cast<float*>(bar_gen + S::kBarBytes + kBlockN * 16);
  uint8_t* slab_gen = smem_gen + kStages * S::kStageBytes;

  ptx::pdl_launch_dependents();   // the next kernel of the stream may start its prologue (PDL launches only)
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
    if constexpr (!kDirect) ptx::prefetch_tensormap(&tm_out);
    if constexpr (S::kSlabLoad) ptx::prefetch_tensormap(&tm_res);
    if constexpr (kDual) ptx::prefetch_tensormap(&tm_out2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), kCtas);  // one producer arrive per CTA of the pair (on the leader's barrier)
      ptx::mbar_init(empty_bar(s), 1);     // one tcgen05.commit
    }
    for (int a = 0; a < S::kAccStages; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);             // one tcgen05.commit
      ptx::mbar_init(tempty_bar(a), kCtas * 128 * S::kGroups);  // every working epilogue thread of the pair
    }
    for (int b = 0; b < 4; ++b) {
      ptx::mbar_init(rfull_bar(b), 1);   // residual producer's arrive.expect_tx
      ptx::mbar_init(rempty_bar(b), 1);  // the storing epilogue thread, once the TMA store has read the slab
    }
    ptx::fence_mbar_init();
  }
  if constexpr (kCtas > 1) ptx::cluster_sync();  // peer barriers exist before anybody signals them / allocs TMEM
  if (warp == 2) ptx::tmem_alloc<kCtas>(tmem_slot, S::kTmemCols);
  ptx::tcgen05_fence_before();
  if constexpr (kCtas > 1) ptx::cluster_sync(); else __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // barriers, TMEM and descriptors are set up: from here on global memory written by the previous kernel is read
  ptx::pdl_wait();

  const int m_tiles = (args.M + kTileM - 1) / kTileM;
  const int n_tiles = (args.N + kBlockN - 1) / kBlockN;
  const int total_tiles = m_tiles * n_tiles;
  const int cluster_id = blockIdx.x / kCtas;
  const int num_clusters = gridDim.x / kCtas;
  const int k_iters = args.num_k_blocks + (args.lora_ksteps > 0 ? 1 : 0);
  // static persistent schedule, `it` = this cluster's running tile counter: tile = cluster + it * clusters with n fastest
  auto tile_at = [&](int it, int& m_blk, int& n_blk) -> bool {
    const int tile = cluster_id + it * num_clusters;
    if (tile >= total_tiles) return false;
    m_blk = tile / n_tiles;
    n_blk = tile - m_blk * n_tiles;
    return true;
  };

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int m_blk, n_blk;
      for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
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
      int m_blk, n_blk;
      for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
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
  } else if (warp == 3) {
    // ======================= residual producer (fp32-residual / activation-gradient epilogues) =======================
    if constexpr (S::kSlabLoad) {
      if (ptx::elect_one()) {
        int slab = 0;  // running slab counter of this CTA
        int m_blk, n_blk;
        for (int it = 0; tile_at(it, m_blk, n_blk); ++it) {
          const int row0 = m_blk * kTileM + int(cta_rank) * kBlockM;
          const int col0 = n_blk * kBlockN;
          for (int s = 0; s < kSlabsPerTile; ++s, ++slab) {
            const int b = slab % kSlabs;
            const uint32_t ph = uint32_t(slab / kSlabs) & 1u;
            ptx::mbar_wait(rempty_bar(b), ph ^ 1u);
            ptx::mbar_arrive_expect_tx(rfull_bar(b), kSlabBytes);
            ptx::tma_load_2d(&tm_res, rfull_bar(b), slab_base + b * kSlabBytes, col0 + s * kSlabCols, row0,
                             ptx::kEvictFirst);
          }
        }
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ======================= epilogue =======================
    constexpr int kGroups = S::kGroups;
    constexpr int kBufPG = S::kBufPerGroup;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;     // epilogue group 0 / 1
    const int r_in_tile = quad * 32 + lane;
    const bool storer = (threadIdx.x & 127) == 0;  // first thread of each group issues that group's TMA stores
    int my_slab = grp;                   // running index (over the whole kernel) of the next slab this group handles
    int m_blk = 0, n_blk = 0;
    for (int it = 0; grp < kGroups && tile_at(it, m_blk, n_blk); ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * kTileM + int(cta_rank) * kBlockM;
      const int row = row0 + r_in_tile;
      const int col0 = n_blk * kBlockN;
      const bool row_ok = row < args.M;

      if constexpr (!kDirect) {
        // stage this tile's bias slice (and, activation epilogues, its slice of the consumer's LoRA-A) while the MMAs of the tile
        // are still running: with ~all of L1 carved out as shared memory a __ldg in the slab loop is an exposed L2 round trip
        // per 32 columns (c_fc in-step 0.93 -> 0.81 ms)
        static_assert(!kAct || 128 * kGroups >= kBlockN, "one staged LoRA-A row per epilogue thread");
        constexpr int kEpiThreads = 128 * kGroups;
        const int t = threadIdx.x - 128;  // 0..kEpiThreads-1
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
        float bv0 = 0.f, bv1 = 0.f;
        if constexpr (kAct) {
          if (args.down_a != nullptr && t < kBlockN && col0 + t < args.N) v0 = __ldg(reinterpret_cast<const float4*>(args.down_a) + col0 + t);
        }
        if (args.bias != nullptr && t < kBlockN && col0 + t < args.N) bv0 = __ldg(args.bias + col0 + t);
        if (args.bias != nullptr && t + kEpiThreads < kBlockN && col0 + t + kEpiThreads < args.N)
          bv1 = __ldg(args.bias + col0 + t + kEpiThreads);
        // previous tile's readers are done with the buffers
        if constexpr (kGroups == 2) epi_bar_sync_all(); else epi_bar_sync(0);
        if constexpr (kAct) { if (t < kBlockN) down_s[t] = v0; }
        if (t < kBlockN) bias_s[t] = bv0;
        if (t + kEpiThreads < kBlockN) bias_s[t + kEpiThreads] = bv1;
        if constexpr (kGroups == 2) epi_bar_sync_all(); else epi_bar_sync(0);
      }
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acc * kBlockN);
      auto release_tmem = [&]() {
        ptx::tcgen05_fence_before();
        if constexpr (kCtas == 1) ptx::mbar_arrive(tempty_bar(acc));
        else ptx::mbar_arrive_cluster(tempty_bar(acc), 0);
      };

      if constexpr (kDirect) {
        // ---- patch embedding: rows are scattered (image boundaries), direct global stores; group g takes chunks c % 2 == g ----
        const int img = row / args.group;
        const long long out_row = (long long)row + img + 1;
        const float* addend = args.residual + size_t(row - img * args.group + 1) * args.N;
        // the positional-embedding values of the NEXT chunk are requested before this chunk's accumulator is read: with L1
        // carved out as shared memory they come from L2 (one exposed round trip per chunk otherwise)
        float4 pe[8], pe_next[8];
        auto load_pe = [&](int c, float4 (&dst)[8]) {
          const int col = col0 + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = (row_ok && col < args.N) ? __ldg(reinterpret_cast<const float4*>(addend + col) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        load_pe(grp, pe);
#pragma unroll 1
        for (int c = grp; c < kBlockN / 32; c += kGroups) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + uint32_t(c * 32), v);
          if (c + kGroups < kBlockN / 32) load_pe(c + kGroups, pe_next);
          ptx::tmem_ld_wait();
          if (c + kGroups >= kBlockN / 32) release_tmem();
          const int col = col0 + c * 32;
          if (row_ok && col < args.N) {
            float* o = reinterpret_cast<float*>(args.out) + size_t(out_row) * args.ldc + col;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 r = pe[i];
              *reinterpret_cast<float4*>(o + 4 * i) =
                  make_float4(__uint_as_float(v[4 * i]) + r.x, __uint_as_float(v[4 * i + 1]) + r.y,
                              __uint_as_float(v[4 * i + 2]) + r.z, __uint_as_float(v[4 * i + 3]) + r.w);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pe[i] = pe_next[i];
        }
      } else {
        uint64_t dacc01 = 0ull, dacc23 = 0ull;   // packed partial sums of the consumer's LoRA down-projection
#pragma unroll 1
        for (int s = grp; s < kSlabsPerTile; s += kGroups, my_slab += kGroups) {
          const int b = grp + kGroups * ((my_slab / kGroups) % kBufPG);   // this group's ring of kBufPG buffers
          const int col = col0 + s * kSlabCols;
          uint8_t* my_row = slab_gen + b * kSlabBytes + r_in_tile * 128;  // this thread's 128-byte row of the slab
          const int sw = r_in_tile & 7;                                    // 128B swizzle: 16B chunk j -> j ^ (row & 7)
          if constexpr (kResidual) {
            // ---- out = acc + bias + residual: the slab already holds the residual (TMA-loaded by warp 3) ----
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(s * 32), v);
            ptx::mbar_wait(rfull_bar(b), uint32_t(my_slab / kSlabs) & 1u);
            ptx::tmem_ld_wait();
            if (s + kGroups >= kSlabsPerTile) release_tmem();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4* p = reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4));
              float4 r = *p;
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              bb = reinterpret_cast<const float4*>(bias_s + s * 32)[j];
              float t0, t1, t2, t3;   // r += (acc + bias), as packed FADD2s (same rounding order as the scalar form)
              add2(t0, t1, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), bb.x, bb.y);
              add2(t2, t3, __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]), bb.z, bb.w);
              add2(r.x, r.y, r.x, r.y, t0, t1);
              add2(r.z, r.w, r.z, r.w, t2, t3);
              *p = r;
            }
            ptx::fence_proxy_async_smem();
            epi_bar_sync(grp);
            if (storer) {
              ptx::tma_store_2d(&tm_out, slab_base + b * kSlabBytes, col, row0);
              ptx::tma_store_commit();
              // this group's previous slab has been read by its store by now (at most 1 group still in flight): hand it back
              if (my_slab >= kGroups) {
                ptx::tma_store_wait_read<1>();
                ptx::mbar_arrive(rempty_bar((my_slab - kGroups) % kSlabs));
              }
            }
          } else if constexpr (kActGrad) {
            // ---- out = acc * act'(u): the slab already holds the 64 pre-activations of this row (TMA-loaded by warp 3) ----
            uint32_t v0[32], v1[32];
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(s * 64), v0);
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(s * 64 + 32), v1);
            ptx::mbar_wait(rfull_bar(b), uint32_t(my_slab / kSlabs) & 1u);
            ptx::tmem_ld_wait();
            if (s + kGroups >= kSlabsPerTile) release_tmem();
            const bool erf_act = args.group == 2;
            auto grad = [&](float u) {
              if (erf_act) {
                const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
                return cdf + u * 0.3989422804014327f * __expf(-0.5f * u * u);
              }
              // QuickGELU: s = sigmoid(1.702 u) = 0.5 + 0.5 tanh(0.851 u);  d/du = s + 1.702 u s (1 - s)
              float t;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * u));
              const float sg = fmaf(0.5f, t, 0.5f);
              return fmaf(1.702f * u * sg, 1.0f - sg, sg);
            };
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4* p = reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4));
              uint4 uu = *p;
              uint32_t* pu = reinterpret_cast<uint32_t*>(&uu);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 u2 = Act<kF16>::unpack(pu[e]);
                const int c = j * 8 + 2 * e;
                const float a0 = __uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]);
                const float a1 = __uint_as_float(c < 32 ? v0[(c + 1) & 31] : v1[(c + 1) & 31]);
                if (erf_act) {
                  pu[e] = Act<kF16>::pack(a0 * grad(u2.x), a1 * grad(u2.y));
                } else {
                  float o0 = a0, o1 = a1;
                  quick_gelu_grad_mul2(o0, o1, u2.x, u2.y);
                  pu[e] = Act<kF16>::pack(o0, o1);
                }
              }
              *p = uu;
            }
            ptx::fence_proxy_async_smem();
            epi_bar_sync(grp);
            if (storer) {
              ptx::tma_store_2d(&tm_out, slab_base + b * kSlabBytes, col, row0);
              ptx::tma_store_commit();
              if (my_slab >= kGroups) {
                ptx::tma_store_wait_read<1>();
                ptx::mbar_arrive(rempty_bar((my_slab - kGroups) % kSlabs));
              }
            }
          } else {
            // ---- 16-bit outputs: 64 columns per slab ----
            uint32_t v0[32], v1[32];
            if constexpr (kDual) {
              // both staging slabs of the group (pre-activation: grp, activation: grp + 2) are reused by every slab
              if (storer) ptx::tma_store_wait_read<0>();
              epi_bar_sync(grp);
            }
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(s * 64), v0);
            ptx::tmem_ld_32x32b_x32(taddr + uint32_t(s * 64 + 32), v1);
            // bias of the first 32 columns rides under the TMEM load latency (the wait below is a compiler barrier)
            float4 bb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) bb[i] = reinterpret_cast<const float4*>(bias_s + s * 64)[i];
            ptx::tmem_ld_wait();
            if (s + kGroups >= kSlabsPerTile) release_tmem();
            uint4 pk[8];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              float f[32];
              if (half == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bb[i] = reinterpret_cast<const float4*>(bias_s + s * 64 + 32)[i];
              }
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = bb[i >> 2];
                add2(f[i], f[i + 1], __uint_as_float(half == 0 ? v0[i] : v1[i]), __uint_as_float(half == 0 ? v0[i + 1] : v1[i + 1]), b4.x, b4.y);
                add2(f[i + 2], f[i + 3], __uint_as_float(half == 0 ? v0[i + 2] : v1[i + 2]), __uint_as_float(half == 0 ? v0[i + 3] : v1[i + 3]), b4.z,
                     b4.w);
              }
              if constexpr (kDual) {
                uint8_t* u_row = slab_gen + grp * kSlabBytes + r_in_tile * 128;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                  uint4 q;
                  q.x = Act<kF16>::pack(f[i], f[i + 1]);
                  q.y = Act<kF16>::pack(f[i + 2], f[i + 3]);
                  q.z = Act<kF16>::pack(f[i + 4], f[i + 5]);
                  q.w = Act<kF16>::pack(f[i + 6], f[i + 7]);
                  *reinterpret_cast<uint4*>(u_row + (((half * 4 + i / 8) ^ sw) << 4)) = q;
                }
                if (args.group == 2) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 32; i += 2) quick_gelu2(f[i], f[i + 1]);
                }
              } else if constexpr (kEpi == kEpiBiasGeluBf16) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) quick_gelu2(f[i], f[i + 1]);
              } else if constexpr (kEpi == kEpiGeluExactBf16) {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = gelu_erf(f[i]);
              }
              if constexpr (kAct) {
                if (args.down_a != nullptr) {
                  // warp-uniform shared-memory address: one broadcast wavefront per column; packed FFMA2 on (a.x,a.y), (a.z,a.w)
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(&down_s[s * 64 + half * 32 + i]);
                    uint64_t ff;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(ff) : "f"(f[i]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dacc01) : "l"(a.x), "l"(ff));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dacc23) : "l"(a.y), "l"(ff));
                  }
                }
              }
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 q;
                q.x = Act<kF16>::pack(f[i], f[i + 1]);
                q.y = Act<kF16>::pack(f[i + 2], f[i + 3]);
                q.z = Act<kF16>::pack(f[i + 4], f[i + 5]);
                q.w = Act<kF16>::pack(f[i + 6], f[i + 7]);
                pk[half * 4 + i / 8] = q;
              }
            }
            if constexpr (kDual) {
              uint8_t* h_row = slab_gen + (grp + 2) * kSlabBytes + r_in_tile * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(h_row + ((j ^ sw) << 4)) = pk[j];
              ptx::fence_proxy_async_smem();
              epi_bar_sync(grp);
              if (storer) {
                ptx::tma_store_2d(&tm_out2, slab_base + grp * kSlabBytes, col, row0);        // pre-activation
                ptx::tma_store_2d(&tm_out, slab_base + (grp + 2) * kSlabBytes, col, row0);   // activation
                ptx::tma_store_commit();
              }
            } else {
              // the slab this group wrote kBufPG steps ago used the same buffer: its TMA store must have read it by now
              if (storer) ptx::tma_store_wait_read<kBufPG - 1>();
              epi_bar_sync(grp);
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = pk[j];
              ptx::fence_proxy_async_smem();
              epi_bar_sync(grp);
              if (storer) {
                ptx::tma_store_2d(&tm_out, slab_base + b * kSlabBytes, col, row0);
                ptx::tma_store_commit();
              }
            }
          }
        }
        if constexpr (kAct) {
          // one partial per (column tile, epilogue group): part[kGroups * n_blk + grp][row][0..3]
          if (args.down_a != nullptr && row_ok)
            *reinterpret_cast<ulonglong2*>(args.down_part + (size_t(kGroups * n_blk + grp) * args.M + row) * 4) =
                make_ulonglong2(dacc01, dacc23);
        }
      }
    }
    // all global writes of this CTA's TMA stores must be complete before the kernel ends
    if constexpr (!kDirect) {
      if (storer && grp < kGroups) {
        ptx::tma_store_wait<0>();
      }
    }
  }

  // ======================= teardown =======================
  ptx::tcgen05_fence_before();
  if constexpr (kCtas > 1) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kCtas>(tmem_base, S::kTmemCols);
}

}  // namespace iic
