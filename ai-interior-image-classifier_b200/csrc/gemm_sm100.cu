// Host side of the tcgen05 GEMM: TMA tensor-map encoding and launch dispatch.
#include "gemm_sm100.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>

#include "kernels.h"

namespace iic {

namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved at run time through the runtime so the library has no link-time dependency on libcuda.so
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// row-major [rows, cols] with row pitch `ld` elements; box = [box_rows, 128 bytes of columns], 128-byte swizzle.
// kind: 0 = bf16, 1 = fp16 (64-column boxes), 2 = fp32 (32-column boxes).
bool make_tile_map_kind(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                        int kind) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return false;
  const uint64_t esz = kind == 2 ? 4 : 2;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * esz};
  cuuint32_t box[2] = {uint32_t(128 / esz), box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
bool make_tile_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                   bool f16) {
  return make_tile_map_kind(map, ptr, rows, cols, ld, box_rows, f16 ? 1 : 0);
}

template <int kCtas, int kBlockN, int kEpi, bool kF16>
int launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal, const CUtensorMap& tbl,
               const CUtensorMap& tout, const CUtensorMap& tres, const CUtensorMap& tln, const GemmArgs& args, int num_sms,
               cudaStream_t stream) {
  using S = GemmSmem<kCtas, kBlockN, kEpi>;
  auto kern = gemm_bf16_tn_kernel<kCtas, kBlockN, kEpi, kF16>;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal) != cudaSuccess) return -2;
    attr_done.mark();
  }
  const int tile_m = kBlockM * kCtas;
  const int m_tiles = (args.M + tile_m - 1) / tile_m;
  const int n_tiles = (args.N + kBlockN - 1) / kBlockN;
  const int total = m_tiles * n_tiles;
  int clusters = num_sms / kCtas;
  if (clusters > total) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(unsigned(clusters * kCtas));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // small-batch latency path (kernels.h: g_pdl)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tal, tbl, tout, tres, tln, args);
  return e == cudaSuccess ? 0 : -2;
}

template <int kCtas, int kBlockN, bool kF16>
int dispatch_narrow(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal, const CUtensorMap& tbl,
                    const CUtensorMap& tout, const CUtensorMap& tres, const CUtensorMap& tln, const GemmArgs& args, int num_sms,
                    cudaStream_t stream) {
  switch (epi) {
    case kEpiBiasBf16: return launch_one<kCtas, kBlockN, kEpiBiasBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasGeluBf16: return launch_one<kCtas, kBlockN, kEpiBiasGeluBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiGeluExactBf16: return launch_one<kCtas, kBlockN, kEpiGeluExactBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasResF32: return launch_one<kCtas, kBlockN, kEpiBiasResF32, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasResF32DeepK: return launch_one<kCtas, kBlockN, kEpiBiasResF32DeepK, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiPosF32: return launch_one<kCtas, kBlockN, kEpiPosF32, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    default: return -1;
  }
}

template <int kCtas, bool kF16>
int dispatch_epi(int epi, int tile_n, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal, const CUtensorMap& tbl,
                 const CUtensorMap& tout, const CUtensorMap& tres, const CUtensorMap& tln, const GemmArgs& args, int num_sms,
                 cudaStream_t stream) {
  // skinny outputs (the LoRA down-projection GEMMs, N = 16): a 64-column tile - the 256-wide MMAs on a zero-filled weight tile
  // were what bounded them (K = 3072: 13 us of tensor time per 256-row tile for 16 useful columns)
  if (epi == kEpiBiasBf16 && tile_n == 64)
    return launch_one<kCtas, 64, kEpiBiasBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
  if constexpr (kCtas == 2) {   // small-batch path (gemm_tile_n)
    if (tile_n == 64) return dispatch_narrow<2, 64, kF16>(epi, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    if (tile_n == 128) return dispatch_narrow<2, 128, kF16>(epi, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
  }
  switch (epi) {
    case kEpiActGradBf16: return launch_one<kCtas, 256, kEpiActGradBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasActDualBf16: return launch_one<kCtas, 256, kEpiBiasActDualBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasBf16: return launch_one<kCtas, 256, kEpiBiasBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasGeluBf16: return launch_one<kCtas, 256, kEpiBiasGeluBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasResF32: return launch_one<kCtas, 256, kEpiBiasResF32, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiBiasResF32DeepK: return launch_one<kCtas, 256, kEpiBiasResF32DeepK, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiPosF32: return launch_one<kCtas, 256, kEpiPosF32, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    case kEpiGeluExactBf16: return launch_one<kCtas, 256, kEpiGeluExactBf16, kF16>(ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
    default: return -1;
  }
}

}  // namespace

int gemm_tile_n(int M, int N, int epilogue, int ctas, int num_sms) {
  if (epilogue == kEpiBiasBf16 && N <= 64) return 64;
  static const int narrow_max = [] { const char* e = getenv("IIC_GEMM_NARROW"); return e ? atoi(e) : 1; }();   // 0: always 256
  const bool inference_epi = epilogue == kEpiBiasBf16 || epilogue == kEpiBiasGeluBf16 || epilogue == kEpiGeluExactBf16 ||
                             epilogue == kEpiBiasResF32 || epilogue == kEpiPosF32;
  if (ctas != 2 || !inference_epi || narrow_max == 0 || M <= 0 || N <= 0) return 256;
  const int clusters = num_sms / 2;
  const int tiles256 = ((M + 255) / 256) * ((N + 255) / 256);
  if (tiles256 * 4 <= clusters) return 64;
  if (tiles256 * 2 <= clusters) return 128;
  return 256;
}
int gemm_down_parts(int M, int N, int epilogue, int ctas, int num_sms) {
  const int tn = gemm_tile_n(M, N, epilogue, ctas, num_sms);
  return (tn >= 128 ? 2 : 1) * ((N + tn - 1) / tn);
}

size_t gemm_smem_bytes(int ctas) {
  return ctas == 2 ? GemmSmem<2, 256, kEpiBiasBf16>::kTotal : GemmSmem<1, 256, kEpiBiasBf16>::kTotal;
}

int launch_gemm(const GemmProblem& p, int ctas, int num_sms, cudaStream_t stream, const char** err) {
  static const char* e_shape = "gemm: unsupported shape (need M,N,K > 0, N % 32 == 0 (N % 8 without bias), pitches % 8 == 0)";
  static const char* e_map = "gemm: cuTensorMapEncodeTiled failed (driver entry point missing or bad pointer/pitch)";
  static const char* e_launch = "gemm: kernel launch failed";
  static const char* e_lora = "gemm: LoRA rank pad must be a multiple of 16 and <= 64";
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.N % (p.bias != nullptr ? 32 : 8) != 0 || p.lda % 8 != 0 || p.ldw % 8 != 0 || p.ldc % 8 != 0 ||
      (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) {
    if (err) *err = e_shape;
    return -1;
  }
  // 5 is a variant the launcher selects itself (deep-K ring); 6 and 7 are retired: not valid as a request
  if (p.epilogue < 0 || p.epilogue > kEpiBiasActDualBf16 || (p.epilogue >= kEpiBiasResF32DeepK && p.epilogue < kEpiActGradBf16)) {
    if (err) *err = "gemm: unknown epilogue";
    return -1;
  }
  const bool lora = p.lora_p != nullptr && p.lora_bt != nullptr && p.r_pad > 0;
  if (lora && (p.r_pad % 16 != 0 || p.r_pad > 64)) {
    if (err) *err = e_lora;
    return -1;
  }
  const int tile_n = p.tile_n == 256 ? ((p.epilogue == kEpiBiasBf16 && p.N <= 64) ? 64 : 256) : gemm_tile_n(p.M, p.N, p.epilogue, ctas, num_sms);
  const uint32_t box_b = uint32_t(tile_n / ctas);
  CUtensorMap ta, tb, tal, tbl;
  const bool f16 = p.f16 != 0;
  bool ok = make_tile_map(&ta, p.a, uint64_t(p.M), uint64_t(p.K), uint64_t(p.lda), kBlockM, f16) &&
            make_tile_map(&tb, p.w, uint64_t(p.N), uint64_t(p.K), uint64_t(p.ldw), box_b, f16);
  if (ok && lora) {
    // [rows, r_pad] operands: columns beyond r_pad are out of bounds for TMA and arrive as zeros
    const uint64_t ld = p.lora_ld > 0 ? uint64_t(p.lora_ld) : uint64_t(p.r_pad);
    const uint64_t cols = ld < uint64_t(kBlockK) ? ld : uint64_t(kBlockK);
    ok = make_tile_map(&tal, p.lora_p, uint64_t(p.M), cols, ld, kBlockM, f16) &&
         make_tile_map(&tbl, p.lora_bt, uint64_t(p.N), cols, ld, box_b, f16);
  } else {
    tal = ta;
    tbl = tb;
  }
  // output (and residual) maps for the TMA-store epilogue; the patch-embedding epilogue stores directly
  CUtensorMap tout = ta, tres = ta;
  if (ok && p.epilogue != kEpiPosF32) {
    const bool f32_out = p.epilogue == kEpiBiasResF32;
    ok = make_tile_map_kind(&tout, p.out, uint64_t(p.M), uint64_t(p.N), uint64_t(p.ldc), kBlockM, f32_out ? 2 : (f16 ? 1 : 0));
    if (ok && f32_out) {
      if (p.residual == nullptr) ok = false;
      else ok = make_tile_map_kind(&tres, p.residual, uint64_t(p.M), uint64_t(p.N), uint64_t(p.ldc), kBlockM, 2);
    }
  }
  if (ok && p.epilogue == kEpiActGradBf16) {   // 16-bit pre-activation tile, same geometry as the output
    if (p.residual == nullptr) ok = false;
    else ok = make_tile_map_kind(&tres, p.residual, uint64_t(p.M), uint64_t(p.N), uint64_t(p.ldc), kBlockM, f16 ? 1 : 0);
  }
  CUtensorMap tln = ta;   // second output (kEpiBiasActDualBf16: the kept pre-activation)
  if (ok && p.epilogue == kEpiBiasActDualBf16) {
    if (p.out2 == nullptr || (reinterpret_cast<uintptr_t>(p.out2) & 15) != 0) ok = false;
    else ok = make_tile_map_kind(&tln, p.out2, uint64_t(p.M), uint64_t(p.N), uint64_t(p.ldc), kBlockM, f16 ? 1 : 0);
  }
  if (!ok) {
    if (err) *err = e_map;
    return -1;
  }
  GemmArgs args;
  args.M = p.M;
  args.N = p.N;
  args.num_k_blocks = (p.K + kBlockK - 1) / kBlockK;
  args.lora_ksteps = lora ? p.r_pad / 16 : 0;
  args.bias = p.bias;
  args.residual = p.residual;
  args.out = p.out;
  args.ldc = p.ldc;
  args.group = p.group > 0 ? p.group : 1;
  args.down_a = p.down_a;
  args.down_part = p.down_part;
  // long reductions get the deep-ring variant of the residual epilogue (measured: c_proj 1268 -> 1322 TFLOP/s; the short-K
  // out_proj is bound by its fp32 residual traffic and prefers the 4-slab residual ring)
  const bool deep = p.epilogue == kEpiBiasResF32 && p.K >= 2048;
  const int epi = deep ? int(kEpiBiasResF32DeepK) : p.epilogue;
  int rc;
  if (f16)
    rc = ctas == 2 ? dispatch_epi<2, true>(epi, tile_n, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream)
                   : dispatch_epi<1, true>(epi, tile_n, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
  else
    rc = ctas == 2 ? dispatch_epi<2, false>(epi, tile_n, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream)
                   : dispatch_epi<1, false>(epi, tile_n, ta, tb, tal, tbl, tout, tres, tln, args, num_sms, stream);
  if (rc != 0 && err) *err = rc == -1 ? e_shape : e_launch;
  return rc;
}

}  // namespace iic
