// uint8 HWC -> normalised bf16 patch matrix (or CHW) preprocessing, bit-compatible with the reference's
// `clip._transform(R)` as called at /root/reference/main.py:201, 438, 489 and train_lora.py:149:
//     Resize(R, BICUBIC) on the shorter side -> CenterCrop(R) -> ToTensor (/255) -> Normalize(mean, std)
//
// Pillow semantics restated here (Pillow 12 `ImagingResample`, 8-bit path):
//   * per axis, scale = in/out; when down-scaling the bicubic (a = -0.5) support 2.0 is stretched by the scale
//     (antialias); taps = [int(center - support + .5), int(center + support + .5)) clipped to the image;
//     coefficients normalised to sum 1 in double, then quantised to int(round(c * 2^22));
//   * horizontal pass first, over only the source rows the vertical pass will touch; each pass accumulates in
//     int32 starting from 2^21, shifts right by 22 and clamps to uint8 (the intermediate image is uint8);
//   * an axis whose size does not change is skipped entirely (expressed below as a 1-tap identity kernel, which
//     reproduces the input byte exactly);
//   * torchvision CenterCrop offsets use Python round() = round-half-to-even.
// Only the RxR crop window is ever computed.  Coefficient tables are built on the host in double precision
// (compiled without FMA contraction so they match Pillow's x86-64 build bit for bit) and cached per
// (in_size, out_size); they are uploaded with one async copy per call.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "act_types.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace iic {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // 22

// CLIP normalisation constants (clip._transform)
const float kMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
const float kStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};

struct AxisTable {
  int ksize = 0;
  std::vector<int32_t> bounds;  // [out][2] = (xmin, count)
  std::vector<int32_t> coef;    // [out][ksize]
};

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size)
AxisTable build_axis(int in_size, int out_size) {
  AxisTable t;
  if (in_size == out_size) {  // Pillow skips the pass: identity
    t.ksize = 1;
    t.bounds.resize(size_t(out_size) * 2);
    t.coef.resize(size_t(out_size));
    for (int i = 0; i < out_size; ++i) {
      t.bounds[2 * i] = i;
      t.bounds[2 * i + 1] = 1;
      t.coef[i] = 1 << kPrecisionBits;
    }
    return t;
  }
  const float in0 = 0.0f, in1 = float(in_size);
  double scale = double(in1 - in0) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const int ksize = int(std::ceil(support)) * 2 + 1;
  t.ksize = ksize;
  t.bounds.assign(size_t(out_size) * 2, 0);
  t.coef.assign(size_t(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = int(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = int(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = k[x];
      t.coef[size_t(xx) * ksize + x] =
          v < 0 ? int(-0.5 + v * (1 << kPrecisionBits)) : int(0.5 + v * (1 << kPrecisionBits));
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

// Python 3 round(): ties to even
int py_round(double v) { return int(std::nearbyint(v)); }

struct ImgDesc {
  const uint8_t* src;
  int src_w, src_h;
  int row_first;      // first source row the vertical pass touches
  int rows;           // number of intermediate rows
  long long tmp_off;  // byte offset of this image's intermediate [rows][R][3] in the scratch
  int hk, vk;         // ksize of the horizontal / vertical tables
  int h_bounds, h_coef, v_bounds, v_coef;  // int32 offsets into the uploaded table blob
};

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return uint8_t(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: thread per (intermediate row, output column) of one image
__global__ void __launch_bounds__(256)
resize_h_kernel(const ImgDesc* __restrict__ descs, const int32_t* __restrict__ tables, uint8_t* __restrict__ scratch,
                int R) {
  const ImgDesc d = descs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.rows * R) return;
  const int r = idx / R, xx = idx - r * R;
  const int xmin = tables[d.h_bounds + 2 * xx], cnt = tables[d.h_bounds + 2 * xx + 1];
  const int32_t* k = tables + d.h_coef + xx * d.hk;
  const uint8_t* s = d.src + (size_t(d.row_first + r) * d.src_w + xmin) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int x = 0; x < cnt; ++x) {
    const int c = k[x];
    a0 += int(s[3 * x]) * c;
    a1 += int(s[3 * x + 1]) * c;
    a2 += int(s[3 * x + 2]) * c;
  }
  uint8_t* o = scratch + d.tmp_off + (size_t(r) * R + xx) * 3;
  o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
}

__device__ __forceinline__ uint16_t to16(float v, int f16) {
  if (f16) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

__device__ __forceinline__ void store_px(void* out, int out_mode, int f16, int b, int c, int yy, int xx, int R, int P,
                                         int k_pad, float v) {
  if (out_mode == 0) {
    const int g = R / P;
    const size_t row = size_t(b) * g * g + (yy / P) * g + xx / P;
    reinterpret_cast<uint16_t*>(out)[row * k_pad + c * P * P + (yy % P) * P + xx % P] = to16(v, f16);
  } else if (out_mode == 1) {
    reinterpret_cast<float*>(out)[((size_t(b) * 3 + c) * R + yy) * R + xx] = v;
  } else {
    reinterpret_cast<uint16_t*>(out)[((size_t(b) * 3 + c) * R + yy) * R + xx] = to16(v, f16);
  }
}

// vertical pass + ToTensor + Normalize + layout: thread per output pixel of one image
__global__ void __launch_bounds__(256)
resize_v_kernel(const ImgDesc* __restrict__ descs, const int32_t* __restrict__ tables,
                const uint8_t* __restrict__ scratch, const float* __restrict__ lut, int R, int P, int k_pad, void* out,
                int out_mode, int f16) {
  const ImgDesc d = descs[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * R) return;
  const int yy = idx / R, xx = idx - yy * R;
  const int ymin = tables[d.v_bounds + 2 * yy] - d.row_first, cnt = tables[d.v_bounds + 2 * yy + 1];
  const int32_t* k = tables + d.v_coef + yy * d.vk;
  const uint8_t* s = scratch + d.tmp_off + (size_t(ymin) * R + xx) * 3;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int y = 0; y < cnt; ++y) {
    const int c = k[y];
    const uint8_t* p = s + size_t(y) * R * 3;
    a0 += int(p[0]) * c;
    a1 += int(p[1]) * c;
    a2 += int(p[2]) * c;
  }
  const int b = blockIdx.y;
  store_px(out, out_mode, f16, b, 0, yy, xx, R, P, k_pad, lut[clip8(a0)]);
  store_px(out, out_mode, f16, b, 1, yy, xx, R, P, k_pad, lut[256 + clip8(a1)]);
  store_px(out, out_mode, f16, b, 2, yy, xx, R, P, k_pad, lut[512 + clip8(a2)]);
}

// Same-size fast path, patch-matrix output, P == 16: one thread per (image, row, patch-x) moves 16 pixels:
// 48 contiguous input bytes -> three 32-byte bf16 runs.
template <bool kF16>
__global__ void __launch_bounds__(256)
preprocess_fast_p16_kernel(const uint8_t* __restrict__ img, const float* __restrict__ lut,
                           uint16_t* __restrict__ out, int B, int R, int k_pad) {
  ptx::pdl_launch_dependents();
  __shared__ float s_lut[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];   // constant table: may precede the dependency wait
  __syncthreads();
  ptx::pdl_wait();
  const int g = R / 16;
  const long long total = (long long)B * R * g;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int px = int(i % g);
  const long long t = i / g;
  const int y = int(t % R);
  const int b = int(t / R);
  const uint4* src = reinterpret_cast<const uint4*>(img + ((size_t(b) * R + y) * R + px * 16) * 3);
  uint32_t w[12];
  {
    const uint4 v0 = src[0], v1 = src[1], v2 = src[2];
    w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w;
    w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
  }
  const uint8_t* bytes = reinterpret_cast<const uint8_t*>(w);
  const int py = y >> 4, ky = y & 15;
  uint16_t* dst = out + (size_t(b) * g * g + py * g + px) * k_pad + ky * 16;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float lo = s_lut[c * 256 + bytes[(2 * j) * 3 + c]];
      const float hi = s_lut[c * 256 + bytes[(2 * j + 1) * 3 + c]];
      pk[j] = Act<kF16>::pack(lo, hi);
    }
    uint4* o = reinterpret_cast<uint4*>(dst + c * 256);
    o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}

// generic same-size path (any P / output mode): thread per pixel
__global__ void __launch_bounds__(256)
preprocess_same_kernel(const uint8_t* __restrict__ img, const float* __restrict__ lut, void* out, int out_mode, int f16,
                       int B, int R, int P, int k_pad) {
  ptx::pdl_launch_dependents();
  ptx::pdl_wait();
  const long long total = (long long)B * R * R;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xx = int(i % R);
  const long long t = i / R;
  const int yy = int(t % R);
  const int b = int(t / R);
  const uint8_t* p = img + size_t(i) * 3;
  store_px(out, out_mode, f16, b, 0, yy, xx, R, P, k_pad, lut[p[0]]);
  store_px(out, out_mode, f16, b, 1, yy, xx, R, P, k_pad, lut[256 + p[1]]);
  store_px(out, out_mode, f16, b, 2, yy, xx, R, P, k_pad, lut[512 + p[2]]);
}

}  // namespace

struct PreprocessPlan {
  std::map<std::pair<int, int>, AxisTable> cache;  // (in, out) -> table
  float* d_lut = nullptr;                          // [3][256] fp32: (v/255 - mean)/std, IEEE fp32 like torch
  // grow-only device scratch + pinned staging
  uint8_t* d_scratch = nullptr; size_t scratch_cap = 0;
  int32_t* d_tables = nullptr;  size_t tables_cap = 0;   // bytes
  ImgDesc* d_descs = nullptr;   size_t descs_cap = 0;    // bytes
  uint8_t* h_stage = nullptr;   size_t stage_cap = 0;
  cudaEvent_t stage_free = nullptr;  // previous call's upload finished -> staging buffer reusable
  bool stage_pending = false;
};

PreprocessPlan* preprocess_plan_create() {
  PreprocessPlan* p = new PreprocessPlan();
  float lut[768];
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      volatile float f = float(v) / 255.0f;  // ToTensor: uint8 -> float32, .div(255)
      volatile float g = f - kMean[c];       // Normalize: sub(mean)
      lut[c * 256 + v] = g / kStd[c];        //            .div(std)
    }
  if (cudaMalloc(&p->d_lut, sizeof(lut)) != cudaSuccess ||
      cudaMemcpy(p->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaEventCreateWithFlags(&p->stage_free, cudaEventDisableTiming) != cudaSuccess) {
    preprocess_plan_destroy(p);
    return nullptr;
  }
  return p;
}

void preprocess_plan_destroy(PreprocessPlan* p) {
  if (p == nullptr) return;
  if (p->d_lut) cudaFree(p->d_lut);
  if (p->d_scratch) cudaFree(p->d_scratch);
  if (p->d_tables) cudaFree(p->d_tables);
  if (p->d_descs) cudaFree(p->d_descs);
  if (p->h_stage) cudaFreeHost(p->h_stage);
  if (p->stage_free) cudaEventDestroy(p->stage_free);
  delete p;
}

namespace {
template <typename T>
bool grow(T*& ptr, size_t& cap, size_t need) {
  if (need <= cap) return true;
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  const size_t want = need + need / 2 + 4096;
  if (cudaMalloc(reinterpret_cast<void**>(&ptr), want) != cudaSuccess) return false;
  cap = want;
  return true;
}
}  // namespace

int launch_preprocess_fast(PreprocessPlan* plan, const uint8_t* imgs, int B, int R, int P, int k_pad, void* out,
                           int out_mode, int f16, cudaStream_t stream) {
  if (B <= 0) return 0;
  if (plan == nullptr || R % P != 0) return -1;
  if (out_mode == 0 && P == 16 && (reinterpret_cast<uintptr_t>(imgs) & 15) == 0 && k_pad % 8 == 0) {
    const long long total = (long long)B * R * (R / 16);
    if (f16)
      launch_k(preprocess_fast_p16_kernel<true>, dim3(unsigned((total + 255) / 256)), dim3(256), 0, stream, imgs, plan->d_lut,
               reinterpret_cast<uint16_t*>(out), B, R, k_pad);
    else
      launch_k(preprocess_fast_p16_kernel<false>, dim3(unsigned((total + 255) / 256)), dim3(256), 0, stream, imgs, plan->d_lut,
               reinterpret_cast<uint16_t*>(out), B, R, k_pad);
  } else {
    const long long total = (long long)B * R * R;
    launch_k(preprocess_same_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, stream, imgs, plan->d_lut, out, out_mode, f16,
             B, R, P, k_pad);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int launch_preprocess(PreprocessPlan* plan, const uint8_t* const* imgs, const int* hw, int B, int R, int P, int k_pad,
                      void* out, int out_mode, int f16, cudaStream_t stream, const char** err) {
  static const char* e_arg = "preprocess: bad argument (null image, non-positive size, or R % patch != 0)";
  static const char* e_mem = "preprocess: device/pinned allocation failed";
  static const char* e_launch = "preprocess: CUDA launch/copy failed";
  if (B <= 0) return 0;
  if (plan == nullptr || R % P != 0) { if (err) *err = e_arg; return -1; }

  std::vector<ImgDesc> descs(B);
  std::vector<int32_t> blob;
  std::map<std::pair<int, int>, std::pair<int, int>> placed;  // (in,out) -> (bounds_off, coef_off) in blob
  auto place = [&](int in_size, int out_size, int& ksize, int& b_off, int& c_off) {
    const auto key = std::make_pair(in_size, out_size);
    auto it = plan->cache.find(key);
    if (it == plan->cache.end()) it = plan->cache.emplace(key, build_axis(in_size, out_size)).first;
    const AxisTable& t = it->second;
    ksize = t.ksize;
    auto pit = placed.find(key);
    if (pit == placed.end()) {
      const int bo = int(blob.size());
      blob.insert(blob.end(), t.bounds.begin(), t.bounds.end());
      const int co = int(blob.size());
      blob.insert(blob.end(), t.coef.begin(), t.coef.end());
      pit = placed.emplace(key, std::make_pair(bo, co)).first;
    }
    b_off = pit->second.first;
    c_off = pit->second.second;
  };

  size_t scratch_need = 0;
  int max_rows = 0;
  for (int i = 0; i < B; ++i) {
    const int h = hw[2 * i], w = hw[2 * i + 1];
    if (imgs[i] == nullptr || h <= 0 || w <= 0) { if (err) *err = e_arg; return -1; }
    // torchvision Resize(R): shorter side -> R, longer side -> int(R * long / short)
    int nw, nh;
    if (w <= h) { nw = R; nh = int((long long)R * h / w); }
    else        { nh = R; nw = int((long long)R * w / h); }
    // torchvision CenterCrop(R): round-half-even offsets (resized sides are >= R so no padding case)
    const int top = py_round((nh - R) / 2.0), left = py_round((nw - R) / 2.0);
    ImgDesc& d = descs[i];
    d.src = imgs[i];
    d.src_w = w;
    d.src_h = h;
    int hb, hc, vb, vc;
    place(w, nw, d.hk, hb, hc);
    place(h, nh, d.vk, vb, vc);
    const AxisTable& tv = plan->cache[std::make_pair(h, nh)];
    // tables are laid out per resized coordinate; the crop just offsets into them
    d.h_bounds = hb + 2 * left;
    d.h_coef = hc + left * d.hk;
    d.v_bounds = vb + 2 * top;
    d.v_coef = vc + top * d.vk;
    const int first = tv.bounds[2 * top];
    const int last = tv.bounds[2 * (top + R - 1)] + tv.bounds[2 * (top + R - 1) + 1];
    d.row_first = first;
    d.rows = last - first;
    d.tmp_off = (long long)scratch_need;
    scratch_need += (size_t(d.rows) * R * 3 + 255) & ~size_t(255);
    if (d.rows > max_rows) max_rows = d.rows;
  }

  const size_t tables_bytes = blob.size() * sizeof(int32_t);
  const size_t descs_bytes = descs.size() * sizeof(ImgDesc);
  const size_t stage_need = tables_bytes + descs_bytes;
  if (!grow(plan->d_scratch, plan->scratch_cap, scratch_need) || !grow(plan->d_tables, plan->tables_cap, tables_bytes) ||
      !grow(plan->d_descs, plan->descs_cap, descs_bytes)) {
    if (err) *err = e_mem;
    return -2;
  }
  if (plan->stage_pending) { cudaEventSynchronize(plan->stage_free); plan->stage_pending = false; }
  if (stage_need > plan->stage_cap) {
    if (plan->h_stage) cudaFreeHost(plan->h_stage);
    plan->h_stage = nullptr;
    plan->stage_cap = 0;
    if (cudaMallocHost(reinterpret_cast<void**>(&plan->h_stage), stage_need * 2) != cudaSuccess) {
      if (err) *err = e_mem;
      return -2;
    }
    plan->stage_cap = stage_need * 2;
  }
  std::memcpy(plan->h_stage, blob.data(), tables_bytes);
  std::memcpy(plan->h_stage + tables_bytes, descs.data(), descs_bytes);
  bool ok = cudaMemcpyAsync(plan->d_tables, plan->h_stage, tables_bytes, cudaMemcpyHostToDevice, stream) == cudaSuccess &&
            cudaMemcpyAsync(plan->d_descs, plan->h_stage + tables_bytes, descs_bytes, cudaMemcpyHostToDevice, stream) ==
                cudaSuccess;
  if (ok) { cudaEventRecord(plan->stage_free, stream); plan->stage_pending = true; }

  if (ok) {
    dim3 gh(unsigned((size_t(max_rows) * R + 255) / 256), unsigned(B));
    resize_h_kernel<<<gh, 256, 0, stream>>>(plan->d_descs, plan->d_tables, plan->d_scratch, R);
    dim3 gv(unsigned((size_t(R) * R + 255) / 256), unsigned(B));
    resize_v_kernel<<<gv, 256, 0, stream>>>(plan->d_descs, plan->d_tables, plan->d_scratch, plan->d_lut, R, P, k_pad, out,
                                            out_mode, f16);
    ok = cudaGetLastError() == cudaSuccess;
  }
  if (!ok) { if (err) *err = e_launch; return -2; }
  return 0;
}

}  // namespace iic
