// 16-bit operand formats of the encoder: bf16 (BASELINE's named dtype) or fp16 (OpenAI CLIP's own GPU dtype; 3 more
// mantissa bits, same tensor-core rate).  One template switch, identical kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace iic {

template <bool kF16>
struct Act;

template <>
struct Act<false> {
  using T = __nv_bfloat16;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  }
  static __device__ __forceinline__ T from_float(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
};

template <>
struct Act<true> {
  using T = __half;
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
  static __device__ __forceinline__ T from_float(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b) {
    __half2 r = __hmul2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
};

}  // namespace iic
