// Micro-benchmarks that size the attention kernel's softmax stage on sm_100a (not part of the product):
//   ldtm   tcgen05.ld 32x32b.x32 throughput per SM with 4 / 8 warps
//   mufu   ex2.approx.ftz.f32 throughput
//   fma2   fma.rn.f32x2 throughput
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_sm100 tools/ubench_sm100.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(384, 1) k_ldtm(long long* out, int iters, int nwarps, int mode) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t taddr = tb + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    uint32_t v[32], w[32];
    t0 = clock64();
    if (mode == 0) {        // one load in flight per warp
      for (int i = 0; i < iters; ++i) {
        ld32(taddr + (i & 7) * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31];
      }
    } else {                // two loads in flight per warp
      for (int i = 0; i < iters; i += 2) {
        ld32(taddr + (i & 7) * 32, v);
        ld32(taddr + ((i + 1) & 7) * 32, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31] ^ w[0] ^ w[31];
      }
    }
    t1 = clock64();
  }
  if (acc == 0x12345678u) out[1000] = 1;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && warp < nwarps) out[warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(256, 1) k_mufu(long long* out, float* sink, int iters, int nwarps) {
  const int warp = threadIdx.x >> 5;
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = -0.001f * (threadIdx.x + j);
  long long t0 = clock64();
  if (warp < nwarps)
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j];
  if (s == 1.2345f) sink[0] = s;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

template <int OP>
__global__ void __launch_bounds__(256, 1) k_mufu_op(long long* out, float* sink, int iters, int nwarps) {
  const int warp = threadIdx.x >> 5;
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = 0.37f + 0.001f * (threadIdx.x + j);
  long long t0 = clock64();
  if (warp < nwarps)
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
        if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
      }
    }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j];
  if (s == 1.2345f) sink[0] = s;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

// accuracy of two QuickGELU formulations against double precision: out[0] = max abs err (ex2+rcp), out[1] = max abs err (tanh),
// out[2], out[3] = the same relative to max(|h|, 2^-6)
__global__ void k_gelu_acc(float* out) {
  const float x = -12.f + 24.f * (blockIdx.x * blockDim.x + threadIdx.x) / float(gridDim.x * blockDim.x);
  const double ref = double(x) / (1.0 + exp(-1.702 * double(x)));
  float e, r, t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.702f * 1.4426950408889634f * x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  const float h0 = x * r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
  const float hx = 0.5f * x;
  const float h1 = fmaf(hx, t, hx);
  const float e0 = fabsf(float(double(h0) - ref)), e1 = fabsf(float(double(h1) - ref));
  const float den = fmaxf(fabsf(float(ref)), 0.015625f);
  atomicMax(reinterpret_cast<int*>(out + 0), __float_as_int(e0));
  atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(e1));
  atomicMax(reinterpret_cast<int*>(out + 2), __float_as_int(e0 / den));
  atomicMax(reinterpret_cast<int*>(out + 3), __float_as_int(e1 / den));
}

__global__ void __launch_bounds__(256, 1) k_fma2(long long* out, float* sink, int iters, int nwarps, int packed) {
  const int warp = threadIdx.x >> 5;
  float2 x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = make_float2(0.001f * (threadIdx.x + j), 0.5f);
  const float2 a = make_float2(0.999f, 1.0001f), b = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  if (warp < nwarps)
    for (int i = 0; i < iters; ++i) {
      if (packed) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint64_t& xx = reinterpret_cast<uint64_t&>(x[j]);
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(xx) : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j].x) : "f"(a.x), "f"(b.x));
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[j].y) : "f"(a.y), "f"(b.y));
        }
      }
    }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j].x + x[j].y;
  if (s == 1.2345f) sink[0] = s;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[warp] = t1 - t0;
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 2048 * sizeof(long long)); cudaMalloc(&sink, 16);
  long long h[16];
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int nw : {1, 4, 8}) {
      k_ldtm<<<148, 384>>>(d, iters, nw, mode);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("ldtm failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      // bytes per SM per clock: nw warps * iters loads * 32 lanes * 32 cols * 4 B
      printf("ldtm x32 mode %d warps %d: %.1f clk per load per warp, %.1f B/clk/SM\n", mode, nw, double(h[0]) / iters,
             double(nw) * iters * 4096.0 / double(h[0]));
    }
  for (int nw : {1, 4, 8}) {
    k_mufu<<<148, 256>>>(d, sink, iters, nw);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mufu.ex2 warps %d: %.2f clk per warp-inst, %.1f lanes/clk/SM\n", nw, double(h[0]) / (iters * 16.0), nw * iters * 16.0 * 32 / double(h[0]));
  }
  for (int nw : {4, 8}) {
    k_mufu_op<0><<<148, 256>>>(d, sink, iters, nw);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mufu.tanh warps %d: %.2f clk per warp-inst\n", nw, double(h[0]) / (iters * 16.0));
    k_mufu_op<1><<<148, 256>>>(d, sink, iters, nw);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mufu.rcp  warps %d: %.2f clk per warp-inst\n", nw, double(h[0]) / (iters * 16.0));
  }
  {
    float* acc; cudaMalloc(&acc, 16); cudaMemset(acc, 0, 16);
    k_gelu_acc<<<4096, 256>>>(acc);
    float ha[4]; cudaMemcpy(ha, acc, 16, cudaMemcpyDeviceToHost);
    printf("quickgelu max abs err: ex2+rcp %.3e  tanh %.3e | rel to max(|h|,2^-6): ex2+rcp %.3e  tanh %.3e\n", ha[0], ha[1], ha[2], ha[3]);
  }
  for (int packed = 0; packed < 2; ++packed)
    for (int nw : {4, 8}) {
      k_fma2<<<148, 256>>>(d, sink, iters, nw, packed);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      printf("fma %s warps %d: %.1f scalar FMA lanes/clk/SM\n", packed ? "f32x2" : "f32  ", nw, nw * iters * 32.0 * 32 / double(h[0]));
    }
  return 0;
}
