// Stand-alone sm_100a micro-benchmarks for the constants the attention kernels are designed around:
//   tcgen05.ld (TMEM -> registers) bandwidth per SM, MUFU.EX2 throughput, the cost of the P-pass inner loop.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qwen-image-finetune_b200/csrc -I include tools/microbench.cu -o gpurun_out/microbench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "sm100.cuh"

using namespace qfx;

// mode 0: x32 loads, wait after each; mode 1: two x32 loads in flight per wait; mode 2: x32 loads + 32 ex2 + pack + 4 STS.128
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(long long* out, int iters, int warps_active, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < warps_active) {
    // zero the columns first so the values are benign
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0x3f000000u;
    for (int c = 0; c < 512; c += 32) tmem_st32(tb + lane_off + c, z);
    tmem_st_wait();
  }
  __syncthreads();
  if (warp < warps_active) {
    t0 = clock64();
    const int col_base = ((warp >> 2) * 64) & 511;
    for (int it = 0; it < iters; ++it) {
      uint32_t r[64];
      const uint32_t col = (col_base + (it & 1) * 256) & 511;
      if (MODE == 0) {
        tmem_ld32(tb + lane_off + col, r);
        tmem_ld_wait();
        acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
      } else if (MODE == 1) {
        tmem_ld32(tb + lane_off + col, r);
        tmem_ld32(tb + lane_off + col + 32, r + 32);
        tmem_ld_wait();
        acc += __uint_as_float(r[0]) + __uint_as_float(r[63]);
      } else {
        tmem_ld32(tb + lane_off + col, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float x0 = __uint_as_float(r[2 * e]) * 1.3f - acc, x1 = __uint_as_float(r[2 * e + 1]) * 1.3f - acc;
          if (MODE == 2) pk[e] = pack_bf16(exp2f(x0), exp2f(x1));
          else if (MODE == 3) pk[e] = pack_bf16(x0 * (x1 - 0.3f), x1 * (x0 - 0.3f));              // dS-pass-like: FMA only
          else if (MODE == 4) pk[e] = pack_bf16(exp2f(x0), exp2_fma(x1));                             // half the exponentials on the FMA pipe
          else pk[e] = pack_bf16(exp2f(x0), (e & 1) ? exp2_fma(x1) : exp2f(x1));                      // a quarter
        }
        const uint32_t row = smem_u32(smem) + threadIdx.x * 128;  // the kernels' 128B-swizzled row layout (bank-conflict free)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((v ^ (threadIdx.x & 7)) * 16)), "r"(pk[4 * v]), "r"(pk[4 * v + 1]),
                       "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                       : "memory");
      }
    }
    t1 = clock64();
  }
  __syncthreads();
  if (lane == 0 && warp < warps_active) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

__global__ void __launch_bounds__(512, 1) ex2_kernel(long long* out, int iters, float* sink, float seed) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (i + 1) * 1e-3f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = exp2f(a[i] * 0.25f - 1.f);  // FFMA + MUFU.EX2
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) sink[0] = s;
  if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
}

template <typename F>
static double run(F launch, long long* d_out, int warps) {
  long long h[16];
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx;
}

int main() {
  long long* d_out;
  float* sink;
  cudaMalloc(&d_out, 16 * sizeof(long long) * 4);
  cudaMalloc(&sink, 16);
  const int iters = 4096;
  cudaFuncSetAttribute(tmem_ld_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(tmem_ld_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(tmem_ld_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(tmem_ld_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int warps : {1, 4, 8, 16}) {
    double c0 = run([&] { tmem_ld_kernel<0><<<1, 512, 1024>>>(d_out, iters, warps, sink); }, d_out, warps);
    double c1 = run([&] { tmem_ld_kernel<1><<<1, 512, 1024>>>(d_out, iters, warps, sink); }, d_out, warps);
    double c2 = run([&] { tmem_ld_kernel<2><<<1, 512, 65536>>>(d_out, iters, warps, sink); }, d_out, warps);
    double c3 = run([&] { tmem_ld_kernel<3><<<1, 512, 65536>>>(d_out, iters, warps, sink); }, d_out, warps);
    double c4 = run([&] { tmem_ld_kernel<4><<<1, 512, 65536>>>(d_out, iters, warps, sink); }, d_out, warps);
    double c5 = run([&] { tmem_ld_kernel<5><<<1, 512, 65536>>>(d_out, iters, warps, sink); }, d_out, warps);
    printf("   warps=%2d  P-pass clk/iter: all-MUFU %.1f | FMA-only %.1f | 1/2 poly %.1f | 1/4 poly %.1f   (elem/clk/SM %.2f %.2f %.2f %.2f)\n", warps,
           c2 / iters, c3 / iters, c4 / iters, c5 / iters, 1024.0 * warps * iters / c2, 1024.0 * warps * iters / c3,
           1024.0 * warps * iters / c4, 1024.0 * warps * iters / c5);
    printf("tmem_ld x32  warps=%2d : 1-in-flight %.1f clk/ld/warp -> %.1f B/clk/SM ; 2-in-flight %.1f clk/pair -> %.1f B/clk/SM ; "
           "ld+32 ex2+pack+STS %.1f clk/iter -> %.2f elem/clk/SM\n",
           warps, c0 / iters, 4096.0 * warps * iters / c0, c1 / iters, 8192.0 * warps * iters / c1, c2 / iters,
           1024.0 * warps * iters / c2);
  }
  for (int warps : {4, 8, 16}) {
    double c = run([&] { ex2_kernel<<<1, warps * 32>>>(d_out, iters, sink, 1.f); }, d_out, warps);
    printf("ex2 (ffma+mufu) warps=%2d : %.2f ex2/clk/SM\n", warps, 16.0 * 32 * warps * iters / c);
  }
  return 0;
}
