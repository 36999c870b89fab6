// Micro-benchmark: how fast can the chip ADD fp32 data into L2-resident / HBM-resident memory?  (What bounds the dQ accumulation of the
// attention backward: 19 key-tile CTAs each add a [2400 x 128] fp32 partial per head = 2.27 GB of atomic traffic per call.)
//   mode 0: red.global.add.f32, warp = 32 consecutive floats (one 128-byte line per instruction)
//   mode 1: red.global.add.v4.f32, warp = 512 consecutive bytes per instruction
//   mode 2: cp.reduce.async.bulk (TMA 1-D bulk reduce-add) of 16 KB shared-memory blocks
//   mode 3: plain st.global.v4 of the same bytes (reference: store bandwidth)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/red_bw tools/experiments/red_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void k_red(float* dst, size_t n_per_pass, int passes, int mode) {
  extern __shared__ __align__(128) float sm[];
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = gridDim.x * (size_t)blockDim.x;
  if (mode == 2) {
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const size_t nblk = n_per_pass / 4096;
      for (int p = 0; p < passes; ++p)
        for (size_t b = blockIdx.x; b < nblk; b += gridDim.x) {
          unsigned s = (unsigned)__cvta_generic_to_shared(sm);
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst + b * 4096), "r"(s), "r"(16384) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    return;
  }
  for (int p = 0; p < passes; ++p) {
    if (mode == 0) {
      for (size_t i = tid; i < n_per_pass; i += nthr) atomicAdd(dst + i, 1.0f);
    } else if (mode == 1) {
      for (size_t i = tid * 4; i < n_per_pass; i += nthr * 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(dst + i), "f"(1.0f) : "memory");
    } else {
      for (size_t i = tid * 4; i < n_per_pass; i += nthr * 4) *reinterpret_cast<float4*>(dst + i) = make_float4(1, 1, 1, 1);
    }
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t n = (size_t)96 * 2400 * 128;  // the dQ tensor: 118 MB fp32 (fits L2 partially)
  const int passes = 19;                      // one pass per key tile
  float* d;
  cudaMalloc(&d, n * 4);
  cudaMemset(d, 0, n * 4);
  cudaFuncSetAttribute(k_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const char* names[4] = {"red.f32 (128B/warp-instr)", "red.v4.f32 (512B/warp-instr)", "TMA bulk reduce 16KB", "st.global.v4 (reference)"};
  for (int threads : {128, 512})
    for (int mode = 0; mode < 4; ++mode) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      k_red<<<sms, threads, 16384>>>(d, n, 1, mode);
      cudaEventRecord(e0);
      k_red<<<sms, threads, 16384>>>(d, n, passes, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("threads/SM %3d  %-30s %8.3f ms  %7.1f GB/s (%.2f GB)  err=%s\n", threads, names[mode], ms, n * 4.0 * passes / ms / 1e6,
             n * 4.0 * passes / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
