// Micro-benchmark: tcgen05.mma issue/throughput for M=128, N in {64,128,256}, K=16 (bf16, SS mode, SW128 K-major).
// One CTA per SM, one thread issues `iters` groups of 4 MMAs (one 64-wide K-block), optionally rotating accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__global__ void __launch_bounds__(128, 1) k(int N, int iters, int nacc, int mode, long long* out) {
  const int same_smem = 0;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t dbar[8];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // 192 KiB of bf16 ~0.0078
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dbar[i], 1); mbar_init(&done_bar, 1); mbar_arrive(&done_bar); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384 * 4;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = a0 + (same_smem ? 0 : (i & 3) * 16384);
      const uint32_t b = b0 + (same_smem ? 0 : (i & 3) * 32768);
      const uint32_t d = tmem + (i % nacc) * N;
      if (mode & 2) { mbar_wait(&done_bar, 0); tc_fence_after(); }
      for (int kk = 0; kk < 4; ++kk) umma_bf16(d, umma_desc_sw128(a + kk * 32), umma_desc_sw128(b + kk * 32), idesc, 1);
      if (mode & 1) umma_commit(&dbar[i & 7]);
      if (mode & 4) { umma_commit(&dbar[(i + 1) & 7]); }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048;
  for (int grid : {1, sms})
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2})
        for (int same : {0, 1, 2, 3, 7}) {
          if (nacc * N > 512) continue;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          k<<<grid, 128, 200 * 1024>>>(N, 64, nacc, same, d);   // warm
          cudaEventRecord(e0);
          k<<<grid, 128, 200 * 1024>>>(N, iters, nacc, same, d);
          cudaEventRecord(e1); cudaDeviceSynchronize();
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          double macs = (double)grid * iters * 4 * 128.0 * N * 16;
          printf("grid %3d N %3d nacc %d mode %d: issue %.1f clk/MMA, total %.1f clk/MMA, %.1f MAC/clk/SM, %.1f TFLOP/s (err %s)\n", grid, N, nacc, same,
                 h[0] / (iters * 4.0), h[1] / (iters * 4.0), 128.0 * N * 16 / (h[1] / (iters * 4.0)), 2 * macs / (ms * 1e-3) / 1e12,
                 cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
