// Micro-benchmark: tcgen05.ld throughput and latency.  W warps (4 .. 16) of one CTA read 32-lane x 32-column pieces of tensor memory in a
// loop, with D loads in flight per wait.  Prints clk per 4 KiB warp-load and the implied bytes/clk per SM; optionally with UMMAs running
// on the tensor pipe at the same time (mode 1) to see whether the accumulator writes compete with the reads.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

template <int D>
__global__ void __launch_bounds__(544, 1) k(int warps, int iters, int mma, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int stop;
  for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); stop = 0; fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 16) {
    // optional UMMA stream into columns [256, 512) while the other warps read [0, 256)
    if (mma && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384;
      int n = 0;
      while (*(volatile int*)&stop < warps && n < 200000) {
        for (int kk = 0; kk < 4; ++kk) umma_bf16(tmem + 256 + (n & 1) * 128, umma_desc_sw128(a0 + kk * 32), umma_desc_sw128(b0 + kk * 32), idesc, 1);
        if ((n & 15) == 15) { umma_commit(&bar); mbar_wait(&bar, (n >> 4) & 1); }
        ++n;
      }
      umma_commit(&bar);
      mbar_wait(&bar, (n >> 4) & 1);
    }
  } else if (warp < warps) {
    const uint32_t t = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      float v[D][32];
#pragma unroll
      for (int d = 0; d < D; ++d) tmem_ld32(t + ((i + d) & 1) * 128, v[d]);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < D; ++d) acc += v[d][0] + v[d][31];
    }
    long long t1 = clock64();
    if (lane == 0) { out[1 + warp] = t1 - t0; atomicAdd(&stop, 1); }
    if (acc == 1234.5f) out[0] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int D>
void run(int warps, int mma, long long* d) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  k<D><<<1, 544, 80 * 1024>>>(warps, 64, mma, d);
  k<D><<<1, 544, 80 * 1024>>>(warps, iters, mma, d);
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
  const double per_load = (double)mx / (iters * D);
  printf("warps %2d  loads in flight %d  umma %d : %.1f clk per 4 KiB warp-load, %.1f B/clk per SM (%.1f per SMSP)  [%s]\n", warps, D, mma, per_load,
         4096.0 * warps / per_load, 4096.0 * warps / per_load / 4, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  cudaMalloc(&d, 32 * 8);
  cudaMemset(d, 0, 32 * 8);
  for (int mma : {0, 1})
    for (int warps : {1, 4, 8, 12, 16}) {
      run<1>(warps, mma, d);
      run<2>(warps, mma, d);
      run<4>(warps, mma, d);
    }
  return 0;
}
