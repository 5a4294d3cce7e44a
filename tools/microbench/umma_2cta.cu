// Micro-benchmark 4: cta_group::2 UMMA (M = 256 across a CTA pair).  Checks the PTX protocol (cluster launch, paired TMEM
// allocation, leader-only issue, multicast commit) and measures clk/MMA for N = 128 / 256 per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst, uint32_t n) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(n) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t t, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(n) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(alo), "r"(blo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int N, int iters, long long* out, float* val) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 32 * 1024; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // 128 KiB of bf16 2^-7
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) tmem_alloc2(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  cluster_sync();
  const uint32_t tmem = slot;
  const uint32_t rank = cluster_ctarank();
  long long t0 = clock64(), t1 = 0;
  if (threadIdx.x < 32 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, N);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem) + 65536);
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma2(tmem + (i & 1) * N, a_lo + (i & 3) * 1024 + 2 * kk, b_lo + (i & 1) * 1024 + 2 * kk, idesc, i >= 2);
      }
      __syncwarp();
    }
    t1 = clock64();
    if (elect_one()) commit2(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  long long t2 = clock64();
  if (threadIdx.x < 32) {
    float v[32];
    tmem_ld32(tmem, v);
    tmem_ld_wait();
    if (threadIdx.x == 0) { val[blockIdx.x * 2] = v[0]; val[blockIdx.x * 2 + 1] = v[31]; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before(); __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* val; cudaMalloc(&val, 8 * 256);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 2048;
  for (int N : {64, 128, 256}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, 128, 160 * 1024>>>(N, 16, d, val);
    cudaEventRecord(e0);
    k<<<sms, 128, 160 * 1024>>>(N, iters, d, val);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    float hv[4]; cudaMemcpy(hv, val, 16, cudaMemcpyDeviceToHost);
    double macs = (double)sms * iters * 4 * 128.0 * N * 16;   // per SM: 128 rows x N x 16 per MMA
    // accumulator i&1: (iters/2 - 1) accumulating groups + 1 overwrite group = iters/2 groups of 4 MMAs x 16 x 2^-14
    printf("cta_group::2 M=256 N=%3d: issue %.1f clk/MMA, total %.1f clk/MMA, %.1f TFLOP/s; D[0][0]=%g D[0][31]=%g (expect %g), peer D=%g  (%s)\n", N,
           h[0] / (iters * 4.0), h[1] / (iters * 4.0), 2 * macs / (ms * 1e-3) / 1e12, hv[0], hv[1], (iters / 2) * 4 * 16 * 6.103515625e-05, hv[2],
           cudaGetErrorString(e));
  }
  return 0;
}
