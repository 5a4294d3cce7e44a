// Micro-benchmark 2: cost of the MMA issue path.  Variants:
//   v0: single thread (lane 0 branch), descriptors rebuilt per MMA        (what the round-1 kernels did first)
//   v1: whole warp converged, elect_one predicate on MMA + commit, descriptors rebuilt
//   v2: as v1 with descriptors precomputed (hi word constant, lo word += offset)
// each with the realistic per-K-block protocol: wait(full) + fence + 4 MMAs + commit(empty).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_raw(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}

template <int V>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, full[4], empty[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_arrive(&full[i]); }   // full[] phase 0 complete
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = umma_idesc_bf16(128, N);
  const uint32_t a0 = smem_u32(smem), b0 = a0 + 16384 * 4;
  long long t0 = 0, t1 = 0, t2 = 0;
  if (V == 0) {
    if (threadIdx.x == 0) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int s = i & 3;
        mbar_wait(&full[s], 0);
        tc_fence_after();
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem + (i & 1) * N, umma_desc_sw128(a0 + s * 16384 + kk * 32), umma_desc_sw128(b0 + s * 32768 + kk * 32), idesc, 1);
        umma_commit(&empty[s]);
      }
      t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t2 = clock64();
    }
  } else if (threadIdx.x < 32) {
    const uint64_t da = umma_desc_sw128(a0), db = umma_desc_sw128(b0);
    const uint32_t ahi = (uint32_t)(da >> 32), bhi = (uint32_t)(db >> 32), alo0 = (uint32_t)da, blo0 = (uint32_t)db;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i & 3;
      mbar_wait(&full[s], 0);
      tc_fence_after();
      if (V == 1) {
        if (elect_one()) {
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + (i & 1) * N, umma_desc_sw128(a0 + s * 16384 + kk * 32), umma_desc_sw128(b0 + s * 32768 + kk * 32), idesc, 1);
          umma_commit(&empty[s]);
        }
      } else {
        const uint32_t alo = alo0 + s * (16384 >> 4), blo = blo0 + s * (32768 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_raw(tmem + (i & 1) * N, alo + kk * 2, ahi, blo + kk * 2, bhi, idesc, 1);
          umma_commit(&empty[s]);
        }
      }
      __syncwarp();
    }
    t1 = clock64();
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t2 = clock64();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int V>
void run(int N, long long* d, int sms) {
  const int iters = 2048;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<sms, 128, 200 * 1024>>>(N, 64, d);
  cudaEventRecord(e0);
  k<V><<<sms, 128, 200 * 1024>>>(N, iters, d);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  double macs = (double)sms * iters * 4 * 128.0 * N * 16;
  printf("v%d N %3d: issue %.1f clk/K-block (ideal TC %.0f), total %.1f clk/K-block, %.1f TFLOP/s (%s)\n", V, N, h[0] / (double)iters, 4.0 * N / 2,
         h[1] / (double)iters, 2 * macs / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int N : {128, 256}) { run<0>(N, d, sms); run<1>(N, d, sms); run<2>(N, d, sms); }
  return 0;
}
