// Micro-benchmark 3: how fast can every SM stream the SAME packed-weight image through a ring of 16 KiB stages with
// 1-D TMA bulk copies (the weight-broadcast pattern of the fused kernels)?  Consumer = one thread that waits + releases.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__global__ void __launch_bounds__(64, 1) k(const uint8_t* img, size_t img_bytes, int stage_bytes, int slots, int iters, int private_copy, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  const uint8_t* base = img + (private_copy ? (size_t)blockIdx.x * img_bytes : 0);
  const int per_img = (int)(img_bytes / stage_bytes);
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % slots, ph = (i / slots) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&full[s], stage_bytes);
      tma_bulk_g2s(smem + (size_t)s * stage_bytes, base + (size_t)(i % per_img) * stage_bytes, stage_bytes, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % slots, ph = (i / slots) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  const size_t img_bytes = 1 << 20;   // 1 MiB "weights" (FFN W1+W2 size)
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* img; cudaMalloc(&img, img_bytes * sms); cudaMemset(img, 1, img_bytes * sms);
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters_bytes = 64 << 20;
  for (int priv : {0, 1})
    for (int stage : {16384, 32768})
      for (int slots : {2, 4, 6, 8, 12}) {
        if ((size_t)stage * slots > 196 * 1024) continue;
        const int iters = iters_bytes / stage;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms, 64, 200 * 1024>>>(img, img_bytes, stage, slots, 64, priv, d);
        cudaEventRecord(e0);
        k<<<sms, 64, 200 * 1024>>>(img, img_bytes, stage, slots, iters, priv, d);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%s stage %5d slots %2d: %.1f clk/stage, %.1f B/clk/SM, %.2f TB/s aggregate (%s)\n", priv ? "private" : "shared ", stage, slots,
               h / (double)iters, (double)stage * iters / h, (double)stage * iters * sms / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
