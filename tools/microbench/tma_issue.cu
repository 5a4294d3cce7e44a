// Micro-benchmark 5: cost of ISSUING 1-D TMA bulk copies from one thread (no ring back-pressure): 12 copies into 12 free
// slots, timed with clock64 around (a) arrive.expect_tx + cp.async.bulk pairs, (b) the copies alone on one barrier.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__global__ void __launch_bounds__(64, 1) k(const uint8_t* img, int bytes, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[12];
  if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < 12; ++i) {
        mbar_arrive_expect_tx(&full[i], bytes);
        tma_bulk_g2s(smem + (size_t)i * 16384, img + (size_t)i * bytes, bytes, &full[i]);
      }
    } else {
      mbar_arrive_expect_tx(&full[0], bytes * 12);
      for (int i = 0; i < 12; ++i) tma_bulk_g2s(smem + (size_t)i * 16384, img + (size_t)i * bytes, bytes, &full[0]);
    }
    long long t1 = clock64();
    if (mode == 0) { for (int i = 0; i < 12; ++i) mbar_wait(&full[i], 0); } else mbar_wait(&full[0], 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
}

int main() {
  uint8_t* img; cudaMalloc(&img, 1 << 20); cudaMemset(img, 1, 1 << 20);
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int grid : {1, sms})
    for (int mode : {0, 1})
      for (int bytes : {2048, 8192, 16384}) {
        k<<<grid, 64, 200 * 1024>>>(img, bytes, mode, d);
        k<<<grid, 64, 200 * 1024>>>(img, bytes, mode, d);
        cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("grid %3d mode %d bytes %5d: issue %.1f clk/copy, all landed after %lld clk (%s)\n", grid, mode, bytes, h[0] / 12.0, h[1],
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
