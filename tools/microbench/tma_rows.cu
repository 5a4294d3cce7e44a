// Micro-benchmark: can bulk-TMA copies of 256-byte fp32 row segments (the H-pass access pattern of qkv_direct_kernel: pass-order row ->
// strided canonical row, one K-block = 64 channels = 256 B per row) keep the HBM pipe full?  One warp per CTA issues the copies (lane l ->
// rows l, l + 32, ...), a slot = one K-block of a 128-row tile (32 KiB), ring of S slots; the consumer only releases.  Also 1 KiB whole-row
// copies (slot = 32 rows) for comparison.  Prints aggregate TB/s over a 145 MB array (larger than L2).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__device__ __forceinline__ int hpass_row(int p, int T, int H, int W) {   // p = ((b*W + w)*T + t)*H + h -> ((b*T + t)*H + h)*W + w
  int h = p % H; int r = p / H; int t = r % T; r /= T; int w = r % W; int b = r / W;
  return ((b * T + t) * H + h) * W + w;
}

__global__ void __launch_bounds__(96, 1) k(const uint8_t* src, int rows, int tiles, int slots, int whole_rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int step = 0; step < 4; ++step) {              // K-block (segments) or 32-row group (whole rows)
        if (lane == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_arrive_expect_tx(&full[s], 32768); }
        __syncwarp();
        uint8_t* dst = smem + s * 32768;
        if (!whole_rows) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = lane + 32 * j;
            const int r = min(tile * 128 + rl, rows - 1);
            tma_bulk_g2s(dst + rl * 256, src + (size_t)hpass_row(r, 2, 41, 41) * 1024 + step * 256, 256, &full[s]);
          }
        } else {
          const int rl = step * 32 + lane;
          const int r = min(tile * 128 + rl, rows - 1);
          tma_bulk_g2s(dst + lane * 1024, src + (size_t)hpass_row(r, 2, 41, 41) * 1024, 1024, &full[s]);
        }
        if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int step = 0; step < 4; ++step) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
      }
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
}

int main() {
  const int rows = 42 * 2 * 41 * 41, tiles = (rows + 127) / 128;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* src; cudaMalloc(&src, (size_t)rows * 1024); cudaMemset(src, 1, (size_t)rows * 1024);
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int whole : {0, 1})
    for (int slots : {1, 2, 3, 4, 6}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      k<<<sms, 96, 200 * 1024>>>(src, rows, tiles, slots, whole, d);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) k<<<sms, 96, 200 * 1024>>>(src, rows, tiles, slots, whole, d);
      cudaEventRecord(e1); cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("%s slots %d (%3d KiB in flight): %.1f us, %.2f TB/s, CTA 0: %.0f clk per tile  [%s]\n", whole ? "1 KiB rows      " : "256 B segments  ", slots, slots * 32, ms * 1e3,
             (double)rows * 1024 / (ms * 1e-3) / 1e12, (double)h / ((tiles + sms - 1) / sms), cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
