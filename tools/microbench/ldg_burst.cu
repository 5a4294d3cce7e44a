// Micro-benchmark: per-SM global-load throughput as a function of the BURST size: every warp issues L independent LDG.128 (half a warp per
// 256-byte row segment, H-pass row order), then consumes them all, and repeats.  W warps per CTA, one CTA per SM, 200 KiB of dynamic
// shared memory requested (so the L1 is as small as in the fused kernels).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ int hpass_row(int p, int T, int H, int W) {
  int h = p % H; int r = p / H; int t = r % T; r /= T; int w = r % W; int b = r / W;
  return ((b * T + t) * H + h) * W + w;
}
template <int W, int L>
__global__ void __launch_bounds__(W * 32, 1) k(const float* src, int rows, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4, c16 = lane & 15;
  // a "burst" = L row pairs x one 256-byte K-block segment; bursts are dealt round-robin to all warps of all CTAs
  const long long total = (long long)rows / 2 * 4 / L;
  float acc = 0.f;
  for (long long b = (long long)blockIdx.x * W + warp; b < total; b += (long long)gridDim.x * W) {
    float4 v[L];
    const long long seg0 = b * L;               // segment index = (row pair, K-block)
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const long long seg = seg0 + j;
      const int pr = (int)(seg >> 2) * 2 + half, kb = (int)(seg & 3);
      v[j] = __ldg(reinterpret_cast<const float4*>(src + (size_t)hpass_row(pr, 2, 41, 41) * 256 + kb * 64) + c16);
    }
#pragma unroll
    for (int j = 0; j < L; ++j) acc += v[j].x + v[j].w;
  }
  if (acc == 1234.5f) out[0] = acc;
}
template <int W, int L>
void run(const float* src, int rows, float* out, int sms) {
  cudaFuncSetAttribute(k<W, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<W, L><<<sms, W * 32, 200 * 1024>>>(src, rows, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<W, L><<<sms, W * 32, 200 * 1024>>>(src, rows, out);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("warps %2d  burst %2d LDG.128 (%3d KiB in flight per SM): %.1f us, %.2f TB/s  [%s]\n", W, L, W * L * 512 / 1024, ms * 1e3, (double)rows * 1024 / (ms * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int rows = 42 * 2 * 41 * 41;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *src, *out;
  cudaMalloc(&src, (size_t)rows * 1024); cudaMemset(src, 0, (size_t)rows * 1024);
  cudaMalloc(&out, 16);
  run<4, 8>(src, rows, out, sms); run<4, 16>(src, rows, out, sms); run<4, 32>(src, rows, out, sms);
  run<8, 4>(src, rows, out, sms); run<8, 8>(src, rows, out, sms); run<8, 16>(src, rows, out, sms); run<8, 32>(src, rows, out, sms);
  run<16, 4>(src, rows, out, sms); run<16, 8>(src, rows, out, sms); run<16, 16>(src, rows, out, sms);
  run<24, 4>(src, rows, out, sms); run<24, 8>(src, rows, out, sms);
  return 0;
}
