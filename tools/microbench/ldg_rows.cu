// Micro-benchmark: the load side of qkv_direct_kernel in isolation.  PW producer warps per CTA (one CTA per SM) read the fp32 residual
// stream in the H-pass order (half a warp per 256-byte row segment, K-block-major) plus the clip-shared positional table, with NS register
// sets of 4 + 4 float4 in flight per lane, and fold the values into a checksum (optionally also writing the bf16 images to shared
// memory as the kernel does).  Prints aggregate TB/s of src bytes: the ceiling of this load scheme without the GEMM around it.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#ifndef NOPOS
#define NOPOS 0
#endif

__device__ __forceinline__ int hpass_row(int p, int T, int H, int W) {
  int h = p % H; int r = p / H; int t = r % T; r /= T; int w = r % W; int b = r / W;
  return ((b * T + t) * H + h) * W + w;
}

template <int MODE> __device__ __forceinline__ float4 ld16(const float4* p) {
  float4 r;
  if (MODE == 0) return __ldg(p);
  if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

template <int PW, int NS, bool STS, int MODE>
__global__ void __launch_bounds__(PW * 32, 1) k(const float* src, const float* pos, int rows, int tiles, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, c16 = lane & 15;
  constexpr int RPW = 128 / PW;            // rows per warp per tile
  constexpr int BPK = RPW / 8;             // batches (4 row pairs = 8 rows) per K-block per warp
  constexpr int NB = 4 * BPK;              // batches per tile
  float4 sv[NS][4], qv[NS][4];
  float acc = 0.f;
  auto issue = [&](int set, int tile, int b) {
    const int kb = b / BPK, sub = b % BPK;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pr = min(tile * 128 + pw * RPW + sub * 8 + 2 * j + half, rows - 1);
      const int c = hpass_row(pr, 2, 41, 41);
      sv[set][j] = ld16<MODE>(reinterpret_cast<const float4*>(src + (size_t)c * 256 + kb * 64) + c16);
      qv[set][j] = NOPOS ? make_float4(0.f, 0.f, 0.f, 0.f) : ld16<MODE>(reinterpret_cast<const float4*>(pos + (size_t)(c % 3362) * 256 + kb * 64) + c16);
    }
  };
  auto consume = [&](int set, int b) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 s = sv[set][j], q = qv[set][j];
      if (STS) {
        __nv_bfloat162 a = __floats2bfloat162_rn(s.x + q.x, s.y + q.y), bb = __floats2bfloat162_rn(s.z + q.z, s.w + q.w);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&bb);
        *reinterpret_cast<uint2*>(smem + ((b & 3) * 16384) + (pw * RPW + 2 * j + half) * 128 + c16 * 8) = u;
        a = __floats2bfloat162_rn(s.x, s.y); bb = __floats2bfloat162_rn(s.z, s.w);
        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&bb);
        *reinterpret_cast<uint2*>(smem + 65536 + ((b & 3) * 16384) + (pw * RPW + 2 * j + half) * 128 + c16 * 8) = u;
      } else {
        acc += s.x + q.y + s.z + q.w;
      }
    }
  };
  // flat batch index over all tiles of this CTA, NS - 1 batches ahead
  const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * NB;
  auto tile_of = [&](int fb) { return (int)blockIdx.x + (fb / NB) * (int)gridDim.x; };
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) if (s < total) issue(s, tile_of(s), s % NB);
  for (int fb0 = 0; fb0 < total; fb0 += NS) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int fb = fb0 + s;
      if (fb < total) {
        const int nx = fb + NS - 1;
        if (nx < total) issue((s + NS - 1) % NS, tile_of(nx), nx % NB);
        consume(s, fb % NB);
      }
    }
  }
  if (STS) acc = smem[threadIdx.x * 4];
  if (acc == 1234.5f) out[0] = acc;
}

template <int PW, int NS, bool STS, int MODE>
void run(const float* src, const float* pos, int rows, int tiles, float* out, int sms, int dyn = 200 * 1024) {
  cudaFuncSetAttribute(k<PW, NS, STS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PW, NS, STS, MODE><<<sms, PW * 32, dyn>>>(src, pos, rows, tiles, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<PW, NS, STS, MODE><<<sms, PW * 32, dyn>>>(src, pos, rows, tiles, out);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<PW, NS, STS, MODE>);
  printf("mode %d dyn smem %3d KiB  warps %2d  sets %d  (%3d KiB of src+pos loads in flight)  smem stores %d  regs %3d: %.1f us, %.2f TB/s of src  [%s]\n", MODE, dyn >> 10, PW, NS, PW * (NS - 1) * 4, (int)STS,
         fa.numRegs, ms * 1e3, (double)rows * 1024 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int rows = 42 * 2 * 41 * 41, tiles = (rows + 127) / 128;
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float *src, *pos, *out;
  cudaMalloc(&src, (size_t)rows * 1024); cudaMemset(src, 0, (size_t)rows * 1024);
  cudaMalloc(&pos, (size_t)3362 * 1024); cudaMemset(pos, 0, (size_t)3362 * 1024);
  cudaMalloc(&out, 16);
  run<4, 2, false, 0>(src, pos, rows, tiles, out, sms);
  run<4, 3, false, 0>(src, pos, rows, tiles, out, sms);
  run<8, 2, false, 0>(src, pos, rows, tiles, out, sms);
  run<8, 3, false, 0>(src, pos, rows, tiles, out, sms);
  run<16, 2, false, 0>(src, pos, rows, tiles, out, sms);
  run<16, 3, false, 0>(src, pos, rows, tiles, out, sms);
  run<32, 2, false, 0>(src, pos, rows, tiles, out, sms);
  run<32, 3, false, 0>(src, pos, rows, tiles, out, sms);
  run<16, 2, true, 0>(src, pos, rows, tiles, out, sms);
  run<32, 2, true, 0>(src, pos, rows, tiles, out, sms, 160 * 1024);
  run<16, 2, false, 0>(src, pos, rows, tiles, out, sms, 160 * 1024);
  run<32, 2, false, 0>(src, pos, rows, tiles, out, sms, 160 * 1024);
  return 0;
}
