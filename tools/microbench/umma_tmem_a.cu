// Micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form: D[tmem] += A[tmem] * B[smem desc]).
// Checks the assumed A layout -- row m in lane m, 32-bit column c of the A region holds bf16 elements k = 2c (low half) and
// 2c + 1 (high half) -- against a CPU reference, and measures the issue/throughput rate next to the smem-A form.
//   M = 128, N = 128, K = 64 (4 MMAs of K = 16; the A column offset advances by 8 per MMA).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// a: [128][64] bf16 row-major, b: [128 n][64 k] bf16 row-major (K-major), d: [128][128] fp32
// use_cp = 0: A written to TMEM with tcgen05.st (thread = row);  use_cp = 1: A staged in shared memory as a SWIZZLE_128B
// K-major tile image and moved with four tcgen05.cp.128x256b (one per K = 16 slice: 32 bytes per row -> 8 TMEM columns)
__global__ void __launch_bounds__(128, 1) k_check(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d, int iters, long long* clk, int use_cp, int N = 128) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // B tile image: SW128 K-major, 128 rows x 128 B
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(smem + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(b + r * 64 + c * 8);
  }
  uint8_t* a_img = smem + 16384;
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_img + sw128_offset(r, c)) = *reinterpret_cast<const uint4*>(a + r * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t t_d = tmem, t_a = tmem + 256;                 // D: columns [0,128), A: columns [256, 288)
  // A -> TMEM: thread = row; 64 bf16 = 32 packed columns
  if (!use_cp) {
    float v[32];
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a + tid * 64);
#pragma unroll
    for (int c = 0; c < 32; ++c) u[c] = src[c];                // element 2c in the low half (little endian)
    tmem_st32(t_a + ((uint32_t)(warp * 32) << 16), v);
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t b0 = smem_u32(smem);
    if (use_cp)
      for (int kk = 0; kk < 4; ++kk) tmem_cp_128x256b(t_a + 8 * kk, umma_desc_sw128(smem_u32(a_img) + kk * 32));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int kk = 0; kk < 4; ++kk) umma_bf16_ts(t_d, t_a + 8 * kk, umma_desc_sw128(b0 + kk * 32), idesc, (it | kk) ? 1u : 0u);
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    clk[0] = t1 - t0; clk[1] = t2 - t0;
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    float v[32];
    tmem_ld32(t_d + ((uint32_t)(warp * 32) << 16) + 32 * c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d[tid * 128 + 32 * c + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  std::vector<__nv_bfloat16> ha(128 * 64), hb(128 * 64);
  std::vector<float> fa(128 * 64), fb(128 * 64);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) {
    ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fa[i] = __bfloat162float(ha[i]);
    hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fb[i] = __bfloat162float(hb[i]);
  }
  __nv_bfloat16 *da, *db; float* dd; long long* dc;
  cudaMalloc(&da, 128 * 64 * 2); cudaMalloc(&db, 128 * 64 * 2); cudaMalloc(&dd, 128 * 128 * 4); cudaMalloc(&dc, 16);
  cudaMemcpy(da, ha.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), 128 * 64 * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int use_cp = 0; use_cp < 2; ++use_cp) {
  k_check<<<1, 128, 64 * 1024>>>(da, db, dd, 1, dc, use_cp);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> hd(128 * 128);
  cudaMemcpy(hd.data(), dd, 128 * 128 * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double r = 0;
      for (int k = 0; k < 64; ++k) r += (double)fa[m * 64 + k] * fb[n * 64 + k];
      maxerr = fmax(maxerr, fabs(r - hd[m * 128 + n])); maxref = fmax(maxref, fabs(r));
    }
  printf("A-from-TMEM via %s: max abs err %.3e (max |ref| %.3f) -> %s\n", use_cp ? "tcgen05.cp.128x256b from a SW128 smem image" : "tcgen05.st (packed bf16 pairs, k = 2c low / 2c+1 high)",
         maxerr, maxref, maxerr < 1e-3 * maxref ? "LAYOUT OK" : "MISMATCH");
  }
  const int iters = 2048;
  for (int N : {32, 64, 96, 128}) {
    k_check<<<1, 128, 64 * 1024>>>(da, db, dd, iters, dc, 0, N);
    cudaDeviceSynchronize();
    long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    printf("rate, N = %3d, A in TMEM: issue %.1f clk/MMA, total %.1f clk/MMA (full rate would be %d)\n", N, hc[0] / (iters * 4.0), hc[1] / (iters * 4.0), N / 2);
  }
  return 0;
}
