// Tensor-pipe rate of the SMALL tcgen05.mma shapes the attention core issues, from a converged warp with an elected lane and
// precomputed descriptors (the issue path of the real kernels), so the numbers are pipe rates, not issue costs:
//   S  : M = 128, N = NP, K = 16, A and B in shared memory (K-major SWIZZLE_64B)
//   PV : M = 128, N = 32 (and 64), K = 16, A in tensor memory, B in shared memory (MN-major SWIZZLE_64B)
// Operand contents are irrelevant (zeros).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
#include "../../axial_vs_b200/csrc/attn_tc.cuh"
using namespace axvs;

// mode 0: S-type, N = n_cols; mode 1: PV-type with N = n_cols; dep: 1 = every instruction accumulates into the SAME D, 0 = 4 different D regions round-robin
__global__ void __launch_bounds__(32, 1) k_rate(int mode, int n_cols, int dep, int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 16; i += 32) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncwarp(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = smem_u32(smem);
  const uint32_t a_lo = ((base & 0x3FFFFu) >> 4) | (1u << 16), b_lo = (((base + 8192) & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t v_lo = (((base + 8192) & 0x3FFFFu) >> 4) | (64u << 16);
  const uint32_t idesc = mode == 0 ? umma_idesc_bf16(128, n_cols) : (umma_idesc_bf16(128, n_cols) | (1u << 16));
  const uint32_t dstep = dep ? 0u : 64u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (elect_one()) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (mode == 0) umma_ss_raw(tmem + (k & 3) * dstep, a_lo + 2 * (k & 1), AT_DESC_HI_SW64, b_lo + 2 * (k & 1), AT_DESC_HI_SW64, idesc, 1u);
        else umma_ts_raw(tmem + (k & 3) * dstep, tmem + 256 + 8 * (k & 3), v_lo + (k & 3) * 64, AT_DESC_HI_SW64, idesc, 1u);
      }
    }
    __syncwarp();
  }
  const long long t1 = clock64();
  if (elect_one()) umma_commit(&bar);
  __syncwarp();
  mbar_wait(&bar, 0);
  const long long t2 = clock64();
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t0; }
  tc_fence_before(); __syncwarp(); tc_fence_after();
  tmem_dealloc(tmem, 512);
}

int main() {
  long long* dc; cudaMalloc(&dc, 16);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 512;
  struct C { int mode, n, dep; const char* what; };
  const C cases[] = {{0, 48, 1, "S  N=48 smem x smem, same D"}, {0, 48, 0, "S  N=48 smem x smem, 4 D regions"}, {0, 32, 0, "S  N=32"}, {0, 96, 0, "S  N=96"}, {0, 128, 0, "S  N=128"}, {0, 224, 0, "S  N=224"},
                     {1, 32, 1, "PV N=32 tmem x MN-major smem, same D"}, {1, 32, 0, "PV N=32, 4 D regions"}, {1, 64, 0, "PV N=64, 4 D regions"}};
  for (const C& c : cases) {
    k_rate<<<1, 32, 64 * 1024>>>(c.mode, c.n, c.dep, iters, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.what, cudaGetErrorString(e)); return 1; }
    long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    printf("%-44s issue %.1f clk/MMA, complete %.1f clk/MMA\n", c.what, hc[0] / (iters * 8.0), hc[1] / (iters * 8.0));
  }
  return 0;
}
