// Micro-benchmark: the steady-state rate of the fused kernels' GEMM core in isolation -- one issuer warp running N = 128 UMMAs with the A
// operand in tensor memory over a ring of 32 KiB weight units that a producer thread refills by 1-D bulk TMA from an L2-resident image
// (every unit is used once, exactly as in traj_ts / ffn / qkv_direct), optionally with an A-tile stream (TMA + tcgen05.cp of four K-block
// images every `a_every` units).  Prints clk per weight unit (ideal: 8 UMMAs x 64 clk = 512).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

constexpr int WU = 32768, KB = 16384;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* img, int units, int slots, int a_every, int ss_mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_ring = smem;
  uint8_t* a_ring = smem + 5 * WU;                 // 4 K-block slots
  __shared__ uint64_t w_full[8], w_empty[8], a_full[4], a_empty[4], done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait(&w_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&w_full[s], WU);
      tma_bulk_g2s(w_ring + s * WU, img + (size_t)(u & 15) * WU, WU, &w_full[s]);
      if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && a_every > 0) {
    uint32_t cnt = 0;
    for (int u = 0; u < units; u += a_every)
      for (int kb = 0; kb < 4; ++kb, ++cnt) {
        const uint32_t s = cnt & 3, ph = (cnt >> 2) & 1;
        mbar_wait(&a_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[s], KB);
        tma_bulk_g2s(a_ring + s * KB, img + (size_t)(16 + (cnt & 15)) * KB, KB, &a_full[s]);
      }
  } else if (warp == 2) {
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t w_addr = smem_u32(w_ring), a_addr = smem_u32(a_ring);
    uint32_t s = 0, ph = 0, a_cnt = 0;
    const long long t0 = clock64();
    for (int u = 0; u < units; ++u) {
      if (a_every > 0 && u % a_every == 0) {
        for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
          const uint32_t as = a_cnt & 3;
          mbar_wait(&a_full[as], (a_cnt >> 2) & 1);
          tc_fence_after();
          if (elect_one()) {
            if (!ss_mode) tmem_cp_kblock(tmem + 128 + 32 * kb, a_addr + as * KB);
            umma_commit(&a_empty[as]);
          }
          __syncwarp();
        }
      }
      mbar_wait(&w_full[s], ph);
      tc_fence_after();
      if (ss_mode) umma_unit_elect(tmem + 256 + (u & 1) * 128, a_addr, a_addr + KB, w_addr + s * WU, idesc, false, &w_empty[s], nullptr, nullptr, nullptr);
      else umma_unit_elect_ts(tmem + 256 + (u & 1) * 128, tmem + 128, tmem + 160, w_addr + s * WU, idesc, false, &w_empty[s], nullptr, nullptr);
      if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
    }
    const long long t1 = clock64();
    umma_commit_elect(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* img; cudaMalloc(&img, 1 << 20); cudaMemset(img, 0x3c, 1 << 20);
  long long* d; cudaMalloc(&d, 16);
  const int smem_bytes = 5 * WU + 4 * KB + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int units = 4096;
  for (int grid : {1, sms})
    for (int ss : {0, 1})
      for (int a_every : {0, 8, 4})
        for (int slots : {2, 3, 4, 5}) {
          if (ss && a_every == 0) continue;
          k<<<grid, 128, smem_bytes>>>(img, 64, slots, a_every, ss, d);
          k<<<grid, 128, smem_bytes>>>(img, units, slots, a_every, ss, d);
          cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("grid %3d  %s  A stream every %d units  W slots %d : %.0f clk per unit (issue %.0f)  [%s]\n", grid, ss ? "A in smem" : "A in TMEM", a_every, slots,
                 (double)h[1] / units, (double)h[0] / units, cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
