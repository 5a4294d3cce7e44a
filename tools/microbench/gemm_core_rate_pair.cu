// Micro-benchmark: the GEMM core of the fused kernels as a CTA PAIR (cta_group::2): M = 256 over two CTAs, N = 128 UMMAs with the A operand
// in each CTA's own tensor memory, every CTA staging only HALF of each 32 KiB weight unit (64 of the 128 B rows; the tensor core reads the
// other half from the peer's shared memory) by bulk TMA, the non-leader forwarding its "half landed" barriers to the leader, the leader
// issuing and committing with a cluster multicast.  Optionally an A-tile stream (each CTA TMA-loads its own tile, the leader's
// tcgen05.cp.cta_group::2 moves both into tensor memory).  Prints clk per weight unit (cta_group::1 core: 670; ideal 512).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

constexpr int WH = 16384, KB = 16384;

__device__ __forceinline__ void umma_ts_pair(uint32_t d, uint32_t ta, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(ta), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_cp_pair(uint32_t taddr, uint32_t desc_lo) {
  asm volatile(
      "{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\t"
      "tcgen05.cp.cta_group::2.128x256b [%0], d;\n\t}" ::"r"(taddr), "r"(desc_lo), "r"(UMMA_DESC_HI)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const uint8_t* img, int units, int slots, int a_every, long long* out, float* val) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_ring = smem;                         // up to 8 half-unit slots
  uint8_t* a_ring = smem + 8 * WH;                // 4 K-block slots
  __shared__ uint64_t w_full[8], w_empty[8], a_full[4], a_empty[4], done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = threadIdx.x; i < (8 * WH + 4 * KB) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    const uint32_t fullc = rank == 0 ? 2 : 1;
    for (int i = 0; i < 8; ++i) { mbar_init(&w_full[i], fullc); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], fullc); mbar_init(&a_empty[i], 1); }
    mbar_init(&done, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_alloc_pair(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = slot;
  if (warp == 0 && lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int u = 0; u < units; ++u) {
      mbar_wait_cluster(&w_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&w_full[s], WH);
      const uint8_t* src = img + (size_t)(u & 15) * 32768 + rank * 8192;
      tma_bulk_g2s(w_ring + s * WH, src, 8192, &w_full[s]);
      tma_bulk_g2s(w_ring + s * WH + 8192, src + 16384, 8192, &w_full[s]);
      if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && a_every > 0) {
    uint32_t cnt = 0;
    for (int u = 0; u < units; u += a_every)
      for (int kb = 0; kb < 4; ++kb, ++cnt) {
        const uint32_t s = cnt & 3, ph = (cnt >> 2) & 1;
        mbar_wait_cluster(&a_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&a_full[s], KB);
        tma_bulk_g2s(a_ring + s * KB, img + (size_t)(16 + (cnt & 15)) * KB, KB, &a_full[s]);
      }
  } else if (warp == 3 && lane == 0 && rank != 0) {
    // relay: forward my full barriers to the leader
    uint32_t s = 0, ph = 0, a_cnt = 0;
    for (int u = 0; u < units; ++u) {
      if (a_every > 0 && u % a_every == 0)
        for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
          mbar_wait_cluster(&a_full[a_cnt & 3], (a_cnt >> 2) & 1);
          mbar_arrive_cluster_relaxed(&a_full[a_cnt & 3], 0);
        }
      mbar_wait_cluster(&w_full[s], ph);
      mbar_arrive_cluster_relaxed(&w_full[s], 0);
      if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
    }
  } else if (warp == 2 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, 128);
    const uint32_t w_addr = smem_u32(w_ring), a_addr = smem_u32(a_ring);
    uint32_t s = 0, ph = 0, a_cnt = 0;
    const long long t0 = clock64();
    for (int u = 0; u < units; ++u) {
      if (a_every > 0 && u % a_every == 0) {
        for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
          const uint32_t as = a_cnt & 3;
          mbar_wait_cluster(&a_full[as], (a_cnt >> 2) & 1);
          tc_fence_after();
          const uint32_t lo = umma_desc_lo(a_addr + as * KB);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tmem_cp_pair(tmem + 128 + 32 * kb + 8 * kk, lo + 2 * kk);
            umma_commit_pair(&a_empty[as]);
          }
          __syncwarp();
        }
      }
      mbar_wait_cluster(&w_full[s], ph);
      tc_fence_after();
      const uint32_t w_lo = umma_desc_lo(w_addr + s * WH);
      const uint32_t d = tmem + 256 + (u & 1) * 128;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ts_pair(d, tmem + 128 + 8 * kk, w_lo + 2 * kk, idesc, kk ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_ts_pair(d, tmem + 160 + 8 * kk, w_lo + (8192 >> 4) + 2 * kk, idesc, 1u);
        umma_commit_pair(&w_empty[s]);
      }
      __syncwarp();
      if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit_pair(&done);
    __syncwarp();
    mbar_wait_cluster(&done, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  if (!(warp == 2 && rank == 0)) mbar_wait_cluster(&done, 0);
  tc_fence_after();
  if (warp == 0) {
    float v[32];
    tmem_ld32(tmem + 256, v);
    tmem_ld_wait();
    if (lane == 0 && blockIdx.x < 2) val[blockIdx.x] = v[0];
  }
  tc_fence_before(); __syncthreads();
  cluster_sync_all();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint8_t* img; cudaMalloc(&img, 1 << 20); cudaMemset(img, 0x3c, 1 << 20);
  long long* d; cudaMalloc(&d, 16);
  float* val; cudaMalloc(&val, 16);
  const int smem_bytes = 8 * WH + 4 * KB + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int units = 4096;
  for (int grid : {2, sms & ~1})
    for (int a_every : {0, 8, 4})
      for (int slots : {2, 3, 4, 6, 8}) {
        k<<<grid, 128, smem_bytes>>>(img, 64, slots, a_every, d, val);
        k<<<grid, 128, smem_bytes>>>(img, units, slots, a_every, d, val);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        float hv[2]; cudaMemcpy(hv, val, 8, cudaMemcpyDeviceToHost);
        printf("grid %3d  pair, A in TMEM  A stream every %d units  W half-slots %d : %.0f clk per unit (issue %.0f)  acc[0] = %g / %g  [%s]\n", grid, a_every, slots,
               (double)h[1] / units, (double)h[0] / units, hv[0], hv[1], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  return 0;
}
