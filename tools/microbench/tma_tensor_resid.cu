// Micro-benchmark behind DESIGN.md section 7 "what would come next (1)": how fast can ONE SM bring in a 128-row x 256-column fp32 residual tile
// (128 KiB; rows are whole 1 KiB token rows of a [rows, 256] matrix)
//   (a) the way the epilogues of traj_pair_kernel / ffn_n256_pair_kernel do it today: LDG.128, a warp instruction = 4 rows x 128 bytes, 32 loads per
//       thread of 256 threads, all issued before the first use (registers);
//   (b) as 2-D tensor-map TMA boxes of 32 columns x 128 rows (16 KiB, SWIZZLE_128B so that a thread-per-row reader is bank-conflict free) into a ring
//       of shared-memory slots, one elected thread issuing;
//   (c) like (b) with boxes of 32 columns x 32 rows (4 KiB, the granularity a per-warp epilogue would use).
// Every CTA streams tiles (persistent, 148 CTAs) of an array larger than L2 and of one that fits (L2-resident, like the prefetched residual rows).
// Prints bytes per clock per SM and aggregate TB/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tensor_resid tma_tensor_resid.cu   (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__device__ __forceinline__ void tma_tensor_2d_g2s(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// (a) register path
__global__ void __launch_bounds__(256, 1) k_ldg(const float* __restrict__ src, int tiles, float* sink, long long* clk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = warp >> 2, sub = lane >> 3, piece = lane & 7;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    float4 rr[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = tile * 128 + (warp & 3) * 32 + i * 4 + sub;
        rr[j][i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * 256 + 128 * g + 32 * j + piece * 4));
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc += rr[j][i].x + rr[j][i].y + rr[j][i].z + rr[j][i].w;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = clock64() - t0;
  if (acc == 123.456f) sink[0] = acc;
}

// (b) / (c) tensor-map TMA boxes of 32 columns x box_rows rows into `slots` slots of 16 KiB
__global__ void __launch_bounds__(64, 1) k_tma(const __grid_constant__ CUtensorMap map, int tiles, int slots, int box_rows, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[8], empty[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } fence_barrier_init(); }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int cg = 0; cg < 8; ++cg) {                      // eight 32-column groups = one 16 KiB slot each
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], 16384);
        for (int r0 = 0; r0 < 128; r0 += box_rows)
          tma_tensor_2d_g2s(smem + s * 16384 + r0 * 128, &map, cg * 32, tile * 128 + r0, &full[s]);
        if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)
      for (int cg = 0; cg < 8; ++cg) {
        mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == (uint32_t)slots) { s = 0; ph ^= 1; }
      }
    if (blockIdx.x == 0) clk[0] = clock64() - t0;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeFn encode = (EncodeFn)fn;
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 16);
  for (int big : {1, 0}) {
    const int rows = big ? 42 * 2 * 41 * 41 / 128 * 128 : 148 * 2 * 128;                  // 145 MB (> L2) / 38 MB (L2-resident after the warm-up)
    const int tiles = rows / 128;
    float* src; cudaMalloc(&src, (size_t)rows * 1024); cudaMemset(src, 0, (size_t)rows * 1024);
    const double bytes = (double)rows * 1024;
    const double per_cta_tiles = (double)((tiles + sms - 1) / sms);
    auto report = [&](const char* name, float ms, long long clk) {
      printf("%-44s %-12s %8.1f us  %6.2f TB/s  %6.1f B/clk per SM (CTA 0: %.0f clk per 128 KiB tile)  [%s]\n", name, big ? "145 MB" : "38 MB (L2)", ms * 1e3,
             bytes / (ms * 1e-3) / 1e12, 131072.0 * per_cta_tiles / (double)clk, (double)clk / per_cta_tiles, cudaGetErrorString(cudaGetLastError()));
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms; long long h;
    k_ldg<<<sms, 256>>>(src, tiles, sink, d);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_ldg<<<sms, 256>>>(src, tiles, sink, d);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    report("(a) LDG.128, 32 loads / thread, 256 threads", ms / 5, h);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024);
    for (int box_rows : {128, 32})
      for (int slots : {2, 4, 8}) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {256, (cuuint64_t)rows}, strides[1] = {1024};
        const cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
        const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }
        k_tma<<<sms, 64, 8 * 16384 + 1024>>>(map, tiles, slots, box_rows, d);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) k_tma<<<sms, 64, 8 * 16384 + 1024>>>(map, tiles, slots, box_rows, d);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        char name[96];
        snprintf(name, sizeof name, "(%c) TMA tensor boxes 32 x %3d, %d x 16 KiB slots", box_rows == 128 ? 'b' : 'c', box_rows, slots);
        report(name, ms / 5, h);
      }
    cudaFree(src);
  }
  return 0;
}
