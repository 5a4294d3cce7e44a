// Micro-test for the operand forms of the tcgen05 attention core (csrc/attn_tc.cuh):
//   (1) S = Q K^T with BOTH operands in shared memory as K-major SWIZZLE_64B tiles of 64-byte rows (one head: 32 bf16 per row),
//       M = 128, N = NP (multiple of 16), K = 32 (two K = 16 instructions, the second 32 bytes further into the row);
//   (2) O = P V with P in TENSOR MEMORY (packed bf16 pairs, as validated in umma_tmem_a.cu) and V in shared memory as an
//       MN-MAJOR SWIZZLE_64B operand: rows = keys (the K dimension), 32 contiguous head channels (the N dimension) per 64-byte
//       row -- i.e. V exactly as the q|k|v projection stores it, no transposition.  One K = 16 instruction reads 16 key rows
//       (1024 bytes); several descriptor variants (LBO / SBO roles) are tried and the one that matches the CPU result is printed.
// Row r, 16-byte chunk c of a tile lives at r * 64 + ((c ^ ((r >> 1) & 3)) << 4)   (Swizzle<2,4,3> on the byte address).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../axial_vs_b200/csrc/ptx.cuh"
using namespace axvs;

__host__ __device__ inline uint32_t sw64_off(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }

__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a),
               "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}

// q: [128][32], k: [NP][32], v: [NP][32] bf16 row-major; p: [128][NP] bf16;  s_out: [128][NP], o_out: [128][32] fp32
__global__ void __launch_bounds__(128, 1) k_check(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* pm,
                                                  float* s_out, float* o_out, int NP, int variant, int iters, long long* clk, int r0 = 0, int use_base_offset = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 128 x 64 B
  uint8_t* sK = smem + 8192;          // NP x 64 B (NP <= 256)
  uint8_t* sV = smem + 8192 + 20480;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 4; i += 128) *reinterpret_cast<uint4*>(sQ + sw64_off(i >> 2, i & 3)) = *reinterpret_cast<const uint4*>(q + (i >> 2) * 32 + (i & 3) * 8);
  for (int i = tid; i < (NP + 64) * 4; i += 128) {   // zero (finite) rows around the tiles
    *reinterpret_cast<uint4*>(sK + i * 16) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(sV + i * 16) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  for (int i = tid; i < NP * 4; i += 128) {          // tile row r lives at buffer row r0 + r, swizzled by its ABSOLUTE row
    *reinterpret_cast<uint4*>(sK + sw64_off(r0 + (i >> 2), i & 3)) = *reinterpret_cast<const uint4*>(k + (i >> 2) * 32 + (i & 3) * 8);
    *reinterpret_cast<uint4*>(sV + sw64_off(r0 + (i >> 2), i & 3)) = *reinterpret_cast<const uint4*>(v + (i >> 2) * 32 + (i & 3) * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t t_s = tmem, t_p = tmem + 256, t_o = tmem + 400;
  // P -> TMEM as packed pairs (thread = row)
  for (int c0 = 0; c0 < NP / 2; c0 += 8) {
    uint32_t u[8];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pm + (size_t)tid * NP) + c0;
#pragma unroll
    for (int c = 0; c < 8; ++c) u[c] = src[c];
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_p + ((uint32_t)(warp * 32) << 16) + c0), "r"(u[0]),
                 "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                 : "memory");
  }
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc_s = umma_idesc_bf16(128, NP);
    const uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);   // B is MN-major
    const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK) + r0 * 64, av = smem_u32(sV) + r0 * 64;
    const uint64_t bo = use_base_offset ? ((uint64_t)((ak >> 7) & 7) << 49) : 0;   // descriptor base_offset field, bits [49,52)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < 2; ++kk) umma_ss(t_s, desc_sw64(aq + kk * 32, 16, 512), desc_sw64(ak + kk * 32, 16, 512) | bo, idesc_s, kk ? 1u : 0u);
      for (int kk = 0; kk < NP / 16; ++kk) {
        uint64_t bd;
        if (variant == 0) bd = desc_sw64(av + kk * 1024, 1024, 512);        // SBO = 8 key rows apart, LBO = next 16-key slab (unused: N = one atom)
        else if (variant == 1) bd = desc_sw64(av + kk * 1024, 512, 1024);   // roles swapped
        else if (variant == 2) bd = desc_sw64(av + kk * 1024, 512, 512);
        else bd = desc_sw64(av + kk * 1024, 16, 512);
        umma_ts(t_o, t_p + 8 * kk, bd | bo, idesc_o, kk ? 1u : 0u);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    clk[0] = t1 - t0; clk[1] = clock64() - t0;
  }
  __syncthreads();
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < NP; c += 16) {
    float x[16];
    tmem_ld16(t_s + lane_base + c, x);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) s_out[tid * NP + c + i] = x[i];
  }
  {
    float x[32];
    tmem_ld32(t_o + lane_base, x);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) o_out[tid * 32 + i] = x[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  for (int NP : {48, 32, 16, 96, 176, 224}) {
    std::vector<__nv_bfloat16> hq(128 * 32), hk(NP * 32), hv(NP * 32), hp(128 * NP);
    srand(NP);
    auto rnd = [] { return (rand() % 2001 - 1000) / 1000.f; };
    for (auto& x : hq) x = __float2bfloat16(rnd());
    for (auto& x : hk) x = __float2bfloat16(rnd());
    for (auto& x : hv) x = __float2bfloat16(rnd());
    for (auto& x : hp) x = __float2bfloat16(fabsf(rnd()));
    __nv_bfloat16 *dq, *dk, *dv, *dp; float *ds, *dout; long long* dc;
    cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dk, hk.size() * 2); cudaMalloc(&dv, hv.size() * 2); cudaMalloc(&dp, hp.size() * 2);
    cudaMalloc(&ds, 128 * NP * 4); cudaMalloc(&dout, 128 * 32 * 4); cudaMalloc(&dc, 16);
    cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dk, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dv, hv.data(), hv.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dp, hp.data(), hp.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int variant = 0; variant < 4; ++variant) {
      k_check<<<1, 128, 80 * 1024>>>(dq, dk, dv, dp, ds, dout, NP, variant, 1, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("NP %d variant %d: launch failed: %s\n", NP, variant, cudaGetErrorString(e)); return 1; }
      std::vector<float> hs(128 * NP), ho(128 * 32);
      cudaMemcpy(hs.data(), ds, hs.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double es = 0, rs = 0, eo = 0, ro = 0;
      for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < NP; ++n) {
          double r = 0;
          for (int c = 0; c < 32; ++c) r += (double)__bfloat162float(hq[m * 32 + c]) * __bfloat162float(hk[n * 32 + c]);
          es = fmax(es, fabs(r - hs[m * NP + n])); rs = fmax(rs, fabs(r));
        }
        for (int c = 0; c < 32; ++c) {
          double r = 0;
          for (int n = 0; n < NP; ++n) r += (double)__bfloat162float(hp[m * NP + n]) * __bfloat162float(hv[n * 32 + c]);
          eo = fmax(eo, fabs(r - ho[m * 32 + c])); ro = fmax(ro, fabs(r));
        }
      }
      printf("NP %3d variant %d: S = Q K^T (SW64 K-major smem x smem) err %.3e / %.3f -> %s | O = P V (TMEM x MN-major SW64 smem) err %.3e / %.3f -> %s\n", NP, variant, es,
             rs, es < 1e-3 * rs ? "OK" : "MISMATCH", eo, ro, eo < 1e-3 * ro ? "OK" : "MISMATCH");
    }
    for (int r0 : {1, 2, 5, 8, 41}) for (int ubo = 0; ubo < 2; ++ubo) {
      k_check<<<1, 128, 80 * 1024>>>(dq, dk, dv, dp, ds, dout, NP, 0, 1, dc, r0, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("NP %d r0 %d: launch failed: %s\n", NP, r0, cudaGetErrorString(e)); return 1; }
      std::vector<float> hs(128 * NP), ho(128 * 32);
      cudaMemcpy(hs.data(), ds, hs.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double es = 0, rs = 0, eo = 0, ro = 0;
      for (int m = 0; m < 128; ++m) {
        for (int n = 0; n < NP; ++n) {
          double r = 0;
          for (int c = 0; c < 32; ++c) r += (double)__bfloat162float(hq[m * 32 + c]) * __bfloat162float(hk[n * 32 + c]);
          es = fmax(es, fabs(r - hs[m * NP + n])); rs = fmax(rs, fabs(r));
        }
        for (int c = 0; c < 32; ++c) {
          double r = 0;
          for (int n = 0; n < NP; ++n) r += (double)__bfloat162float(hp[m * NP + n]) * __bfloat162float(hv[n * 32 + c]);
          eo = fmax(eo, fabs(r - ho[m * 32 + c])); ro = fmax(ro, fabs(r));
        }
      }
      printf("NP %3d tiles start %2d rows into the buffer (absolute-row swizzle), base_offset field %s: S %s (%.2e)  O %s (%.2e)\n", NP, r0, ubo ? "set" : "0  ",
             es < 1e-3 * rs ? "OK" : "MISMATCH", es, eo < 1e-3 * ro ? "OK" : "MISMATCH", eo);
    }
    k_check<<<1, 128, 80 * 1024>>>(dq, dk, dv, dp, ds, dout, NP, 0, 1024, dc);
    cudaDeviceSynchronize();
    long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    printf("NP %3d rate: %d UMMAs per sub-item (2 x N=%d + %d x N=32): issue %.0f clk, total %.0f clk per sub-item\n", NP, 2 + NP / 16, NP, NP / 16, hc[0] / 1024.0, hc[1] / 1024.0);
  }
  return 0;
}
