// FFN tail with 256-column hidden chunks (default at fusion level >= 4): same math, interface and final epilogue as
// ffn_fused_kernel (ffn_fused.cuh), different GEMM schedule:
//   GEMM 1 runs as N = 256 UMMAs (A = LayerNorm1 tile in shared memory, B = one 32 KiB unit of 256 hidden rows x 64 K) --
//   128 clk per instruction, the full tensor rate, where the N = 128 form is held to ~90 clk by the shared-memory port;
//   the 256-column accumulator is drained by both epilogue groups (128 columns each), written back in place as bf16 pairs,
//   and GEMM 2 reads it from tensor memory at 64 clk per N = 128 instruction.
// One accumulator stage (TMEM: acc2 256 + stage 256), so GEMM 1 (j+1) follows GEMM 2 (j) in issue order and there are
// half as many hand-shakes per tile; 16.4 k clk of UMMA time per tile instead of 20.2 k.
#pragma once
#include "ffn_fused.cuh"

namespace axvs {

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_n256_kernel(const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need a 1024 B aligned base (no static smem in this kernel)
  uint8_t* a_ring = smem;
  uint8_t* h_buf = a_ring + FF_A_SLOTS * TF_KB;
  uint8_t* w_ring = h_buf + FF_H_BYTES;
  float2* xchg = reinterpret_cast<float2*>(w_ring + FF_W_SLOTS * TF_WU);   // [2 parity][2 group][128] (sum, sumsq)
  float* sb1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xchg) + FF_XCHG_BYTES);   // b1 [d_ffn]
  float* sb2 = sb1 + FF_MAX_DFFN;                                                           // b2 [256]
  float* sg2 = sb2 + 256;                                                                   // ln2 gamma
  float* sbe2 = sg2 + 256;                                                                  // ln2 beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbe2 + 256);
  uint64_t* a_full = bars;                    // [FF_A_SLOTS]
  uint64_t* a_empty = a_full + FF_A_SLOTS;    // [FF_A_SLOTS]
  uint64_t* w_full = a_empty + FF_A_SLOTS;    // [6]
  uint64_t* w_empty = w_full + FF_W_SLOTS;    // [6]
  uint64_t* s_full = w_empty + FF_W_SLOTS;    // [2]
  uint64_t* h_ready = s_full + 2;             // [2] epilogue -> MMA: the bf16 hidden chunk is in place in TMEM stage j & 1
  uint64_t* acc_full = h_ready + 2;           // MMA -> epilogue
  uint64_t* acc_free = acc_full + 1;          // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NJ = p.d_ffn / 256;                 // hidden chunks of 256 columns

  if (threadIdx.x == 0) {
    for (int i = 0; i < FF_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FF_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(&s_full[0], 1);
    mbar_init(&h_ready[0], 8);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  // biases / LayerNorm2 affine -> shared memory (the ~10 KiB of L1 left beside 217 KiB of smem cannot keep them hot)
  for (int i = threadIdx.x; i < p.d_ffn; i += FF_THREADS) sb1[i] = p.b1[i];
  for (int i = threadIdx.x; i < 256; i += FF_THREADS) { sb2[i] = p.b2[i]; sg2[i] = p.ln2_g[i]; sbe2[i] = p.ln2_b[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue groups
    setmaxnreg_inc<224>();   // 256*224 + 128*56 = 64512 = the CTA register pool at launch (384 x 168)
    const int g = warp >> 2;
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t t_acc = tmem + lane_base + 128 * g;          // my 128 output columns of acc2
    const int sub = lane >> 3, piece = lane & 7;
    uint8_t* stg = h_buf + warp * 4096;                         // per-warp transpose staging (final epilogue only)
    uint32_t it = 0;
    AXVS_PROF_DECL(7)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // ---- hidden chunks of 256 columns on ONE 256-column stage: group g drains its 128 columns (bias + ReLU -> bf16 pairs)
      // and writes them back in place over the first 64 columns of its region: the tensor-memory A operand of GEMM 2
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        const uint32_t hc = it * NJ + j;                         // global chunk counter
        AXVS_PROF_WAIT(0, mbar_wait(&s_full[0], hc & 1))
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + 128 * g;
        uint32_t hpk[64];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v0[32], v1[32];
          tmem_ld32(t_s + 64 * c, v0);
          tmem_ld32(t_s + 64 * c + 32, v1);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(sb1 + j * 256 + 128 * g + 64 * c);
#pragma unroll
          for (int i = 0; i < 8; ++i) {                         // packed fp32 adds, ReLU folded into the bf16x2 conversion
            const float4 bb = b4[i], bc = b4[8 + i];
            const float2 a0 = add_f32x2(make_float2(v0[4 * i], v0[4 * i + 1]), make_float2(bb.x, bb.y));
            const float2 a1 = add_f32x2(make_float2(v0[4 * i + 2], v0[4 * i + 3]), make_float2(bb.z, bb.w));
            const float2 c0 = add_f32x2(make_float2(v1[4 * i], v1[4 * i + 1]), make_float2(bc.x, bc.y));
            const float2 c1 = add_f32x2(make_float2(v1[4 * i + 2], v1[4 * i + 3]), make_float2(bc.z, bc.w));
            hpk[32 * c + 2 * i] = pack_bf16x2_relu(a0.x, a0.y);
            hpk[32 * c + 2 * i + 1] = pack_bf16x2_relu(a1.x, a1.y);
            hpk[32 * c + 16 + 2 * i] = pack_bf16x2_relu(c0.x, c0.y);
            hpk[32 * c + 16 + 2 * i + 1] = pack_bf16x2_relu(c1.x, c1.y);
          }
        }
        tmem_st32u(t_s, *reinterpret_cast<const uint32_t(*)[32]>(&hpk[0]));
        tmem_st32u(t_s + 32, *reinterpret_cast<const uint32_t(*)[32]>(&hpk[32]));
        AXVS_PROF_WAIT(6, tmem_st_wait())
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[0]);
      }
      // ---- final: t = acc2 + b2 + s, LayerNorm2, store.  Rows are one-per-thread in TMEM; a per-warp transpose through
      // shared memory (h_buf is idle: every GEMM 2 of this tile has retired) makes the global traffic row-segment
      // coalesced.  In the transposed domain lane (sub, piece) owns 4 columns of rows {4*i + sub}.
      // The residual (s + b2) is fetched BEFORE waiting for the accumulator so its latency hides behind the last GEMMs.
      const int row0 = tile * 128 + wq * 32;
      float4 t[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 128 * g + 32 * c + piece * 4;
        const float4 bb = *reinterpret_cast<const float4*>(sb2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + i * 4 + sub;
          float4 sres = (r < p.rows) ? __ldg(reinterpret_cast<const float4*>(p.s32 + (size_t)r * 256 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[c][i] = make_float4(sres.x + bb.x, sres.y + bb.y, sres.z + bb.z, sres.w + bb.w);
        }
      }
      AXVS_PROF_WAIT(2, mbar_wait(acc_full, it & 1))
      tc_fence_after();
#ifdef AXVS_WAIT_PROFILE
      const long long tf0_ = clock64();
#endif
      float ps[8], pq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        {
          float v[32];
          tmem_ld32(t_acc + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) {                                            // acc2 fully read: release it for the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + sub;
          const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
          float4 tv = t[c][i];
          tv.x += a.x; tv.y += a.y; tv.z += a.z; tv.w += a.w;
          t[c][i] = tv;
          ps[i] += tv.x + tv.y + tv.z + tv.w;
          pq[i] += tv.x * tv.x + tv.y * tv.y + tv.z * tv.z + tv.w * tv.w;
        }
        __syncwarp();
      }
#ifdef AXVS_WAIT_PROFILE
      prof_acc_[4] += clock64() - tf0_;
#endif
      // row statistics: reduce over the 8 lanes sharing a row, then combine with the other column half (other group)
      float2* xc = xchg + (it & 1) * 256;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
          pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], o);
        }
        if (piece == 0) xc[g * 128 + wq * 32 + i * 4 + sub] = make_float2(ps[i], pq[i]);
      }
      AXVS_PROF_WAIT(3, asm volatile("bar.sync 1, 256;" ::: "memory"))
      float mean[8], rstd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 other = xc[(g ^ 1) * 128 + wq * 32 + i * 4 + sub];
        mean[i] = (ps[i] + other.x) * (1.f / 256.f);
        const float var = fmaxf((pq[i] + other.y) * (1.f / 256.f) - mean[i] * mean[i], 0.f);
        rstd[i] = rsqrtf(var + p.eps);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 128 * g + 32 * c + piece * 4;
        const float4 gg = *reinterpret_cast<const float4*>(sg2 + col), be = *reinterpret_cast<const float4*>(sbe2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + i * 4 + sub;
          if (r < p.rows) {
            const float4 tv = t[c][i];
            *reinterpret_cast<float4*>(p.out + (size_t)r * 256 + col) =
                make_float4((tv.x - mean[i]) * rstd[i] * gg.x + be.x, (tv.y - mean[i]) * rstd[i] * gg.y + be.y,
                            (tv.z - mean[i]) * rstd[i] * gg.z + be.z, (tv.w - mean[i]) * rstd[i] * gg.w + be.w);
          }
        }
      }
#ifdef AXVS_WAIT_PROFILE
      prof_acc_[5] += clock64() - tf0_;
#endif
    }
    AXVS_PROF_FLUSH(8 + 8 * g, 7, (warp & 3) == 0 && lane == 0)
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t slot = cnt % FF_A_SLOTS, phase = (cnt / FF_A_SLOTS) & 1;
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, p.s_img + ((size_t)tile * 4 + kb) * TF_KB, TF_KB, &a_full[slot]);
        }
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer (32 KiB units)
      uint32_t slot = 0, phase = 0;
      AXVS_PROF_DECL(1)
      auto push = [&](const uint8_t* img, int unit) {
        AXVS_PROF_WAIT(0, mbar_wait(&w_empty[slot], phase ^ 1))
        mbar_arrive_expect_tx(&w_full[slot], TF_WU);
        tma_bulk_g2s(w_ring + slot * TF_WU, img + (size_t)unit * TF_WU, TF_WU, &w_full[slot]);
        if (++slot == FF_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      // consumption order of the issuer: GEMM 1 (0); then per chunk GEMM 2 (j) [4 units] and GEMM 1 (j + 1) [4 units]
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) push(p.w1, kb);             // N = 256 units: (chunk j, K-block kb) = 4 j + kb
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
#pragma unroll 1
          for (int u = 0; u < 4; ++u) push(p.w2, 2 * (2 * j + (u >> 1)) + (u & 1));      // (K group 2j + kg, output half)
          if (j + 1 < NJ) {
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) push(p.w1, 4 * (j + 1) + kb);
          }
        }
      }
      AXVS_PROF_FLUSH(32, 1, true)
    } else if (warp == 10) {
      // =============================================================== MMA issuer (converged warp, elected lane issues)
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, it = 0;
      AXVS_PROF_DECL(5)
      auto w_wait = [&]() -> uint32_t {
        AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == FF_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      const uint32_t idesc256 = umma_idesc_bf16(128, 256);
      // GEMM 1 of one chunk: four K-blocks, each one N = 256 unit (4 UMMAs at 128 clk: the full rate with A in shared memory)
      auto gemm1 = [&](int j) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t ac = a_cnt + kb, sl = ac % FF_A_SLOTS;
          if (j == 0) {
            AXVS_PROF_WAIT(2, mbar_wait(&a_full[sl], (ac / FF_A_SLOTS) & 1))
            tc_fence_after();
          }
          const uint32_t ws = w_wait();
          umma_kblock_elect(tmem + 256, a_ring_addr + sl * TF_KB, w_ring_addr + ws * TF_WU, idesc256, kb != 0,
                            &w_empty[ws], j == NJ - 1 ? &a_empty[sl] : nullptr);
        }
        umma_commit_elect(&s_full[0]);
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        gemm1(0);
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
          // GEMM 2, K-chunk j: acc2 += h (128 x 256, bf16 pairs in TMEM) * W2[:, 256 j : 256 (j + 1)]^T
          const uint32_t hc = it * NJ + j;
          if (j == 0) AXVS_PROF_WAIT(4, mbar_wait(acc_free, (it & 1) ^ 1))   // previous tile's final epilogue has drained acc2
          AXVS_PROF_WAIT(3, mbar_wait(&h_ready[0], hc & 1))
          tc_fence_after();
#pragma unroll 1
          for (int u = 0; u < 4; ++u) {
            const int kg = u >> 1, half = u & 1;
            const uint32_t t_h = tmem + 256 + 128 * kg;                      // group kg's 128 hidden columns as 64 packed columns
            const uint32_t ws = w_wait();
            umma_unit_elect_ts(tmem + half * 128, t_h, t_h + 32, w_ring_addr + ws * TF_WU, idesc, (j | kg) != 0,
                               &w_empty[ws], (u == 3 && j == NJ - 1) ? acc_full : nullptr, nullptr);
          }
          // the stage is overwritten only after GEMM 2 (j) above: the tensor pipe executes in issue order
          if (j + 1 < NJ) gemm1(j + 1);
        }
        a_cnt += 4;
      }
      AXVS_PROF_FLUSH(0, 5, lane == 0)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
