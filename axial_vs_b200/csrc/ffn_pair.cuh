// CTA-pair (cta_group::2) version of ffn_fused_kernel: two CTAs of a cluster process two adjacent 128-token tiles with
// ONE stream of M = 256 tensor-core instructions issued by the leader CTA.  Each CTA stages only HALF of every weight unit
// (the hardware reads the other half from the peer's shared memory), which halves the L2 -> SM weight traffic and the
// shared memory spent on the weight ring, and N = 128 UMMAs run at the full 64 clk rate (88 clk with cta_group::1;
// tools/microbench/umma_2cta.cu).
//
// Protocol (per CTA, identical barrier offsets in both CTAs):
//   * A-tile and weight TMA producers run in BOTH CTAs and signal their local full barriers; the non-leader's warp 10
//     ("relay") forwards each completed local full barrier to the leader's barrier (count 2 there) with a remote arrive;
//   * the leader's MMA warp waits its full barriers, issues tcgen05.mma.cta_group::2 and commits with a cluster multicast,
//     so empty / stage-full / h_free / acc_full barriers fire in both CTAs;
//   * epilogue -> MMA hand-offs (stage drained, h ready, acc drained) are remote arrives on the LEADER's barriers (count 16).
#pragma once
#include "ffn_fused.cuh"

namespace axvs {

constexpr int FP_A_SLOTS = 6;
constexpr int FP_W_SLOTS = 5;                  // half units: 16 KiB each
constexpr int FP_WH = 16384;                   // bytes of half a weight unit: [2 K-blocks][64 rows x 128 B]
constexpr int FP_SMEM_BYTES = FP_A_SLOTS * TF_KB + FF_H_BYTES + FP_W_SLOTS * FP_WH + FF_XCHG_BYTES + FF_BIAS_BYTES + 512;
static_assert(FP_SMEM_BYTES <= 232448, "ffn_pair_kernel exceeds the 227 KiB shared-memory limit");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FF_THREADS, 1) ffn_pair_kernel(const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* h_buf = a_ring + FP_A_SLOTS * TF_KB;
  uint8_t* w_ring = h_buf + FF_H_BYTES;
  float2* xchg = reinterpret_cast<float2*>(w_ring + FP_W_SLOTS * FP_WH);
  float* sb1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xchg) + FF_XCHG_BYTES);
  float* sb2 = sb1 + FF_MAX_DFFN;
  float* sg2 = sb2 + 256;
  float* sbe2 = sg2 + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbe2 + 256);
  uint64_t* a_full = bars;                    // [FP_A_SLOTS]
  uint64_t* a_empty = a_full + FP_A_SLOTS;
  uint64_t* w_full = a_empty + FP_A_SLOTS;    // [FP_W_SLOTS]
  uint64_t* w_empty = w_full + FP_W_SLOTS;
  uint64_t* s_full = w_empty + FP_W_SLOTS;    // [2]
  uint64_t* s_empty = s_full + 2;             // [2]   (leader's copy is the live one)
  uint64_t* h_ready = s_empty + 2;            //       (leader)
  uint64_t* h_free = h_ready + 1;
  uint64_t* acc_full = h_free + 1;
  uint64_t* acc_free = acc_full + 1;          //       (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NJ = p.d_ffn / 128;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pair_tiles = (p.tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    const uint32_t fullc = rank == 0 ? 2 : 1;              // leader: own producer + the peer's relay
    for (int i = 0; i < FP_A_SLOTS; ++i) { mbar_init(&a_full[i], fullc); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FP_W_SLOTS; ++i) { mbar_init(&w_full[i], fullc); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 16); }
    mbar_init(h_ready, 16);
    mbar_init(h_free, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.d_ffn; i += FF_THREADS) sb1[i] = p.b1[i];
  for (int i = threadIdx.x; i < 256; i += FF_THREADS) { sb2[i] = p.b2[i]; sg2[i] = p.ln2_g[i]; sbe2[i] = p.ln2_b[i]; }
  __syncthreads();
  cluster_sync_all();                                       // both CTAs' barriers are initialised before any remote arrive
  if (warp == 10) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue groups (both CTAs, own 128 rows)
    setmaxnreg_inc<224>();
    const int g = warp >> 2;
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t t_acc = tmem + lane_base + 128 * g;
    const int sub = lane >> 3, piece = lane & 7;
    uint8_t* stg = h_buf + warp * 4096;
    uint32_t it = 0;
    AXVS_PROF_DECL(7)
    for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
      const int tile = 2 * pt + (int)rank;                     // may be == p.tiles (odd tile count): every row masked
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        const uint32_t hc = it * NJ + j;
        const int stage = j & 1;
        AXVS_PROF_WAIT(0, mbar_wait_cluster(&s_full[stage], (hc >> 1) & 1))
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + stage * 128 + 64 * g;
        uint32_t hpk[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
          tmem_ld32(t_s + 32 * c, v); tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(sb1 + j * 128 + 64 * g + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i];
            hpk[c * 16 + 2 * i] = pack_bf16x2(fmaxf(v[4 * i] + bb.x, 0.f), fmaxf(v[4 * i + 1] + bb.y, 0.f));
            hpk[c * 16 + 2 * i + 1] = pack_bf16x2(fmaxf(v[4 * i + 2] + bb.z, 0.f), fmaxf(v[4 * i + 3] + bb.w, 0.f));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[stage], 0);
        AXVS_PROF_WAIT(1, mbar_wait_cluster(h_free, (hc & 1) ^ 1))
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 u = make_uint4(hpk[4 * q], hpk[4 * q + 1], hpk[4 * q + 2], hpk[4 * q + 3]);
          *reinterpret_cast<uint4*>(h_buf + g * TF_KB + sw128_offset(row_in_tile, q)) = u;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(h_ready, 0);
      }
      // ---- final: t = acc2 + b2 + s, LayerNorm2, store (see ffn_fused_kernel)
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
      AXVS_PROF_WAIT(2, mbar_wait_cluster(acc_full, it & 1))
      tc_fence_after();
      float ps[8], pq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        {
          float v[32];
          tmem_ld32(t_acc + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(acc_free, 0);
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
      asm volatile("bar.sync 1, 256;" ::: "memory");
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
    }
    AXVS_PROF_FLUSH(8 + 8 * g, 7, (warp & 3) == 0 && lane == 0 && rank == 0)
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer (own tile, both CTAs)
      uint32_t cnt = 0;
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
        int tile = 2 * pt + (int)rank;
        if (tile >= p.tiles) tile = p.tiles - 1;                 // dummy tile of an odd count: load something valid
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t slot = cnt % FP_A_SLOTS, phase = (cnt / FP_A_SLOTS) & 1;
          mbar_wait_cluster(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, p.s_img + ((size_t)tile * 4 + kb) * TF_KB, TF_KB, &a_full[slot]);
        }
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer: my half (64 rows) of every unit
      uint32_t slot = 0, phase = 0;
      auto push = [&](const uint8_t* img, int unit) {
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], FP_WH);
        const uint8_t* src = img + (size_t)unit * TF_WU + rank * 64 * 128;
        tma_bulk_g2s(w_ring + slot * FP_WH, src, 8192, &w_full[slot]);
        tma_bulk_g2s(w_ring + slot * FP_WH + 8192, src + TF_KB, 8192, &w_full[slot]);
        if (++slot == FP_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) { push(p.w1, 2 * j); push(p.w1, 2 * j + 1); }
          if (j >= 1) { push(p.w2, 2 * (j - 1)); push(p.w2, 2 * (j - 1) + 1); }
        }
      }
    } else if (warp == 10 && rank != 0) {
      // =============================================================== relay (non-leader): forward my full barriers to the leader
      if (lane == 0) {
        uint32_t a_cnt = 0, w_slot = 0, w_phase = 0;
        auto fwd_w = [&]() {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          mbar_arrive_cluster_relaxed(&w_full[w_slot], 0);
          if (++w_slot == FP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        };
        for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
          for (int j = 0; j <= NJ; ++j) {
            if (j < NJ) {
#pragma unroll 1
              for (int kg = 0; kg < 2; ++kg) {
                if (j == 0) {
#pragma unroll 1
                  for (int e = 0; e < 2; ++e) {
                    const uint32_t ac = a_cnt + 2 * kg + e;
                    const uint32_t s = ac % FP_A_SLOTS;
                    mbar_wait_cluster(&a_full[s], (ac / FP_A_SLOTS) & 1);
                    mbar_arrive_cluster_relaxed(&a_full[s], 0);
                  }
                }
                fwd_w();
              }
            }
            if (j >= 1) { fwd_w(); fwd_w(); }
          }
          a_cnt += 4;
        }
      }
    } else if (warp == 10) {
      // =============================================================== MMA issuer (leader CTA; converged warp, elected lane)
      const uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), h_addr = smem_u32(h_buf), w_ring_addr = smem_u32(w_ring);
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, s_cnt0 = 0, s_cnt1 = 0, it = 0;
      AXVS_PROF_DECL(5)
      auto w_wait = [&]() -> uint32_t {
        AXVS_PROF_WAIT(0, mbar_wait_cluster(&w_full[w_slot], w_phase))
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == FP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
            const int g = j & 1;
            const uint32_t sc = g ? s_cnt1 : s_cnt0;
            AXVS_PROF_WAIT(1, mbar_wait_cluster(&s_empty[g], (sc & 1) ^ 1))
            if (g) ++s_cnt1; else ++s_cnt0;
            tc_fence_after();
#pragma unroll 1
            for (int kg = 0; kg < 2; ++kg) {
              const uint32_t ac0 = a_cnt + 2 * kg, ac1 = ac0 + 1;
              const uint32_t s0 = ac0 % FP_A_SLOTS, s1 = ac1 % FP_A_SLOTS;
              if (j == 0) {
                AXVS_PROF_WAIT(2, mbar_wait_cluster(&a_full[s0], (ac0 / FP_A_SLOTS) & 1); mbar_wait_cluster(&a_full[s1], (ac1 / FP_A_SLOTS) & 1))
                tc_fence_after();
              }
              const uint32_t ws = w_wait();
              umma_unit_elect_pair(tmem + 256 + g * 128, a_ring_addr + s0 * TF_KB, a_ring_addr + s1 * TF_KB, w_ring_addr + ws * FP_WH, idesc, kg != 0,
                                   &w_empty[ws], j == NJ - 1 ? &a_empty[s0] : nullptr, j == NJ - 1 ? &a_empty[s1] : nullptr,
                                   kg == 1 ? &s_full[g] : nullptr);
            }
          }
          if (j >= 1) {
            const int jj = j - 1;
            const uint32_t hc = it * NJ + jj;
            if (jj == 0) AXVS_PROF_WAIT(4, mbar_wait_cluster(acc_free, (it & 1) ^ 1))
            AXVS_PROF_WAIT(3, mbar_wait_cluster(h_ready, hc & 1))
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              const uint32_t ws = w_wait();
              umma_unit_elect_pair(tmem + half * 128, h_addr, h_addr + TF_KB, w_ring_addr + ws * FP_WH, idesc, jj != 0, &w_empty[ws],
                                   half ? h_free : nullptr, (half && j == NJ) ? acc_full : nullptr, nullptr);
            }
          }
        }
        a_cnt += 4;
      }
      AXVS_PROF_FLUSH(0, 5, lane == 0)
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

}  // namespace axvs
