// CTA-pair (cta_group::2) version of traj_ts_kernel: the two CTAs of a cluster process two adjacent 128-row tiles with ONE stream of
// M = 256 tensor-core instructions issued by the leader CTA.  Same math, TMEM plan and epilogue as traj_ts_kernel (traj_ts.cuh; reference
// WC/temporal_attention.py:61-75 + residual :204/:213 + norm1 :217).
//
// Why: a single CTA that streams every weight unit through shared memory ONCE per 128-row tile is bound by the shared-memory port, not by
// the tensor pipe: each N = 128 instruction reads 4 KiB of B and the TMA refills 4 KiB behind it = 128 B/clk at the 64 clk rate.  Measured
// (tools/microbench/gemm_core_rate.cu): 670 clk per 32 KiB unit instead of 512, 889 with the per-frame tcgen05.cp of the A tile.  With
// cta_group::2 each CTA stages only HALF of every unit (64 of its 128 rows; the tensor core reads the other half from the peer), so the
// port load halves: 512 clk per unit, 616 with the A copies (tools/microbench/gemm_core_rate_pair.cu) -- and the weight ring shrinks to
// 16 KiB slots, which pays for an A ring of 6 K-blocks.
//
// Protocol (identical barrier offsets in both CTAs):
//   * A-tile and weight TMA producers run in BOTH CTAs (own tile / own half of the unit) and signal their local full barriers; the
//     non-leader's warp 10 ("relay") forwards each completed local full barrier to the leader's (count 2 there) with a remote arrive;
//   * the leader's warp 10 waits its full barriers, issues tcgen05.cp.cta_group::2 / tcgen05.mma.cta_group::2 (both CTAs' shared
//     memory -> both CTAs' tensor memory at the same addresses) and commits with a cluster multicast, so a_empty / w_empty / s_full
//     fire in both CTAs;
//   * epilogue -> issuer hand-offs (stage drained, o ready) are remote arrives on the LEADER's barriers (count 8 / 16).
#pragma once
#include "traj_ts.cuh"

namespace axvs {

#ifndef TP_PREFETCH
#define TP_PREFETCH 1
#endif
constexpr int TP_A_SLOTS = 6;
constexpr int TP_W_SLOTS = 5;                   // half units
constexpr int TP_WH = 16384;                    // half a weight unit: [2 K-blocks][64 rows x 128 B]
constexpr int TP_SMEM_BYTES = TT_STG_BYTES + TP_A_SLOTS * TF_KB + TP_W_SLOTS * TP_WH + TF_BIAS_BYTES + 512;
static_assert(TP_SMEM_BYTES <= 232448, "traj_pair_kernel exceeds the 227 KiB shared-memory limit");

__device__ __forceinline__ void umma_bf16_ts_lo_pair(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One weight unit for a CTA pair with A in tensor memory (each CTA's own, same address): each CTA holds half of the unit's rows as
// [2 K-blocks][64 rows x 128 B] (K-blocks 8 KiB apart); 8 UMMAs with M = 256, N = 128.  Elected lane of a converged warp in the leader.
__device__ __forceinline__ void umma_unit_elect_ts_pair(uint32_t tmem_d, uint32_t ta0, uint32_t ta1, uint32_t w, uint32_t idesc, bool accumulate,
                                                        uint64_t* c0, uint64_t* c1) {
  const uint32_t w_lo = umma_desc_lo(w);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ts_lo_pair(tmem_d, ta0 + 8 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ts_lo_pair(tmem_d, ta1 + 8 * k, w_lo + (8192 >> 4) + 2 * k, idesc, 1u);
    if (c0) umma_commit_pair(c0);
    if (c1) umma_commit_pair(c1);
  }
  __syncwarp();
}
// One K-block image (128 rows x 64 bf16) of EACH CTA's shared memory (same offset) -> 32 columns of each CTA's tensor memory.
__device__ __forceinline__ void tmem_cp_kblock_pair(uint32_t taddr, uint32_t smem_addr) {
  const uint32_t lo = umma_desc_lo(smem_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    asm volatile(
        "{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\t"
        "tcgen05.cp.cta_group::2.128x256b [%0], d;\n\t}" ::"r"(taddr + 8 * k), "r"(lo + 2 * k), "r"(UMMA_DESC_HI)
        : "memory");
}

// Frame-major row order: both tiles of pair `pt` hold frame-0 tokens (the tile's x_diag image is its x_0 image) -> x_0 is neither loaded
// nor copied a second time.  Evaluated identically by the producers, the relay and the issuer.
#define TP_SKIP0(pt) (p.tm_rpad != 0 && 2 * (pt) + 1 < (p.tm_rpad >> 7))
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1) traj_pair_kernel(const TrajParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* stg_all = smem;
  uint8_t* a_ring = smem + TT_STG_BYTES;
  uint8_t* w_ring = a_ring + TP_A_SLOTS * TF_KB;
  float* sb_pq = reinterpret_cast<float*>(w_ring + TP_W_SLOTS * TP_WH);
  float* sb_v2 = sb_pq + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb_v2 + 256);
  uint64_t* a_full = bars;                       // [TP_A_SLOTS]
  uint64_t* a_empty = a_full + TP_A_SLOTS;       // [TP_A_SLOTS]
  uint64_t* w_full = a_empty + TP_A_SLOTS;       // [TP_W_SLOTS]
  uint64_t* w_empty = w_full + TP_W_SLOTS;       // [TP_W_SLOTS]
  uint64_t* s_full = w_empty + TP_W_SLOTS;       // [2]
  uint64_t* s_empty = s_full + 2;                // [2]  (the leader's copy is the live one)
  uint64_t* o_ready = s_empty + 2;               //      (leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_ready + 2);   // o_ready[g]: group g's half of o (heads 4g .. 4g+3) is in tensor memory
  uint32_t* tile_flag = tmem_slot + 2;     // tiles whose frames phase has started (paces the residual prefetcher, warp 11)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pair_tiles = (p.tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    const uint32_t fullc = rank == 0 ? 2 : 1;                // leader: own producer + the peer's relay
    for (int i = 0; i < TP_A_SLOTS; ++i) { mbar_init(&a_full[i], fullc); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TP_W_SLOTS; ++i) { mbar_init(&w_full[i], fullc); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    mbar_init(&o_ready[0], 8);
    mbar_init(&o_ready[1], 8);
    *tile_flag = 0;
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += TF_THREADS) { sb_pq[i] = p.b_pq[i]; sb_v2[i] = p.b_v2[i]; }
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' barriers are initialised before any remote arrive
  if (warp == 10) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = *tmem_slot;
  const int F = p.F;

  if (warp < 8) {
    // =============================================================== epilogue groups (both CTAs, own 128 rows)
    setmaxnreg_inc<224>();   // 256*224 + 128*56 = 64512 = the CTA register pool at launch (384 x 168)
    const int g = warp >> 2;                                     // group = TMEM stage = head quad
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_qp = tmem + lane_base + 64 * g;             // my 4 heads of q2 / o as bf16 pairs (16 columns per head)
    const uint32_t t_s = tmem + lane_base + 256 + 128 * g;       // my accumulator stage
    uint8_t* stg = stg_all + warp * 4096;
    uint32_t s_cnt = 0;                                          // items consumed on my stage
    uint32_t it = 0;                                             // tile iteration
    for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
      const int tile = 2 * pt + (int)rank;                       // may be == p.tiles (odd tile count): every row masked
      const bool trc_ = it == 3 && lane == 0 && (warp & 3) == 0 && pair == 0;
      const int tb_ = 100 + 100 * g + 200 * (int)rank;
      AXVS_TRACE(trc_, tb_ + 0)
      // ---- q2 of my 4 heads: (acc + bias) * scale*log2e -> bf16 pairs in Q2P; the stage is then free for the frame chunks
      mbar_wait_cluster(&s_full[g], s_cnt & 1);
      AXVS_TRACE(trc_, tb_ + 1)
      ++s_cnt;
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float v[32];
        tmem_ld32(t_s + 32 * j, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_bf16x2((v[2 * i] + sb_pq[128 * g + 32 * j + 2 * i]) * p.scale_log2e, (v[2 * i + 1] + sb_pq[128 * g + 32 * j + 2 * i + 1]) * p.scale_log2e);
        tmem_st16u(t_qp + 16 * j, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[g], 0);
      AXVS_TRACE(trc_, tb_ + 2)
#if TP_PREFETCH
      if (warp == 0 && lane == 0) atomicExch(tile_flag, it + 1);   // the frames phase of this tile starts: ~10 k clk until the residual is needed
                                                                    // (a pacing hint only, no data hangs on it; atomics keep racecheck quiet)
#endif

      float m_run[4], l_run[4], o[4][32];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        m_run[h] = -INFINITY;
        l_run[h] = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h][i] = 0.f;
      }
      // ---- frames: two 128-column chunks (2 heads each) per frame on my stage
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          AXVS_TRACE(trc_, tb_ + 10 + 4 * (2 * f + cc))
          mbar_wait_cluster(&s_full[g], s_cnt & 1);
          AXVS_TRACE(trc_, tb_ + 11 + 4 * (2 * f + cc))
          ++s_cnt;
          tc_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            // two TMEM round trips per head: (k2, q2) then v2, the v2 load issued before the logit arithmetic
            const int lh = cc * 2 + hh;
            float k2[32], q2f[16], v2[32];
            tmem_ld32(t_s + 32 * hh, k2);
            tmem_ld16(t_qp + 16 * lh, q2f);
            tmem_ld_wait();
            tmem_ld32(t_s + 64 + 32 * hh, v2);
            float2 sacc = make_float2(0.f, 0.f);                // packed fp32 FMAs: two channels per instruction
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t u = __float_as_uint(q2f[i]);
              sacc = fma_f32x2(make_float2(bf16lo_to_f32(u), bf16hi_to_f32(u)), make_float2(k2[2 * i], k2[2 * i + 1]), sacc);
            }
            const float s = sacc.x + sacc.y;
            const float mn = fmaxf(m_run[lh], s);
            const float corr = exp2f(m_run[lh] - mn);
            const float pe = exp2f(s - mn);
            l_run[lh] = l_run[lh] * corr + pe;
            m_run[lh] = mn;
            tmem_ld_wait();
            if (hh == 1) {                                       // both heads of the chunk are in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[g], 0);
              AXVS_TRACE(trc_, tb_ + 12 + 4 * (2 * f + cc))
            }
            const float2 pe2 = make_float2(pe, pe), corr2 = make_float2(corr, corr);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 r = fma_f32x2(pe2, make_float2(v2[2 * i], v2[2 * i + 1]), mul_f32x2(make_float2(o[lh][2 * i], o[lh][2 * i + 1]), corr2));
              o[lh][2 * i] = r.x;
              o[lh][2 * i + 1] = r.y;
            }
          }
        }
      }
      // ---- o = o / l + bv2 -> bf16 pairs over my (now dead) q2 columns: the tensor-memory A operand of the output projection
#pragma unroll
      for (int lh = 0; lh < 4; ++lh) {
        const float inv = 1.f / l_run[lh];
        const int col0 = 128 * g + 32 * lh;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_bf16x2(fmaf(o[lh][2 * i], inv, sb_v2[col0 + 2 * i]), fmaf(o[lh][2 * i + 1], inv, sb_v2[col0 + 2 * i + 1]));
        tmem_st16u(t_qp + 16 * lh, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(&o_ready[g], 0);
      AXVS_TRACE(trc_, tb_ + 41)
      // The residual loads below fill the load/store queue for ~2k clk; issued while the OTHER group still reads its bias from shared
      // memory for o (same queue) they delayed its o_ready arrive -- the start of GEMM 3 -- by that much.  So: wait until both groups
      // have handed o over (the loads then overlap GEMM 3 instead of the hand-off).
      // ---- output projection item on my stage: out = resid + acc + bproj (my 128 output columns).
      // TMEM rows are one-per-thread; a 4 KiB per-warp transpose through shared memory turns the global accesses into full
      // 128-byte row segments.  resid + bias are fetched BEFORE waiting for the accumulator (latency hides behind GEMM 3).
      {
        const int r = tile * 128 + row_in_tile;
        const int my_orow = traj_row_canonical(r, p);
        const int sub = lane >> 3, piece = lane & 7;
        int orow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) orow[i] = __shfl_sync(0xffffffffu, my_orow, i * 4 + sub);
        float4 rr[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 128 * g + 32 * j + piece * 4;
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b_proj + col));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x = (p.resid && orow[i] >= 0) ? __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)orow[i] * 256 + col))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            rr[j][i] = make_float4(x.x + bb.x, x.y + bb.y, x.z + bb.z, x.w + bb.w);
          }
        }
        AXVS_TRACE(trc_, tb_ + 42)
        mbar_wait_cluster(&s_full[g], s_cnt & 1);
        AXVS_TRACE(trc_, tb_ + 43)
        ++s_cnt;
        tc_fence_after();
        if (p.ln_g == nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
              if (j == 3) {                                        // accumulator fully read: the stage is free for the next tile
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[g], 0);
                AXVS_TRACE(trc_, tb_ + 44)
              }
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
            __syncwarp();
            const int col = 128 * g + 32 * j + piece * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub;
              const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
              if (orow[i] >= 0)
                *reinterpret_cast<float4*>(p.out + (size_t)orow[i] * 256 + col) =
                    make_float4(a.x + rr[j][i].x, a.y + rr[j][i].y, a.z + rr[j][i].z, a.w + rr[j][i].w);
            }
            __syncwarp();
          }
        } else {
          // ---- fused LayerNorm: keep the row values in registers, combine the statistics of the two column halves through
          // shared memory (the partner warp's staging area), then write fp32 rows + the bf16 tile image of the normalised row
          float ps[8], pq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
              if (j == 3) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[g], 0);
                AXVS_TRACE(trc_, tb_ + 44)
              }
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub;
              const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
              float4 tv = rr[j][i];
              tv.x += a.x; tv.y += a.y; tv.z += a.z; tv.w += a.w;
              rr[j][i] = tv;
              ps[i] += tv.x + tv.y + tv.z + tv.w;
              pq[i] += tv.x * tv.x + tv.y * tv.y + tv.z * tv.z + tv.w * tv.w;
            }
            __syncwarp();
          }
          // statistics exchange: my 32 rows' (sum, sumsq) go to the head of my staging area (parity-alternating halves), the
          // partner warp (same rows, other column half) reads them after the 256-thread barrier
          float2* xc_mine = reinterpret_cast<float2*>(stg + (it & 1) * 256);
          const float2* xc_other = reinterpret_cast<const float2*>(stg_all + (warp ^ 4) * 4096 + (it & 1) * 256);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
              pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], o);
            }
            if (piece == 0) xc_mine[i * 4 + sub] = make_float2(ps[i], pq[i]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          float mean[8], rstd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 other = xc_other[i * 4 + sub];
            mean[i] = (ps[i] + other.x) * (1.f / 256.f);
            const float var = fmaxf((pq[i] + other.y) * (1.f / 256.f) - mean[i] * mean[i], 0.f);
            rstd[i] = rsqrtf(var + p.ln_eps);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = 128 * g + 32 * j + piece * 4;
            const float4 gg = __ldg(reinterpret_cast<const float4*>(p.ln_g + col)), be = __ldg(reinterpret_cast<const float4*>(p.ln_b + col));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 tv = rr[j][i];
              const float4 y = make_float4((tv.x - mean[i]) * rstd[i] * gg.x + be.x, (tv.y - mean[i]) * rstd[i] * gg.y + be.y,
                                           (tv.z - mean[i]) * rstd[i] * gg.z + be.z, (tv.w - mean[i]) * rstd[i] * gg.w + be.w);
              // 16-byte image chunk = 8 channels = this lane (even piece) + its odd neighbour
              const uint32_t lo = pack_bf16x2(y.x, y.y), hi = pack_bf16x2(y.z, y.w);
              const uint32_t nlo = __shfl_down_sync(0xffffffffu, lo, 1), nhi = __shfl_down_sync(0xffffffffu, hi, 1);
              if (orow[i] >= 0) {
                *reinterpret_cast<float4*>(p.out + (size_t)orow[i] * 256 + col) = y;
                if ((piece & 1) == 0) {
                  const uint32_t orw = (uint32_t)orow[i];
                  *reinterpret_cast<uint4*>(p.ln_img + ((size_t)(orw >> 7) * 4 + (col >> 6)) * TF_KB + sw128_offset(orw & 127u, (col & 63) >> 3)) =
                      make_uint4(lo, hi, nlo, nhi);
                }
              }
            }
          }
          // (no second barrier: my next write to this staging area is the next tile's projection transpose, which cannot start before
          //  the partner warp has arrived on o_ready for that tile -- i.e. long after it has read these statistics)
        }
        AXVS_TRACE(trc_, tb_ + 45)
      }
    }
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer (own tile: x_diag, x_0 .. x_{F-1})
      uint32_t slot = 0, phase = 0;
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
        int tile = 2 * pt + (int)rank;
        if (tile >= p.tiles) tile = p.tiles - 1;                 // dummy tile of an odd count: load something valid
        const bool skip0 = TP_SKIP0(pt);
#pragma unroll 1
        for (int item = 0; item < 4 * (F + 1); ++item) {
          if (skip0 && item >= 4 && item < 8) continue;          // x_0 of a frame-0 tile is its x_diag: already in tensor memory
          const uint8_t* src = (item < 4) ? traj_xd_tile(tile, p) + (size_t)item * TF_KB
                                          : p.x_img + (((size_t)((item >> 2) - 1) * p.tiles + tile) * 4 + (item & 3)) * TF_KB;
          mbar_wait_cluster(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, src, TF_KB, &a_full[slot]);
          if (++slot == TP_A_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer: my half (64 rows) of every unit
      uint32_t slot = 0, phase = 0;
      auto push = [&](const uint8_t* img, int unit) {
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], TP_WH);
        const uint8_t* src = img + (size_t)unit * TF_WU + rank * 64 * 128;
        tma_bulk_g2s(w_ring + slot * TP_WH, src, 8192, &w_full[slot]);
        tma_bulk_g2s(w_ring + slot * TP_WH + 8192, src + TF_KB, 8192, &w_full[slot]);
        if (++slot == TP_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
        for (int u = 0; u < 4; ++u) push(p.w_pq, u);               // (half, kg) = (u >> 1, u & 1)
#pragma unroll 1
        for (int i = 0; i < 8 * F; ++i) {
          const int ci = (i >> 1) & 3;
          const int c = ((ci & 1) << 1) | (ci >> 1);               // chunk order 0,2,1,3: stages alternate
          push(p.w_pkv, c * 2 + (i & 1));
        }
#pragma unroll 1
        for (int u = 0; u < 4; ++u) push(p.w_proj, ((u & 1) << 1) | (u >> 1));   // K group 0 of both row tiles first (see GEMM 3)
      }
#if TP_PREFETCH
    } else if (warp == 11 && p.resid != nullptr) {
      // =============================================================== residual prefetcher: pulls the tile's 128 fp32 rows into L2 while the
      // frames are computed, so that the epilogue's loads in front of GEMM 3 are L2 hits (no registers or shared memory needed)
      uint32_t it = 0;
      for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
        const int tile = 2 * pt + (int)rank;
        while (atomicAdd(tile_flag, 0u) < it + 1) __nanosleep(200);
#pragma unroll 1
        for (int rr = lane; rr < 128; rr += 32) {
          const int r = tile * 128 + rr;
          const int cr = traj_row_canonical(r, p);
          if (cr >= 0) {
            const char* src = reinterpret_cast<const char*>(p.resid + (size_t)cr * 256);
#pragma unroll
            for (int l = 0; l < 8; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 128 * l));
          }
        }
      }
#endif
    } else if (warp == 10 && rank != 0) {
      // =============================================================== relay (non-leader): forward my full barriers to the leader,
      // in the order the leader consumes them
      if (lane == 0) {
        uint32_t a_slot = 0, a_phase = 0, w_slot = 0, w_phase = 0;
        auto fwd_a = [&]() {
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait_cluster(&a_full[a_slot], a_phase);
            mbar_arrive_cluster_relaxed(&a_full[a_slot], 0);
            if (++a_slot == TP_A_SLOTS) { a_slot = 0; a_phase ^= 1; }
          }
        };
        auto fwd_w = [&](int n) {
#pragma unroll 1
          for (int i = 0; i < n; ++i) {
            mbar_wait_cluster(&w_full[w_slot], w_phase);
            mbar_arrive_cluster_relaxed(&w_full[w_slot], 0);
            if (++w_slot == TP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          }
        };
        for (int pt = pair; pt < pair_tiles; pt += npairs) {
          fwd_a();
          fwd_w(4);
#pragma unroll 1
          for (int f = 0; f < F; ++f) { if (!(f == 0 && TP_SKIP0(pt))) fwd_a(); fwd_w(8); }
          fwd_w(4);
        }
      }
    } else if (warp == 10) {
      // =============================================================== tcgen05.cp + MMA issuer (leader CTA; converged warp, elected lane)
      const uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
      const uint32_t t_xa = tmem + 128, t_qp = tmem;
      uint32_t a_slot = 0, a_phase = 0, w_slot = 0, w_phase = 0, s_cnt0 = 0, s_cnt1 = 0, it = 0;
      auto w_wait = [&]() -> uint32_t {
        mbar_wait_cluster(&w_full[w_slot], w_phase);
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == TP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      auto stage_wait = [&](int g) {
        const uint32_t sc = g ? s_cnt1 : s_cnt0;
        mbar_wait_cluster(&s_empty[g], (sc & 1) ^ 1);
        if (g) ++s_cnt1; else ++s_cnt0;
      };
      // the next A tile of BOTH CTAs (4 K-block images each) -> XA; ordered by the tensor pipe behind the UMMAs still reading the old tile
      auto copy_tile = [&]() {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait_cluster(&a_full[a_slot], a_phase);
          tc_fence_after();
          if (elect_one()) {
            tmem_cp_kblock_pair(t_xa + 32 * kb, a_ring_addr + a_slot * TF_KB);
            umma_commit_pair(&a_empty[a_slot]);
          }
          __syncwarp();
          if (++a_slot == TP_A_SLOTS) { a_slot = 0; a_phase ^= 1; }
        }
      };
      for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
        // ---- GEMM 1: q2 halves -> the two stages (free once the previous tile's projection has been drained in both CTAs)
        const bool trc_ = it == 3 && lane == 0 && pair == 0;
        AXVS_TRACE(trc_, 0)
        copy_tile();
        AXVS_TRACE(trc_, 3)
        AXVS_TRACE(trc_, 1)
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u >> 1, kg = u & 1;
          if (kg == 0) { stage_wait(half); tc_fence_after(); }  // half 0 starts as soon as group 0 has drained ITS accumulator of the previous tile
          const uint32_t ws = w_wait();
          umma_unit_elect_ts_pair(tmem + 256 + half * 128, t_xa + 64 * kg, t_xa + 64 * kg + 32, w_ring_addr + ws * TP_WH, idesc, kg != 0,
                                  &w_empty[ws], kg == 1 ? &s_full[half] : nullptr);
        }
        AXVS_TRACE(trc_, 2)
        // ---- GEMM 2: per frame, four 128-column chunks alternating between the two TMEM stages
#pragma unroll 1
        for (int f = 0; f < F; ++f) {
          AXVS_TRACE(trc_, 60 + 2 * f)
          if (!(f == 0 && TP_SKIP0(pt))) copy_tile();            // frame-0 tiles: XA still holds x_diag = x_0
          AXVS_TRACE(trc_, 61 + 2 * f)
#pragma unroll 1
          for (int ci = 0; ci < 4; ++ci) {
            const int g = ci & 1;                               // chunk order 0,2,1,3 -> stage 0,1,0,1
            AXVS_TRACE(trc_, 10 + 4 * (4 * f + ci))
            stage_wait(g);
            tc_fence_after();
            AXVS_TRACE(trc_, 11 + 4 * (4 * f + ci))
#pragma unroll 1
            for (int kg = 0; kg < 2; ++kg) {
              const uint32_t ws = w_wait();
              umma_unit_elect_ts_pair(tmem + 256 + g * 128, t_xa + 64 * kg, t_xa + 64 * kg + 32, w_ring_addr + ws * TP_WH, idesc, kg != 0,
                                      &w_empty[ws], kg == 1 ? &s_full[g] : nullptr);
              AXVS_TRACE(trc_, 12 + kg + 4 * (4 * f + ci))
            }
          }
        }
        // ---- GEMM 3: output projection, A = o (bf16 pairs written over q2 by both CTAs' epilogues), accumulators = both stages
        AXVS_TRACE(trc_, 50)
        // K group 0 (channels 0-127 = heads 0-3) is group 0's half of o, K group 1 group 1's: the first K group of both accumulators is issued
        // as soon as group 0 has handed its half over, while group 1 (whose last chunk of the last frame completes later) still packs its
        // own.  Per accumulator the order stays K group 0, then 1: bit-identical to the single-barrier schedule.
        mbar_wait_cluster(&o_ready[0], it & 1);
        AXVS_TRACE(trc_, 51)
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u & 1, kg = u >> 1;
          if (kg == 0) { stage_wait(half); tc_fence_after(); }
          else if (half == 0) { mbar_wait_cluster(&o_ready[1], it & 1); tc_fence_after(); AXVS_TRACE(trc_, 52) }
          const uint32_t ws = w_wait();
          umma_unit_elect_ts_pair(tmem + 256 + half * 128, t_qp + 64 * kg, t_qp + 64 * kg + 32, w_ring_addr + ws * TP_WH, idesc, kg != 0,
                                  &w_empty[ws], kg == 1 ? &s_full[half] : nullptr);
        }
        AXVS_TRACE(trc_, 53)
      }
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
