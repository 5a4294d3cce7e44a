// CTA-pair (cta_group::2) version of ffn_n256_kernel (same math, schedule and epilogue; reference WC/temporal_attention.py:181-185,218):
// the leader CTA issues ONE stream of M = 256 instructions for two adjacent 128-row tiles, each CTA stages half of every weight unit
// (GEMM 1: 128 of the 256 hidden rows of a unit; GEMM 2: 64 of the 128 output rows of each K-block), so the shared-memory traffic of the
// GEMM core per CTA halves (see traj_pair.cuh).  Protocol as there: local full barriers forwarded by the non-leader's warp 10, multicast
// commits, remote arrives for "hidden chunk ready" / "accumulator drained".
#pragma once
#include "ffn_n256.cuh"
#include "traj_pair.cuh"

namespace axvs {

constexpr int FQ_W_SLOTS = 6;                   // half units
constexpr int FQ_WH = 16384;
constexpr int FQ_SMEM_BYTES = FF_A_SLOTS * TF_KB + FF_H_BYTES + FQ_W_SLOTS * FQ_WH + FF_XCHG_BYTES + FF_BIAS_BYTES + 512;
static_assert(FQ_SMEM_BYTES <= 232448, "ffn_n256_pair_kernel exceeds the 227 KiB shared-memory limit");

// One K-block (4 UMMAs, M = 256 over the CTA pair, both operands in shared memory) + up to two multicast commits.
__device__ __forceinline__ void umma_kblock_elect_pair(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool accumulate,
                                                       uint64_t* commit_a, uint64_t* commit_b) {
  const uint32_t a_lo = umma_desc_lo(a_addr), w_lo = umma_desc_lo(w_addr);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo_pair(tmem_d, a_lo + 2 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
    if (commit_a) umma_commit_pair(commit_a);
    if (commit_b) umma_commit_pair(commit_b);
  }
  __syncwarp();
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FF_THREADS, 1) ffn_n256_pair_kernel(const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need a 1024 B aligned base (no static smem in this kernel)
  uint8_t* a_ring = smem;
  uint8_t* h_buf = a_ring + FF_A_SLOTS * TF_KB;
  uint8_t* w_ring = h_buf + FF_H_BYTES;
  float2* xchg = reinterpret_cast<float2*>(w_ring + FQ_W_SLOTS * FQ_WH);   // [2 parity][2 group][128] (sum, sumsq)
  float* sb1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xchg) + FF_XCHG_BYTES);   // b1 [d_ffn]
  float* sb2 = sb1 + FF_MAX_DFFN;                                                           // b2 [256]
  float* sg2 = sb2 + 256;                                                                   // ln2 gamma
  float* sbe2 = sg2 + 256;                                                                  // ln2 beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbe2 + 256);
  uint64_t* a_full = bars;                    // [FF_A_SLOTS]
  uint64_t* a_empty = a_full + FF_A_SLOTS;    // [FF_A_SLOTS]
  uint64_t* w_full = a_empty + FF_A_SLOTS;    // [6]
  uint64_t* w_empty = w_full + FQ_W_SLOTS;
  uint64_t* s_full = w_empty + FQ_W_SLOTS;    // [2]
  uint64_t* h_ready = s_full + 2;             // [2] epilogue -> MMA: the bf16 hidden chunk is in place in TMEM stage j & 1
  uint64_t* acc_full = h_ready + 2;           // MMA -> epilogue
  uint64_t* acc_free = acc_full + 1;          // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NJ = p.d_ffn / 256;                 // hidden chunks of 256 columns
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pair_tiles = (p.tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    const uint32_t fullc = rank == 0 ? 2 : 1;              // leader: own producer + the peer's relay
    for (int i = 0; i < FF_A_SLOTS; ++i) { mbar_init(&a_full[i], fullc); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FQ_W_SLOTS; ++i) { mbar_init(&w_full[i], fullc); mbar_init(&w_empty[i], 1); }
    mbar_init(&s_full[0], 1);
    mbar_init(&h_ready[0], 16);                              // the leader's copy is live: 8 warps x 2 CTAs
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);                                 // (leader)
    fence_barrier_init();
  }
  // biases / LayerNorm2 affine -> shared memory (the ~10 KiB of L1 left beside 217 KiB of smem cannot keep them hot)
  for (int i = threadIdx.x; i < p.d_ffn; i += FF_THREADS) sb1[i] = p.b1[i];
  for (int i = threadIdx.x; i < 256; i += FF_THREADS) { sb2[i] = p.b2[i]; sg2[i] = p.ln2_g[i]; sbe2[i] = p.ln2_b[i]; }
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' barriers are initialised before any remote arrive
  if (warp == 10) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
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
    for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
      const int tile = 2 * pt + (int)rank;                     // may be == p.tiles (odd tile count): every row masked
      const bool trc_ = it == 3 && lane == 0 && wq == 0;
      const int tb_ = 100 + 100 * g;
      // ---- the residual rows (128 KiB per tile = ~6 k clk of this SM's HBM share, and about as long to ISSUE: the load/store unit takes
      // them at the memory system's pace) are requested in four 32-column batches, one after each hidden chunk's drain, so that both the
      // issue time and the latency hide behind the next chunk's UMMAs; they wait in registers until the final epilogue (the drains work
      // on 32 columns at a time to leave room).  The single-CTA kernel issues all of them after the last chunk, in front of the epilogue.
      const int row0 = tile * 128 + wq * 32;
      float4 t[4][8];
      auto load_resid = [&](int c) {
        const int col = 128 * g + 32 * c + piece * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {                              // no arithmetic on the loaded values here: it would wait for them
          const int r = row0 + i * 4 + sub;
          t[c][i] = (r < p.rows) ? __ldg(reinterpret_cast<const float4*>(p.s32 + (size_t)r * 256 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // ---- hidden chunks of 256 columns on ONE 256-column stage: group g drains its 128 columns (bias + ReLU -> bf16 pairs)
      // and writes them back in place over the first 64 columns of its region: the tensor-memory A operand of GEMM 2
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        const uint32_t hc = it * NJ + j;                         // global chunk counter
        AXVS_TRACE(trc_, tb_ + 4 * j)
        mbar_wait_cluster(&s_full[0], hc & 1);
        AXVS_TRACE(trc_, tb_ + 4 * j + 1)
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + 128 * g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                              // 32 columns at a time: 16 packed words back over the stage's head
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(sb1 + j * 256 + 128 * g + 32 * c);
          uint32_t hpk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {                           // packed fp32 adds, ReLU folded into the bf16x2 conversion
            const float4 bb = b4[i];
            const float2 a0 = add_f32x2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(bb.x, bb.y));
            const float2 a1 = add_f32x2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(bb.z, bb.w));
            hpk[2 * i] = pack_bf16x2_relu(a0.x, a0.y);
            hpk[2 * i + 1] = pack_bf16x2_relu(a1.x, a1.y);
          }
          tmem_st16u(t_s + 16 * c, hpk);                           // columns [16c, 16c + 16) <= the columns already read
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(&h_ready[0], 0);
        AXVS_TRACE(trc_, tb_ + 4 * j + 2)
        // NJ is a runtime value: spell the four batches out so that t[][] keeps compile-time indices (registers, not local memory)
        if (j == NJ - 4) load_resid(0);
        else if (j == NJ - 3) load_resid(1);
        else if (j == NJ - 2) load_resid(2);
        else if (j == NJ - 1) load_resid(3);
      }
      if (NJ < 4) {                                                // fewer chunks than batches: the rest now
        if (NJ < 2) load_resid(2);
        if (NJ < 3) load_resid(1);
        if (NJ < 4) load_resid(0);
      }
      // ---- final: t = acc2 + b2 + s, LayerNorm2, store.  Rows are one-per-thread in TMEM; a per-warp transpose through
      // shared memory (h_buf is idle: every GEMM 2 of this tile has retired) makes the global traffic row-segment
      // coalesced.  In the transposed domain lane (sub, piece) owns 4 columns of rows {4*i + sub}.
      // The residual (s + b2) is fetched BEFORE waiting for the accumulator so its latency hides behind the last GEMMs.
      AXVS_TRACE(trc_, tb_ + 40)
      mbar_wait_cluster(acc_full, it & 1);
      AXVS_TRACE(trc_, tb_ + 41)
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
          if (c == 3) {                                            // acc2 fully read: release it for the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(acc_free, 0);
            AXVS_TRACE(trc_, tb_ + 42)
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        __syncwarp();
        const float4 bb2 = *reinterpret_cast<const float4*>(sb2 + 128 * g + 32 * c + piece * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + sub;
          const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
          float4 tv = t[c][i];
          tv.x = (tv.x + bb2.x) + a.x; tv.y = (tv.y + bb2.y) + a.y; tv.z = (tv.z + bb2.z) + a.z; tv.w = (tv.w + bb2.w) + a.w;
          t[c][i] = tv;
          ps[i] += tv.x + tv.y + tv.z + tv.w;
          pq[i] += tv.x * tv.x + tv.y * tv.y + tv.z * tv.z + tv.w * tv.w;
        }
        __syncwarp();
      }
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
      AXVS_TRACE(trc_, tb_ + 43)
    }
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
          const uint32_t slot = cnt % FF_A_SLOTS, phase = (cnt / FF_A_SLOTS) & 1;
          mbar_wait_cluster(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, p.s_img + ((size_t)tile * 4 + kb) * TF_KB, TF_KB, &a_full[slot]);
        }
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer: my half of every unit (16 KiB)
      uint32_t slot = 0, phase = 0;
      auto push1 = [&](int unit) {                                 // W1, N = 256 unit [256 rows x 128 B]: my 128 rows are contiguous
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], FQ_WH);
        tma_bulk_g2s(w_ring + slot * FQ_WH, p.w1 + (size_t)unit * TF_WU + rank * FQ_WH, FQ_WH, &w_full[slot]);
        if (++slot == FQ_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      auto push2 = [&](int unit) {                                 // W2, unit [2 K-blocks][128 rows x 128 B]: my 64 rows of each K-block
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], FQ_WH);
        const uint8_t* src = p.w2 + (size_t)unit * TF_WU + rank * 64 * 128;
        tma_bulk_g2s(w_ring + slot * FQ_WH, src, 8192, &w_full[slot]);
        tma_bulk_g2s(w_ring + slot * FQ_WH + 8192, src + TF_KB, 8192, &w_full[slot]);
        if (++slot == FQ_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      // consumption order of the issuer: GEMM 1 (0); then per chunk GEMM 2 (j) [4 units] and GEMM 1 (j + 1) [4 units]
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) push1(kb);                  // N = 256 units: (chunk j, K-block kb) = 4 j + kb
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
#pragma unroll 1
          for (int u = 0; u < 4; ++u) push2(2 * (2 * j + (u >> 1)) + (u & 1));      // (K group 2j + kg, output half)
          if (j + 1 < NJ) {
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) push1(4 * (j + 1) + kb);
          }
        }
      }
    } else if (warp == 10 && rank != 0) {
      // =============================================================== relay (non-leader): forward my full barriers to the leader, in
      // the order the leader consumes them
      if (lane == 0) {
        uint32_t a_cnt = 0, w_slot = 0, w_phase = 0;
        auto fwd_w = [&]() {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          mbar_arrive_cluster_relaxed(&w_full[w_slot], 0);
          if (++w_slot == FQ_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        };
        for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
            const uint32_t sl = a_cnt % FF_A_SLOTS;
            mbar_wait_cluster(&a_full[sl], (a_cnt / FF_A_SLOTS) & 1);
            mbar_arrive_cluster_relaxed(&a_full[sl], 0);
            fwd_w();
          }
#pragma unroll 1
          for (int j = 0; j < NJ; ++j) {
            fwd_w(); fwd_w(); fwd_w(); fwd_w();
            if (j + 1 < NJ) { fwd_w(); fwd_w(); fwd_w(); fwd_w(); }
          }
        }
      }
    } else if (warp == 10) {
      // =============================================================== MMA issuer (leader CTA; converged warp, elected lane issues)
      const uint32_t idesc = umma_idesc_bf16(256, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, it = 0;
      auto w_wait = [&]() -> uint32_t {
        mbar_wait_cluster(&w_full[w_slot], w_phase);
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == FQ_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      const uint32_t idesc256 = umma_idesc_bf16(256, 256);
      // GEMM 1 of one chunk: four K-blocks, each one N = 256 unit (4 UMMAs at 128 clk: the full rate with A in shared memory)
      auto gemm1 = [&](int j) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t ac = a_cnt + kb, sl = ac % FF_A_SLOTS;
          if (j == 0) {
            mbar_wait_cluster(&a_full[sl], (ac / FF_A_SLOTS) & 1);
            tc_fence_after();
          }
          const uint32_t ws = w_wait();
          umma_kblock_elect_pair(tmem + 256, a_ring_addr + sl * TF_KB, w_ring_addr + ws * FQ_WH, idesc256, kb != 0,
                                 &w_empty[ws], j == NJ - 1 ? &a_empty[sl] : nullptr);
        }
        if (elect_one()) umma_commit_pair(&s_full[0]);
        __syncwarp();
      };
      for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
        const bool trc_ = it == 3 && lane == 0;
        AXVS_TRACE(trc_, 0)
        gemm1(0);
        AXVS_TRACE(trc_, 1)
#pragma unroll 1
        for (int j = 0; j < NJ; ++j) {
          // GEMM 2, K-chunk j: acc2 += h (128 x 256, bf16 pairs in TMEM) * W2[:, 256 j : 256 (j + 1)]^T
          const uint32_t hc = it * NJ + j;
          AXVS_TRACE(trc_, 10 + 4 * j)
          if (j == 0) mbar_wait_cluster(acc_free, (it & 1) ^ 1);   // previous tile's final epilogue has drained acc2
          mbar_wait_cluster(&h_ready[0], hc & 1);
          AXVS_TRACE(trc_, 11 + 4 * j)
          tc_fence_after();
#pragma unroll 1
          for (int u = 0; u < 4; ++u) {
            const int kg = u >> 1, half = u & 1;
            const uint32_t t_h = tmem + 256 + 128 * kg;                      // group kg's 128 hidden columns as 64 packed columns
            const uint32_t ws = w_wait();
            umma_unit_elect_ts_pair(tmem + half * 128, t_h, t_h + 32, w_ring_addr + ws * FQ_WH, idesc, (j | kg) != 0,
                                    &w_empty[ws], (u == 3 && j == NJ - 1) ? acc_full : nullptr);
          }
          // the stage is overwritten only after GEMM 2 (j) above: the tensor pipe executes in issue order
          AXVS_TRACE(trc_, 12 + 4 * j)
          if (j + 1 < NJ) gemm1(j + 1);
          AXVS_TRACE(trc_, 13 + 4 * j)
        }
        a_cnt += 4;
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
