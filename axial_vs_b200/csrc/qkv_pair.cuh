// CTA-pair (cta_group::2) version of qkv_direct_kernel (same math, layouts and epilogue; reference WC/temporal_attention.py:42-44 with the
// layer's `src + pos` :200 and axis permutes :197,206 folded into the loads).  The leader CTA issues ONE stream of M = 256 instructions for
// two adjacent 128-row tiles; each CTA stages half of every weight unit, which halves the shared-memory traffic of the GEMM core -- in the
// single-CTA kernel that traffic (B reads + TMA refills = 128 B/clk) saturates the shared-memory port, and the converting producers' and
// the epilogue's shared-memory accesses queue behind it (epilogue busy 85 % of the tile, issuer waiting 43 % for drained stages).
// Protocol as in traj_pair.cuh: local full barriers forwarded to the leader by the non-leader's warp 17, multicast commits, remote
// arrives for "stage drained".
#pragma once
#include "qkv_direct.cuh"
#include "traj_pair.cuh"

namespace axvs {

constexpr int QP_W_SLOTS = 5;                   // half units
constexpr int QP_WH = 16384;
constexpr int QP_STAGE_BYTES = 8 * 4096;        // per epilogue warp: two 32 rows x 64 B staging buffers (bulk stores read one while the next is written)
// qkv_pair_kernel's own ring depths (the MSDeformAttn kernels that share this header keep QD_A_SLOTS / QP_W_SLOTS): FOUR A pair-slots hold a whole
// tile (4 K-blocks), so the producers finish the next tile while the current one is computed and drained -- with three, the last K-block
// could only be loaded after the issuer had started the tile, and the issuer waited for it (17 % of its time) -- paid for with a shorter weight ring
#ifndef QQ_A_SLOTS
#define QQ_A_SLOTS 4
#endif
#ifndef QQ_W_SLOTS
#define QQ_W_SLOTS 3
#endif
constexpr int QQ_SMEM_BYTES = QQ_A_SLOTS * 2 * TF_KB + QQ_W_SLOTS * QP_WH + QP_STAGE_BYTES + QK_BIAS_BYTES + 512;
static_assert(QQ_SMEM_BYTES <= 232448, "qkv_pair_kernel exceeds the 227 KiB shared-memory limit");
constexpr int QP_SMEM_BYTES = QD_A_SLOTS * 2 * TF_KB + QP_W_SLOTS * QP_WH + QP_STAGE_BYTES + QK_BIAS_BYTES + 512;
static_assert(QP_SMEM_BYTES <= 232448, "qkv_pair_kernel exceeds the 227 KiB shared-memory limit");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QD_THREADS, 1) qkv_pair_kernel(const QkvDirectParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QQ_A_SLOTS * 2 * TF_KB;
  uint8_t* stage_all = w_ring + QQ_W_SLOTS * QP_WH;
  float* sbias = reinterpret_cast<float*>(stage_all + QP_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QQ_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QQ_A_SLOTS;      // tcgen05.commit after the slot's copies
  uint64_t* w_full = a_empty + QQ_A_SLOTS;      // [QQ_W_SLOTS]
  uint64_t* w_empty = w_full + QQ_W_SLOTS;
  uint64_t* s_full = w_empty + QQ_W_SLOTS;      // [2] accumulator stage of group g complete
  uint64_t* s_empty = s_full + 2;               // [2] drained by the 4 warps of group g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pair_tiles = (p.tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    const uint32_t extra = rank == 0 ? 1 : 0;                 // leader: + the peer's relay
    for (int i = 0; i < QQ_A_SLOTS; ++i) { mbar_init(&a_full[i], QD_PRODUCER_WARPS + extra); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QQ_W_SLOTS; ++i) { mbar_init(&w_full[i], 1 + extra); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 16); }   // s_empty: the leader's copy is live (8 warps x 2 CTAs)
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 768; i += QD_THREADS) sbias[i] = p.bias[i];
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' barriers are initialised before any remote arrive
  if (warp == 17) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue: group g drains heads 2g, 2g+1 of every chunk
    const int g = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* stg = stage_all + warp * 4096;
    uint32_t cnt = 0;                                          // chunks consumed (stage = cnt & 1)
    uint32_t rnd = 0;                                          // head rounds done (staging buffer = rnd & 1)
    AXVS_PROF_DECL(6)
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      const int tile = 2 * pt + (int)rank;                     // may be == p.tiles (odd tile count): every row masked
      const int row0 = tile * 128 + (warp & 3) * 32;           // first row of this warp
      if (p.swz_N > 0) {
        // Unit-major destination: lane = row.  The lane's 64-byte row goes to the staging buffer already permuted the way the attention
        // kernel wants it in global memory ((pos >> 1) & 3), so a run of rows of one frame is one contiguous piece of the (sequence, head)
        // region: the first lane of every run writes it with ONE bulk copy (shared -> global, async proxy).  The previous version read the
        // staging tile back and issued 16-byte global stores: 8 more LSU instructions per round on a load/store pipe that the converting
        // producers keep full (wait profile: 750 clk per round in that phase).
        const int r = row0 + lane;
        const int seq = r / p.swz_N, ii = r - seq * p.swz_N;
        const int f = ii / p.swz_n, j = ii - f * p.swz_n;
        const uint32_t seq_row = (uint32_t)seq * 24u * p.swz_N;                       // first region row of the sequence (head 0)
        const uint32_t pos_q = (uint32_t)ii, pos_k = (uint32_t)(p.swz_N + 2 * f * p.swz_n + j);
        int run = min(min(p.swz_n - j, 32 - lane), p.rows - r);                       // rows of my run if I am its first lane
        if (lane != 0 && j != 0) run = 0;
#pragma unroll 1
        for (int rt = 0; rt < 6; ++rt, ++cnt) {
          const int st = cnt & 1;                              // 6 chunks per tile: even, so st == rt & 1
          AXVS_PROF_WAIT(0, mbar_wait_cluster(&s_full[st], (cnt >> 1) & 1))
          tc_fence_after();
          const uint32_t t_s = tmem + lane_base + 256 + st * 128;
          const int which = rt >> 1;
          const uint32_t pos = which == 0 ? pos_q : pos_k + (which == 2 ? (uint32_t)p.swz_n : 0u);
          const uint32_t key = (pos >> 1) & 3u;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc, ++rnd) {              // one head (32 columns) at a time
            const int c = 2 * g + cc;
            float v[32];
            float4 bb[8];                                      // the round's 32 biases, requested before the accumulator load so that the
            {                                                  // shared-memory latency (a busy load/store pipe) overlaps it
              const float4* b4 = reinterpret_cast<const float4*>(sbias + rt * 128 + c * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) bb[q] = b4[q];
            }
            AXVS_PROF_MARK(t_ld_)
            tmem_ld32(t_s + 32 * c, v);
            tmem_ld_wait();
            AXVS_PROF_SPAN(1, t_ld_)
            if (cc == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[st], 0);
            }
            AXVS_PROF_MARK(t_cv_)
            uint8_t* sb = stg + (rnd & 1) * 2048;
            bulk_wait_group_read<1>();                         // the copies of round rnd - 2 have read this buffer
            __syncwarp();
            AXVS_PROF_SPAN(4, t_cv_)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = bb[2 * q], b1 = bb[2 * q + 1];
              uint4 u;
              const float2 t0 = add_f32x2(make_float2(v[8 * q], v[8 * q + 1]), make_float2(b0.x, b0.y));
              const float2 t1 = add_f32x2(make_float2(v[8 * q + 2], v[8 * q + 3]), make_float2(b0.z, b0.w));
              const float2 t2 = add_f32x2(make_float2(v[8 * q + 4], v[8 * q + 5]), make_float2(b1.x, b1.y));
              const float2 t3 = add_f32x2(make_float2(v[8 * q + 6], v[8 * q + 7]), make_float2(b1.z, b1.w));
              u.x = pack_bf16x2(t0.x, t0.y);
              u.y = pack_bf16x2(t1.x, t1.y);
              u.z = pack_bf16x2(t2.x, t2.y);
              u.w = pack_bf16x2(t3.x, t3.y);
              *reinterpret_cast<uint4*>(sb + lane * 64 + (((uint32_t)q ^ key) << 4)) = u;
            }
            AXVS_PROF_MARK(t_fn_)
            fence_proxy_async_smem();
            __syncwarp();
            AXVS_PROF_SPAN(5, t_fn_)
            AXVS_PROF_SPAN(2, t_cv_)
            AXVS_PROF_MARK(t_st_)
            if (run > 0) {
              const int head = (rt & 1) * 4 + c;
              tma_bulk_s2g(reinterpret_cast<uint8_t*>(p.qkv) + (size_t)(seq_row + (uint32_t)head * 3u * p.swz_N + pos) * 64, sb + lane * 64,
                           (uint32_t)run * 64u);
            }
            bulk_commit_group();
            AXVS_PROF_SPAN(3, t_st_)
          }
        }
        continue;
      }
      // head-major destination (mma.sync attention baseline): 2 KiB per-warp transpose so every store instruction writes 512 contiguous bytes
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++cnt) {
        const int st = cnt & 1;
        mbar_wait_cluster(&s_full[st], (cnt >> 1) & 1);
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + st * 128;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * g + cc;
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          if (cc == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[st], 0);
          }
          {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + rt * 128 + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = b4[2 * q], b1 = b4[2 * q + 1];
              uint4 u;
              const float2 t0 = add_f32x2(make_float2(v[8 * q], v[8 * q + 1]), make_float2(b0.x, b0.y));
              const float2 t1 = add_f32x2(make_float2(v[8 * q + 2], v[8 * q + 3]), make_float2(b0.z, b0.w));
              const float2 t2 = add_f32x2(make_float2(v[8 * q + 4], v[8 * q + 5]), make_float2(b1.x, b1.y));
              const float2 t3 = add_f32x2(make_float2(v[8 * q + 6], v[8 * q + 7]), make_float2(b1.z, b1.w));
              u.x = pack_bf16x2(t0.x, t0.y);
              u.y = pack_bf16x2(t1.x, t1.y);
              u.z = pack_bf16x2(t2.x, t2.y);
              u.w = pack_bf16x2(t3.x, t3.y);
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = u;
            }
          }
          __syncwarp();
          {
            const int which = rt >> 1, head = (rt & 1) * 4 + c;
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.qkv + ((size_t)(which * 8 + head) * p.rows + row0) * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + (lane >> 2), piece = lane & 3;
              const uint4 u = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
              if (row0 + rl < p.rows) *reinterpret_cast<uint4*>(dst + rl * 64 + piece * 16) = u;
            }
          }
          __syncwarp();
        }
      }
    }
    bulk_wait_group<0>();                                      // every bulk store of this thread has completed
    AXVS_PROF_FLUSH(24, 6, rank == 0 && warp == 0 && lane == 0)
  } else if (warp < 8 + QD_PRODUCER_WARPS) {
    // =============================================================== converting A producers
    // Work unit = one K-block of the lane's 8 rows: a BURST of 8 src + 8 pos loads, then convert + store all of them.  (The first version
    // software-pipelined two half-size register sets; the consumer of set b then waited on a scoreboard shared with the loads of set b+1
    // issued just before it, so only one set was ever in flight per warp: tools/microbench/ldg_rows.cu 2.8 TB/s against 4.4 TB/s for
    // plain bursts of the same register footprint, tools/microbench/ldg_burst.cu.)  The other seven producer warps cover the gap
    // between a warp's bursts.
    const int pw = warp - 8;
    const int half = lane >> 4, c16 = lane & 15;               // row of the pair, 16-byte piece (4 channels) of the 256-byte segment
    uint32_t cnt = 0;
    float4 sv[8], qv[8];
    uint32_t crow[8];                                          // canonical token of this lane's 8 rows (0xFFFFFFFF = past the end)
    AXVS_PROF_DECL(2)
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      const int tile = 2 * pt + (int)rank;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pr = tile * 128 + pw * 16 + 2 * j + half;
        crow[j] = pr < p.rows ? (uint32_t)pass_to_canonical(pr, p.map_mode, p.dims) : 0xFFFFFFFFu;
      }
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++cnt) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t c = crow[j];
          sv[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(p.src + (size_t)c * C256 + kb * 64) + c16) : z;
        }
        if (p.pos) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t c = crow[j];
            qv[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)pos_row(c, p.dims) * C256 + kb * 64) + c16) : z;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) qv[j] = z;
        }
        const uint32_t slot = cnt % QQ_A_SLOTS, phase = (cnt / QQ_A_SLOTS) & 1;
        AXVS_PROF_WAIT(0, mbar_wait_cluster(&a_empty[slot], phase ^ 1))
        uint8_t* dst = a_ring + slot * 2 * TF_KB;
        AXVS_PROF_MARK(t_cv_)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = pw * 16 + 2 * j + half;
          const uint32_t off = sw128_offset(r, c16 >> 1) + (c16 & 1) * 8;
          uint2 u;
          u.x = pack_bf16x2(sv[j].x + qv[j].x, sv[j].y + qv[j].y);
          u.y = pack_bf16x2(sv[j].z + qv[j].z, sv[j].w + qv[j].w);
          *reinterpret_cast<uint2*>(dst + off) = u;                                   // A1 K-block image
          u.x = pack_bf16x2(sv[j].x, sv[j].y);
          u.y = pack_bf16x2(sv[j].z, sv[j].w);
          *reinterpret_cast<uint2*>(dst + TF_KB + off) = u;                           // A2 K-block image
        }
        fence_proxy_async_smem();
        __syncwarp();
        AXVS_PROF_SPAN(1, t_cv_)
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
    }
    AXVS_PROF_FLUSH(48, 2, rank == 0 && pw == 0 && lane == 0)
  } else if (warp == 16 && lane == 0) {
    // =============================================================== weight producer: my half (64 rows) of each of the 12 units per tile
    uint32_t slot = 0, phase = 0;
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
      for (int u = 0; u < 12; ++u) {
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], QP_WH);
        const uint8_t* src = p.w + (size_t)u * TF_WU + rank * 64 * 128;
        tma_bulk_g2s(w_ring + slot * QP_WH, src, 8192, &w_full[slot]);
        tma_bulk_g2s(w_ring + slot * QP_WH + 8192, src + TF_KB, 8192, &w_full[slot]);
        if (++slot == QQ_W_SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 17 && rank != 0) {
    // =============================================================== relay (non-leader): forward my full barriers to the leader in
    // the order the leader consumes them
    if (lane == 0) {
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0;
      for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
          const uint32_t slot = a_cnt % QQ_A_SLOTS;
          mbar_wait_cluster(&a_full[slot], (a_cnt / QQ_A_SLOTS) & 1);
          mbar_arrive_cluster(&a_full[slot], 0);               // release: my producers' generic-proxy writes were fenced before their arrive
        }
#pragma unroll 1
        for (int u = 0; u < 12; ++u) {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          mbar_arrive_cluster_relaxed(&w_full[w_slot], 0);
          if (++w_slot == QQ_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        }
      }
    }
  } else if (warp == 17) {
    // =============================================================== tcgen05.cp + MMA issuer (leader CTA; converged warp, elected lane)
    const uint32_t idesc = umma_idesc_bf16(256, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;     // ccnt: chunks issued
    AXVS_PROF_DECL(3)
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      // both CTAs' A operands: four pair slots each -> TMEM.  These copies are ordered by the tensor pipe behind every UMMA of the
      // previous tile (issued earlier by this thread), which still reads the old contents.
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
        const uint32_t slot = a_cnt % QQ_A_SLOTS;
        AXVS_PROF_WAIT(2, mbar_wait_cluster(&a_full[slot], (a_cnt / QQ_A_SLOTS) & 1))
        tc_fence_after();
        const uint32_t sa = a_ring_addr + slot * 2 * TF_KB;
        if (elect_one()) {
          tmem_cp_kblock_pair(tmem + 32 * kb, sa);
          tmem_cp_kblock_pair(tmem + 128 + 32 * kb, sa + TF_KB);
          umma_commit_pair(&a_empty[slot]);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++ccnt) {
        const int g = rt & 1;
        const uint32_t gc = ccnt >> 1;                         // chunks already issued to group g (6 per tile: even count)
        AXVS_PROF_WAIT(1, mbar_wait_cluster(&s_empty[g], (gc & 1) ^ 1))
        tc_fence_after();
        const uint32_t t_a = tmem + (rt < 4 ? 0 : 128);        // A1 for the q / k chunks, A2 for the v chunks
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          AXVS_PROF_WAIT(0, mbar_wait_cluster(&w_full[w_slot], w_phase))
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QQ_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          umma_unit_elect_ts_pair(tmem + 256 + g * 128, t_a + 64 * kg, t_a + 64 * kg + 32, w_ring_addr + ws * QP_WH, idesc, kg != 0,
                                  &w_empty[ws], kg == 1 ? &s_full[g] : nullptr);
        }
      }
    }
    AXVS_PROF_FLUSH(40, 3, lane == 0)
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc_pair(tmem, 512);
  }
}

}  // namespace axvs
