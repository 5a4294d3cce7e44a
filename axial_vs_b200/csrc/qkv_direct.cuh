// q | k | v projections reading the fp32 residual stream directly (fusion level 4): the tile-image pack kernel and its
// 2 KiB/token round trip through HBM disappear.  Same GEMM schedule, TMEM staging and head-major epilogue as
// qkv_fused_kernel; the single TMA A-producer warp is replaced by eight converting producer warps:
//
//   item 0..3 : A1 K-block kb = bf16(q_in[c(row)][64 kb ..] + pos[c(row)][64 kb ..])      (pos optional)
//   item 4..7 : A2 K-block kb = bf16(v_in[c(row)][64 kb ..])                              (second read of the rows: L2 hit)
//
// c(row) = canonical token of pass-order row `row` (the reference's axis permutes, WC/temporal_attention.py:197,206, as an
// index map).  A producer warp owns 16 rows of the item; half a warp reads one 256-byte row segment per instruction,
// converts and writes 8-byte pieces of the SWIZZLE_128B K-major image the MMA consumes, then fence.proxy.async + one mbarrier
// arrive per warp.  Loads are software-pipelined over two register sets (up to 16 x LDG.128 in flight per lane).
//
// Warp roles (576 threads, 112 registers each, no setmaxnreg): warps 0-7 epilogue, 8-15 A producers, 16 weight TMA, 17 MMA.
#pragma once
#include "qkv_fused.cuh"

namespace axvs {

constexpr int QD_THREADS = 576;
constexpr int QD_PRODUCER_WARPS = 8;

struct QkvDirectParams {
  const float* q_in;       // fp32 [tokens, 256] canonical order
  const float* v_in;       // fp32 [tokens, 256] (may equal q_in)
  const float* pos;        // fp32 [tokens, 256] or null
  const uint8_t* w;        // unit format of [Wq; Wk; Wv]
  const float* bias;       // [768]
  __nv_bfloat16* qkv;      // head-major [3][8][rows][32]
  int rows, tiles, map_mode;
  AxialDims dims;
};

__global__ void __launch_bounds__(QD_THREADS, 1) qkv_direct_kernel(const QkvDirectParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QK_A_SLOTS * TF_KB;
  uint8_t* stage_all = w_ring + QK_W_SLOTS * TF_WU;
  float* sbias = reinterpret_cast<float*>(stage_all + QK_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QK_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QK_A_SLOTS;
  uint64_t* w_full = a_empty + QK_A_SLOTS;      // [QK_W_SLOTS]
  uint64_t* w_empty = w_full + QK_W_SLOTS;
  uint64_t* s_full = w_empty + QK_W_SLOTS;      // [4]
  uint64_t* s_empty = s_full + 4;               // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QK_A_SLOTS; ++i) { mbar_init(&a_full[i], QD_PRODUCER_WARPS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QK_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 768; i += QD_THREADS) sbias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue: group g drains chunks rt with rt & 1 == g
    const int g = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* stg = stage_all + warp * 2048;
    uint32_t cnt = 0;                                          // chunks consumed by this group (stage = 2 * (cnt & 1) + g)
    AXVS_PROF_DECL(1)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int row0 = tile * 128 + (warp & 3) * 32;           // first row of this warp
#pragma unroll 1
      for (int rt = g; rt < 6; rt += 2, ++cnt) {
        const int stage = 2 * (cnt & 1) + g;
        AXVS_PROF_WAIT(0, mbar_wait(&s_full[stage], (cnt >> 1) & 1))
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + stage * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                          // one head (32 columns) at a time
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
          }
          {
            const float4* b4 = reinterpret_cast<const float4*>(sbias + rt * 128 + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = b4[2 * q], b1 = b4[2 * q + 1];
              uint4 u;
              u.x = pack_bf16x2(v[8 * q] + b0.x, v[8 * q + 1] + b0.y);
              u.y = pack_bf16x2(v[8 * q + 2] + b0.z, v[8 * q + 3] + b0.w);
              u.z = pack_bf16x2(v[8 * q + 4] + b1.x, v[8 * q + 5] + b1.y);
              u.w = pack_bf16x2(v[8 * q + 6] + b1.z, v[8 * q + 7] + b1.w);
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
    AXVS_PROF_FLUSH(44 + 2 * g, 1, (warp & 3) == 0 && lane == 0)
  } else if (warp < 8 + QD_PRODUCER_WARPS) {
    // =============================================================== converting A producers
    // Work unit = batch: 4 of the lane's 8 rows of one item (4 src + 4 pos loads).  Two register sets alternate so the loads
    // of batch b+1 (also across item and tile boundaries) are in flight while batch b is converted and stored.
    const int pw = warp - 8;
    const int half = lane >> 4, c16 = lane & 15;               // row of the pair, 16-byte piece (4 channels) of the 256-byte segment
    uint32_t cnt = 0;
    AXVS_PROF_DECL(2)
    float4 sv[2][4], qv[2][4];
    uint32_t crow[8], crow_n[8];                               // canonical token of this lane's 8 rows (0xFFFFFFFF = past the end)
    auto rows_of_tile = [&](int tile, uint32_t (&cr)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pr = tile * 128 + pw * 16 + 2 * j + half;
        cr[j] = pr < p.rows ? (uint32_t)pass_to_canonical(pr, p.map_mode, p.dims) : 0xFFFFFFFFu;
      }
    };
    auto issue = [&](float4 (&s)[4], float4 (&q)[4], const uint32_t (&cr)[8], int item, int hf) {
      const int kb = item & 3;
      const float* src = item < 4 ? p.q_in : p.v_in;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = cr[4 * hf + j];
        s[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(src + (size_t)c * C256 + kb * 64) + c16) : z;
      }
      if (item < 4 && p.pos) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c = cr[4 * hf + j];
          q[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)c * C256 + kb * 64) + c16) : z;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = z;
      }
    };
    auto consume = [&](float4 (&s)[4], float4 (&q)[4], int hf) {
      const uint32_t slot = cnt % QK_A_SLOTS, phase = (cnt / QK_A_SLOTS) & 1;
      if (hf == 0) { AXVS_PROF_WAIT(0, mbar_wait(&a_empty[slot], phase ^ 1)) }
      uint8_t* dst = a_ring + slot * TF_KB;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = pw * 16 + 2 * (4 * hf + j) + half;
        uint2 u;
        u.x = pack_bf16x2(s[j].x + q[j].x, s[j].y + q[j].y);
        u.y = pack_bf16x2(s[j].z + q[j].z, s[j].w + q[j].w);
        *reinterpret_cast<uint2*>(dst + sw128_offset(r, c16 >> 1) + (c16 & 1) * 8) = u;
      }
      if (hf == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
        ++cnt;
      }
    };
    if ((int)blockIdx.x < p.tiles) {
      rows_of_tile(blockIdx.x, crow);
      issue(sv[0], qv[0], crow, 0, 0);
    }
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const bool has_next = tile + (int)gridDim.x < p.tiles;
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        if (b < 15) {
          issue(sv[(b + 1) & 1], qv[(b + 1) & 1], crow, (b + 1) >> 1, (b + 1) & 1);
        } else if (has_next) {
          rows_of_tile(tile + gridDim.x, crow_n);
          issue(sv[0], qv[0], crow_n, 0, 0);
        }
        consume(sv[b & 1], qv[b & 1], b & 1);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) crow[j] = crow_n[j];
    }
    AXVS_PROF_FLUSH(48, 2, pw == 0 && lane == 0)
  } else if (warp == 16 && lane == 0) {
    // =============================================================== weight producer (12 units per tile)
    uint32_t slot = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int u = 0; u < 12; ++u) {
        mbar_wait(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], TF_WU);
        tma_bulk_g2s(w_ring + slot * TF_WU, p.w + (size_t)u * TF_WU, TF_WU, &w_full[slot]);
        if (++slot == QK_W_SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 17) {
    // =============================================================== MMA issuer
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;     // ccnt: chunks issued (stage = 2 * ((ccnt >> 1) & 1) + (ccnt & 1))
    AXVS_PROF_DECL(3)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++ccnt) {
        const int g = rt & 1;
        const uint32_t gc = ccnt >> 1;
        const int stage = 2 * (gc & 1) + g;
        AXVS_PROF_WAIT(1, mbar_wait(&s_empty[stage], ((gc >> 1) & 1) ^ 1))
        tc_fence_after();
        const uint32_t abase = a_cnt + (rt < 4 ? 0 : 4);       // A1 items for q/k chunks, A2 items for v chunks
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          const uint32_t ac0 = abase + 2 * kg, ac1 = ac0 + 1;
          const uint32_t s0 = ac0 % QK_A_SLOTS, s1 = ac1 % QK_A_SLOTS;
          if (rt == 0 || rt == 4) {
            AXVS_PROF_WAIT(2, mbar_wait(&a_full[s0], (ac0 / QK_A_SLOTS) & 1); mbar_wait(&a_full[s1], (ac1 / QK_A_SLOTS) & 1))
            tc_fence_after();
          }
          AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QK_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          const bool last_use = (rt == 3 || rt == 5);
          umma_unit_elect(tmem + stage * 128, a_ring_addr + s0 * TF_KB, a_ring_addr + s1 * TF_KB, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                          &w_empty[ws], last_use ? &a_empty[s0] : nullptr, last_use ? &a_empty[s1] : nullptr, kg == 1 ? &s_full[stage] : nullptr);
        }
      }
      a_cnt += 8;
    }
    AXVS_PROF_FLUSH(40, 3, lane == 0)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
