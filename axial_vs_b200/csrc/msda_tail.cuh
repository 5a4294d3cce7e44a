// Tail of the MSDeformAttn attention block as ONE CTA-pair kernel (WC/ops/modules/ms_deform_attn.py:124 with
// the layer's residual + norm1, WC/msdeformattn.py:208-209):
//
//   y = src + sampled W_out^T + b_out,   s = LayerNorm1(y)   ->  fp32 rows s (the FFN's residual) + the bf16 tile image the FFN kernel loads
//
// It replaces a generic-GEMM launch (bf16 A, fp32 residual epilogue) and ln_image_kernel, i.e. one fp32 round trip of y through HBM.
// Structure of msda_front_pair_kernel (msda_front.cuh): the A producers copy the sampler's bf16 rows into SWIZZLE_128B K-block images, the
// issuer moves them into tensor memory (tcgen05.cp.cta_group::2) and issues M = 256 / N = 128 UMMAs for the two 128-column chunks of
// W_out (four half-staged weight units per tile).  LayerNorm needs whole rows, and a row's 256 columns are drained by TWO warps (column
// groups 2g, 2g+1 of each chunk): pass 1 adds bias and residual in the transposed (row-contiguous) layout, writes the un-normalised y to
// the output rows and accumulates (sum, sum of squares) per row; the two warps of a row quarter exchange their partial statistics through
// shared memory; pass 2 re-reads y (L2-resident: written a few microseconds earlier by the same lanes), normalises and writes s and the
// image.  The accumulator stages are released after pass 1's tcgen05.ld, so the next tile's UMMAs overlap the LayerNorm.
#pragma once
#include "msda_front.cuh"

namespace axvs {

struct MsdaTailParams {
  const __nv_bfloat16* samp;   // bf16 [rows, 256] (msda_sample output)
  const float* resid;          // fp32 [rows, 256] (src)
  const uint8_t* w;            // unit image (pack_weight_units, k_major 0) of output_proj.weight [256, 256]: 4 units
  const float* bias;           // [256]
  const float* ln_g;           // [256]
  const float* ln_b;           // [256]
  float* out;                  // fp32 [rows, 256] = LayerNorm1(y)
  uint8_t* img;                // bf16 tile image of the same (per 128-row tile: 4 K-block images of 16 KiB, SWIZZLE_128B)
  int rows, tiles;
  float eps;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QD_THREADS, 1) msda_tail_pair_kernel(const MsdaTailParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;                       // pair slots of 2 x 16 KiB; only the first image of a slot carries an operand here ...
  float2* xc_all = reinterpret_cast<float2*>(a_ring + TF_KB);   // ... the second image of slot 0 holds the statistics exchange: [2 parities][8 warps][32 rows]
  uint8_t* w_ring = a_ring + QD_A_SLOTS * 2 * TF_KB;
  uint8_t* stage_all = w_ring + QP_W_SLOTS * QP_WH;
  float* sbias = reinterpret_cast<float*>(stage_all + QK_STAGE_BYTES);        // [256] bias | [256] gamma | [256] beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QD_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QD_A_SLOTS;      // tcgen05.commit after the slot's copies
  uint64_t* w_full = a_empty + QD_A_SLOTS;      // [QP_W_SLOTS]
  uint64_t* w_empty = w_full + QP_W_SLOTS;
  uint64_t* s_full = w_empty + QP_W_SLOTS;      // [2] accumulator stage (= chunk) complete
  uint64_t* s_empty = s_full + 2;               // [2] drained by the 8 epilogue warps of both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int pair_tiles = (p.tiles + 1) >> 1;

  if (threadIdx.x == 0) {
    const uint32_t extra = rank == 0 ? 1 : 0;                 // leader: + the peer's relay
    for (int i = 0; i < QD_A_SLOTS; ++i) { mbar_init(&a_full[i], QD_PRODUCER_WARPS + extra); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QP_W_SLOTS; ++i) { mbar_init(&w_full[i], 1 + extra); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 16); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += QD_THREADS) { sbias[i] = p.bias[i]; sbias[256 + i] = p.ln_g[i]; sbias[512 + i] = p.ln_b[i]; }
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' barriers are initialised before any remote arrive
  if (warp == 17) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue: warp (g, quarter) owns rows quarter*32.. and the column
    // groups 2g, 2g+1 of both chunks, i.e. columns [64 g, 64 g + 64) and [128 + 64 g, 128 + 64 g + 64)
    const int g = warp >> 2, quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint8_t* stg = stage_all + warp * 2048;
    const int sub = lane >> 2, piece = lane & 3;               // after the transpose: rows 8 i + sub, 16-byte piece (4 floats) of a 64-byte half group
    uint32_t it = 0;
    for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
      const int tile = 2 * pt + (int)rank;                     // may be == p.tiles (odd tile count): every row masked
      const int row0 = tile * 128 + quarter * 32;              // first row of this warp
      // residual block of this warp (32 rows x 2 segments of 256 B) -> L2 while the accumulators complete
      if (row0 + lane < p.rows) {
        const float* rp = p.resid + (size_t)(row0 + lane) * C256 + 64 * g;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 32));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 128));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + 160));
      }
      float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
      // ---- pass 1: y = acc + bias + resid -> output rows (un-normalised), per-row partial statistics
#pragma unroll 1
      for (int rt = 0; rt < 2; ++rt) {
        mbar_wait_cluster(&s_full[rt], it & 1);                // chunk rt always lands in stage rt (two chunks per tile)
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + rt * 128;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * g + cc;
          const int col = rt * 128 + c * 32;
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          if (cc == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[rt], 0);
          }
          const float4* b4 = reinterpret_cast<const float4*>(sbias + col);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 rr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {                      // residual pieces first: in flight during the transpose
              const int row = row0 + 8 * i + sub;
              rr[i] = row < p.rows ? __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)row * C256 + col + 16 * h) + piece) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = b4[4 * h + q];
              const float2 t0 = add_f32x2(make_float2(v[16 * h + 4 * q], v[16 * h + 4 * q + 1]), make_float2(b.x, b.y));
              const float2 t1 = add_f32x2(make_float2(v[16 * h + 4 * q + 2], v[16 * h + 4 * q + 3]), make_float2(b.z, b.w));
              *reinterpret_cast<float4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = make_float4(t0.x, t0.y, t1.x, t1.y);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + sub;
              float4 y = *reinterpret_cast<const float4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
              y.x += rr[i].x; y.y += rr[i].y; y.z += rr[i].z; y.w += rr[i].w;
              ps[i] += (y.x + y.y) + (y.z + y.w);
              pq[i] += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
              if (row0 + rl < p.rows) *(reinterpret_cast<float4*>(p.out + (size_t)(row0 + rl) * C256 + col + 16 * h) + piece) = y;
            }
            __syncwarp();
          }
        }
      }
      // ---- statistics: the four lanes of a row, then the partner warp (same rows, other 128 columns)
      float2* xc_mine = xc_all + ((it & 1) * 8 + warp) * 32;
      const float2* xc_other = xc_all + ((it & 1) * 8 + (warp ^ 4)) * 32;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], 1); pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], 1);
        ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], 2); pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], 2);
        if (piece == 0) xc_mine[8 * i + sub] = make_float2(ps[i], pq[i]);
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      float mean[4], rstd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 o = xc_other[8 * i + sub];
        mean[i] = (ps[i] + o.x) * (1.f / 256.f);
        const float var = fmaxf((pq[i] + o.y) * (1.f / 256.f) - mean[i] * mean[i], 0.f);
        rstd[i] = rsqrtf(var + p.eps);
      }
      // ---- pass 2: s = (y - mean) rstd gamma + beta -> fp32 rows and the bf16 tile image
#pragma unroll 1
      for (int gi = 0; gi < 8; ++gi) {                         // (rt, cc, h)
        const int col = (gi >> 2) * 128 + (2 * g + ((gi >> 1) & 1)) * 32 + 16 * (gi & 1) + piece * 4;
        const float4 gg = *reinterpret_cast<const float4*>(sbias + 256 + col), be = *reinterpret_cast<const float4*>(sbias + 512 + col);
        float4 y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = row0 + 8 * i + sub;
          y[i] = row < p.rows ? *reinterpret_cast<const float4*>(p.out + (size_t)row * C256 + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = row0 + 8 * i + sub;
          const float4 s = make_float4((y[i].x - mean[i]) * rstd[i] * gg.x + be.x, (y[i].y - mean[i]) * rstd[i] * gg.y + be.y,
                                       (y[i].z - mean[i]) * rstd[i] * gg.z + be.z, (y[i].w - mean[i]) * rstd[i] * gg.w + be.w);
          // 16-byte image chunk = 8 channels = this lane (even piece) + its odd neighbour
          const uint32_t lo = pack_bf16x2(s.x, s.y), hi = pack_bf16x2(s.z, s.w);
          const uint32_t nlo = __shfl_down_sync(0xffffffffu, lo, 1), nhi = __shfl_down_sync(0xffffffffu, hi, 1);
          if (row < p.rows) {
            *reinterpret_cast<float4*>(p.out + (size_t)row * C256 + col) = s;
            if ((piece & 1) == 0)
              *reinterpret_cast<uint4*>(p.img + ((size_t)(row >> 7) * 4 + (col >> 6)) * TF_KB + sw128_offset(row & 127, (col & 63) >> 3)) =
                  make_uint4(lo, hi, nlo, nhi);
          }
        }
      }
    }
  } else if (warp < 8 + QD_PRODUCER_WARPS) {
    // =============================================================== A producers: the sampler's bf16 rows -> K-block images (no conversion).
    // Lane = (row of a pair, 8-byte piece of the row's 128-byte K-block segment); one K-block of the lane's 8 rows per burst.
    const int pw = warp - 8;
    const int half = lane >> 4, c16 = lane & 15;
    uint32_t cnt = 0;
    uint2 sv[8];
    uint32_t crow[8];
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      const int tile = 2 * pt + (int)rank;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pr = tile * 128 + pw * 16 + 2 * j + half;
        crow[j] = pr < p.rows ? (uint32_t)pr : 0xFFFFFFFFu;
      }
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++cnt) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t c = crow[j];
          sv[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const uint2*>(p.samp + (size_t)c * C256 + kb * 64) + c16) : make_uint2(0u, 0u);
        }
        const uint32_t slot = cnt % QD_A_SLOTS, phase = (cnt / QD_A_SLOTS) & 1;
        mbar_wait_cluster(&a_empty[slot], phase ^ 1);
        uint8_t* dst = a_ring + slot * 2 * TF_KB;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = pw * 16 + 2 * j + half;
          *reinterpret_cast<uint2*>(dst + sw128_offset(r, c16 >> 1) + (c16 & 1) * 8) = sv[j];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
    }
  } else if (warp == 16 && lane == 0) {
    // =============================================================== weight producer: my half (64 rows) of each of the 4 units per tile
    uint32_t slot = 0, phase = 0;
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        mbar_wait_cluster(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], QP_WH);
        const uint8_t* src = p.w + (size_t)u * TF_WU + rank * 64 * 128;
        tma_bulk_g2s(w_ring + slot * QP_WH, src, 8192, &w_full[slot]);
        tma_bulk_g2s(w_ring + slot * QP_WH + 8192, src + TF_KB, 8192, &w_full[slot]);
        if (++slot == QP_W_SLOTS) { slot = 0; phase ^= 1; }
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
          const uint32_t slot = a_cnt % QD_A_SLOTS;
          mbar_wait_cluster(&a_full[slot], (a_cnt / QD_A_SLOTS) & 1);
          mbar_arrive_cluster(&a_full[slot], 0);               // release: my producers' generic-proxy writes were fenced before their arrive
        }
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          mbar_arrive_cluster_relaxed(&w_full[w_slot], 0);
          if (++w_slot == QP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        }
      }
    }
  } else if (warp == 17) {
    // =============================================================== tcgen05.cp + MMA issuer (leader CTA; converged warp, elected lane)
    const uint32_t idesc = umma_idesc_bf16(256, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, it = 0;
    for (int pt = pair; pt < pair_tiles; pt += npairs, ++it) {
      // both CTAs' A operand -> TMEM columns [0, 128); ordered by the tensor pipe behind every UMMA of the previous tile
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
        const uint32_t slot = a_cnt % QD_A_SLOTS;
        mbar_wait_cluster(&a_full[slot], (a_cnt / QD_A_SLOTS) & 1);
        tc_fence_after();
        if (elect_one()) {
          tmem_cp_kblock_pair(tmem + 32 * kb, a_ring_addr + slot * 2 * TF_KB);
          umma_commit_pair(&a_empty[slot]);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int rt = 0; rt < 2; ++rt) {
        mbar_wait_cluster(&s_empty[rt], (it & 1) ^ 1);         // stage rt drained by the previous tile's pass 1
        tc_fence_after();
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          umma_unit_elect_ts_pair(tmem + 256 + rt * 128, tmem + 64 * kg, tmem + 64 * kg + 32, w_ring_addr + ws * QP_WH, idesc, kg != 0,
                                  &w_empty[ws], kg == 1 ? &s_full[rt] : nullptr);
        }
      }
    }
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
