// Front end of the MSDeformAttn spatial layer as ONE CTA-pair kernel (WC/ops/modules/ms_deform_attn.py:98-103
// with the layer's `query = src + pos`, WC/msdeformattn.py:207):
//
//   A1 = bf16(src + pos),  A2 = bf16(src)
//   [sampling_offsets | attention_weights] = A1 W_oa^T + b_oa     fp32 [rows, n_oa]               (n_oa = 8 L P 3 <= 384)
//   value = A2 W_value^T + b_value                                 bf16 HEAD-MAJOR [image][head][len][32]
//
// Same structure as qkv_pair_kernel (qkv_pair.cuh: converting producers that read the fp32 stream once, both images of a K-block into
// tensor memory with tcgen05.cp.cta_group::2, M = 256 / N = 128 UMMAs with A from tensor memory, each CTA staging half of every weight
// unit) with five 128-column chunks per tile -- three of W_oa (zero rows above n_oa) from A1, two of W_value from A2 -- and its own
// epilogue.  It replaces two launches of the generic GEMM: the value projection, and the 288-wide offsets / logits projection, which as
// a 512-column GEMM staged `src + pos` twice (two 256-column chunks, the second one nearly empty) and took 3 x the time of the former.
// Head-major value rows put the 64 bytes a sampling tap reads from one head next to the same head's x-neighbour (msda.cuh).
#pragma once
#include "qkv_pair.cuh"

namespace axvs {

constexpr int MF_CHUNKS = 5;                    // 3 x 128 columns of offsets | logits, 2 x 128 columns of value
constexpr int MF_OA_CHUNKS = 3;

struct MsdaFrontParams {
  const float* src;        // fp32 [rows, 256]
  const float* pos;        // fp32 [rows or len, 256] or null (dims.pos_mod = len: one table shared by the images)
  const uint8_t* w;        // unit image (pack_weight_units, k_major 0) of [W_oa zero-padded to 384 rows ; W_value]: 10 units
  const float* bias;       // [640] = b_oa zero-padded to 384 | b_value
  float* oa;               // fp32 [rows, n_oa]
  __nv_bfloat16* value;    // bf16 [images][8][len][32]
  int rows, tiles, len, n_oa;
  AxialDims dims;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(QD_THREADS, 1) msda_front_pair_kernel(const MsdaFrontParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QD_A_SLOTS * 2 * TF_KB;
  uint8_t* stage_all = w_ring + QP_W_SLOTS * QP_WH;
  float* sbias = reinterpret_cast<float*>(stage_all + QK_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QD_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QD_A_SLOTS;      // tcgen05.commit after the slot's copies
  uint64_t* w_full = a_empty + QD_A_SLOTS;      // [QP_W_SLOTS]
  uint64_t* w_empty = w_full + QP_W_SLOTS;
  uint64_t* s_full = w_empty + QP_W_SLOTS;      // [2] accumulator stage complete
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
  for (int i = threadIdx.x; i < MF_CHUNKS * 128; i += QD_THREADS) sbias[i] = p.bias[i];
  __syncthreads();
  cluster_sync_all();                                          // both CTAs' barriers are initialised before any remote arrive
  if (warp == 17) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue: group g drains column groups 2g, 2g+1 of every chunk
    const int g = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* stg = stage_all + warp * 2048;
    uint32_t cnt = 0;                                          // chunks consumed (stage = cnt & 1; five per tile: the stage parity alternates between tiles)
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      const int tile = 2 * pt + (int)rank;                     // may be == p.tiles (odd tile count): every row masked
      const int row0 = tile * 128 + (warp & 3) * 32;           // first row of this warp
      // head-major value row (head 0) of the FOUR rows this lane stores after the transpose (rows 8 i + lane / 4 of the warp's 32)
      uint32_t vrow4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + 8 * i + (lane >> 2);
        const int img = r / p.len;
        vrow4[i] = (uint32_t)img * 8u * (uint32_t)p.len + (uint32_t)(r - img * p.len);
      }
#pragma unroll 1
      for (int rt = 0; rt < MF_CHUNKS; ++rt, ++cnt) {
        const int st = cnt & 1;
        mbar_wait_cluster(&s_full[st], (cnt >> 1) & 1);
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + st * 128;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {                       // 32 columns at a time
          const int c = 2 * g + cc;
          const int col = rt * 128 + c * 32;                   // column of [oa (384) | value (256)]
          const bool live = rt >= MF_OA_CHUNKS || col < p.n_oa;        // zero-weight padding columns of the offsets | logits block are not stored
          float v[32];
          if (live) {
            tmem_ld32(t_s + 32 * c, v);
            tmem_ld_wait();
          }
          if (cc == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(&s_empty[st], 0);
          }
          if (!live) continue;
          const float4* b4 = reinterpret_cast<const float4*>(sbias + col);
          if (rt < MF_OA_CHUNKS) {
            // fp32 offsets | logits: two halves of 16 columns through the 2 KiB per-warp transpose (32 rows x 64 B), so every store
            // instruction writes 8 row pieces of 64 contiguous bytes
#pragma unroll
            for (int h = 0; h < 2; ++h) {
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
                const int rl = 8 * i + (lane >> 2), piece = lane & 3;
                const float4 u = *reinterpret_cast<const float4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
                if (row0 + rl < p.rows) *reinterpret_cast<float4*>(p.oa + (size_t)(row0 + rl) * p.n_oa + col + 16 * h + piece * 4) = u;
              }
              __syncwarp();
            }
          } else {
            // bf16 value, one head (32 channels = 64 B per row)
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
            __syncwarp();
            const uint32_t head_off = (uint32_t)((rt - MF_OA_CHUNKS) * 4 + c) * (uint32_t)p.len;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + (lane >> 2), piece = lane & 3;
              const uint4 u = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
              if (row0 + rl < p.rows)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.value) + (size_t)(vrow4[i] + head_off) * 64 + piece * 16) = u;
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 8 + QD_PRODUCER_WARPS) {
    // =============================================================== converting A producers (as qkv_pair_kernel: bursts of one K-block)
    const int pw = warp - 8;
    const int half = lane >> 4, c16 = lane & 15;               // row of the pair, 16-byte piece (4 channels) of the 256-byte segment
    uint32_t cnt = 0;
    float4 sv[8], qv[8];
    uint32_t crow[8];                                          // token of this lane's 8 rows (0xFFFFFFFF = past the end)
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      const int tile = 2 * pt + (int)rank;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int pr = tile * 128 + pw * 16 + 2 * j + half;
        crow[j] = pr < p.rows ? (uint32_t)pr : 0xFFFFFFFFu;
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
        const uint32_t slot = cnt % QD_A_SLOTS, phase = (cnt / QD_A_SLOTS) & 1;
        mbar_wait_cluster(&a_empty[slot], phase ^ 1);
        uint8_t* dst = a_ring + slot * 2 * TF_KB;
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
        if (lane == 0) mbar_arrive(&a_full[slot]);
      }
    }
  } else if (warp == 16 && lane == 0) {
    // =============================================================== weight producer: my half (64 rows) of each of the 10 units per tile
    uint32_t slot = 0, phase = 0;
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
#pragma unroll 1
      for (int u = 0; u < 2 * MF_CHUNKS; ++u) {
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
        for (int u = 0; u < 2 * MF_CHUNKS; ++u) {
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
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;     // ccnt: chunks issued
    for (int pt = pair; pt < pair_tiles; pt += npairs) {
      // both CTAs' A operands -> TMEM; ordered by the tensor pipe behind every UMMA of the previous tile (issued earlier by this thread)
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
        const uint32_t slot = a_cnt % QD_A_SLOTS;
        mbar_wait_cluster(&a_full[slot], (a_cnt / QD_A_SLOTS) & 1);
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
      for (int rt = 0; rt < MF_CHUNKS; ++rt, ++ccnt) {
        const int g = ccnt & 1;                                // accumulator stage
        const uint32_t gc = ccnt >> 1;                         // chunks already issued to that stage
        mbar_wait_cluster(&s_empty[g], (gc & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_a = tmem + (rt < MF_OA_CHUNKS ? 0 : 128);     // A1 for the offsets | logits chunks, A2 for the value chunks
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          mbar_wait_cluster(&w_full[w_slot], w_phase);
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QP_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          umma_unit_elect_ts_pair(tmem + 256 + g * 128, t_a + 64 * kg, t_a + 64 * kg + 32, w_ring_addr + ws * QP_WH, idesc, kg != 0,
                                  &w_empty[ws], kg == 1 ? &s_full[g] : nullptr);
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
