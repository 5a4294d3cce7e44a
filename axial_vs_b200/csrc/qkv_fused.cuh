// q | k | v projections (WC/temporal_attention.py:42-44; CC:98) as one persistent tcgen05 kernel per 128-row tile:
//
//   [q | k] = A1 [Wq; Wk]^T + b,   v = A2 Wv^T + b      A1 = bf16(q_in + pos), A2 = bf16(v_in)   (tile images, TMA-fed)
//
// Six 128-column chunks (q heads 0-3, q heads 4-7, k ..., v ...) rotate through four TMEM accumulator stages; the
// epilogue adds the bias and writes bf16 in a HEAD-MAJOR layout  qkv[which][head][row][32]  so that (a) a warp's 32
// threads (one row each) write 2 KiB contiguous per head -- fully coalesced straight from registers -- and (b) the
// attention kernel reads each (sequence, head) operand as one contiguous block.
//
// Warp roles (384 threads): warps 0-3 / 4-7 = epilogue groups (even / odd chunks), warp 8 = A-tile TMA producer,
// warp 9 = weight TMA producer (32 KiB units), warp 10 = MMA issuer (converged warp, elected lane).
#pragma once
#include "traj_fused.cuh"

namespace axvs {

constexpr int QK_THREADS = 384;
constexpr int QK_A_SLOTS = 6;
constexpr int QK_W_SLOTS = 3;
constexpr int QK_BIAS_BYTES = 768 * 4;
constexpr int QK_STAGE_BYTES = 8 * 2048;     // per-warp 32 rows x 64 B transpose tile for the head-major stores
constexpr int QK_SMEM_BYTES = QK_A_SLOTS * TF_KB + QK_W_SLOTS * TF_WU + QK_STAGE_BYTES + QK_BIAS_BYTES + 512;

struct QkvParams {
  const uint8_t* a1_img;   // [tiles][4][16 KiB]   q/k input (+pos)
  const uint8_t* a2_img;   // [tiles][4][16 KiB]   v input (may equal a1_img)
  const uint8_t* w;        // unit format of [Wq; Wk; Wv] (768 x 256): unit = 2 * rt + kg, rt = 0..5
  const float* bias;       // [768]
  __nv_bfloat16* qkv;      // head-major [3][8][rows][32]
  int rows, tiles;
};

__global__ void __launch_bounds__(QK_THREADS, 1) qkv_fused_kernel(const QkvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QK_A_SLOTS * TF_KB;
  uint8_t* stage_all = w_ring + QK_W_SLOTS * TF_WU;
  float* sbias = reinterpret_cast<float*>(stage_all + QK_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QK_A_SLOTS]
  uint64_t* a_empty = a_full + QK_A_SLOTS;
  uint64_t* w_full = a_empty + QK_A_SLOTS;      // [QK_W_SLOTS]
  uint64_t* w_empty = w_full + QK_W_SLOTS;
  uint64_t* s_full = w_empty + QK_W_SLOTS;      // [4]
  uint64_t* s_empty = s_full + 4;               // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QK_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QK_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 768; i += QK_THREADS) sbias[i] = p.bias[i];
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
          // bias + bf16, then a 2 KiB per-warp transpose so every store instruction writes 512 contiguous bytes
          // (one row per thread would emit 16-byte pieces at a 64-byte stride: 32 half-filled sectors per instruction)
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
  } else if (warp == 8 && lane == 0) {
    // =============================================================== A-tile producer: A1 kb0..3, then A2 kb0..3
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int item = 0; item < 8; ++item, ++cnt) {
        const uint32_t slot = cnt % QK_A_SLOTS, phase = (cnt / QK_A_SLOTS) & 1;
        const uint8_t* src = (item < 4 ? p.a1_img : p.a2_img) + ((size_t)tile * 4 + (item & 3)) * TF_KB;
        mbar_wait(&a_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&a_full[slot], TF_KB);
        tma_bulk_g2s(a_ring + slot * TF_KB, src, TF_KB, &a_full[slot]);
      }
    }
  } else if (warp == 9 && lane == 0) {
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
  } else if (warp == 10) {
    // =============================================================== MMA issuer
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;     // ccnt: chunks issued (stage = 2 * ((ccnt >> 1) & 1) + (ccnt & 1))
    AXVS_PROF_DECL(3)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++ccnt) {
        const int g = rt & 1;
        const uint32_t gc = ccnt >> 1;                         // chunks already issued to group g (6 per tile: even count)
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
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// a1 = bf16(src + pos), a2 = bf16(src) written as tile images (pass-order rows); replaces pack_kq_kernel on the fused path.
// lane l owns columns 8l..8l+7 = K-block l/8, 16-byte chunk l%8 of the image row.
__global__ void pack_image_kernel(const float* __restrict__ src, const float* __restrict__ pos, uint8_t* __restrict__ a1, uint8_t* __restrict__ a2,
                                  int rows, int map_mode, AxialDims d) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int pr = blockIdx.x * wpb + (threadIdx.x >> 5); pr < rows; pr += gridDim.x * wpb) {
    const size_t c = (size_t)pass_to_canonical(pr, map_mode, d);
    const float4* s4 = reinterpret_cast<const float4*>(src + c * C256) + lane * 2;
    float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1);
    const size_t off = ((size_t)(pr >> 7) * 4 + (lane >> 3)) * TF_KB + sw128_offset(pr & 127, lane & 7);
    if (a2) {
      uint4 u;
      u.x = pack_bf16x2(s0.x, s0.y); u.y = pack_bf16x2(s0.z, s0.w);
      u.z = pack_bf16x2(s1.x, s1.y); u.w = pack_bf16x2(s1.z, s1.w);
      *reinterpret_cast<uint4*>(a2 + off) = u;
    }
    if (pos) {
      const float4* p4 = reinterpret_cast<const float4*>(pos + (size_t)pos_row((uint32_t)c, d) * C256) + lane * 2;
      const float4 q0 = __ldg(p4), q1 = __ldg(p4 + 1);
      s0.x += q0.x; s0.y += q0.y; s0.z += q0.z; s0.w += q0.w;
      s1.x += q1.x; s1.y += q1.y; s1.z += q1.z; s1.w += q1.w;
    }
    uint4 u;
    u.x = pack_bf16x2(s0.x, s0.y); u.y = pack_bf16x2(s0.z, s0.w);
    u.z = pack_bf16x2(s1.x, s1.y); u.w = pack_bf16x2(s1.z, s1.w);
    *reinterpret_cast<uint4*>(a1 + off) = u;
  }
}

}  // namespace axvs
static_assert(axvs::QK_SMEM_BYTES <= 232448, "qkv_fused_kernel exceeds the 227 KiB shared-memory limit");
