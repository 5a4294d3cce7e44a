// Per-frame-softmax spatial attention of TrajectoryAttention on the 5th-generation tensor cores (Appendix A steps 2-5):
//   x[s, q, f, head*32 + :] = softmax_i( scale * Q[s,q,head,:] . K[s, f*n+i, head, :] ) @ V[s, f*n+i, head, :]
// Reference: WC/temporal_attention.py:47-60 (the softmax is taken independently inside every key frame).
//
// Work unit = (sequence, head, block of 128 queries, chunk of FC key frames):
//   S_f = Q K_f^T     tcgen05.mma, both operands in shared memory (K-major SWIZZLE_64B tiles of 64-byte rows = one head),
//                     M = 128, N = NP = n rounded up to 16, K = 32 (two instructions) -> NP fp32 columns of tensor memory
//   softmax           one thread per query row (TMEM lane): tcgen05.ld of its score row, max, 2^x, sum in registers;
//                     the un-normalised probabilities go back IN PLACE as bf16 pairs (tcgen05.st) = the A operand of P V
//   O_f = P_f V_f     tcgen05.mma with A in tensor memory and V_f in shared memory as an MN-major SWIZZLE_64B operand
//                     (rows = keys, 32 contiguous channels per row: exactly how the q|k|v kernel stores v), N = 32
//   epilogue          O_f / l_f -> bf16 -> the x_f (and, for the query's own frame, x_diag) rows of the SWIZZLE_128B
//                     tile images traj_ts_kernel loads by TMA
// q | k | v arrive head-major ([which][head][row][32] bf16) with the 16-byte chunks of every row pre-permuted by the
// q|k|v kernel (chunk c of a row at position c ^ ((i >> 1) & 3), i = row index inside the sequence for q, inside the key
// frame for k / v), so plain 1-D TMA bulk copies land ready-to-use SWIZZLE_64B tiles.  Operand forms validated in
// tools/microbench/umma_sw64.cu (K-major and MN-major SWIZZLE_64B descriptors: SBO = 512 B, LBO unused).
//
// The per-unit tensor work is tiny (5-10 instructions), so every fixed cost is paid once per UNIT, not per frame: one
// shared-memory slot = Q + the chunk's K_f | V_f, filled by ONE warp-wide cp.async.bulk (lane 0: Q, lanes 1 + 2f / 2 + 2f:
// K_f / V_f) against one mbarrier; one commit per phase.  G softmax / epilogue groups of 4 warps (warps 0 .. 4G-1, TMEM
// lane quarter = warp & 3) work on G units at once (unit k -> TMEM buffer and group k % G); the issuer runs the S phase up to
// G - 1 units ahead of the P V phase.  Then one TMA producer warp and one MMA issuer warp.
#pragma once
#include "attn.cuh"

namespace axvs {

constexpr int AT_Q_BYTES = 128 * 64;
constexpr int AT_MAX_SLOTS = 12;
constexpr int AT_MAX_G = 4;

struct AttnTcParams {
  const __nv_bfloat16* qkv;   // head-major [3][8][rows_total][32], chunks pre-permuted (see above)
  size_t rows_total;
  uint8_t* x_img;             // [F][tiles][4][16 KiB]
  uint8_t* xd_img;            // [tiles][4][16 KiB]
  int tiles, N, n, F, NP, QB; // QB = query blocks per sequence
  int FC, NCH;                // frames per chunk, chunks per item (NCH = ceil(F / FC))
  int G, buf_cols;            // softmax groups = TMEM buffers, columns per buffer (>= FC * (NP + 32))
  int num_units;              // num_seq * 8 * QB * NCH
  int slots, slot_bytes;      // shared-memory ring
  float scale_log2e;
};

constexpr uint32_t AT_DESC_HI_SW64 = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B (8 rows of 64 B), version 1, SWIZZLE_64B

__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void umma_ss_raw(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_raw(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// unit index -> (item, chunk); item -> (sequence * 8 + head, query block)
struct AtUnit { int sh, qb, f0, fc; };
__device__ __forceinline__ AtUnit at_decode(int unit, const AttnTcParams& p) {
  AtUnit u;
  const int ch = unit % p.NCH;
  const int item = unit / p.NCH;
  u.qb = item % p.QB;
  u.sh = item / p.QB;
  u.f0 = ch * p.FC;
  u.fc = min(p.FC, p.F - u.f0);
  return u;
}

// One frame's softmax for this thread's query row.  NT16 > 0: the score row (NP = 16 * NT16 <= 64 columns) is held in registers
// (one TMEM round trip); NT16 == 0: two passes over 16-column pieces (any NP).  Returns the row sum of the probabilities.
template <int NT16>
__device__ __forceinline__ float at_softmax_frame(uint32_t t_s, int NP, int n, float sc) {
  float sum = 0.f;
  if constexpr (NT16 > 0) {
    float v[NT16][16];
#pragma unroll
    for (int c = 0; c < NT16; ++c) tmem_ld16(t_s + 16 * c, v[c]);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NT16; ++c) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (16 * c + i >= n) v[c][i] = -INFINITY;             // resolved per (c, i) at run time only for the straddling piece
        mx = fmaxf(mx, v[c][i]);
      }
    }
    const float mxs = -mx * sc;
#pragma unroll
    for (int c = 0; c < NT16; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e0 = ex2_approx(fmaf(v[c][2 * i], sc, mxs)), e1 = ex2_approx(fmaf(v[c][2 * i + 1], sc, mxs));   // 2^-inf = 0 past the frame
        sum += e0 + e1;
        pk[i] = pack_bf16x2(e0, e1);
      }
      tmem_st8u(t_s + 8 * c, pk);
    }
  } else {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < NP; c0 += 16) {
      float v[16];
      tmem_ld16(t_s + c0, v);
      tmem_ld_wait();
      if (c0 + 16 <= n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (c0 + i < n) mx = fmaxf(mx, v[i]);
      }
    }
    const float mxs = -mx * sc;
#pragma unroll 1
    for (int c0 = 0; c0 < NP; c0 += 16) {                     // writes columns [c0/2, c0/2 + 8): always behind the reads
      float v[16];
      tmem_ld16(t_s + c0, v);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e0 = (c0 + 2 * i < n) ? ex2_approx(fmaf(v[2 * i], sc, mxs)) : 0.f;
        const float e1 = (c0 + 2 * i + 1 < n) ? ex2_approx(fmaf(v[2 * i + 1], sc, mxs)) : 0.f;
        sum += e0 + e1;
        pk[i] = pack_bf16x2(e0, e1);
      }
      tmem_st8u(t_s + (c0 >> 1), pk);
    }
  }
  return sum;
}

constexpr int AT_MAX_FC = 4;   // frames per chunk the epilogue keeps row sums for

template <int NT16>
__global__ void __launch_bounds__(128 * AT_MAX_G + 64, 1) spatial_attn_tc_kernel(const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.slots * p.slot_bytes);
  uint64_t* full = bars;                            // [AT_MAX_SLOTS]
  uint64_t* empty = full + AT_MAX_SLOTS;
  uint64_t* s_full = empty + AT_MAX_SLOTS;          // [G] scores of the buffer's unit complete
  uint64_t* p_full = s_full + AT_MAX_G;             // [G] probabilities written (4 warps)
  uint64_t* o_full = p_full + AT_MAX_G;             // [G] P V complete
  uint64_t* buf_free = o_full + AT_MAX_G;           // [G] O read back (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(buf_free + AT_MAX_G);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int F = p.F, n = p.n, N = p.N, NP = p.NP, G = p.G;

  // zero the ring once: the pad rows [n, NP) of every K_f | V_f tile and the rows of a Q tile past the end of a sequence are
  // never written by the bulk copies and must stay finite (their products meet zero probabilities / are ignored)
  {
    const int total16 = (p.slots * p.slot_bytes) >> 4;
    for (int i = threadIdx.x; i < total16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < AT_MAX_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < AT_MAX_G; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&buf_free[i], 4); }
    fence_barrier_init();
  }
  if (warp == 4 * G + 1) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4 * G) {
    // =============================================================== softmax + epilogue groups
    const int g = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_buf = tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * p.buf_cols;
    const float sc = p.scale_log2e;
    uint32_t use = 0;                                         // units this group has processed
    for (int unit = blockIdx.x + g * gridDim.x; unit < p.num_units; unit += G * gridDim.x, ++use) {
      const AtUnit u = at_decode(unit, p);
      const int head = u.sh & 7;
      const size_t seq_row0 = (size_t)(u.sh >> 3) * N;
      const int qi = u.qb * 128 + row;                        // query index inside the sequence
      const bool valid = qi < N;
      const uint32_t par = use & 1;
      const uint32_t t_o = t_buf + u.fc * NP;
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      float inv[AT_MAX_FC];
#pragma unroll
      for (int j = 0; j < AT_MAX_FC; ++j)
        if (j < u.fc) inv[j] = at_softmax_frame<NT16>(t_buf + j * NP, NP, n, sc);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // ---- epilogue: O_f / l_f -> bf16 -> tile images
      const int kb = head >> 1, ch0 = (head & 1) * 4;
      const size_t r = seq_row0 + (valid ? qi : 0);
      const size_t img_off = ((r >> 7) * 4 + kb) * (size_t)ATT2_KB;
      const uint32_t rl = (uint32_t)(r & 127);
      mbar_wait(&o_full[g], par);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < AT_MAX_FC; ++j) {
        if (j < u.fc) {
          float o[32];
          tmem_ld32(t_o + 32 * j, o);
          tmem_ld_wait();
          if (j == u.fc - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&buf_free[g]);
          }
          if (valid) {
            const int f = u.f0 + j;
            const float il = __frcp_rn(inv[j]);
            uint8_t* dst = p.x_img + (size_t)f * p.tiles * 4 * ATT2_KB + img_off;
            const bool diag = (unsigned)(qi - f * n) < (unsigned)n;   // qi / n == f
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 w;
              w.x = pack_bf16x2(o[8 * c] * il, o[8 * c + 1] * il);
              w.y = pack_bf16x2(o[8 * c + 2] * il, o[8 * c + 3] * il);
              w.z = pack_bf16x2(o[8 * c + 4] * il, o[8 * c + 5] * il);
              w.w = pack_bf16x2(o[8 * c + 6] * il, o[8 * c + 7] * il);
              const uint32_t off = sw128_offset(rl, ch0 + c);
              *reinterpret_cast<uint4*>(dst + off) = w;
              if (diag) *reinterpret_cast<uint4*>(p.xd_img + img_off + off) = w;
            }
          }
        }
      }
    }
  } else if (warp == 4 * G) {
    // =============================================================== TMA producer: one warp-wide bulk copy per unit
    uint32_t cnt = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++cnt) {
      const AtUnit u = at_decode(unit, p);
      const int head = u.sh & 7;
      const size_t seq_row0 = (size_t)(u.sh >> 3) * N;
      const int rows_q = min(128, N - u.qb * 128);
      const uint32_t slot = cnt % p.slots;
      uint8_t* dst = ring + (size_t)slot * p.slot_bytes;
      if (lane == 0) {
        mbar_wait(&empty[slot], ((cnt / p.slots) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[slot], rows_q * 64 + u.fc * 2 * n * 64);
      }
      __syncwarp();
      // lane 0: Q rows; lane 1 + 2j: K of frame f0 + j; lane 2 + 2j: V of frame f0 + j
      if (lane < 1 + 2 * u.fc) {
        const int j = (lane - 1) >> 1, isv = (lane - 1) & 1;
        const void* src;
        uint8_t* d;
        uint32_t bytes;
        if (lane == 0) {
          src = p.qkv + ((size_t)head * p.rows_total + seq_row0 + (size_t)u.qb * 128) * 32;
          d = dst;
          bytes = rows_q * 64;
        } else {
          src = p.qkv + ((size_t)((isv ? 16 : 8) + head) * p.rows_total + seq_row0 + (size_t)(u.f0 + j) * n) * 32;
          d = dst + AT_Q_BYTES + (2 * j + isv) * NP * 64;
          bytes = n * 64;
        }
        tma_bulk_g2s(d, src, bytes, &full[slot]);
      }
      __syncwarp();
    }
  } else {
    // =============================================================== MMA issuer (converged warp, elected lane)
    const uint32_t idesc_s = umma_idesc_bf16(128, NP);
    const uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);        // B (= V) is MN-major
    const uint32_t ring_addr = smem_u32(ring);
    uint32_t k = 0;                                            // units whose S phase has been issued
    uint32_t pv_k = 0;                                         // units whose P V phase has been issued
    int pv_unit = blockIdx.x;
    auto issue_pv = [&]() {
      const uint32_t b = pv_k % G, slot = pv_k % p.slots;
      const int fc = at_decode(pv_unit, p).fc;
      mbar_wait(&p_full[b], (pv_k / G) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_p = tmem + b * p.buf_cols;
        for (int j = 0; j < fc; ++j) {
          const uint32_t v_lo = (((ring_addr + slot * p.slot_bytes + AT_Q_BYTES + (2 * j + 1) * NP * 64) & 0x3FFFFu) >> 4) | (64u << 16);
          for (int kk = 0; kk < (NP >> 4); ++kk)
            umma_ts_raw(t_p + fc * NP + 32 * j, t_p + j * NP + 8 * kk, v_lo + kk * 64, AT_DESC_HI_SW64, idesc_o, kk ? 1u : 0u);
        }
        umma_commit(&o_full[b]);
        umma_commit(&empty[slot]);
      }
      __syncwarp();
      ++pv_k;
      pv_unit += gridDim.x;
    };
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++k) {
      const int fc = at_decode(unit, p).fc;
      const uint32_t b = k % G, slot = k % p.slots;
      mbar_wait(&full[slot], (k / p.slots) & 1);
      mbar_wait(&buf_free[b], ((k / G) & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t qa = (((ring_addr + slot * p.slot_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
        for (int j = 0; j < fc; ++j) {
          const uint32_t ka = qa + ((AT_Q_BYTES + 2 * j * NP * 64) >> 4);
          umma_ss_raw(tmem + b * p.buf_cols + j * NP, qa, AT_DESC_HI_SW64, ka, AT_DESC_HI_SW64, idesc_s, 0u);
          umma_ss_raw(tmem + b * p.buf_cols + j * NP, qa + 2, AT_DESC_HI_SW64, ka + 2, AT_DESC_HI_SW64, idesc_s, 1u);
        }
        umma_commit(&s_full[b]);
      }
      __syncwarp();
      if (k + 1 - pv_k >= (uint32_t)G) issue_pv();              // keep the S phase at most G - 1 units ahead of the P V phase
    }
    while (pv_k < k) issue_pv();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 * G + 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
