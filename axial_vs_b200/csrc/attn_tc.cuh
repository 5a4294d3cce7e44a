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
//
// Operand layout in HBM ("unit-major", written by qkv_direct_kernel when swz_N > 0): one contiguous region of 3 N rows of
// 64 bytes per (sequence, head):  Q rows [0, N), then per key frame f its K rows at N + 2 f n and its V rows at N + (2 f + 1) n;
// the 16-byte chunks of the row at position `pos` of the region are permuted by (pos >> 1) & 3.  The tensor core applies
// the SWIZZLE_64B XOR to ABSOLUTE shared-memory address bits (tools/microbench/umma_sw64.cu: operand tiles may start at any
// row of a 512-byte-aligned buffer, base_offset = 0), so ONE 1-D TMA bulk copy of the region lands Q and every K_f / V_f
// ready to use, with no padding between the frames: an instruction that reads NP > n rows runs into the next tile's rows,
// whose scores are never read and whose probabilities are zero.  Units that do not cover a whole (sequence, head) -- more
// than 128 queries, or more frames than fit in tensor memory -- take two copies (the Q block, the K_f | V_f rows of the chunk).
//
// The per-unit tensor work is tiny (5-10 instructions), so every fixed cost is paid once per unit: one shared-memory slot and
// one mbarrier per unit, one commit per phase, descriptors computed by the converged issuer warp outside the elected branch
// (a lone diverged lane pays ~150 clk per instruction).  G softmax / epilogue groups of 4 warps (warps 0 .. 4G-1, TMEM lane
// quarter = warp & 3) work on G units at once (unit k -> TMEM buffer and group k % G); the issuer runs the S phase up to G - 1
// units ahead of the P V phase.  Then one TMA producer warp and one MMA issuer warp.
#pragma once
#include "attn.cuh"

namespace axvs {

constexpr int AT_Q_BYTES = 128 * 64;
constexpr int AT_MAX_SLOTS = 12;
constexpr int AT_MAX_G = 4;

struct AttnTcParams {
  const __nv_bfloat16* qkv;   // unit-major: [(sequence, head)][3 N rows][32], chunks permuted (see above)
  size_t rows_total;
  uint8_t* x_img;             // [F][tiles][4][16 KiB]
  uint8_t* xd_img;            // [tiles][4][16 KiB]
  int tiles, N, n, F, NP, QB; // QB = query blocks per sequence
  int FC, NCH;                // frames per chunk, chunks per item (NCH = ceil(F / FC))
  int G, buf_cols;            // softmax groups = TMEM buffers, columns per buffer
  int p_stride, o_off;        // buffer layout: S_j at j NP (fp32 scores); P_j (bf16 pairs) at j p_stride; O_j at o_off + 32 j.  Plain: p_stride = NP
                              // (P_j in place over S_j), o_off = FC NP.  Compact: p_stride = NP / 2, o_off = FC NP / 2 rounded up to 32 -- the
                              // outputs land on score columns that are dead by then, 128 columns hold two frames of 48 keys (4 groups, not 3)
  int num_units;              // num_seq * 8 * QB * NCH
  int slots, slot_bytes;      // shared-memory ring
  int single;                 // 1: a unit is a whole (sequence, head) region -> one bulk copy
  int tm_rpad;                // > 0: frame-major output rows (row = t * tm_rpad + sequence * n + j, see TrajParams), no x_diag image
  uint32_t n_magic;           // floor(2^32 / n) + 1: qi / n == umulhi(qi, n_magic) for qi < 2^16
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

__device__ __forceinline__ float max3_f32(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// unit index -> (item, chunk); item -> (sequence * 8 + head, query block)
struct AtUnit { int sh, qb, f0, fc; };
__device__ __forceinline__ AtUnit at_decode(int unit, const AttnTcParams& p) {
  AtUnit u;
  if (p.single) { u.sh = unit; u.qb = 0; u.f0 = 0; u.fc = p.F; return u; }
  const int ch = unit % p.NCH;
  const int item = unit / p.NCH;
  u.qb = item % p.QB;
  u.sh = item / p.QB;
  u.f0 = ch * p.FC;
  u.fc = min(p.FC, p.F - u.f0);
  return u;
}

// One frame's softmax for this thread's query row.  NT16 > 0: the score row (NP = 16 * NT16 <= 64 columns) is held in registers
// (one TMEM round trip; NP <= 96); NT16 == 0: two passes over 32-column pieces with prefetch (any NP).  Returns the row sum of the probabilities.
template <int NT16>
__device__ __forceinline__ float at_softmax_frame(uint32_t t_s, uint32_t t_p, int NP, int n, float sc) {
  float sum = 0.f;
  if constexpr (NT16 > 0) {
    float v[NT16][16];
#pragma unroll
    for (int c = 0; c < NT16; ++c) tmem_ld16(t_s + 16 * c, v[c]);
    tmem_ld_wait();
    {
      // NP = n rounded up to 16, so only the LAST 16-column piece can reach past the end of the frame
      const int tail = n - 16 * (NT16 - 1);                    // valid columns of the last piece, 1 .. 16
#pragma unroll
      for (int i = 1; i < 16; ++i) if (i >= tail) v[NT16 - 1][i] = -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NT16; ++c) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) mx = max3_f32(mx, v[c][i], v[c][i + 1]);
    }
    const float2 sc2 = make_float2(sc, sc), mxs2 = make_float2(-mx * sc, -mx * sc);
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NT16; ++c) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 a = fma_f32x2(make_float2(v[c][2 * i], v[c][2 * i + 1]), sc2, mxs2);   // 2^-inf = 0 past the frame
        const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        sum2 = add_f32x2(sum2, e);
        pk[i] = pack_bf16x2(e.x, e.y);
      }
      tmem_st8u(t_p + 8 * c, pk);                              // every score of the frame is in registers: any column at or below t_s + NP is free
    }
    sum = sum2.x + sum2.y;
  } else {
    // long frames (NP > 96): two passes over 32-column pieces, the next piece's tcgen05.ld in flight while the current one is
    // processed (tcgen05.wait::ld covers every load issued before it, so the prefetch is issued right after the wait)
    const int NP32 = NP & ~31;
    float va[32], vb[32];
    auto max32 = [&](const float (&v)[32], int c0, float m) {
      if (c0 + 32 <= n) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) m = max3_f32(m, v[i], v[i + 1]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (c0 + i < n) m = fmaxf(m, v[i]);
      }
      return m;
    };
    float mx = -INFINITY;
    if (NP32 > 0) tmem_ld32(t_s, va);
#pragma unroll 1
    for (int c0 = 0; c0 < NP32; c0 += 64) {
      tmem_ld_wait();
      if (c0 + 32 < NP32) tmem_ld32(t_s + c0 + 32, vb);
      mx = max32(va, c0, mx);
      if (c0 + 32 < NP32) {
        tmem_ld_wait();
        if (c0 + 64 < NP32) tmem_ld32(t_s + c0 + 64, va);
        mx = max32(vb, c0 + 32, mx);
      }
    }
    if (NP & 16) {
      float v[16];
      tmem_ld16(t_s + NP32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) if (NP32 + i < n) mx = fmaxf(mx, v[i]);
    }
    const float mxs = -mx * sc;
    const float2 sc2 = make_float2(sc, sc), mxs2 = make_float2(mxs, mxs);
    float2 sum2 = make_float2(0.f, 0.f);
    auto exp32 = [&](const float (&v)[32], int c0) {           // probabilities of columns [c0, c0 + 32) -> packed columns [c0/2, c0/2 + 16)
      uint32_t pk[16];
      if (c0 + 32 <= n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 a = fma_f32x2(make_float2(v[2 * i], v[2 * i + 1]), sc2, mxs2);
          const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
          sum2 = add_f32x2(sum2, e);
          pk[i] = pack_bf16x2(e.x, e.y);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = (c0 + 2 * i < n) ? ex2_approx(fmaf(v[2 * i], sc, mxs)) : 0.f;
          const float e1 = (c0 + 2 * i + 1 < n) ? ex2_approx(fmaf(v[2 * i + 1], sc, mxs)) : 0.f;
          sum2 = add_f32x2(sum2, make_float2(e0, e1));
          pk[i] = pack_bf16x2(e0, e1);
        }
      }
      tmem_st16u(t_s + (c0 >> 1), pk);                         // always behind the columns still to be read
    };
    if (NP32 > 0) tmem_ld32(t_s, va);
#pragma unroll 1
    for (int c0 = 0; c0 < NP32; c0 += 64) {
      tmem_ld_wait();
      if (c0 + 32 < NP32) tmem_ld32(t_s + c0 + 32, vb);
      exp32(va, c0);
      if (c0 + 32 < NP32) {
        tmem_ld_wait();
        if (c0 + 64 < NP32) tmem_ld32(t_s + c0 + 64, va);
        exp32(vb, c0 + 32);
      }
    }
    if (NP & 16) {
      float v[16];
      tmem_ld16(t_s + NP32, v);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e0 = (NP32 + 2 * i < n) ? ex2_approx(fmaf(v[2 * i], sc, mxs)) : 0.f;
        const float e1 = (NP32 + 2 * i + 1 < n) ? ex2_approx(fmaf(v[2 * i + 1], sc, mxs)) : 0.f;
        sum2 = add_f32x2(sum2, make_float2(e0, e1));
        pk[i] = pack_bf16x2(e0, e1);
      }
      tmem_st8u(t_s + (NP32 >> 1), pk);
    }
    sum = sum2.x + sum2.y;
  }
  return sum;
}

constexpr int AT_MAX_FC = 4;   // frames per chunk the epilogue keeps row sums for

template <int NT16>
__global__ void __launch_bounds__((NT16 >= 1 && NT16 <= 3) ? 128 * 4 + 96 : (NT16 == 5 || NT16 == 6) ? 128 * 2 + 96 : 128 * 3 + 96, 1) spatial_attn_tc_kernel(const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ring = smem;
  uint8_t* stg_all = ring + (size_t)p.slots * p.slot_bytes;   // [4 G warps][32 rows x 64 B] transpose staging of the epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_all + 4 * AT_MAX_G * 2048);
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

  // zero the ring once: rows an instruction reads beyond what the bulk copies wrote (past the last V_f, past a short Q block)
  // must be finite -- their products meet zero probabilities or land in score rows / columns nobody reads
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
    uint8_t* stg = stg_all + warp * 2048;
    const uint32_t t_buf = tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * p.buf_cols;
    const float sc = p.scale_log2e;
    uint32_t use = 0;                                         // units this group has processed
    AXVS_PROF_DECL(2)
    for (int unit = blockIdx.x + g * gridDim.x; unit < p.num_units; unit += G * gridDim.x, ++use) {
      const AtUnit u = at_decode(unit, p);
      const int head = u.sh & 7;
      const size_t seq_row0 = (size_t)(u.sh >> 3) * N;
      const uint32_t par = use & 1;
      const uint32_t t_o = t_buf + p.o_off;
      AXVS_PROF_WAIT(0, mbar_wait(&s_full[g], par))
      tc_fence_after();
      float inv[AT_MAX_FC];
#pragma unroll
      for (int j = 0; j < AT_MAX_FC; ++j)
        if (j < u.fc) inv[j] = at_softmax_frame<NT16>(t_buf + j * NP, t_buf + j * p.p_stride, NP, n, sc);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g]);
      // ---- epilogue: O_f / l_f -> bf16 -> tile images.  A thread holds one row (64 bytes per frame); a 2 KiB per-warp transpose
      // lets every store instruction write whole 64-byte row segments (4 lanes per row) instead of 32 scattered 16-byte pieces
      // (those cost one L1 wavefront each and bound the first version of this kernel).
      const int kb = head >> 1, ch0 = (head & 1) * 4;
      const int q_w0 = u.qb * 128 + (warp & 3) * 32;          // query index of this warp's first row
      uint32_t orow4[4];                                      // tile-order row of the four rows this lane stores (0xFFFFFFFF: past the sequence)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q_w0 + 8 * i + (lane >> 2);
        if (qi >= N) orow4[i] = 0xFFFFFFFFu;
        else if (p.tm_rpad) { const int t = (int)__umulhi((uint32_t)qi, p.n_magic); orow4[i] = (uint32_t)(t * p.tm_rpad + (u.sh >> 3) * n + (qi - t * n)); }
        else orow4[i] = (uint32_t)(seq_row0 + qi);
      }
      AXVS_PROF_WAIT(1, mbar_wait(&o_full[g], par))
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
          const float il = __frcp_rn(inv[j]);
          const float2 il2 = make_float2(il, il);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 w;
            const float2 m0 = mul_f32x2(make_float2(o[8 * c], o[8 * c + 1]), il2), m1 = mul_f32x2(make_float2(o[8 * c + 2], o[8 * c + 3]), il2);
            const float2 m2 = mul_f32x2(make_float2(o[8 * c + 4], o[8 * c + 5]), il2), m3 = mul_f32x2(make_float2(o[8 * c + 6], o[8 * c + 7]), il2);
            w.x = pack_bf16x2(m0.x, m0.y);
            w.y = pack_bf16x2(m1.x, m1.y);
            w.z = pack_bf16x2(m2.x, m2.y);
            w.w = pack_bf16x2(m3.x, m3.y);
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = w;
          }
          __syncwarp();
          const int f = u.f0 + j;
          uint8_t* dst = p.x_img + (size_t)f * p.tiles * 4 * ATT2_KB;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + (lane >> 2), piece = lane & 3;
            const uint4 w = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
            const int qi = q_w0 + rr;                          // query index inside the sequence
            if (orow4[i] != 0xFFFFFFFFu) {
              const size_t r = orow4[i];
              const size_t off = ((r >> 7) * 4 + kb) * (size_t)ATT2_KB + sw128_offset((uint32_t)(r & 127), ch0 + piece);
              *reinterpret_cast<uint4*>(dst + off) = w;
              if (!p.tm_rpad && (unsigned)(qi - f * n) < (unsigned)n) *reinterpret_cast<uint4*>(p.xd_img + off) = w;   // qi / n == f
            }
          }
          __syncwarp();
        }
      }
    }
    AXVS_PROF_FLUSH(52, 2, warp == 0 && lane == 0)
  } else if (warp == 4 * G) {
    // =============================================================== TMA producer: one (or two) bulk copies per unit
    if (lane == 0) {
      uint32_t slot = 0, sphase = 0;                           // ring position and its phase bit (no divisions on this path)
      AXVS_PROF_DECL(1)
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const AtUnit u = at_decode(unit, p);
        uint8_t* dst = ring + (size_t)slot * p.slot_bytes;
        const __nv_bfloat16* region = p.qkv + (size_t)u.sh * 3 * N * 32;
        AXVS_PROF_WAIT(0, mbar_wait(&empty[slot], sphase ^ 1))
        if (p.single) {
          mbar_arrive_expect_tx(&full[slot], 3 * N * 64);
          tma_bulk_g2s(dst, region, 3 * N * 64, &full[slot]);
        } else {
          const int rows_q = min(128, N - u.qb * 128);
          const int kpos = N + 2 * u.f0 * n;                  // region row of the chunk's first K row
          mbar_arrive_expect_tx(&full[slot], rows_q * 64 + u.fc * 2 * n * 64);
          tma_bulk_g2s(dst, region + (size_t)u.qb * 128 * 32, rows_q * 64, &full[slot]);
          tma_bulk_g2s(dst + AT_Q_BYTES + (kpos & 7) * 64, region + (size_t)kpos * 32, u.fc * 2 * n * 64, &full[slot]);
        }
        if (++slot == (uint32_t)p.slots) { slot = 0; sphase ^= 1; }
      }
      AXVS_PROF_FLUSH(60, 1, true)
    }
  } else {
    // =============================================================== MMA issuers (converged warps, elected lane): warp 4G + 1 issues
    // the S phase of every unit, warp 4G + 2 the P V phase -- the per-instruction issue cost (~150 clk per elected block) is
    // the bottleneck of this kernel, and the two phases are ordered by mbarriers only (S of unit k + G waits for buf_free, which
    // follows o_full of unit k), so they can be issued from two warps in parallel
    const uint32_t idesc_s = umma_idesc_bf16(128, NP);
    const uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);        // B (= V) is MN-major
    const uint32_t ring_addr = smem_u32(ring);
    // shared-memory row of the unit's first K row inside its slot
    auto k_row0 = [&](const AtUnit& u) { return p.single ? N : 128 + ((N + 2 * u.f0 * n) & 7); };
    uint32_t slot = 0, sphase = 0, b = 0, bphase = 0;          // ring slot / TMEM buffer of the current unit and their phase bits
    auto advance = [&]() {
      if (++slot == (uint32_t)p.slots) { slot = 0; sphase ^= 1; }
      if (++b == (uint32_t)G) { b = 0; bphase ^= 1; }
    };
    if (warp == 4 * G + 1) {
      AXVS_PROF_DECL(3)
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const AtUnit u = at_decode(unit, p);
        const uint32_t base = ring_addr + slot * p.slot_bytes;
        const uint32_t qa = ((base & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t k_base = base + k_row0(u) * 64;
        AXVS_PROF_WAIT(0, mbar_wait(&full[slot], sphase))
        AXVS_PROF_WAIT(1, mbar_wait(&buf_free[b], bphase ^ 1))
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < u.fc; ++j) {
          const uint32_t ka = (((k_base + 2 * j * n * 64) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t t_d = tmem + b * p.buf_cols + j * NP;
          const bool last = j == u.fc - 1;
          if (elect_one()) {
            umma_ss_raw(t_d, qa, AT_DESC_HI_SW64, ka, AT_DESC_HI_SW64, idesc_s, 0u);
            umma_ss_raw(t_d, qa + 2, AT_DESC_HI_SW64, ka + 2, AT_DESC_HI_SW64, idesc_s, 1u);
            if (last) umma_commit(&s_full[b]);
          }
          __syncwarp();
        }
        advance();
      }
      AXVS_PROF_FLUSH(56, 3, lane == 0)
    } else {
      AXVS_PROF_DECL(1)
      for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
        const AtUnit u = at_decode(unit, p);
        const uint32_t t_p = tmem + b * p.buf_cols;
        const uint32_t v_base = ring_addr + slot * p.slot_bytes + (k_row0(u) + n) * 64;
        AXVS_PROF_WAIT(0, mbar_wait(&p_full[b], bphase))
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < u.fc; ++j) {
          const uint32_t v_lo = (((v_base + 2 * j * n * 64) & 0x3FFFFu) >> 4) | (64u << 16);
          const uint32_t t_d = t_p + p.o_off + 32 * j, t_a = t_p + j * p.p_stride;
          const bool last = j == u.fc - 1;
          if constexpr (NT16 > 0) {
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < NT16; ++kk) umma_ts_raw(t_d, t_a + 8 * kk, v_lo + kk * 64, AT_DESC_HI_SW64, idesc_o, kk ? 1u : 0u);
              if (last) { umma_commit(&o_full[b]); umma_commit(&empty[slot]); }
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int k0 = 0; k0 < (NP >> 4); k0 += 4) {
              const int rem = (NP >> 4) - k0;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  if (kk < rem) umma_ts_raw(t_d, t_a + 8 * (k0 + kk), v_lo + (k0 + kk) * 64, AT_DESC_HI_SW64, idesc_o, (k0 + kk) ? 1u : 0u);
                if (last && rem <= 4) { umma_commit(&o_full[b]); umma_commit(&empty[slot]); }
              }
              __syncwarp();
            }
          }
        }
        advance();
      }
      AXVS_PROF_FLUSH(36, 1, lane == 0)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 * G + 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
