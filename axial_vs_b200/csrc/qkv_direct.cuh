// q | k | v projections reading the fp32 residual stream directly (fusion level 4): the tile-image pack kernel and its
// 2 KiB/token round trip through HBM disappear.
//
//   A1 = bf16(src[c(row)] + pos[c(row)]),  A2 = bf16(src[c(row)])        c(row) = canonical token of pass-order row `row`
//   [q | k] = A1 [Wq; Wk]^T + b,   v = A2 Wv^T + b                        (the reference's axis permutes, WC/temporal_attention.py:197,206,
//                                                                          exist only as this index map)
//
// Eight converting producer warps read 256-byte row segments (half a warp per segment, software-pipelined over two register
// sets), and write BOTH images of a K-block from the one src load into a 32 KiB "pair slot" (SWIZZLE_128B K-major).  The MMA
// warp moves each image into TENSOR MEMORY with tcgen05.cp (columns [0,128) = A1, [128,256) = A2) and frees the slot at once,
// so the shared-memory ring only decouples producers and tensor pipe -- the whole tile's A operand lives in TMEM, and the
// N = 128 UMMAs read it from there at the full 64 clk rate (A in shared memory: ~90 clk, operand reads saturate the smem port).
// Six 128-column chunks (q heads 0-3, 4-7, k, k, v, v) alternate between two TMEM accumulator stages; BOTH epilogue groups
// drain every chunk (two heads each), which halves the time a stage stays occupied -- the issuer's stage waits would otherwise
// delay the next tile's tcgen05.cp and stall the producers.  The epilogue adds the bias and writes bf16 HEAD-MAJOR
// qkv[which][head][row][32]  through a per-warp transpose.
//
// Warp roles (576 threads, no setmaxnreg): warps 0-7 epilogue, 8-15 A producers, 16 weight TMA, 17 MMA / tcgen05.cp issuer.
#pragma once
#include "qkv_fused.cuh"

namespace axvs {

constexpr int QD_THREADS = 576;
constexpr int QD_PRODUCER_WARPS = 8;
constexpr int QD_A_SLOTS = 3;                 // pair slots: [A1 K-block image | A2 K-block image] = 32 KiB
constexpr int QD_SMEM_BYTES = QD_A_SLOTS * 2 * TF_KB + QK_W_SLOTS * TF_WU + QK_STAGE_BYTES + QK_BIAS_BYTES + 512;
static_assert(QD_SMEM_BYTES <= 232448, "qkv_direct_kernel exceeds the 227 KiB shared-memory limit");

struct QkvDirectParams {
  const float* src;        // fp32 [tokens, 256] canonical order: q = k input (before the positional term) and v input
  const float* pos;        // fp32 [tokens, 256] or null
  const uint8_t* w;        // unit format of [Wq; Wk; Wv]
  const float* bias;       // [768]
  __nv_bfloat16* qkv;      // head-major [3][8][rows][32]
  int rows, tiles, map_mode;
  AxialDims dims;
  int swz_N, swz_n;        // > 0: "unit-major" output for the tcgen05 attention kernel (attn_tc.cuh) instead of head-major: one region of
                           // 3 N rows per (sequence, head) -- Q rows, then K_f | V_f per key frame (N = swz_N tokens per sequence, n = swz_n per
                           // frame) -- with the 16-byte chunks of the row at region position pos permuted by (pos >> 1) & 3
};

__global__ void __launch_bounds__(QD_THREADS, 1) qkv_direct_kernel(const QkvDirectParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QD_A_SLOTS * 2 * TF_KB;
  uint8_t* stage_all = w_ring + QK_W_SLOTS * TF_WU;
  float* sbias = reinterpret_cast<float*>(stage_all + QK_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QD_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QD_A_SLOTS;      // tcgen05.commit after the slot's copies
  uint64_t* w_full = a_empty + QD_A_SLOTS;      // [QK_W_SLOTS]
  uint64_t* w_empty = w_full + QK_W_SLOTS;
  uint64_t* s_full = w_empty + QK_W_SLOTS;      // [2] accumulator stage of group g complete
  uint64_t* s_empty = s_full + 2;               // [2] drained by the 4 warps of group g
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QD_A_SLOTS; ++i) { mbar_init(&a_full[i], QD_PRODUCER_WARPS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QK_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 768; i += QD_THREADS) sbias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue: group g drains heads 2g, 2g+1 of every chunk
    const int g = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* stg = stage_all + warp * 2048;
    uint32_t cnt = 0;                                          // chunks consumed (stage = cnt & 1)
    AXVS_PROF_DECL(1)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int row0 = tile * 128 + (warp & 3) * 32;           // first row of this warp
      uint32_t um_q = 0, um_k = 0, pos_q = 0, pos_k = 0;       // unit-major: rows of q / k of row (row0 + lane) for head 0 (v = k + n), positions inside the region
      if (p.swz_N > 0) {
        const int r = row0 + lane;
        const int seq = r / p.swz_N, i = r - seq * p.swz_N;
        const int f = i / p.swz_n, j = i - f * p.swz_n;
        pos_q = (uint32_t)i;
        pos_k = (uint32_t)(p.swz_N + 2 * f * p.swz_n + j);
        um_q = (uint32_t)seq * 24u * p.swz_N + pos_q;
        um_k = (uint32_t)seq * 24u * p.swz_N + pos_k;
      }
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++cnt) {
        const int st = cnt & 1;                                // 6 chunks per tile: even, so st == rt & 1
        AXVS_PROF_WAIT(0, mbar_wait(&s_full[st], (cnt >> 1) & 1))
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + st * 128;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {                       // one head (32 columns) at a time
          const int c = 2 * g + cc;
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          if (cc == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
          }
          // bias + bf16, then a 2 KiB per-warp transpose so every store instruction writes 512 contiguous bytes
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
            if (p.swz_N > 0) {
              const uint32_t my_row = (which == 0 ? um_q : um_k + (which == 2 ? p.swz_n : 0)) + (uint32_t)head * 3u * p.swz_N;
              const uint32_t my_key = ((which == 0 ? pos_q : pos_k + (which == 2 ? p.swz_n : 0)) >> 1) & 3u;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rl = 8 * i + (lane >> 2), piece = lane & 3;
                const uint4 u = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
                const uint32_t drow = __shfl_sync(0xffffffffu, my_row, rl), key = __shfl_sync(0xffffffffu, my_key, rl);
                if (row0 + rl < p.rows) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.qkv) + (size_t)drow * 64 + ((piece ^ key) << 4)) = u;
              }
            } else {
            uint8_t* dst = reinterpret_cast<uint8_t*>(p.qkv + ((size_t)(which * 8 + head) * p.rows + row0) * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rl = 8 * i + (lane >> 2), piece = lane & 3;
              const uint4 u = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
              if (row0 + rl < p.rows) *reinterpret_cast<uint4*>(dst + rl * 64 + piece * 16) = u;
            }
            }
          }
          __syncwarp();
        }
      }
    }
    AXVS_PROF_FLUSH(44 + 2 * g, 1, (warp & 3) == 0 && lane == 0)
  } else if (warp < 8 + QD_PRODUCER_WARPS) {
    // =============================================================== converting A producers
    // Work unit = batch: 4 of the lane's 8 rows of one K-block (4 src + 4 pos loads).  Two register sets alternate so the
    // loads of batch b+1 (also across K-block and tile boundaries) are in flight while batch b is converted and stored.
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
    auto issue = [&](float4 (&s)[4], float4 (&q)[4], const uint32_t (&cr)[8], int kb, int hf) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = cr[4 * hf + j];
        s[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(p.src + (size_t)c * C256 + kb * 64) + c16) : z;
      }
      if (p.pos) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t c = cr[4 * hf + j];
          q[j] = c != 0xFFFFFFFFu ? __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)pos_row(c, p.dims) * C256 + kb * 64) + c16) : z;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = z;
      }
    };
    auto consume = [&](float4 (&s)[4], float4 (&q)[4], int hf) {
      const uint32_t slot = cnt % QD_A_SLOTS, phase = (cnt / QD_A_SLOTS) & 1;
      if (hf == 0) { AXVS_PROF_WAIT(0, mbar_wait(&a_empty[slot], phase ^ 1)) }
      uint8_t* dst = a_ring + slot * 2 * TF_KB;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = pw * 16 + 2 * (4 * hf + j) + half;
        const uint32_t off = sw128_offset(r, c16 >> 1) + (c16 & 1) * 8;
        uint2 u;
        u.x = pack_bf16x2(s[j].x + q[j].x, s[j].y + q[j].y);
        u.y = pack_bf16x2(s[j].z + q[j].z, s[j].w + q[j].w);
        *reinterpret_cast<uint2*>(dst + off) = u;                                   // A1 K-block image
        u.x = pack_bf16x2(s[j].x, s[j].y);
        u.y = pack_bf16x2(s[j].z, s[j].w);
        *reinterpret_cast<uint2*>(dst + TF_KB + off) = u;                           // A2 K-block image
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
      for (int b = 0; b < 8; ++b) {
        if (b < 7) {
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
    // =============================================================== tcgen05.cp + MMA issuer (converged warp, elected lane)
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;     // ccnt: chunks issued
    AXVS_PROF_DECL(3)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      // the tile's A operand: four pair slots -> TMEM.  These copies are ordered by the tensor pipe behind every UMMA of the
      // previous tile (issued earlier by this thread), which still reads the old contents.
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
        const uint32_t slot = a_cnt % QD_A_SLOTS;
        AXVS_PROF_WAIT(2, mbar_wait(&a_full[slot], (a_cnt / QD_A_SLOTS) & 1))
        tc_fence_after();
        const uint32_t sa = a_ring_addr + slot * 2 * TF_KB;
        if (elect_one()) {
          tmem_cp_kblock(tmem + 32 * kb, sa);
          tmem_cp_kblock(tmem + 128 + 32 * kb, sa + TF_KB);
          umma_commit(&a_empty[slot]);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int rt = 0; rt < 6; ++rt, ++ccnt) {
        const int g = rt & 1;
        const uint32_t gc = ccnt >> 1;                         // chunks already issued to group g (6 per tile: even count)
        AXVS_PROF_WAIT(1, mbar_wait(&s_empty[g], (gc & 1) ^ 1))
        tc_fence_after();
        const uint32_t t_a = tmem + (rt < 4 ? 0 : 128);        // A1 for the q / k chunks, A2 for the v chunks
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QK_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          umma_unit_elect_ts(tmem + 256 + g * 128, t_a + 64 * kg, t_a + 64 * kg + 32, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                             &w_empty[ws], kg == 1 ? &s_full[g] : nullptr, nullptr);
        }
      }
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
