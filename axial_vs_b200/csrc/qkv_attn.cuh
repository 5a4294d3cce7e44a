// Fused front end of TrajectoryAttention (fusion level 5): q | k | v projections AND the per-frame-softmax spatial attention
// in one persistent kernel, so q, k and v (1.5 KiB/token written + 1.5 KiB/token re-read by a separate attention kernel)
// never touch HBM.  Reference: WC/temporal_attention.py:42-60 (q/k/v Linear, per-frame softmax over the n keys of each
// frame, x = A V) with the layer's `src + pos` (:200) and axis permutes (:197,206) folded into the loads.
//
// A tile = S = floor(128 / N) whole sequences (N = F*n tokens each), i.e. S*N <= 128 pass-order rows:
//   producers (8 warps)  : as qkv_direct_kernel -- fp32 src (+pos) rows -> bf16 SWIZZLE_128B K-block images A1 | A2
//   warp 17              : tcgen05.cp images -> TMEM (A1 = columns [0,128), A2 = [128,256)); six 128-column chunks in the
//                          order q,k,v of heads 0-3 then q,k,v of heads 4-7, alternating between two accumulator stages;
//                          UMMAs read A from tensor memory (N = 128 at the full 64 clk rate)
//   warps 0-7            : drain each chunk (+bias, bf16) into a shared-memory q|k|v buffer of the head group laid out for
//                          ldmatrix, then run the attention of that head group with mma.sync.m16n8k16: work item =
//                          (sequence, head, 16-query block, key frame): S = Q K_f^T, one-shot softmax, P V_f, and write the
//                          x_f / x_diag tile images traj_ts_kernel consumes (16-byte pieces of its SWIZZLE_128B rows)
//   warp 16              : weight units by TMA
// The next tile's loads, copies and first two chunks overlap the attention of the current tile's second head group.
//
// Sequences longer than a tile (N > 128) or frames longer than 64 tokens use the unfused kernels (qkv_direct + attention v2).
#pragma once
#include "qkv_direct.cuh"

namespace axvs {

constexpr int QA_THREADS = 576;
constexpr int QA_PRODUCER_WARPS = 8;
constexpr int QA_A_SLOTS = 3;                      // single K-block images (16 KiB)
constexpr int QA_W_SLOTS = 2;
constexpr int QA_HEAD_BYTES = 128 * 64;            // one head of one operand: 128 rows x 32 bf16
constexpr int QA_QKV_BYTES = 3 * 4 * QA_HEAD_BYTES;   // q | k | v of one head group
constexpr int QA_STG_BYTES = 8 * 1024;             // per-warp 16 rows x 64 B staging of the image writes
constexpr int QA_SMEM_BYTES = QA_A_SLOTS * TF_KB + QA_W_SLOTS * TF_WU + QA_QKV_BYTES + QA_STG_BYTES + QK_BIAS_BYTES + 512;
static_assert(QA_SMEM_BYTES <= 232448, "qkv_attn_kernel exceeds the 227 KiB shared-memory limit");

struct QkvAttnParams {
  const float* src;        // fp32 [tokens, 256] canonical order
  const float* pos;        // fp32 [tokens, 256] or null
  const uint8_t* w;        // unit format of [Wq; Wk; Wv]
  const float* bias;       // [768]
  uint8_t* x_img;          // [F][img_tiles][4][16 KiB]
  uint8_t* xd_img;         // [img_tiles][4][16 KiB]
  int rows, num_seq, N, n, F, S, tiles, img_tiles, map_mode;
  AxialDims dims;
  float scale_log2e;
};

template <int NT16>
__global__ void __launch_bounds__(QA_THREADS, 1) qkv_attn_kernel(const QkvAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* a_ring = smem;
  uint8_t* w_ring = a_ring + QA_A_SLOTS * TF_KB;
  uint8_t* qkv_buf = w_ring + QA_W_SLOTS * TF_WU;      // [which 3][head 4][128 rows][64 B], rows swizzled with att_off
  uint8_t* stg_all = qkv_buf + QA_QKV_BYTES;
  float* sbias = reinterpret_cast<float*>(stg_all + QA_STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 768);
  uint64_t* a_full = bars;                      // [QA_A_SLOTS], one arrive per producer warp
  uint64_t* a_empty = a_full + QA_A_SLOTS;      // tcgen05.commit after the slot's copy
  uint64_t* w_full = a_empty + QA_A_SLOTS;      // [QA_W_SLOTS]
  uint64_t* w_empty = w_full + QA_W_SLOTS;
  uint64_t* s_full = w_empty + QA_W_SLOTS;      // [2]
  uint64_t* s_empty = s_full + 2;               // [2], all 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rows_per_tile = p.S * p.N;

  if (threadIdx.x == 0) {
    for (int i = 0; i < QA_A_SLOTS; ++i) { mbar_init(&a_full[i], QA_PRODUCER_WARPS); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < QA_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 768; i += QA_THREADS) sbias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== chunk drains + attention
    const int g = warp >> 2;                                   // drains heads 2g, 2g+1 of every chunk
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int my_row = (warp & 3) * 32 + lane;                 // tile-local row of this thread in the drains
    uint8_t* stg = stg_all + warp * 1024;
    const int n_mblk = (p.N + 15) >> 4;
    const int gq = lane >> 2, t4 = lane & 3;
    uint32_t cnt = 0;                                          // chunks consumed (stage = cnt & 1)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int seqs = min(p.S, p.num_seq - tile * p.S);       // sequences of this tile
      const size_t tile_row0 = (size_t)tile * rows_per_tile;   // first pass-order row of the tile
#pragma unroll 1
      for (int hg = 0; hg < 2; ++hg) {
        // ---- q, k, v of heads 4hg..4hg+3 -> qkv_buf
#pragma unroll 1
        for (int which = 0; which < 3; ++which, ++cnt) {
          const int st = cnt & 1;
          mbar_wait(&s_full[st], (cnt >> 1) & 1);
          tc_fence_after();
          const uint32_t t_s = tmem + lane_base + 256 + st * 128;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int hl = 2 * g + cc;                         // head inside the group
            float v[32];
            tmem_ld32(t_s + 32 * hl, v);
            tmem_ld_wait();
            if (cc == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_empty[st]);
            }
            const float4* b4 = reinterpret_cast<const float4*>(sbias + which * 256 + (hg * 4 + hl) * 32);
            uint8_t* dst = qkv_buf + (which * 4 + hl) * QA_HEAD_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b0 = b4[2 * q], b1 = b4[2 * q + 1];
              uint4 u;
              u.x = pack_bf16x2(v[8 * q] + b0.x, v[8 * q + 1] + b0.y);
              u.y = pack_bf16x2(v[8 * q + 2] + b0.z, v[8 * q + 3] + b0.w);
              u.z = pack_bf16x2(v[8 * q + 4] + b1.x, v[8 * q + 5] + b1.y);
              u.w = pack_bf16x2(v[8 * q + 6] + b1.z, v[8 * q + 7] + b1.w);
              *reinterpret_cast<uint4*>(dst + att_off(my_row, q)) = u;
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");          // the head group's q | k | v are complete
        // ---- attention of the head group: items (sequence, head, 16-query block, key frame) round-robin over the 8 warps
        {
          const int per_seq = 4 * n_mblk * p.F;
          const int items = seqs * per_seq;
#pragma unroll 1
          for (int item = warp; item < items; item += 8) {
            int r0 = item;
            const int s = r0 / per_seq; r0 -= s * per_seq;
            const int hl = r0 / (n_mblk * p.F); r0 -= hl * (n_mblk * p.F);
            const int mb = r0 / p.F, f = r0 - mb * p.F;
            const uint8_t* sQ = qkv_buf + (0 * 4 + hl) * QA_HEAD_BYTES;
            const uint8_t* sK = qkv_buf + (1 * 4 + hl) * QA_HEAD_BYTES;
            const uint8_t* sV = qkv_buf + (2 * 4 + hl) * QA_HEAD_BYTES;
            const int seq_l0 = s * p.N;                          // tile-local first row of the sequence
            const int key_l0 = seq_l0 + f * p.n;                 // tile-local first key row of the frame
            uint32_t qa[2][4];
            {
              const int r = min(seq_l0 + mb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, 127);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(qa[ks], sQ + att_off(r, ks * 2 + (lane >> 4)));
            }
            float sc[2 * NT16][4];
#pragma unroll
            for (int j = 0; j < 2 * NT16; ++j) {
              sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
              uint32_t kf[4];
              ldmatrix_x4(kf, sK + att_off(min(key_l0 + j * 8 + (lane & 7), 127), lane >> 3));
              mma_bf16_16816(sc[j], qa[0], kf[0], kf[1]);
              mma_bf16_16816(sc[j], qa[1], kf[2], kf[3]);
            }
            float mx[2] = {-INFINITY, -INFINITY};          // same arithmetic as spatial_attn_v2_kernel (bit-identical results)
#pragma unroll
            for (int j = 0; j < 2 * NT16; ++j) {
              if (j * 8 + 8 > p.n) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (j * 8 + t4 * 2 + (e & 1) >= p.n) sc[j][e] = -INFINITY;
              }
              mx[0] = fmaxf(mx[0], fmaxf(sc[j][0], sc[j][1]));
              mx[1] = fmaxf(mx[1], fmaxf(sc[j][2], sc[j][3]));
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
              mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
              mx[h] *= -p.scale_log2e;
            }
            float rs[2] = {0.f, 0.f};
            uint32_t pa[NT16][4];
#pragma unroll
            for (int j = 0; j < 2 * NT16; ++j) {
              const float p0 = ex2_approx(fmaf(sc[j][0], p.scale_log2e, mx[0])), p1 = ex2_approx(fmaf(sc[j][1], p.scale_log2e, mx[0]));
              const float p2 = ex2_approx(fmaf(sc[j][2], p.scale_log2e, mx[1])), p3 = ex2_approx(fmaf(sc[j][3], p.scale_log2e, mx[1]));
              rs[0] += p0 + p1;
              rs[1] += p2 + p3;
              pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
              pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
            }
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < NT16; ++ks) {
#pragma unroll
              for (int jn = 0; jn < 4; jn += 2) {
                uint32_t vb[4];
                const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                // keys past the frame carry probability 0; their V rows (next frame / sequence, or the clamped last row) are finite
                ldmatrix_x4_trans(vb, sV + att_off(min(key_l0 + key, 127), jn + (lane >> 4)));
                mma_bf16_16816(acc[jn], pa[ks], vb[0], vb[1]);
                mma_bf16_16816(acc[jn + 1], pa[ks], vb[2], vb[3]);
              }
            }
            // normalise -> per-warp staging (16 rows x 64 B) -> 16-byte stores into the tile images
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float l = rs[h];
              l += __shfl_xor_sync(0xffffffffu, l, 1);
              l += __shfl_xor_sync(0xffffffffu, l, 2);
              const float inv = __frcp_rn(l);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint32_t*>(stg + att_off(gq + h * 8, j) + t4 * 4) = pack_bf16x2(acc[j][h * 2] * inv, acc[j][h * 2 + 1] * inv);
            }
            __syncwarp();
            const int head = hg * 4 + hl, kb = head >> 1, ch0 = (head & 1) * 4;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int idx = lane + 32 * i;
              const int row = idx >> 2, chunk = idx & 3;
              const int qi = mb * 16 + row;                     // query index inside the sequence
              if (qi < p.N) {
                const uint4 v = *reinterpret_cast<const uint4*>(stg + att_off(row, chunk));
                const size_t r = tile_row0 + seq_l0 + qi;
                const size_t off = ((r >> 7) * 4 + kb) * ATT2_KB + sw128_offset((uint32_t)(r & 127), ch0 + chunk);
                *reinterpret_cast<uint4*>(p.x_img + (size_t)f * p.img_tiles * 4 * ATT2_KB + off) = v;
                if ((unsigned)(qi - f * p.n) < (unsigned)p.n) *reinterpret_cast<uint4*>(p.xd_img + off) = v;
              }
            }
            __syncwarp();
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");          // qkv_buf may be overwritten by the next head group
      }
    }
  } else if (warp < 8 + QA_PRODUCER_WARPS) {
    // =============================================================== converting A producers (see qkv_direct_kernel)
    const int pw = warp - 8;
    const int half = lane >> 4, c16 = lane & 15;
    uint32_t cnt = 0;                                          // K-blocks produced: images 2 cnt (A1) and 2 cnt + 1 (A2)
    float4 sv[2][4], qv[2][4];
    uint32_t crow[8], crow_n[8];
    auto rows_of_tile = [&](int tile, uint32_t (&cr)[8]) {
      const int valid = min(p.S, p.num_seq - tile * p.S) * p.N;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int l = pw * 16 + 2 * j + half;
        cr[j] = l < valid ? (uint32_t)pass_to_canonical(tile * rows_per_tile + l, p.map_mode, p.dims) : 0xFFFFFFFFu;
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
      const uint32_t i0 = 2 * cnt, i1 = i0 + 1;
      const uint32_t s0 = i0 % QA_A_SLOTS, s1 = i1 % QA_A_SLOTS;
      if (hf == 0) {
        mbar_wait(&a_empty[s0], ((i0 / QA_A_SLOTS) & 1) ^ 1);
        mbar_wait(&a_empty[s1], ((i1 / QA_A_SLOTS) & 1) ^ 1);
      }
      uint8_t* d1 = a_ring + s0 * TF_KB;
      uint8_t* d2 = a_ring + s1 * TF_KB;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = pw * 16 + 2 * (4 * hf + j) + half;
        const uint32_t off = sw128_offset(r, c16 >> 1) + (c16 & 1) * 8;
        uint2 u;
        u.x = pack_bf16x2(s[j].x + q[j].x, s[j].y + q[j].y);
        u.y = pack_bf16x2(s[j].z + q[j].z, s[j].w + q[j].w);
        *reinterpret_cast<uint2*>(d1 + off) = u;                                    // A1 K-block image
        u.x = pack_bf16x2(s[j].x, s[j].y);
        u.y = pack_bf16x2(s[j].z, s[j].w);
        *reinterpret_cast<uint2*>(d2 + off) = u;                                    // A2 K-block image
      }
      if (hf == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_full[s0]); mbar_arrive(&a_full[s1]); }
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
  } else if (warp == 16 && lane == 0) {
    // =============================================================== weight producer: chunk order q,k,v (heads 0-3), q,k,v (4-7)
    uint32_t slot = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int ci = 0; ci < 6; ++ci) {
        const int rt = (ci % 3) * 2 + ci / 3;                  // row tile of [Wq; Wk; Wv]: which * 2 + head group
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          mbar_wait(&w_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&w_full[slot], TF_WU);
          tma_bulk_g2s(w_ring + slot * TF_WU, p.w + (size_t)(2 * rt + kg) * TF_WU, TF_WU, &w_full[slot]);
          if (++slot == QA_W_SLOTS) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 17) {
    // =============================================================== tcgen05.cp + MMA issuer (converged warp, elected lane)
    const uint32_t idesc = umma_idesc_bf16(128, 128);
    const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
    uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, ccnt = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      // the tile's A operand -> TMEM; ordered by the tensor pipe behind the UMMAs of the previous tile
#pragma unroll 1
      for (int i = 0; i < 8; ++i, ++a_cnt) {
        const uint32_t slot = a_cnt % QA_A_SLOTS;
        mbar_wait(&a_full[slot], (a_cnt / QA_A_SLOTS) & 1);
        tc_fence_after();
        if (elect_one()) {
          tmem_cp_kblock(tmem + (i & 1) * 128 + 32 * (i >> 1), a_ring_addr + slot * TF_KB);
          umma_commit(&a_empty[slot]);
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int ci = 0; ci < 6; ++ci, ++ccnt) {
        const int st = ccnt & 1;
        const uint32_t gc = ccnt >> 1;                         // fills already issued on this stage
        mbar_wait(&s_empty[st], (gc & 1) ^ 1);
        tc_fence_after();
        const uint32_t t_a = tmem + ((ci % 3) < 2 ? 0 : 128);  // A1 for q / k, A2 for v
#pragma unroll 1
        for (int kg = 0; kg < 2; ++kg) {
          mbar_wait(&w_full[w_slot], w_phase);
          tc_fence_after();
          const uint32_t ws = w_slot;
          if (++w_slot == QA_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
          umma_unit_elect_ts(tmem + 256 + st * 128, t_a + 64 * kg, t_a + 64 * kg + 32, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                             &w_empty[ws], kg == 1 ? &s_full[st] : nullptr, nullptr);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
