// Fused "temporal half" of TrajectoryAttention (Appendix A steps 6-8 + residual), one persistent tcgen05 kernel:
//
//   q2  = (x_diag Wpq^T + bpq) * scale                     (GEMM 1, accumulator resident in TMEM columns [0,256))
//   for every frame f:  [k2 | v2]_f = x_f Wpkv^T + bpkv     (GEMM 2, 128-column chunks = 2 heads, double-buffered TMEM)
//        a_f = softmax_f(q2 . k2_f)   online over f,  o += a_f v2_f      (epilogue warps, fp32 registers)
//   out[canonical row] = resid + o Wproj^T + bproj          (GEMM 3, A operand = o written to smem by the epilogue)
//
// so k2|v2 (the largest intermediate of the reference, [B', N, F, 2C]), q2 and o never touch HBM.
// Reference: WC/temporal_attention.py:61-75 (+ residual add :204 / :213).  The k2 bias is dropped: it adds the same
// constant q2.bk2 to every frame's logit and softmax over frames is shift-invariant.
//
// Inputs are "tile images": x_f and x_diag stored as [tile][K-block][128 rows x 128 B, SWIZZLE_128B] so that one
// 16 KiB TMA bulk copy lands a ready-to-use UMMA A-operand K-block (written in that layout by the attention kernel).
//
// Warp roles (384 threads): warps 0-3 = epilogue group 0 (heads 0-3, TMEM stage 0), warps 4-7 = epilogue group 1
// (heads 4-7, stage 1), warp 8 = A-tile TMA producer, warp 9 = weight TMA producer, warp 10 = MMA issuer.
#pragma once
#include "gemm.cuh"

namespace axvs {

constexpr int TF_THREADS = 384;
constexpr int TF_A_SLOTS = 4;
constexpr int TF_W_SLOTS = 3;
constexpr int TF_KB = 16384;                       // one K-block tile: 128 rows x 64 bf16
constexpr int TF_WU = 32768;                       // one weight unit: 128 rows x 128 K (two K-blocks), one TMA bulk copy
constexpr int TF_O_BYTES = 4 * TF_KB;              // o operand, 128 x 256 bf16
constexpr int TF_BIAS_BYTES = 2 * 256 * 4;          // b_pq | b_v2 staged in shared memory
constexpr int TF_SMEM_BYTES = TF_O_BYTES + TF_A_SLOTS * TF_KB + TF_W_SLOTS * TF_WU + TF_BIAS_BYTES + 512;

struct TrajParams {
  const uint8_t* x_img;    // [F][tiles][4][16 KiB]
  const uint8_t* xd_img;   // [tiles][4][16 KiB]
  const uint8_t* w_pq;     // unit format (row tile, K group): 4 units
  const uint8_t* w_pkv;    // unit format, rows re-ordered per head pair: chunk c = [k2 heads 2c,2c+1 | v2 heads 2c,2c+1]: 8 units
  const uint8_t* w_proj;   // unit format: 4 units
  const float* b_pq;
  const float* b_v2;       // proj_kv.bias[256:512]
  const float* b_proj;
  const float* resid;      // fp32 canonical, may be null
  float* out;              // fp32 canonical
  int rows, tiles, F;
  int map_mode;
  AxialDims dims;
  float scale_log2e;       // head_dim^-0.5 * log2(e): logits are produced directly in the exp2 domain
  // optional fused LayerNorm of the output row (norm1 of the layer, WC/temporal_attention.py:217): when ln_g != null the
  // kernel writes out = LN(resid + proj(o) + b) as fp32 rows AND as the bf16 tile image the FFN kernel consumes
  const float* ln_g;
  const float* ln_b;
  uint8_t* ln_img;         // [ceil(rows/128)][4][16 KiB], indexed by CANONICAL row
  float ln_eps;
  // Frame-major row order (tm_rpad > 0; traj_ts / traj_pair only): tile-order row r = t * tm_rpad + seq * tm_n + j, i.e. all tokens of
  // frame 0 first (padded to a multiple of 128 rows), then frame 1, ...  The temporal stage is row-wise, so any order is valid -- and in
  // this one every tile has ONE frame index t, whose x_t tile IS the tile's x_diag: the attention kernel writes no x_diag image
  // (0.5 KiB per token less to write and to read back), `xd_img` is unused, rows = F * tm_rpad, tiles = rows / 128.
  int tm_rpad;             // rows per frame group, multiple of 128 (0 = pass order)
  int tm_rt;               // valid rows per frame group = sequences * tm_n
  int tm_n;                // tokens of one frame in a sequence
};

// canonical token of tile-order row r, -1 for a padding row / a row past the end
__device__ __forceinline__ int traj_row_canonical(int r, const TrajParams& p) {
  if (p.tm_rpad == 0) return r < p.rows ? pass_to_canonical(r, p.map_mode, p.dims) : -1;
  const int t = r / p.tm_rpad, rem = r - t * p.tm_rpad;
  if (t >= p.F || rem >= p.tm_rt) return -1;
  const int seq = rem / p.tm_n, j = rem - seq * p.tm_n;
  return pass_to_canonical((seq * p.F + t) * p.tm_n + j, p.map_mode, p.dims);
}
// the four K-block images of a tile's x_diag operand
__device__ __forceinline__ const uint8_t* traj_xd_tile(int tile, const TrajParams& p) {
  if (p.tm_rpad == 0) return p.xd_img + (size_t)tile * 4 * TF_KB;
  const int t = tile / (p.tm_rpad >> 7);
  return p.x_img + ((size_t)t * p.tiles + tile) * 4 * TF_KB;
}

__global__ void __launch_bounds__(TF_THREADS, 1) traj_fused_kernel(const TrajParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need a 1024 B aligned base (no static smem in this kernel)
  uint8_t* o_buf = smem;
  uint8_t* a_ring = smem + TF_O_BYTES;
  uint8_t* w_ring = a_ring + TF_A_SLOTS * TF_KB;
  float* sb_pq = reinterpret_cast<float*>(w_ring + TF_W_SLOTS * TF_WU);
  float* sb_v2 = sb_pq + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb_v2 + 256);
  uint64_t* a_full = bars;                       // [TF_A_SLOTS]
  uint64_t* a_empty = a_full + TF_A_SLOTS;       // [TF_A_SLOTS]
  uint64_t* w_full = a_empty + TF_A_SLOTS;       // [TF_W_SLOTS]
  uint64_t* w_empty = w_full + TF_W_SLOTS;       // [TF_W_SLOTS]
  uint64_t* s_full = w_empty + TF_W_SLOTS;       // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* q2_full = s_empty + 2;
  uint64_t* q2_free = q2_full + 1;
  uint64_t* o_ready = q2_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TF_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TF_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    mbar_init(q2_full, 1);
    mbar_init(q2_free, 8);
    mbar_init(o_ready, 8);
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 256; i += TF_THREADS) { sb_pq[i] = p.b_pq[i]; sb_v2[i] = p.b_v2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int F = p.F;

  if (warp < 8) {
    // =============================================================== epilogue groups
    setmaxnreg_inc<224>();   // 256*224 + 128*56 = 64512 = the CTA register pool at launch (384 x 168)
    const int g = warp >> 2;                                     // group = TMEM stage = head quad
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_q2 = tmem + lane_base + 128 * g;            // my 4 heads of q2
    const uint32_t t_s = tmem + lane_base + 256 + 128 * g;       // my accumulator stage
    uint32_t s_cnt = 0;                                          // items consumed on my stage
    uint32_t it = 0;                                             // tile iteration
    AXVS_PROF_DECL(3)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // ---- finalise q2: (acc + bias) * scale*log2e, written back to TMEM
      AXVS_PROF_WAIT(0, mbar_wait(q2_full, it & 1))
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float v[32];
        tmem_ld32(t_q2 + 32 * j, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] + sb_pq[128 * g + 32 * j + i]) * p.scale_log2e;
        tmem_st32(t_q2 + 32 * j, v);
      }
      tmem_st_wait();

      float m_run[4], l_run[4], o[4][32];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        m_run[h] = -INFINITY;
        l_run[h] = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h][i] = 0.f;
      }
      // ---- frames: two 128-column chunks (2 heads each) per frame on my stage
      for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          AXVS_PROF_WAIT(1, mbar_wait(&s_full[g], s_cnt & 1))
          ++s_cnt;
          tc_fence_after();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int lh = cc * 2 + hh;
            float s = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float k2[16], q2[16];
              tmem_ld16(t_s + 32 * hh + 16 * hf, k2);
              tmem_ld16(t_q2 + 32 * lh + 16 * hf, q2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) s = fmaf(q2[i], k2[i], s);
            }
            const float mn = fmaxf(m_run[lh], s);
            const float corr = exp2f(m_run[lh] - mn);
            const float pe = exp2f(s - mn);
            l_run[lh] = l_run[lh] * corr + pe;
            m_run[lh] = mn;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float v2[16];
              tmem_ld16(t_s + 64 + 32 * hh + 16 * hf, v2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[lh][16 * hf + i] = fmaf(pe, v2[i], o[lh][16 * hf + i] * corr);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[g]);
        }
      }
      // ---- o = o / l + bv2  -> bf16 A operand of the output projection (K-blocks 2g, 2g+1 of o_buf)
#pragma unroll
      for (int lh = 0; lh < 4; ++lh) {
        const float inv = 1.f / l_run[lh];
        const int col0 = 128 * g + 32 * lh;
        uint8_t* kb_base = o_buf + (col0 >> 6) * TF_KB;
        const int chunk0 = (col0 & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) t[i] = fmaf(o[lh][8 * q + i], inv, sb_v2[col0 + 8 * q + i]);
          uint4 u;
          u.x = pack_bf16x2(t[0], t[1]); u.y = pack_bf16x2(t[2], t[3]);
          u.z = pack_bf16x2(t[4], t[5]); u.w = pack_bf16x2(t[6], t[7]);
          *reinterpret_cast<uint4*>(kb_base + sw128_offset(row_in_tile, chunk0 + q)) = u;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(o_ready);
        mbar_arrive(q2_free);
      }
      // ---- output projection item on my stage: out = resid + acc + bproj (my 128 output columns).
      // TMEM rows are one-per-thread; a 4 KiB per-warp transpose through shared memory (o_buf is idle here: the
      // projection MMAs that read it have retired) turns the global accesses into full 128-byte row segments.
      // resid + bias are fetched BEFORE waiting for the accumulator, so their latency hides behind the projection GEMM.
      {
        const int r = tile * 128 + row_in_tile;
        const int my_orow = (r < p.rows) ? pass_to_canonical(r, p.map_mode, p.dims) : -1;
        uint8_t* stg = o_buf + warp * 4096;
        const int sub = lane >> 3, piece = lane & 7;
        int orow[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) orow[i] = __shfl_sync(0xffffffffu, my_orow, i * 4 + sub);
        float4 rr[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 128 * g + 32 * j + piece * 4;
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b_proj + col));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 x = (p.resid && orow[i] >= 0) ? __ldg(reinterpret_cast<const float4*>(p.resid + (size_t)orow[i] * 256 + col))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            rr[j][i] = make_float4(x.x + bb.x, x.y + bb.y, x.z + bb.z, x.w + bb.w);
          }
        }
        AXVS_PROF_WAIT(2, mbar_wait(&s_full[g], s_cnt & 1))
        ++s_cnt;
        tc_fence_after();
        if (p.ln_g == nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
            __syncwarp();
            const int col = 128 * g + 32 * j + piece * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub;
              const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
              if (orow[i] >= 0)
                *reinterpret_cast<float4*>(p.out + (size_t)orow[i] * 256 + col) =
                    make_float4(a.x + rr[j][i].x, a.y + rr[j][i].y, a.z + rr[j][i].z, a.w + rr[j][i].w);
            }
            __syncwarp();
          }
        } else {
          // ---- fused LayerNorm: keep the row values in registers, combine the statistics of the two column halves through
          // shared memory (upper half of o_buf), then write fp32 rows + the bf16 tile image of the normalised row
          float ps[8], pq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rl = i * 4 + sub;
              const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
              float4 tv = rr[j][i];
              tv.x += a.x; tv.y += a.y; tv.z += a.z; tv.w += a.w;
              rr[j][i] = tv;
              ps[i] += tv.x + tv.y + tv.z + tv.w;
              pq[i] += tv.x * tv.x + tv.y * tv.y + tv.z * tv.z + tv.w * tv.w;
            }
            __syncwarp();
          }
          float2* xc = reinterpret_cast<float2*>(o_buf + 32768) + (it & 1) * 256;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
              pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], o);
            }
            if (piece == 0) xc[g * 128 + (warp & 3) * 32 + i * 4 + sub] = make_float2(ps[i], pq[i]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          float mean[8], rstd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 other = xc[(g ^ 1) * 128 + (warp & 3) * 32 + i * 4 + sub];
            mean[i] = (ps[i] + other.x) * (1.f / 256.f);
            const float var = fmaxf((pq[i] + other.y) * (1.f / 256.f) - mean[i] * mean[i], 0.f);
            rstd[i] = rsqrtf(var + p.ln_eps);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = 128 * g + 32 * j + piece * 4;
            const float4 gg = __ldg(reinterpret_cast<const float4*>(p.ln_g + col)), be = __ldg(reinterpret_cast<const float4*>(p.ln_b + col));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 tv = rr[j][i];
              const float4 y = make_float4((tv.x - mean[i]) * rstd[i] * gg.x + be.x, (tv.y - mean[i]) * rstd[i] * gg.y + be.y,
                                           (tv.z - mean[i]) * rstd[i] * gg.z + be.z, (tv.w - mean[i]) * rstd[i] * gg.w + be.w);
              // 16-byte image chunk = 8 channels = this lane (even piece) + its odd neighbour
              const uint32_t lo = pack_bf16x2(y.x, y.y), hi = pack_bf16x2(y.z, y.w);
              const uint32_t nlo = __shfl_down_sync(0xffffffffu, lo, 1), nhi = __shfl_down_sync(0xffffffffu, hi, 1);
              if (orow[i] >= 0) {
                *reinterpret_cast<float4*>(p.out + (size_t)orow[i] * 256 + col) = y;
                if ((piece & 1) == 0) {
                  const uint32_t orw = (uint32_t)orow[i];
                  *reinterpret_cast<uint4*>(p.ln_img + ((size_t)(orw >> 7) * 4 + (col >> 6)) * TF_KB + sw128_offset(orw & 127u, (col & 63) >> 3)) =
                      make_uint4(lo, hi, nlo, nhi);
                }
              }
            }
          }
        }
      }
      // released only now: the staging area aliases o_buf, which the other group refills once this stage's next
      // chunks have been drained (see the ordering argument in DESIGN.md)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);
    }
    AXVS_PROF_FLUSH(8 + 8 * g, 3, (warp & 3) == 0 && lane == 0)
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer (x_diag, x_0 .. x_{F-1})
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int item = 0; item < 4 * (F + 1); ++item, ++cnt) {
          const uint32_t slot = cnt % TF_A_SLOTS, phase = (cnt / TF_A_SLOTS) & 1;
          const uint8_t* src = (item < 4) ? p.xd_img + ((size_t)tile * 4 + item) * TF_KB
                                          : p.x_img + (((size_t)((item >> 2) - 1) * p.tiles + tile) * 4 + (item & 3)) * TF_KB;
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, src, TF_KB, &a_full[slot]);
        }
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer (32 KiB units)
      uint32_t slot = 0, phase = 0;
      AXVS_PROF_DECL(1)
      auto push = [&](const uint8_t* img, int unit) {
        AXVS_PROF_WAIT(0, mbar_wait(&w_empty[slot], phase ^ 1))
        mbar_arrive_expect_tx(&w_full[slot], TF_WU);
        tma_bulk_g2s(w_ring + slot * TF_WU, img + (size_t)unit * TF_WU, TF_WU, &w_full[slot]);
        if (++slot == TF_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int u = 0; u < 4; ++u) push(p.w_pq, u);               // (half, kg) = (u >> 1, u & 1)
#pragma unroll 1
        for (int i = 0; i < 8 * F; ++i) {
          const int ci = (i >> 1) & 3;
          const int c = ((ci & 1) << 1) | (ci >> 1);               // chunk order 0,2,1,3: stages alternate
          push(p.w_pkv, c * 2 + (i & 1));
        }
#pragma unroll 1
        for (int u = 0; u < 4; ++u) push(p.w_proj, u);
      }
      AXVS_PROF_FLUSH(32, 1, true)
    } else if (warp == 10) {
      // =============================================================== MMA issuer (converged warp, elected lane issues)
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring), o_addr = smem_u32(o_buf);
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, s_cnt0 = 0, s_cnt1 = 0, it = 0;
      AXVS_PROF_DECL(5)
      auto w_wait = [&]() -> uint32_t {
        AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == TF_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        // ---- GEMM 1: q2 accumulators (columns [0,256)); the 4 x_diag K-blocks stay resident for both column halves
        AXVS_PROF_WAIT(4, mbar_wait(q2_free, (it & 1) ^ 1))
        tc_fence_after();
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u >> 1, kg = u & 1;
          const uint32_t ac0 = a_cnt + 2 * kg, ac1 = ac0 + 1;
          const uint32_t s0 = ac0 % TF_A_SLOTS, s1 = ac1 % TF_A_SLOTS;
          if (half == 0) {
            AXVS_PROF_WAIT(2, mbar_wait(&a_full[s0], (ac0 / TF_A_SLOTS) & 1); mbar_wait(&a_full[s1], (ac1 / TF_A_SLOTS) & 1))
            tc_fence_after();
          }
          const uint32_t ws = w_wait();
          umma_unit_elect(tmem + half * 128, a_ring_addr + s0 * TF_KB, a_ring_addr + s1 * TF_KB, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                          &w_empty[ws], half ? &a_empty[s0] : nullptr, half ? &a_empty[s1] : nullptr, u == 3 ? q2_full : nullptr);
        }
        a_cnt += 4;
        // ---- GEMM 2: per frame, four 128-column chunks alternating between the two TMEM stages
#pragma unroll 1
        for (int f = 0; f < F; ++f) {
#pragma unroll 1
          for (int ci = 0; ci < 4; ++ci) {
            const int g = ci & 1;                               // chunk order 0,2,1,3 -> stage 0,1,0,1
            const uint32_t sc = g ? s_cnt1 : s_cnt0;
            AXVS_PROF_WAIT(1, mbar_wait(&s_empty[g], (sc & 1) ^ 1))
            if (g) ++s_cnt1; else ++s_cnt0;
            tc_fence_after();
#pragma unroll 1
            for (int kg = 0; kg < 2; ++kg) {
              const uint32_t ac0 = a_cnt + 2 * kg, ac1 = ac0 + 1;
              const uint32_t s0 = ac0 % TF_A_SLOTS, s1 = ac1 % TF_A_SLOTS;
              if (ci == 0) {
                AXVS_PROF_WAIT(2, mbar_wait(&a_full[s0], (ac0 / TF_A_SLOTS) & 1); mbar_wait(&a_full[s1], (ac1 / TF_A_SLOTS) & 1))
                tc_fence_after();
              }
              const uint32_t ws = w_wait();
              umma_unit_elect(tmem + 256 + g * 128, a_ring_addr + s0 * TF_KB, a_ring_addr + s1 * TF_KB, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                              &w_empty[ws], ci == 3 ? &a_empty[s0] : nullptr, ci == 3 ? &a_empty[s1] : nullptr, kg == 1 ? &s_full[g] : nullptr);
            }
          }
          a_cnt += 4;
        }
        // ---- GEMM 3: output projection, A = o (written by the epilogue), accumulators = both stages
        AXVS_PROF_WAIT(3, mbar_wait(o_ready, it & 1))
        AXVS_PROF_WAIT(1, mbar_wait(&s_empty[0], (s_cnt0 & 1) ^ 1))
        ++s_cnt0;
        AXVS_PROF_WAIT(1, mbar_wait(&s_empty[1], (s_cnt1 & 1) ^ 1))
        ++s_cnt1;
        tc_fence_after();
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u >> 1, kg = u & 1;
          const uint32_t ws = w_wait();
          // both stages are published only after the LAST read of o_buf: the epilogue reuses o_buf as transpose staging
          umma_unit_elect(tmem + 256 + half * 128, o_addr + (2 * kg) * TF_KB, o_addr + (2 * kg + 1) * TF_KB, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                          &w_empty[ws], u == 3 ? &s_full[0] : nullptr, u == 3 ? &s_full[1] : nullptr, nullptr);
        }
      }
      AXVS_PROF_FLUSH(0, 5, lane == 0)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// x [rows, F, 256] bf16 (row-major, v1 attention output) -> tile images for traj_fused_kernel (test / bridge path)
__global__ void x_to_image_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ x_img, uint8_t* __restrict__ xd_img,
                                  int rows, int tiles, int F, int N, int n) {
  const size_t total = (size_t)rows * F * 32;   // 16-byte chunks
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int ch = (int)(idx & 31);
    const size_t rf = idx >> 5;
    const int f = (int)(rf % F);
    const int r = (int)(rf / F);
    const uint4 v = *reinterpret_cast<const uint4*>(x + (rf * 256) + ch * 8);
    const int tile = r >> 7, rr = r & 127, kb = ch >> 3, c8 = ch & 7;
    const size_t off = (size_t)kb * TF_KB + sw128_offset(rr, c8);
    *reinterpret_cast<uint4*>(x_img + ((size_t)f * tiles + tile) * 4 * TF_KB + off) = v;
    if (f == (r % N) / n) *reinterpret_cast<uint4*>(xd_img + (size_t)tile * 4 * TF_KB + off) = v;
  }
}

}  // namespace axvs
