// Fused "temporal half" of TrajectoryAttention, tensor-memory-operand version (fusion level 4).  Same math and interface
// as traj_fused_kernel (traj_fused.cuh; reference WC/temporal_attention.py:61-75 + residual :204/:213 + norm1 :217), but
// every UMMA reads its A operand from TENSOR MEMORY, which lets the N = 128 instructions run at the full 64 clk rate
// (A in shared memory: ~90 clk, the operand reads saturate the shared-memory port) and removes every shared-memory
// hand-off between the epilogue and the tensor pipe:
//
//   TMEM columns [  0,128)  Q2P : q2 = (x_diag Wpq^T + bpq) * scale as bf16 pairs, later overwritten IN PLACE by o (A of GEMM 3)
//                [128,256)  XA  : the current A tile (x_diag, then x_f per frame), copied from the TMA-landed K-block
//                                 images with tcgen05.cp; the shared-memory slot is released as soon as the copy retires
//                [256,512)  two 128-column accumulator stages (stage g <-> epilogue group g <-> heads 4g..4g+3)
//
//   GEMM 1: q2 halves -> stages, finalised (bias, scale, bf16) into Q2P by the epilogue
//   GEMM 2: per frame four [k2 | v2] chunks of two heads; online softmax over frames and o accumulation in registers
//           (a variant in which both groups drain every chunk, one head each, removed the issuer's stage waits but cost
//            more in epilogue hand-shakes than it saved: 150 us vs 132 us per res4 launch)
//   GEMM 3: out = resid + o Wproj^T + bproj (+ fused LayerNorm), A = o from Q2P, accumulators = both stages
//
// Hazards between copies / UMMAs on the same TMEM columns are resolved by the tensor pipe executing in issue order.
// q2 is held as bf16 (the same rounding every other GEMM operand of the path gets); measured effect on the layer output
// < 2e-5 of its max-abs (oracle emulation), inside the 1e-2 parity tolerance.
//
// Warp roles (384 threads): warps 0-3 / 4-7 = epilogue groups, warp 8 = A-tile TMA producer, warp 9 = weight TMA producer,
// warp 10 = tcgen05.cp + MMA issuer (converged warp, elected lane).
#pragma once
#include "traj_fused.cuh"

namespace axvs {

#ifndef TT_HOIST
#define TT_HOIST 0
#endif
#ifndef TT_OWAIT
#define TT_OWAIT 0
#endif
#ifndef TT_PREFETCH
#define TT_PREFETCH 0
#endif
constexpr int TT_A_SLOTS = 4;
constexpr int TT_W_SLOTS = 4;
constexpr int TT_STG_BYTES = 8 * 4096;             // per-warp transpose staging of the output epilogue (+ LayerNorm statistics exchange)
constexpr int TT_SMEM_BYTES = TT_STG_BYTES + TT_A_SLOTS * TF_KB + TT_W_SLOTS * TF_WU + TF_BIAS_BYTES + 512;
static_assert(TT_SMEM_BYTES <= 232448, "traj_ts_kernel exceeds the 227 KiB shared-memory limit");

__device__ __forceinline__ float bf16lo_to_f32(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__global__ void __launch_bounds__(TF_THREADS, 1) traj_ts_kernel(const TrajParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* stg_all = smem;
  uint8_t* a_ring = smem + TT_STG_BYTES;
  uint8_t* w_ring = a_ring + TT_A_SLOTS * TF_KB;
  float* sb_pq = reinterpret_cast<float*>(w_ring + TT_W_SLOTS * TF_WU);
  float* sb_v2 = sb_pq + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb_v2 + 256);
  uint64_t* a_full = bars;                       // [TT_A_SLOTS]
  uint64_t* a_empty = a_full + TT_A_SLOTS;       // [TT_A_SLOTS]
  uint64_t* w_full = a_empty + TT_A_SLOTS;       // [TT_W_SLOTS]
  uint64_t* w_empty = w_full + TT_W_SLOTS;       // [TT_W_SLOTS]
  uint64_t* s_full = w_empty + TT_W_SLOTS;       // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* o_ready = s_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_ready + 1);
  volatile uint32_t* tile_flag = tmem_slot + 2;     // tiles started by the issuer (paces the residual prefetcher, warp 11)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef AXVS_WAIT_PROFILE
  if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[500] = gt; g_trace[501] = clock64(); }
#endif

  if (threadIdx.x == 0) {
    for (int i = 0; i < TT_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < TT_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    mbar_init(o_ready, 8);
    *tile_flag = 0;
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
    const uint32_t t_qp = tmem + lane_base + 64 * g;             // my 4 heads of q2 / o as bf16 pairs (16 columns per head)
    const uint32_t t_s = tmem + lane_base + 256 + 128 * g;       // my accumulator stage
    uint8_t* stg = stg_all + warp * 4096;
    uint32_t s_cnt = 0;                                          // items consumed on my stage
    uint32_t it = 0;                                             // tile iteration
    AXVS_PROF_DECL(7)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // ---- q2 of my 4 heads: (acc + bias) * scale*log2e -> bf16 pairs in Q2P; the stage is then free for the frame chunks
      const bool trc_ = it == 3 && lane == 0 && (warp & 3) == 0;
      const int tb_ = 100 + 100 * g;
      AXVS_TRACE(trc_, tb_ + 0)
      AXVS_PROF_WAIT(0, mbar_wait(&s_full[g], s_cnt & 1))
      AXVS_TRACE(trc_, tb_ + 1)
      ++s_cnt;
      tc_fence_after();
      AXVS_PROF_MARK(tq2_)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float v[32];
        tmem_ld32(t_s + 32 * j, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_bf16x2((v[2 * i] + sb_pq[128 * g + 32 * j + 2 * i]) * p.scale_log2e, (v[2 * i + 1] + sb_pq[128 * g + 32 * j + 2 * i + 1]) * p.scale_log2e);
        tmem_st16u(t_qp + 16 * j, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[g]);
      AXVS_TRACE(trc_, tb_ + 2)
      AXVS_PROF_SPAN(3, tq2_)

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
          AXVS_TRACE(trc_, tb_ + 10 + 4 * (2 * f + cc))
          AXVS_PROF_WAIT(1, mbar_wait(&s_full[g], s_cnt & 1))
          AXVS_TRACE(trc_, tb_ + 11 + 4 * (2 * f + cc))
          ++s_cnt;
          tc_fence_after();
          AXVS_PROF_MARK(tch_)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            // two TMEM round trips per head: (k2, q2) then v2, the v2 load issued before the logit arithmetic
            const int lh = cc * 2 + hh;
            float k2[32], q2f[16], v2[32];
            tmem_ld32(t_s + 32 * hh, k2);
            tmem_ld16(t_qp + 16 * lh, q2f);
            tmem_ld_wait();
            tmem_ld32(t_s + 64 + 32 * hh, v2);
            float2 sacc = make_float2(0.f, 0.f);                // packed fp32 FMAs: two channels per instruction
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t u = __float_as_uint(q2f[i]);
              sacc = fma_f32x2(make_float2(bf16lo_to_f32(u), bf16hi_to_f32(u)), make_float2(k2[2 * i], k2[2 * i + 1]), sacc);
            }
            const float s = sacc.x + sacc.y;
            const float mn = fmaxf(m_run[lh], s);
            const float corr = exp2f(m_run[lh] - mn);
            const float pe = exp2f(s - mn);
            l_run[lh] = l_run[lh] * corr + pe;
            m_run[lh] = mn;
            tmem_ld_wait();
            if (hh == 1) {                                       // both heads of the chunk are in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&s_empty[g]);
              AXVS_PROF_SPAN(4, tch_)
              AXVS_TRACE(trc_, tb_ + 12 + 4 * (2 * f + cc))
            }
            const float2 pe2 = make_float2(pe, pe), corr2 = make_float2(corr, corr);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 r = fma_f32x2(pe2, make_float2(v2[2 * i], v2[2 * i + 1]), mul_f32x2(make_float2(o[lh][2 * i], o[lh][2 * i + 1]), corr2));
              o[lh][2 * i] = r.x;
              o[lh][2 * i + 1] = r.y;
            }
          }
          AXVS_TRACE(trc_, tb_ + 13 + 4 * (2 * f + cc))
        }
      }
      // ---- o = o / l + bv2 -> bf16 pairs over my (now dead) q2 columns: the tensor-memory A operand of the output projection
      AXVS_PROF_MARK(to_)
      AXVS_TRACE(trc_, tb_ + 40)
#pragma unroll
      for (int lh = 0; lh < 4; ++lh) {
        const float inv = 1.f / l_run[lh];
        const int col0 = 128 * g + 32 * lh;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          pk[i] = pack_bf16x2(fmaf(o[lh][2 * i], inv, sb_v2[col0 + 2 * i]), fmaf(o[lh][2 * i + 1], inv, sb_v2[col0 + 2 * i + 1]));
        tmem_st16u(t_qp + 16 * lh, pk);
      }
      AXVS_PROF_SPAN(6, to_)
      AXVS_PROF_MARK(to2_)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_ready);
      AXVS_TRACE(trc_, tb_ + 41)
      // The residual loads below fill the load/store queue for ~2k clk; issued while the OTHER group still reads its bias from shared
      // memory for o (same queue) they delayed its o_ready arrive -- the start of GEMM 3 -- by that much.  So: wait until both groups
      // have handed o over (the loads then overlap GEMM 3 instead of the hand-off).
#if TT_OWAIT
      mbar_wait(o_ready, it & 1);
#endif
      AXVS_PROF_SPAN(5, to2_)
      // ---- output projection item on my stage: out = resid + acc + bproj (my 128 output columns).
      // TMEM rows are one-per-thread; a 4 KiB per-warp transpose through shared memory turns the global accesses into full
      // 128-byte row segments.  resid + bias are fetched BEFORE waiting for the accumulator (latency hides behind GEMM 3).
      {
        const int r = tile * 128 + row_in_tile;
        const int my_orow = traj_row_canonical(r, p);
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
        AXVS_TRACE(trc_, tb_ + 42)
        AXVS_PROF_WAIT(2, mbar_wait(&s_full[g], s_cnt & 1))
        AXVS_TRACE(trc_, tb_ + 43)
        ++s_cnt;
        tc_fence_after();
        if (p.ln_g == nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
              if (j == 3) {                                        // accumulator fully read: the stage is free for the next tile
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[g]);
                AXVS_TRACE(trc_, tb_ + 44)
              }
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
          // shared memory (the partner warp's staging area), then write fp32 rows + the bf16 tile image of the normalised row
          float ps[8], pq[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            {
              float v[32];
              tmem_ld32(t_s + 32 * j, v);
              tmem_ld_wait();
              if (j == 3) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[g]);
                AXVS_TRACE(trc_, tb_ + 44)
              }
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
          // statistics exchange: my 32 rows' (sum, sumsq) go to the head of my staging area (parity-alternating halves), the
          // partner warp (same rows, other column half) reads them after the 256-thread barrier
          float2* xc_mine = reinterpret_cast<float2*>(stg + (it & 1) * 256);
          const float2* xc_other = reinterpret_cast<const float2*>(stg_all + (warp ^ 4) * 4096 + (it & 1) * 256);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
              pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], o);
            }
            if (piece == 0) xc_mine[i * 4 + sub] = make_float2(ps[i], pq[i]);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          float mean[8], rstd[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 other = xc_other[i * 4 + sub];
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
          // (no second barrier: my next write to this staging area is the next tile's projection transpose, which cannot start before
          //  the partner warp has arrived on o_ready for that tile -- i.e. long after it has read these statistics)
        }
        AXVS_TRACE(trc_, tb_ + 45)
      }
    }
    AXVS_PROF_FLUSH(8 + 8 * g, 7, (warp & 3) == 0 && lane == 0)
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer (x_diag, x_0 .. x_{F-1})
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int item = 0; item < 4 * (F + 1); ++item, ++cnt) {
          const uint32_t slot = cnt % TT_A_SLOTS, phase = (cnt / TT_A_SLOTS) & 1;
          const uint8_t* src = (item < 4) ? traj_xd_tile(tile, p) + (size_t)item * TF_KB
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
        if (++slot == TT_W_SLOTS) { slot = 0; phase ^= 1; }
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
      // =============================================================== tcgen05.cp + MMA issuer
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
      const uint32_t t_xa = tmem + 128, t_qp = tmem;
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, s_cnt0 = 0, s_cnt1 = 0, it = 0;
      AXVS_PROF_DECL(7)
      auto w_wait = [&]() -> uint32_t {
        AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == TT_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      auto stage_wait = [&](int g) {
        const uint32_t sc = g ? s_cnt1 : s_cnt0;
        AXVS_PROF_WAIT(1, mbar_wait(&s_empty[g], (sc & 1) ^ 1))
        if (g) ++s_cnt1; else ++s_cnt0;
      };
      // the next A tile (4 K-block images) -> XA; ordered by the tensor pipe behind the UMMAs still reading the old tile
      auto copy_tile = [&]() {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb, ++a_cnt) {
          const uint32_t slot = a_cnt % TT_A_SLOTS;
          AXVS_PROF_WAIT(2, mbar_wait(&a_full[slot], (a_cnt / TT_A_SLOTS) & 1))
          tc_fence_after();
          if (elect_one()) {
            tmem_cp_kblock(t_xa + 32 * kb, a_ring_addr + slot * TF_KB);
            umma_commit(&a_empty[slot]);
          }
          __syncwarp();
        }
      };
#if TT_HOIST
      if ((int)blockIdx.x < p.tiles) copy_tile();               // x_diag of the first tile
#endif
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        // ---- GEMM 1: q2 halves -> the two stages (free once the previous tile's projection has been drained); x_diag is already in XA
        const bool trc_ = it == 3 && lane == 0;
        AXVS_PROF_MARK(tg1_)
        AXVS_TRACE(trc_, 0)
#if TT_PREFETCH
        if (lane == 0) *tile_flag = it + 1;
#endif
#if !TT_HOIST
        copy_tile();
#endif
        stage_wait(0);
        stage_wait(1);
        tc_fence_after();
        AXVS_TRACE(trc_, 1)
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u >> 1, kg = u & 1;
          const uint32_t ws = w_wait();
          umma_unit_elect_ts(tmem + 256 + half * 128, t_xa + 64 * kg, t_xa + 64 * kg + 32, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                             &w_empty[ws], kg == 1 ? &s_full[half] : nullptr, nullptr);
        }
        AXVS_PROF_SPAN(6, tg1_)
        AXVS_TRACE(trc_, 2)
        AXVS_PROF_MARK(tfr_)
        // ---- GEMM 2: per frame, four 128-column chunks alternating between the two TMEM stages
#pragma unroll 1
        for (int f = 0; f < F; ++f) {
          AXVS_TRACE(trc_, 60 + 2 * f)
          copy_tile();
          AXVS_TRACE(trc_, 61 + 2 * f)
#pragma unroll 1
          for (int ci = 0; ci < 4; ++ci) {
            const int g = ci & 1;                               // chunk order 0,2,1,3 -> stage 0,1,0,1
            AXVS_TRACE(trc_, 10 + 4 * (4 * f + ci))
            stage_wait(g);
            tc_fence_after();
            AXVS_TRACE(trc_, 11 + 4 * (4 * f + ci))
#pragma unroll 1
            for (int kg = 0; kg < 2; ++kg) {
              const uint32_t ws = w_wait();
              umma_unit_elect_ts(tmem + 256 + g * 128, t_xa + 64 * kg, t_xa + 64 * kg + 32, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                                 &w_empty[ws], kg == 1 ? &s_full[g] : nullptr, nullptr);
              AXVS_TRACE(trc_, 12 + kg + 4 * (4 * f + ci))
            }
          }
        }
        AXVS_PROF_SPAN(5, tfr_)
        // XA is dead once the last frame's UMMAs have run (tensor pipe order): the NEXT tile's x_diag goes in now, under the o hand-off
#if TT_HOIST
        if (tile + (int)gridDim.x < p.tiles) copy_tile();
#endif
        AXVS_TRACE(trc_, 3)
        // ---- GEMM 3: output projection, A = o (bf16 pairs written over q2 by the epilogue), accumulators = both stages
        AXVS_TRACE(trc_, 50)
        AXVS_PROF_WAIT(3, mbar_wait(o_ready, it & 1))
        AXVS_TRACE(trc_, 51)
        stage_wait(0);
        stage_wait(1);
        tc_fence_after();
        AXVS_TRACE(trc_, 52)
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {
          const int half = u >> 1, kg = u & 1;
          const uint32_t ws = w_wait();
          umma_unit_elect_ts(tmem + 256 + half * 128, t_qp + 64 * kg, t_qp + 64 * kg + 32, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                             &w_empty[ws], kg == 1 ? &s_full[half] : nullptr, nullptr);
        }
        AXVS_TRACE(trc_, 53)
      }
      AXVS_PROF_FLUSH(0, 7, lane == 0)
    }
#if TT_PREFETCH
    else if (warp == 11 && p.resid != nullptr) {
      // =============================================================== residual prefetcher: the epilogue reads 128 fp32 rows (128 KiB) per tile
      // right before GEMM 3 and then waits for them; pulling the lines into L2 while the tile's frames are still being computed turns a
      // DRAM round trip on the tile's critical path into an L2 hit (no registers or shared memory needed)
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        while (*tile_flag < it + 1) __nanosleep(200);
#pragma unroll 1
        for (int rr = lane; rr < 128; rr += 32) {
          const int r = tile * 128 + rr;
          const int cr = traj_row_canonical(r, p);
          if (cr >= 0) {
            const char* src = reinterpret_cast<const char*>(p.resid + (size_t)cr * 256);
#pragma unroll
            for (int l = 0; l < 8; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 128 * l));
          }
        }
      }
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
#ifdef AXVS_WAIT_PROFILE
  if (threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_trace[502] = gt; g_trace[503] = clock64(); }
#endif
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
