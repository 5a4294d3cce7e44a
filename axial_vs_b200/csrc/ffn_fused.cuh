// Fused FFN tail of the layer (WC/temporal_attention.py:181-185, 218), one persistent tcgen05 kernel per 128-token tile:
//
//   s   = LayerNorm1(x)            -- produced by ln_image_kernel as fp32 rows + a bf16 tile image (UMMA A operand)
//   h_j = relu(s W1[j]^T + b1[j])   j = 0..d_ffn/128-1   (GEMM 1, 128-column chunks into two alternating TMEM stages;
//                                                          epilogue: bias, ReLU, bf16 pairs written back IN PLACE over the
//                                                          stage with tcgen05.st -> tensor-memory A operand of GEMM 2)
//   acc2 += h_j W2[:, j]^T                               (GEMM 2, K-chunk j, A from TMEM: N = 128 UMMAs at the full 64 clk
//                                                          rate; accumulator resident in TMEM columns [0,256))
//   out = LayerNorm2(s + acc2 + b2)                      (final epilogue, coalesced through a shared-memory transpose)
//
// The d_ffn-wide hidden activation never leaves the SM.
// Warp roles (384 threads): warps 0-3 / 4-7 = epilogue groups (even / odd chunks; output columns 0-127 / 128-255),
// warp 8 = A-tile TMA producer, warp 9 = weight TMA producer, warp 10 = MMA issuer.
#pragma once
#include "traj_fused.cuh"

namespace axvs {

constexpr int FF_THREADS = 384;
constexpr int FF_A_SLOTS = 5;
constexpr int FF_W_SLOTS = 3;                  // 32 KiB weight units
constexpr int FF_H_BYTES = 2 * TF_KB;     // per-warp 4 KiB transpose staging of the final epilogue
constexpr int FF_XCHG_BYTES = 2 * 2 * 128 * 8;
constexpr int FF_MAX_DFFN = 1024;
constexpr int FF_BIAS_BYTES = (FF_MAX_DFFN + 3 * 256) * 4;   // b1 | b2 | ln2 gamma | ln2 beta staged in shared memory
constexpr int FF_SMEM_BYTES = FF_A_SLOTS * TF_KB + FF_H_BYTES + FF_W_SLOTS * TF_WU + FF_XCHG_BYTES + FF_BIAS_BYTES + 512;

struct FfnParams {
  const uint8_t* s_img;  // LayerNorm1 output, bf16 tile image [tiles][4][16 KiB]
  const float* s32;      // LayerNorm1 output, fp32 [rows, 256] (residual)
  float* out;            // [rows, 256] fp32
  const float *ln2_g, *ln2_b;
  const uint8_t* w1;     // unit format, (row tile j, K group): unit = 2 j + kg
  const uint8_t* w2;     // unit format, K-major: unit = 2 jj + half
  const float *b1, *b2;
  int rows, tiles, d_ffn;
  float eps;
};

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_fused_kernel(const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need a 1024 B aligned base (no static smem in this kernel)
  uint8_t* a_ring = smem;
  uint8_t* h_buf = a_ring + FF_A_SLOTS * TF_KB;
  uint8_t* w_ring = h_buf + FF_H_BYTES;
  float2* xchg = reinterpret_cast<float2*>(w_ring + FF_W_SLOTS * TF_WU);   // [2 parity][2 group][128] (sum, sumsq)
  float* sb1 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xchg) + FF_XCHG_BYTES);   // b1 [d_ffn]
  float* sb2 = sb1 + FF_MAX_DFFN;                                                           // b2 [256]
  float* sg2 = sb2 + 256;                                                                   // ln2 gamma
  float* sbe2 = sg2 + 256;                                                                  // ln2 beta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbe2 + 256);
  uint64_t* a_full = bars;                    // [FF_A_SLOTS]
  uint64_t* a_empty = a_full + FF_A_SLOTS;    // [FF_A_SLOTS]
  uint64_t* w_full = a_empty + FF_A_SLOTS;    // [6]
  uint64_t* w_empty = w_full + FF_W_SLOTS;    // [6]
  uint64_t* s_full = w_empty + FF_W_SLOTS;    // [2]
  uint64_t* h_ready = s_full + 2;             // [2] epilogue -> MMA: the bf16 hidden chunk is in place in TMEM stage j & 1
  uint64_t* acc_full = h_ready + 2;           // MMA -> epilogue
  uint64_t* acc_free = acc_full + 1;          // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NJ = p.d_ffn / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FF_A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FF_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&h_ready[i], 8); }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  // biases / LayerNorm2 affine -> shared memory (the ~10 KiB of L1 left beside 217 KiB of smem cannot keep them hot)
  for (int i = threadIdx.x; i < p.d_ffn; i += FF_THREADS) sb1[i] = p.b1[i];
  for (int i = threadIdx.x; i < 256; i += FF_THREADS) { sb2[i] = p.b2[i]; sg2[i] = p.ln2_g[i]; sbe2[i] = p.ln2_b[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue groups
    setmaxnreg_inc<224>();   // 256*224 + 128*56 = 64512 = the CTA register pool at launch (384 x 168)
    const int g = warp >> 2;
    const int wq = warp & 3;
    const int row_in_tile = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t t_acc = tmem + lane_base + 128 * g;          // my 128 output columns of acc2
    const int sub = lane >> 3, piece = lane & 7;
    uint8_t* stg = h_buf + warp * 4096;                         // per-warp transpose staging (final epilogue only)
    uint32_t it = 0;
    AXVS_PROF_DECL(7)
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // ---- hidden chunks: BOTH groups drain every chunk (group g takes its 64 columns = K-block g of h_buf), which halves
      // the drain latency that sits on the GEMM 1 -> GEMM 2 critical path: bias + ReLU -> bf16 -> h_buf
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        const uint32_t hc = it * NJ + j;                         // global chunk counter (NJ is even: stage = hc & 1)
        const int stage = j & 1;
        AXVS_PROF_WAIT(0, mbar_wait(&s_full[stage], (hc >> 1) & 1))
        tc_fence_after();
        const uint32_t t_s = tmem + lane_base + 256 + stage * 128 + 64 * g;
        uint32_t hpk[32];
        {
          float v0[32], v1[32];                                    // both halves in flight: one TMEM round trip per chunk
          tmem_ld32(t_s, v0);
          tmem_ld32(t_s + 32, v1);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(sb1 + j * 128 + 64 * g);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = b4[i], bc = b4[8 + i];
            hpk[2 * i] = pack_bf16x2(fmaxf(v0[4 * i] + bb.x, 0.f), fmaxf(v0[4 * i + 1] + bb.y, 0.f));
            hpk[2 * i + 1] = pack_bf16x2(fmaxf(v0[4 * i + 2] + bb.z, 0.f), fmaxf(v0[4 * i + 3] + bb.w, 0.f));
            hpk[16 + 2 * i] = pack_bf16x2(fmaxf(v1[4 * i] + bc.x, 0.f), fmaxf(v1[4 * i + 1] + bc.y, 0.f));
            hpk[16 + 2 * i + 1] = pack_bf16x2(fmaxf(v1[4 * i + 2] + bc.z, 0.f), fmaxf(v1[4 * i + 3] + bc.w, 0.f));
          }
        }
        // the 64 fp32 columns this thread just read become 32 columns of bf16 pairs at the start of the same region
        tmem_st32u(t_s, hpk);
        AXVS_PROF_WAIT(6, tmem_st_wait())
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[stage]);
      }
      // ---- final: t = acc2 + b2 + s, LayerNorm2, store.  Rows are one-per-thread in TMEM; a per-warp transpose through
      // shared memory (h_buf is idle: every GEMM 2 of this tile has retired) makes the global traffic row-segment
      // coalesced.  In the transposed domain lane (sub, piece) owns 4 columns of rows {4*i + sub}.
      // The residual (s + b2) is fetched BEFORE waiting for the accumulator so its latency hides behind the last GEMMs.
      const int row0 = tile * 128 + wq * 32;
      float4 t[4][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 128 * g + 32 * c + piece * 4;
        const float4 bb = *reinterpret_cast<const float4*>(sb2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + i * 4 + sub;
          float4 sres = (r < p.rows) ? __ldg(reinterpret_cast<const float4*>(p.s32 + (size_t)r * 256 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          t[c][i] = make_float4(sres.x + bb.x, sres.y + bb.y, sres.z + bb.z, sres.w + bb.w);
        }
      }
      AXVS_PROF_WAIT(2, mbar_wait(acc_full, it & 1))
      tc_fence_after();
#ifdef AXVS_WAIT_PROFILE
      const long long tf0_ = clock64();
#endif
      float ps[8], pq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ps[i] = pq[i] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        {
          float v[32];
          tmem_ld32(t_acc + 32 * c, v);
          tmem_ld_wait();
          if (c == 3) {                                            // acc2 fully read: release it for the next tile
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + sub;
          const float4 a = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
          float4 tv = t[c][i];
          tv.x += a.x; tv.y += a.y; tv.z += a.z; tv.w += a.w;
          t[c][i] = tv;
          ps[i] += tv.x + tv.y + tv.z + tv.w;
          pq[i] += tv.x * tv.x + tv.y * tv.y + tv.z * tv.z + tv.w * tv.w;
        }
        __syncwarp();
      }
#ifdef AXVS_WAIT_PROFILE
      prof_acc_[4] += clock64() - tf0_;
#endif
      // row statistics: reduce over the 8 lanes sharing a row, then combine with the other column half (other group)
      float2* xc = xchg + (it & 1) * 256;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], o);
          pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], o);
        }
        if (piece == 0) xc[g * 128 + wq * 32 + i * 4 + sub] = make_float2(ps[i], pq[i]);
      }
      AXVS_PROF_WAIT(3, asm volatile("bar.sync 1, 256;" ::: "memory"))
      float mean[8], rstd[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 other = xc[(g ^ 1) * 128 + wq * 32 + i * 4 + sub];
        mean[i] = (ps[i] + other.x) * (1.f / 256.f);
        const float var = fmaxf((pq[i] + other.y) * (1.f / 256.f) - mean[i] * mean[i], 0.f);
        rstd[i] = rsqrtf(var + p.eps);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = 128 * g + 32 * c + piece * 4;
        const float4 gg = *reinterpret_cast<const float4*>(sg2 + col), be = *reinterpret_cast<const float4*>(sbe2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + i * 4 + sub;
          if (r < p.rows) {
            const float4 tv = t[c][i];
            *reinterpret_cast<float4*>(p.out + (size_t)r * 256 + col) =
                make_float4((tv.x - mean[i]) * rstd[i] * gg.x + be.x, (tv.y - mean[i]) * rstd[i] * gg.y + be.y,
                            (tv.z - mean[i]) * rstd[i] * gg.z + be.z, (tv.w - mean[i]) * rstd[i] * gg.w + be.w);
          }
        }
      }
#ifdef AXVS_WAIT_PROFILE
      prof_acc_[5] += clock64() - tf0_;
#endif
    }
    AXVS_PROF_FLUSH(8 + 8 * g, 7, (warp & 3) == 0 && lane == 0)
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 && lane == 0) {
      // =============================================================== A-tile producer
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int kb = 0; kb < 4; ++kb, ++cnt) {
          const uint32_t slot = cnt % FF_A_SLOTS, phase = (cnt / FF_A_SLOTS) & 1;
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], TF_KB);
          tma_bulk_g2s(a_ring + slot * TF_KB, p.s_img + ((size_t)tile * 4 + kb) * TF_KB, TF_KB, &a_full[slot]);
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
        if (++slot == FF_W_SLOTS) { slot = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) { push(p.w1, 2 * j); push(p.w1, 2 * j + 1); }
          if (j >= 1) { push(p.w2, 2 * (j - 1)); push(p.w2, 2 * (j - 1) + 1); }
        }
      }
      AXVS_PROF_FLUSH(32, 1, true)
    } else if (warp == 10) {
      // =============================================================== MMA issuer (converged warp, elected lane issues)
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_ring_addr = smem_u32(a_ring), w_ring_addr = smem_u32(w_ring);
      uint32_t a_cnt = 0, w_slot = 0, w_phase = 0, it = 0;
      AXVS_PROF_DECL(5)
      auto w_wait = [&]() -> uint32_t {
        AXVS_PROF_WAIT(0, mbar_wait(&w_full[w_slot], w_phase))
        tc_fence_after();
        const uint32_t ws = w_slot;
        if (++w_slot == FF_W_SLOTS) { w_slot = 0; w_phase ^= 1; }
        return ws;
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
            // GEMM 1, chunk j -> stage j & 1
            // (the stage's previous occupant, chunk j-2, was consumed by GEMM 2 of that chunk, issued earlier by this thread:
            //  the tensor pipe executes in issue order, so no barrier is needed before overwriting it)
            const int g = j & 1;
#pragma unroll 1
            for (int kg = 0; kg < 2; ++kg) {
              const uint32_t ac0 = a_cnt + 2 * kg, ac1 = ac0 + 1;
              const uint32_t s0 = ac0 % FF_A_SLOTS, s1 = ac1 % FF_A_SLOTS;
              if (j == 0) {
                AXVS_PROF_WAIT(2, mbar_wait(&a_full[s0], (ac0 / FF_A_SLOTS) & 1); mbar_wait(&a_full[s1], (ac1 / FF_A_SLOTS) & 1))
                tc_fence_after();
              }
              const uint32_t ws = w_wait();
              // last chunk: final use of the A K-blocks; last K group: the chunk accumulator is complete
              umma_unit_elect(tmem + 256 + g * 128, a_ring_addr + s0 * TF_KB, a_ring_addr + s1 * TF_KB, w_ring_addr + ws * TF_WU, idesc, kg != 0,
                              &w_empty[ws], j == NJ - 1 ? &a_empty[s0] : nullptr, j == NJ - 1 ? &a_empty[s1] : nullptr, kg == 1 ? &s_full[g] : nullptr);
            }
          }
          if (j >= 1) {
            // GEMM 2, K-chunk j-1: acc2 += h (128 x 128) * W2[:, 128(j-1) : 128j]^T, one unit per output-column half
            const int jj = j - 1;
            const uint32_t hc = it * NJ + jj;
            if (jj == 0) AXVS_PROF_WAIT(4, mbar_wait(acc_free, (it & 1) ^ 1))   // previous tile's final epilogue has drained acc2
            AXVS_PROF_WAIT(3, mbar_wait(&h_ready[jj & 1], (hc >> 1) & 1))
            tc_fence_after();
            const uint32_t t_h = tmem + 256 + (jj & 1) * 128;                 // K 0..63 at columns [0,32), K 64..127 at [64,96)
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
              const uint32_t ws = w_wait();
              umma_unit_elect_ts(tmem + half * 128, t_h, t_h + 64, w_ring_addr + ws * TF_WU, idesc, jj != 0,
                                 &w_empty[ws], (half && j == NJ) ? acc_full : nullptr, nullptr);
            }
          }
        }
        a_cnt += 4;
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

// s = LayerNorm(x): fp32 rows [rows,256] (residual of the FFN) and the bf16 tile image consumed by ffn_fused_kernel.
// One warp per row; lane l owns columns 8l..8l+7 = K-block l/8, 16-byte chunk l%8 of the image row.
__global__ void ln_image_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                float* __restrict__ y32, uint8_t* __restrict__ img, int rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2 + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2), b1 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2 + 1);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)r * C256) + lane * 2;
    float4 a = __ldg(x4), c = __ldg(x4 + 1);
    const float mu = warp_sum(a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w) * (1.f / C256);
    a.x -= mu; a.y -= mu; a.z -= mu; a.w -= mu; c.x -= mu; c.y -= mu; c.z -= mu; c.w -= mu;
    const float rstd = rsqrtf(warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / C256) + eps);
    a.x = a.x * rstd * g0.x + b0.x; a.y = a.y * rstd * g0.y + b0.y; a.z = a.z * rstd * g0.z + b0.z; a.w = a.w * rstd * g0.w + b0.w;
    c.x = c.x * rstd * g1.x + b1.x; c.y = c.y * rstd * g1.y + b1.y; c.z = c.z * rstd * g1.z + b1.z; c.w = c.w * rstd * g1.w + b1.w;
    float4* o = reinterpret_cast<float4*>(y32 + (size_t)r * C256) + lane * 2;
    o[0] = a; o[1] = c;
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(c.x, c.y); u.w = pack_bf16x2(c.z, c.w);
    *reinterpret_cast<uint4*>(img + ((size_t)(r >> 7) * 4 + (lane >> 3)) * TF_KB + sw128_offset(r & 127, lane & 7)) = u;
  }
}

}  // namespace axvs
static_assert(axvs::FF_SMEM_BYTES <= 232448, "ffn_fused_kernel exceeds the 227 KiB shared-memory limit");
static_assert(axvs::TF_SMEM_BYTES <= 232448, "traj_fused_kernel exceeds the 227 KiB shared-memory limit");
static_assert(axvs::GEMM_SMEM_BYTES <= 232448, "gemm_bf16_kernel exceeds the 227 KiB shared-memory limit");
