// Fused tail of the layer (WC/temporal_attention.py:181-185, 217-218), one persistent tcgen05 kernel per 128-token tile:
//
//   s   = LayerNorm1(x)                                  (producer warps: fp32 row -> bf16 SWIZZLE_128B A operand + row stats)
//   h_j = relu(s W1[j]^T + b1[j])   j = 0..d_ffn/128-1   (GEMM 1, 128-column chunks into two alternating TMEM stages;
//                                                          epilogue: bias, ReLU, bf16 -> shared-memory A operand of GEMM 2)
//   acc2 += h_j W2[:, j]^T                               (GEMM 2, K-chunk j, accumulator resident in TMEM columns [0,256))
//   out = LayerNorm2(s + acc2 + b2)                      (final epilogue; s is recomputed in fp32 from x and the row stats)
//
// The d_ffn-wide hidden activation never leaves the SM.
// Warp roles (384 threads): warps 0-3 / 4-7 = epilogue groups (even / odd chunks; output columns 0-127 / 128-255),
// warps 8 and 11 = LayerNorm1 producers, warp 9 = weight TMA producer, warp 10 = MMA issuer.
#pragma once
#include "traj_fused.cuh"

namespace axvs {

constexpr int FF_THREADS = 384;
constexpr int FF_W_SLOTS = 6;
constexpr int FF_A_BYTES = 4 * TF_KB;     // 128 x 256 bf16
constexpr int FF_H_BYTES = 2 * TF_KB;     // 128 x 128 bf16
constexpr int FF_STATS_BYTES = 128 * 8;   // (mean, rstd) per row
constexpr int FF_XCHG_BYTES = 2 * 2 * 128 * 8;
constexpr int FF_SMEM_BYTES = FF_A_BYTES + FF_H_BYTES + FF_W_SLOTS * TF_KB + FF_STATS_BYTES + FF_XCHG_BYTES + 1024 + 512;

struct FfnParams {
  const float* x;        // [rows, 256] fp32
  float* out;            // [rows, 256] fp32
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const uint8_t* w1;     // packed [4 kb][d_ffn rows][128 B]
  const uint8_t* w2;     // packed [d_ffn/64 kb][256 rows][128 B]
  const float *b1, *b2;
  int rows, tiles, d_ffn;
  float eps;
};

__global__ void __launch_bounds__(FF_THREADS, 1) ffn_fused_kernel(const FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_buf = smem;
  uint8_t* h_buf = a_buf + FF_A_BYTES;
  uint8_t* w_ring = h_buf + FF_H_BYTES;
  float2* stats = reinterpret_cast<float2*>(w_ring + FF_W_SLOTS * TF_KB);            // [128]
  float2* xchg = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats) + FF_STATS_BYTES);   // [2 parity][2 group][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + FF_XCHG_BYTES);
  uint64_t* w_full = bars;                    // [6]
  uint64_t* w_empty = w_full + FF_W_SLOTS;    // [6]
  uint64_t* s_full = w_empty + FF_W_SLOTS;    // [2]
  uint64_t* s_empty = s_full + 2;             // [2]
  uint64_t* a_ready = s_empty + 2;            // LN1 producers -> MMA / epilogue
  uint64_t* a_free = a_ready + 1;             // MMA (GEMM 1 of the tile retired) -> LN1 producers
  uint64_t* h_ready = a_free + 1;             // epilogue -> MMA
  uint64_t* h_free = h_ready + 1;             // MMA -> epilogue
  uint64_t* acc_full = h_free + 1;            // MMA -> epilogue
  uint64_t* acc_free = acc_full + 1;          // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NJ = p.d_ffn / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < FF_W_SLOTS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    mbar_init(a_ready, 2);
    mbar_init(a_free, 1);
    mbar_init(h_ready, 4);
    mbar_init(h_free, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 8);
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =============================================================== epilogue groups
    setmaxnreg_inc<224>();   // 256*224 + 128*56 = 64512 = the CTA register pool at launch (384 x 168)
    const int g = warp >> 2;
    const int row_in_tile = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t t_acc = tmem + lane_base + 128 * g;          // my 128 output columns of acc2
    const uint32_t t_s = tmem + lane_base + 256 + 128 * g;      // my GEMM-1 stage
    uint32_t s_cnt = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      // row statistics of this tile (read now: the producers may refill `stats` once GEMM 1 of this tile has retired)
      mbar_wait(a_ready, it & 1);
      const float2 st = stats[row_in_tile];
      // ---- hidden chunks j = g, g+2, ...: bias + ReLU -> bf16 -> h_buf
      for (int j = g; j < NJ; j += 2) {
        mbar_wait(&s_full[g], s_cnt & 1);
        ++s_cnt;
        tc_fence_after();
        uint32_t hpk[64];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld32(t_s + 32 * c, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(p.b1 + j * 128 + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = __ldg(b4 + i);
            hpk[c * 16 + 2 * i] = pack_bf16x2(fmaxf(v[4 * i] + bb.x, 0.f), fmaxf(v[4 * i + 1] + bb.y, 0.f));
            hpk[c * 16 + 2 * i + 1] = pack_bf16x2(fmaxf(v[4 * i + 2] + bb.z, 0.f), fmaxf(v[4 * i + 3] + bb.w, 0.f));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[g]);                 // TMEM stage drained
        const uint32_t hc = it * NJ + j;                         // global chunk counter: h_buf is free once GEMM 2 of chunk hc-1 retired
        mbar_wait(h_free, (hc & 1) ^ 1);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          uint4 u = make_uint4(hpk[4 * q], hpk[4 * q + 1], hpk[4 * q + 2], hpk[4 * q + 3]);
          *reinterpret_cast<uint4*>(h_buf + (q >> 3) * TF_KB + sw128_offset(row_in_tile, q & 7)) = u;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready);
      }
      // ---- final: t = acc2 + b2 + s (s recomputed in fp32), LayerNorm2, store
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const int r = tile * 128 + row_in_tile;
      const bool valid = r < p.rows;
      float t[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld32(t_acc + 32 * c, v);
        tmem_ld_wait();
        const int col = 128 * g + 32 * c;
        const float4* x4 = reinterpret_cast<const float4*>(p.x + (size_t)(valid ? r : 0) * 256 + col);
        const float4* g4 = reinterpret_cast<const float4*>(p.ln1_g + col);
        const float4* be4 = reinterpret_cast<const float4*>(p.ln1_b + col);
        const float4* b4 = reinterpret_cast<const float4*>(p.b2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 xx = __ldg(x4 + i), gg = __ldg(g4 + i), be = __ldg(be4 + i), bb = __ldg(b4 + i);
          t[32 * c + 4 * i] = v[4 * i] + bb.x + ((xx.x - st.x) * st.y * gg.x + be.x);
          t[32 * c + 4 * i + 1] = v[4 * i + 1] + bb.y + ((xx.y - st.x) * st.y * gg.y + be.y);
          t[32 * c + 4 * i + 2] = v[4 * i + 2] + bb.z + ((xx.z - st.x) * st.y * gg.z + be.z);
          t[32 * c + 4 * i + 3] = v[4 * i + 3] + bb.w + ((xx.w - st.x) * st.y * gg.w + be.w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);                      // acc2 columns drained into registers
      // partial statistics over my 128 columns, combined with the sibling thread (other group, same row)
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 128; ++i) sum += t[i];
      const float mean_p = sum * (1.f / 128.f);
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < 128; ++i) { const float d = t[i] - mean_p; m2 = fmaf(d, d, m2); }
      float2* xc = xchg + (it & 1) * 256;
      xc[g * 128 + row_in_tile] = make_float2(mean_p, m2);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float2 other = xc[(g ^ 1) * 128 + row_in_tile];
      const float mean = 0.5f * (mean_p + other.x);
      const float dm = mean_p - other.x;
      const float var = (m2 + other.y + dm * dm * 64.f) * (1.f / 256.f);
      const float rstd = rsqrtf(var + p.eps);
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col = 128 * g + 32 * c;
          float4* o4 = reinterpret_cast<float4*>(p.out + (size_t)r * 256 + col);
          const float4* g4 = reinterpret_cast<const float4*>(p.ln2_g + col);
          const float4* be4 = reinterpret_cast<const float4*>(p.ln2_b + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 gg = __ldg(g4 + i), be = __ldg(be4 + i);
            o4[i] = make_float4((t[32 * c + 4 * i] - mean) * rstd * gg.x + be.x, (t[32 * c + 4 * i + 1] - mean) * rstd * gg.y + be.y,
                                (t[32 * c + 4 * i + 2] - mean) * rstd * gg.z + be.z, (t[32 * c + 4 * i + 3] - mean) * rstd * gg.w + be.w);
          }
        }
      }
    }
  } else {
    setmaxnreg_dec<56>();
    if (warp == 8 || warp == 11) {
      // =============================================================== LayerNorm1 producers (one warp per row)
      const int pw = (warp == 8) ? 0 : 1;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln1_g) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(p.ln1_g) + lane * 2 + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.ln1_b) + lane * 2), b1 = __ldg(reinterpret_cast<const float4*>(p.ln1_b) + lane * 2 + 1);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_free, (it & 1) ^ 1);                         // GEMM 1 of the previous tile has retired
#pragma unroll 2
        for (int rr = pw; rr < 128; rr += 2) {
          const int r = tile * 128 + rr;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
          if (r < p.rows) {
            const float4* x4 = reinterpret_cast<const float4*>(p.x + (size_t)r * 256) + lane * 2;
            a = __ldg(x4);
            c = __ldg(x4 + 1);
          }
          const float mu = warp_sum(a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w) * (1.f / 256.f);
          a.x -= mu; a.y -= mu; a.z -= mu; a.w -= mu; c.x -= mu; c.y -= mu; c.z -= mu; c.w -= mu;
          const float rstd = rsqrtf(warp_sum(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w) * (1.f / 256.f) + p.eps);
          uint4 u;
          u.x = pack_bf16x2(a.x * rstd * g0.x + b0.x, a.y * rstd * g0.y + b0.y);
          u.y = pack_bf16x2(a.z * rstd * g0.z + b0.z, a.w * rstd * g0.w + b0.w);
          u.z = pack_bf16x2(c.x * rstd * g1.x + b1.x, c.y * rstd * g1.y + b1.y);
          u.w = pack_bf16x2(c.z * rstd * g1.z + b1.z, c.w * rstd * g1.w + b1.w);
          // lane l holds columns 8l..8l+7 = K-block l/8, 16-byte chunk l%8
          *reinterpret_cast<uint4*>(a_buf + (lane >> 3) * TF_KB + sw128_offset(rr, lane & 7)) = u;
          if (lane == 0) stats[rr] = make_float2(mu, rstd);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
    } else if (warp == 9 && lane == 0) {
      // =============================================================== weight producer
      uint32_t cnt = 0;
      auto push = [&](const uint8_t* img, int rows_total, int kb, int row0) {
        const uint32_t slot = cnt % FF_W_SLOTS, phase = (cnt / FF_W_SLOTS) & 1;
        mbar_wait(&w_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&w_full[slot], TF_KB);
        tma_bulk_g2s(w_ring + slot * TF_KB, img + ((size_t)kb * rows_total + row0) * 128, TF_KB, &w_full[slot]);
        ++cnt;
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb) push(p.w1, p.d_ffn, kb, j * 128);
          }
          if (j >= 1) {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) push(p.w2, 256, 2 * (j - 1) + (i >> 1), (i & 1) * 128);
          }
        }
      }
    } else if (warp == 10 && lane == 0) {
      // =============================================================== MMA issuer
      const uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t a_addr = smem_u32(a_buf), h_addr = smem_u32(h_buf), w_ring_addr = smem_u32(w_ring);
      uint32_t w_cnt = 0, s_cnt0 = 0, s_cnt1 = 0, it = 0;
      auto w_wait = [&]() -> uint32_t {
        const uint32_t slot = w_cnt % FF_W_SLOTS, phase = (w_cnt / FF_W_SLOTS) & 1;
        mbar_wait(&w_full[slot], phase);
        tc_fence_after();
        return slot;
      };
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        mbar_wait(a_ready, it & 1);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j <= NJ; ++j) {
          if (j < NJ) {
            // GEMM 1, chunk j -> stage j & 1
            const int g = j & 1;
            const uint32_t sc = g ? s_cnt1 : s_cnt0;
            mbar_wait(&s_empty[g], (sc & 1) ^ 1);
            if (g) ++s_cnt1; else ++s_cnt0;
            tc_fence_after();
#pragma unroll 1
            for (int kb = 0; kb < 4; ++kb, ++w_cnt) {
              const uint32_t ws = w_wait();
              umma_kblock(tmem + 256 + g * 128, a_addr + kb * TF_KB, w_ring_addr + ws * TF_KB, idesc, kb != 0);
              umma_commit(&w_empty[ws]);
            }
            umma_commit(&s_full[g]);
            if (j == NJ - 1) umma_commit(a_free);                // a_buf may be overwritten by the next tile's LayerNorm1
          }
          if (j >= 1) {
            // GEMM 2, K-chunk j-1: acc2 += h (128 x 128) * W2[:, 128(j-1) : 128j]^T
            const int jj = j - 1;
            const uint32_t hc = it * NJ + jj;
            if (jj == 0) {
              mbar_wait(acc_free, (it & 1) ^ 1);                 // previous tile's final epilogue has drained acc2
            }
            mbar_wait(h_ready, hc & 1);
            tc_fence_after();
#pragma unroll 1
            for (int i = 0; i < 4; ++i, ++w_cnt) {
              const int kb2 = i >> 1, half = i & 1;
              const uint32_t ws = w_wait();
              umma_kblock(tmem + half * 128, h_addr + kb2 * TF_KB, w_ring_addr + ws * TF_KB, idesc, (jj | kb2) != 0);
              umma_commit(&w_empty[ws]);
            }
            umma_commit(h_free);
          }
        }
        umma_commit(acc_full);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace axvs
