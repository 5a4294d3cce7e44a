// kMaX pixel-decoder axial attention core.
// Reference: AxialAttention.forward, Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:128-157 (eval mode):
//   logits[l, m] = BN0(q_l . k_m) + BN1(q_l . rq[l, m]) + BN2(k_m . rk[l, m])        rq / rk / rv[l, m] = embedding[m - l + 254]
//   w = softmax_m(logits)                                                              (fp32, :147-148)
//   y[c, l] = BNa(sum_m w[l, m] v_m[c]) + BNb(sum_m w[l, m] rv[l, m][c])
// The 1x1 qkv convolution with its folded batch norm is the tcgen05 GEMM (gemm.cuh, fp32 token rows [q | k | v] out); this kernel is
// everything after it, in fp32.  Persistent CTAs walk the (sequence, head) items: q, k, v of the head and the 2L-1 rows of each embedding table that a
// length-L axis can address live in shared memory (rows padded by 4 words: 16-byte row alignment, and 8 consecutive rows tile the 32 banks,
// so float4 reads of consecutive rows are conflict-free), the L x L weights too.  Both contraction phases are register-tiled over four
// positions so that a shared-memory wavefront feeds 40-55 FMAs instead of 20-24.
// Sequences are addressed by strides so the same kernel serves the 1-D module ([N, C, L]) and both passes of AxialAttention2D
// (height axis: token rows at stride W, token-major output feeding the width-axis GEMM; width axis: NCHW output).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "attn.cuh"   // mma_bf16_16816
#include "ptx.cuh"    // pack_bf16x2 / unpack_bf16x2

namespace axvs {

constexpr int KA_MAX_L = 128;          // SIMT kernel: any axis whose tables fit shared memory (the softmax walks the row in 32-lane pieces)
constexpr int KA_MAX_SPAN = 255;
constexpr int KA_THREADS = 256;
constexpr int KA_TC_THREADS = 512;     // tensor-core variant: 16 warps hide the fragment-load / mma latencies of the small tiles

struct KmaxAxialParams {
  const float* qkv;            // fp32 token rows [rows, 2 * H * dk + H * dv]
  int ld;                      // row length
  int L, heads, dk, dv;
  int n_items;                 // heads x sequences, item = s * heads + h
  int seq_inner;               // sequence s -> first row (s / seq_inner) * row_outer + (s % seq_inner) * row_inner; position l adds l * row_pos
  long long row_outer, row_inner, row_pos;
  const float* emb_q;          // [2 * 255 - 1, dk]
  const float* emb_k;          // [509, dk]
  const float* emb_v;          // [509, dv]
  const float* sim_s;          // [3 * heads] folded batch-norm scale / shift of the similarity logits (content, query-rpe, key-rpe)
  const float* sim_t;
  const float* out_s;          // [2 * heads * dv] folded batch norm of the retrieved output (content channels, then rpe channels)
  const float* out_t;
  float* out;                  // element (s, c, l) at (s / seq_inner) * out_outer + (s % seq_inner) * out_inner + c * out_chan + l * out_pos
  long long out_outer, out_inner, out_chan, out_pos;  int overlay;                 // SIMT kernel, long axes: the value-side operands (v, rv) share shared memory with the key-side ones (q, k, rq, rk)
                               // and every table is re-staged per item (see kmax_axial_smem_bytes_overlay)
};

// fp32 activations -> bf16 [rows][hi (K) | lo (K)] token rows: the A operand of the split-precision qkv GEMM, converted ONCE (the GEMM's
// converting producers would redo it for each of the eight 256-column chunks of the 2048 outputs).
// Token-row source: one thread per four channels.
__global__ void __launch_bounds__(256) kmax_split_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int K) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= rows * (K / 4)) return;
  const long long r = e / (K / 4);
  const int c4 = (int)(e - r * (K / 4));
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * K) + c4);
  uint2 h, l;
  h.x = pack_bf16x2(v.x, v.y); h.y = pack_bf16x2(v.z, v.w);
  const float2 h0 = unpack_bf16x2(h.x), h1 = unpack_bf16x2(h.y);
  l.x = pack_bf16x2(v.x - h0.x, v.y - h0.y); l.y = pack_bf16x2(v.z - h1.x, v.w - h1.y);
  uint2* o = reinterpret_cast<uint2*>(out + r * 2 * K);
  o[c4] = h;
  o[K / 4 + c4] = l;
}
// NCHW source [images][K][HW]: a 64-channel x 32-pixel tile goes through shared memory (reads along pixels, writes along channels).
__global__ void __launch_bounds__(256) kmax_split_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int K, int HW) {
  __shared__ float tile[64][33];
  const int img = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < 64; c += 8) tile[c][lane] = p0 + lane < HW ? __ldg(x + ((size_t)img * K + c0 + c) * HW + p0 + lane) : 0.f;
  __syncthreads();
  for (int px = w; px < 32; px += 8) {
    if (p0 + px >= HW) break;
    const float a = tile[2 * lane][px], b = tile[2 * lane + 1][px];
    const uint32_t h = pack_bf16x2(a, b);
    const float2 hf = unpack_bf16x2(h);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + ((size_t)img * HW + p0 + px) * 2 * K);
    o[c0 / 2 + lane] = h;
    o[(K + c0) / 2 + lane] = pack_bf16x2(a - hf.x, b - hf.y);
  }
}

__host__ __device__ inline size_t kmax_axial_smem_bytes(int L, int dk, int dv) {
  const int pk = dk + 4, pv = dv + 4, R = 2 * L - 1;
  return sizeof(float) * ((size_t)2 * L * pk + (size_t)L * pv + (size_t)2 * R * pk + (size_t)R * pv + (size_t)L * (L + 1));
}

// Long axes (L > 64 at the default depths dk = 64, dv = 128): the logits phase needs q, k, rq, rk and the retrieval phase v, rv -- never
// both -- so the two operand sets share one region (the larger of the two) beside the L x (L + 1) weights; the price is re-staging the
// embedding rows for every item (from L2) instead of once per CTA.
__host__ __device__ inline size_t kmax_axial_smem_bytes_overlay(int L, int dk, int dv) {
  const int pk = dk + 4, pv = dv + 4, R = 2 * L - 1;
  const size_t a = (size_t)2 * L * pk + (size_t)2 * R * pk, b = (size_t)L * pv + (size_t)R * pv;
  return sizeof(float) * ((a > b ? a : b) + (size_t)L * (L + 1));
}

__global__ void __launch_bounds__(KA_THREADS) kmax_axial_attn_kernel(const KmaxAxialParams p) {
  extern __shared__ float ka_smem[];
  const int L = p.L, dk = p.dk, dv = p.dv, pk = dk + 4, pv = dv + 4, R = 2 * L - 1;
  float* sq = ka_smem;                    // [L][pk]
  float* sk = sq + L * pk;                // [L][pk]
  float* sv = sk + L * pk;                // [L][pv]
  float* rq = sv + L * pv;                // [R][pk]   row r = relative distance m - l + (L - 1)
  float* rk = rq + R * pk;                // [R][pk]
  float* rv = rk + R * pk;                // [R][pv]
  float* sw = rv + R * pv;                // [L][L + 1] logits -> weights
  if (p.overlay) {                        // key side [sq | sk | rq | rk] and value side [sv | rv] on the same words, the weights behind the larger
    rq = sk + L * pk;
    rk = rq + R * pk;
    sv = ka_smem;
    rv = sv + L * pv;
    const size_t a = (size_t)2 * L * pk + (size_t)2 * R * pk, b = (size_t)L * pv + (size_t)R * pv;
    sw = ka_smem + (a > b ? a : b);
  }
  const int tid = threadIdx.x;
  const int Kd = p.heads * dk;

  // ---- the addressable rows of the three embedding tables: shared by every head and sequence, staged once per (persistent) CTA
  const int e0 = KA_MAX_SPAN - 1 - (L - 1);                        // embedding row of relative distance -(L - 1)
  auto stage_key_tables = [&]() {
    for (int e = tid; e < R * (dk / 4); e += KA_THREADS) {
      const int r = e / (dk / 4), d4 = e - r * (dk / 4);
      reinterpret_cast<float4*>(rq + r * pk)[d4] = __ldg(reinterpret_cast<const float4*>(p.emb_q + (size_t)(e0 + r) * dk) + d4);
      reinterpret_cast<float4*>(rk + r * pk)[d4] = __ldg(reinterpret_cast<const float4*>(p.emb_k + (size_t)(e0 + r) * dk) + d4);
    }
  };
  auto stage_value_table = [&]() {
    for (int e = tid; e < R * (dv / 4); e += KA_THREADS) {
      const int r = e / (dv / 4), d4 = e - r * (dv / 4);
      reinterpret_cast<float4*>(rv + r * pv)[d4] = __ldg(reinterpret_cast<const float4*>(p.emb_v + (size_t)(e0 + r) * dv) + d4);
    }
  };
  if (!p.overlay) { stage_key_tables(); stage_value_table(); }

  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
  const int h = item % p.heads, s = item / p.heads;
  const long long row0 = (long long)(s / p.seq_inner) * p.row_outer + (long long)(s % p.seq_inner) * p.row_inner;

  // ---- stage q, k, v of this head (16-byte loads: a head's slice of a token row is contiguous)
  for (int e = tid; e < L * (dk / 4); e += KA_THREADS) {
    const int l = e / (dk / 4), d4 = e - l * (dk / 4);
    const float* row = p.qkv + (size_t)(row0 + l * p.row_pos) * p.ld;
    reinterpret_cast<float4*>(sq + l * pk)[d4] = __ldg(reinterpret_cast<const float4*>(row + h * dk) + d4);
    reinterpret_cast<float4*>(sk + l * pk)[d4] = __ldg(reinterpret_cast<const float4*>(row + Kd + h * dk) + d4);
  }
  auto stage_v = [&]() {
    for (int e = tid; e < L * (dv / 4); e += KA_THREADS) {
      const int l = e / (dv / 4), d4 = e - l * (dv / 4);
      reinterpret_cast<float4*>(sv + l * pv)[d4] =
          __ldg(reinterpret_cast<const float4*>(p.qkv + (size_t)(row0 + l * p.row_pos) * p.ld + 2 * Kd + h * dv) + d4);
    }
  };
  if (p.overlay) stage_key_tables(); else stage_v();
  __syncthreads();

  // ---- similarity logits: three dot products per (l, m), each through its own batch-norm affine        (:137-145)
  const float s0 = p.sim_s[h], t0 = p.sim_t[h], s1 = p.sim_s[p.heads + h], t1 = p.sim_t[p.heads + h];
  const float s2 = p.sim_s[2 * p.heads + h], t2 = p.sim_t[2 * p.heads + h];
  // item = (relative distance r, block of four query positions): the four entries (l0 + j, l0 + j + r - (L - 1)) share the rows rq[r] and
  // rk[r]; lanes take consecutive r, so the k / rq / rk rows of a warp are consecutive (conflict-free float4 reads) and q is a broadcast
  const int LB = (L + 3) / 4;
  for (int e = tid; e < R * LB; e += KA_THREADS) {
    const int lb = e / R, r = e - lb * R;
    const int l0 = lb * 4, m0 = l0 + r - (L - 1);
    if (m0 + 3 < 0 || m0 >= L) continue;                          // the whole strip lies outside the L x L square
    const float4* a4 = reinterpret_cast<const float4*>(rq + r * pk);
    const float4* b4 = reinterpret_cast<const float4*>(rk + r * pk);
    const float4* q4[4];
    const float4* k4[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + j, m = m0 + j;
      ok[j] = l < L && m >= 0 && m < L;
      q4[j] = reinterpret_cast<const float4*>(sq + (ok[j] ? l : 0) * pk);
      k4[j] = reinterpret_cast<const float4*>(sk + (ok[j] ? m : 0) * pk);
    }
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d = 0; d < dk / 4; ++d) {
      const float4 A = a4[d], B = b4[d];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 Q = q4[j][d], K = k4[j][d];
        c0[j] = fmaf(Q.x, K.x, fmaf(Q.y, K.y, fmaf(Q.z, K.z, fmaf(Q.w, K.w, c0[j]))));
        c1[j] = fmaf(Q.x, A.x, fmaf(Q.y, A.y, fmaf(Q.z, A.z, fmaf(Q.w, A.w, c1[j]))));
        c2[j] = fmaf(K.x, B.x, fmaf(K.y, B.y, fmaf(K.z, B.z, fmaf(K.w, B.w, c2[j]))));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ok[j]) sw[(l0 + j) * (L + 1) + m0 + j] = fmaf(c0[j], s0, t0) + fmaf(c1[j], s1, t1) + fmaf(c2[j], s2, t2);
  }
  __syncthreads();

  // ---- softmax over m, one warp per row                                                              (:147-148)
  for (int l = tid >> 5; l < L; l += KA_THREADS / 32) {
    float* row = sw + l * (L + 1);
    const int lane = tid & 31;
    float x[KA_MAX_L / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < KA_MAX_L / 32; ++u) {
      x[u] = lane + 32 * u < L ? row[lane + 32 * u] : -INFINITY;
      mx = fmaxf(mx, x[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < KA_MAX_L / 32; ++u) {
      x[u] = lane + 32 * u < L ? expf(x[u] - mx) : 0.f;
      sum += x[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.f / sum;
#pragma unroll
    for (int u = 0; u < KA_MAX_L / 32; ++u)
      if (lane + 32 * u < L) row[lane + 32 * u] = x[u] * inv;
  }
  __syncthreads();
  if (p.overlay) {                                                 // the key-side operands are dead: bring in v and the value embedding rows
    stage_v();
    stage_value_table();
    __syncthreads();
  }

  // ---- retrieval: content and positional parts, their batch-norm affines, sum                        (:150-157)
  const float* os = p.out_s;
  const float* ot = p.out_t;
  const int Vd = p.heads * dv;
  const long long out0 = (long long)(s / p.seq_inner) * p.out_outer + (long long)(s % p.seq_inner) * p.out_inner;
  // item = (block of four channels, 32 positions): lanes take consecutive l, so w[l][m] is one conflict-free read, v_m a 16-byte
  // broadcast and the rv rows of a warp are consecutive; channels-first outputs are stored directly (lanes = consecutive positions)
  const int LC = (L + 31) / 32, DB = dv / 4;
  for (int item = tid >> 5; item < DB * LC; item += KA_THREADS / 32) {
    const int db = item / LC, l = (item - db * LC) * 32 + (tid & 31);
    const bool ok = l < L;
    const int lc = ok ? l : L - 1;
    const float* wrow = sw + lc * (L + 1);
    const float4* v4 = reinterpret_cast<const float4*>(sv) + db;
    const float4* r4 = reinterpret_cast<const float4*>(rv + (L - 1 - lc) * pv) + db;        // row m - l + L - 1 at m = 0
    float4 yc = make_float4(0.f, 0.f, 0.f, 0.f), yr = yc;
    for (int m = 0; m < L; ++m) {
      const float wm = wrow[m];
      const float4 V = v4[m * (pv / 4)], Rr = r4[m * (pv / 4)];
      yc.x = fmaf(wm, V.x, yc.x); yc.y = fmaf(wm, V.y, yc.y); yc.z = fmaf(wm, V.z, yc.z); yc.w = fmaf(wm, V.w, yc.w);
      yr.x = fmaf(wm, Rr.x, yr.x); yr.y = fmaf(wm, Rr.y, yr.y); yr.z = fmaf(wm, Rr.z, yr.z); yr.w = fmaf(wm, Rr.w, yr.w);
    }
    if (!ok) continue;
    const int c = h * dv + db * 4;
    float4 y;
    y.x = fmaf(yc.x, os[c], ot[c]) + fmaf(yr.x, os[Vd + c], ot[Vd + c]);
    y.y = fmaf(yc.y, os[c + 1], ot[c + 1]) + fmaf(yr.y, os[Vd + c + 1], ot[Vd + c + 1]);
    y.z = fmaf(yc.z, os[c + 2], ot[c + 2]) + fmaf(yr.z, os[Vd + c + 2], ot[Vd + c + 2]);
    y.w = fmaf(yc.w, os[c + 3], ot[c + 3]) + fmaf(yr.w, os[Vd + c + 3], ot[Vd + c + 3]);
    float* o = p.out + out0 + (long long)c * p.out_chan + (long long)l * p.out_pos;
    if (p.out_chan == 1) {
      *reinterpret_cast<float4*>(o) = y;                           // token rows: four consecutive channels
    } else {
      o[0] = y.x; o[p.out_chan] = y.y; o[2 * p.out_chan] = y.z; o[3 * p.out_chan] = y.w;
    }
  }
  __syncthreads();                                                 // q / k / v / weights are overwritten by the next item
  }
}

// ------------------------------------------------------------------------------------------------------------------------------------
// Tensor-core variant (axis length <= 48): the five small contractions of one (sequence, head) item as mma.sync m16n8k16 tiles with
// SPLIT bf16 operands (x = hi + lo, three MMAs per product: hi.hi + lo.hi + hi.lo), i.e. fp32-grade results from the bf16 tensor cores:
//   S0 = Q K^T [L x L],  QR = Q Rq^T [L x R],  KR = K Rk^T [L x R]       (R = 2L - 1 relative distances)
//   logits[l, m] = BN0(S0[l, m]) + BN1(QR[l, m - l + L - 1]) + BN2(KR[m, m - l + L - 1]),  w = softmax_m
//   yc = W V [L x dv],  yr = W' RV [L x dv]  with  W'[l, r] = w[l, r + l - (L - 1)]  (zero outside the square)
// Every operand is split ONCE when it is staged (the embedding tables once per CTA) into hi / lo images of packed bf16 pairs along the
// contraction index (V and RV transposed, W' materialised by the softmax pass), so a fragment register is one 32-bit shared-memory load.
// Row strides are odd multiples of 4 words: the (row g, word t) fragment pattern then covers the 32 banks exactly once.
__host__ __device__ inline int ka_pad4o(int n) { return n + ((4 - n % 8) + 8) % 8; }          // smallest m >= n with m % 8 == 4
__host__ __device__ inline int ka_pad8(int n) { return n + ((8 - n % 32) + 32) % 32; }        // smallest m >= n with m % 32 == 8

struct KaTcLayout { int ML, RT, RB, SK, SM, SR, LW, RW; size_t x_words, words; };
__host__ __device__ inline KaTcLayout kmax_axial_tc_layout(int L, int dk, int dv) {
  KaTcLayout y;
  y.ML = (L + 15) / 16 * 16;
  y.RT = (2 * L - 1 + 15) / 16 * 16;                   // contraction length of W' RV
  y.RB = (2 * L - 1 + 7) / 8 * 8;                      // rows of the rq / rk images (n-tiles of QR / KR)
  y.SK = ka_pad4o(dk / 2); y.SM = ka_pad4o(y.ML / 2); y.SR = ka_pad4o(y.RT / 2);
  y.LW = ka_pad8(y.ML); y.RW = ka_pad8(y.RB);
  const size_t qr = (size_t)2 * y.ML * y.RW, vt = (size_t)2 * dv * y.SM;
  y.x_words = qr > vt ? qr : vt;                       // QR | KR (fp32) during the logits, then V^T hi | lo
  y.words = (size_t)4 * y.ML * y.SK + (size_t)4 * y.RB * y.SK + (size_t)2 * dv * y.SR + y.x_words + (size_t)2 * y.ML * y.SM +
            (size_t)2 * y.ML * y.SR + (size_t)y.ML * y.LW;
  return y;
}

__device__ __forceinline__ void ka_split(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x, y);
  const float2 h = unpack_bf16x2(hi);
  lo = pack_bf16x2(x - h.x, y - h.y);
}
// c += A B^T for one 16 x 8 tile over `ksteps` 16-wide steps; A / B are hi (and lo = + *_lo words further) images with row stride S
__device__ __forceinline__ void ka_tile(float (&c)[4], const uint32_t* Ah, size_t a_lo, const uint32_t* Bh, size_t b_lo, int SA, int SB, int ksteps) {
  float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};      // three independent accumulation chains
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t* a = Ah + ks * 8;
    const uint32_t* b = Bh + ks * 8;
    const uint32_t ah[4] = {a[0], a[8 * SA], a[4], a[8 * SA + 4]};
    const uint32_t al[4] = {a[a_lo], a[a_lo + 8 * SA], a[a_lo + 4], a[a_lo + 8 * SA + 4]};
    const uint32_t bh0 = b[0], bh1 = b[4], bl0 = b[b_lo], bl1 = b[b_lo + 4];
    mma_bf16_16816(c, ah, bh0, bh1);
    mma_bf16_16816(c1, al, bh0, bh1);
    mma_bf16_16816(c2, ah, bl0, bl1);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) c[u] += c1[u] + c2[u];
}

__global__ void __launch_bounds__(KA_TC_THREADS) kmax_axial_tc_kernel(const KmaxAxialParams p) {
  extern __shared__ uint32_t ka_words[];
  const int L = p.L, dk = p.dk, dv = p.dv, R = 2 * L - 1;
  const KaTcLayout y = kmax_axial_tc_layout(L, dk, dv);
  const int ML = y.ML, RT = y.RT, RB = y.RB, SK = y.SK, SM = y.SM, SR = y.SR, LW = y.LW, RW = y.RW;
  // hi image first, lo image right behind it (same shape)
  uint32_t* qh = ka_words;                         const size_t q_lo = (size_t)ML * SK;     // [ML][SK] x 2
  uint32_t* kh = qh + 2 * q_lo;                                                              // [ML][SK] x 2
  uint32_t* rqh = kh + 2 * q_lo;                   const size_t r_lo = (size_t)RB * SK;     // [RB][SK] x 2
  uint32_t* rkh = rqh + 2 * r_lo;
  uint32_t* rvh = rkh + 2 * r_lo;                  const size_t rv_lo = (size_t)dv * SR;    // RV^T [dv][SR] x 2
  uint32_t* xw = rvh + 2 * rv_lo;                                                            // QR | KR fp32, then V^T hi | lo
  float* sQR = reinterpret_cast<float*>(xw);
  float* sKR = sQR + (size_t)ML * RW;
  uint32_t* vth = xw;                              const size_t vt_lo = (size_t)dv * SM;    // V^T [dv][SM] x 2
  uint32_t* wh = xw + y.x_words;                   const size_t w_lo = (size_t)ML * SM;     // W [ML][SM] x 2
  uint32_t* wph = wh + 2 * w_lo;                   const size_t wp_lo = (size_t)ML * SR;    // W' [ML][SR] x 2
  float* sw = reinterpret_cast<float*>(wph + 2 * wp_lo);                                     // [ML][LW] fp32 scores / weights
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int Kd = p.heads * dk, Vd = p.heads * dv;

  for (size_t e = tid; e < y.words; e += KA_TC_THREADS) ka_words[e] = 0u;
  __syncthreads();
  // ---- embedding tables, once per CTA: rq / rk rows as pairs along d; RV transposed as pairs along r
  const int e0 = KA_MAX_SPAN - 1 - (L - 1);
  for (int e = tid; e < R * (dk / 4); e += KA_TC_THREADS) {
    const int r = e / (dk / 4), d4 = e - r * (dk / 4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.emb_q + (size_t)(e0 + r) * dk) + d4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.emb_k + (size_t)(e0 + r) * dk) + d4);
    uint2 h, l;
    ka_split(a.x, a.y, h.x, l.x); ka_split(a.z, a.w, h.y, l.y);
    *reinterpret_cast<uint2*>(rqh + r * SK + 2 * d4) = h; *reinterpret_cast<uint2*>(rqh + r_lo + r * SK + 2 * d4) = l;
    ka_split(b.x, b.y, h.x, l.x); ka_split(b.z, b.w, h.y, l.y);
    *reinterpret_cast<uint2*>(rkh + r * SK + 2 * d4) = h; *reinterpret_cast<uint2*>(rkh + r_lo + r * SK + 2 * d4) = l;
  }
  for (int e = tid; e < (RT / 2) * (dv / 4); e += KA_TC_THREADS) {          // lanes = consecutive r pairs: conflict-free transposed stores
    const int d4 = e / (RT / 2), rp = e - d4 * (RT / 2);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a = 2 * rp < R ? __ldg(reinterpret_cast<const float4*>(p.emb_v + (size_t)(e0 + 2 * rp) * dv) + d4) : z;
    const float4 b = 2 * rp + 1 < R ? __ldg(reinterpret_cast<const float4*>(p.emb_v + (size_t)(e0 + 2 * rp + 1) * dv) + d4) : z;
    uint32_t h, l;
    ka_split(a.x, b.x, h, l); rvh[(4 * d4 + 0) * SR + rp] = h; rvh[rv_lo + (4 * d4 + 0) * SR + rp] = l;
    ka_split(a.y, b.y, h, l); rvh[(4 * d4 + 1) * SR + rp] = h; rvh[rv_lo + (4 * d4 + 1) * SR + rp] = l;
    ka_split(a.z, b.z, h, l); rvh[(4 * d4 + 2) * SR + rp] = h; rvh[rv_lo + (4 * d4 + 2) * SR + rp] = l;
    ka_split(a.w, b.w, h, l); rvh[(4 * d4 + 3) * SR + rp] = h; rvh[rv_lo + (4 * d4 + 3) * SR + rp] = l;
  }

  const int MT = ML / 16, NLt = (L + 7) / 8, NRt = RB / 8, NDt = dv / 8, MP = ML / 2;
  float4 pq[2], pk_[2];                                            // next item's q / k slices (two 16-byte pieces per thread)
  auto prefetch_qk = [&](int it) {
    const int hh = it % p.heads, ss = it / p.heads;
    const long long r0 = (long long)(ss / p.seq_inner) * p.row_outer + (long long)(ss % p.seq_inner) * p.row_inner;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + u * KA_TC_THREADS;
      if (e < L * (dk / 4)) {
        const int l = e / (dk / 4), d4 = e - l * (dk / 4);
        const float* row = p.qkv + (size_t)(r0 + l * p.row_pos) * p.ld;
        pq[u] = __ldg(reinterpret_cast<const float4*>(row + hh * dk) + d4);
        pk_[u] = __ldg(reinterpret_cast<const float4*>(row + Kd + hh * dk) + d4);
      }
    }
  };
  if ((int)blockIdx.x < p.n_items) prefetch_qk(blockIdx.x);
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int h = item % p.heads, s = item / p.heads;
    const long long row0 = (long long)(s / p.seq_inner) * p.row_outer + (long long)(s % p.seq_inner) * p.row_inner;
    // q / k of this item were prefetched into registers during the previous item's retrieval: split and store them now
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + u * KA_TC_THREADS;
      if (e < L * (dk / 4)) {
        const int l = e / (dk / 4), d4 = e - l * (dk / 4);
        uint2 hh, ll;
        ka_split(pq[u].x, pq[u].y, hh.x, ll.x); ka_split(pq[u].z, pq[u].w, hh.y, ll.y);
        *reinterpret_cast<uint2*>(qh + l * SK + 2 * d4) = hh; *reinterpret_cast<uint2*>(qh + q_lo + l * SK + 2 * d4) = ll;
        ka_split(pk_[u].x, pk_[u].y, hh.x, ll.x); ka_split(pk_[u].z, pk_[u].w, hh.y, ll.y);
        *reinterpret_cast<uint2*>(kh + l * SK + 2 * d4) = hh; *reinterpret_cast<uint2*>(kh + q_lo + l * SK + 2 * d4) = ll;
      }
    }
    __syncthreads();
    // V of this item: issued now, consumed after the softmax (two rows of a pair per register set)
    float4 va[2], vb[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + u * KA_TC_THREADS;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      va[u] = z; vb[u] = z;
      if (e < MP * (dv / 4)) {
        const int d4 = e / MP, mp = e - d4 * MP;
        const float* base = p.qkv + 2 * Kd + h * dv;
        if (2 * mp < L) va[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(row0 + (2 * mp) * p.row_pos) * p.ld) + d4);
        if (2 * mp + 1 < L) vb[u] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(row0 + (2 * mp + 1) * p.row_pos) * p.ld) + d4);
      }
    }

    // ---- S0, QR, KR: one 16 x 8 output tile per job
    const int T0 = MT * NLt, T1 = MT * NRt;
    for (int job = warp; job < T0 + 2 * T1; job += KA_TC_THREADS / 32) {
      int sel, i, j;
      if (job < T0) { sel = 0; i = job / NLt; j = job - i * NLt; }
      else { const int q = job - T0; sel = 1 + q / T1; const int rem = q - (sel - 1) * T1; i = rem / NRt; j = rem - i * NRt; }
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t* A = (sel == 2 ? kh : qh) + (i * 16 + g) * SK + t;
      if (sel == 0) ka_tile(c, A, q_lo, kh + (j * 8 + g) * SK + t, q_lo, SK, SK, dk / 16);
      else ka_tile(c, A, q_lo, (sel == 1 ? rqh : rkh) + (j * 8 + g) * SK + t, r_lo, SK, SK, dk / 16);
      const int r0 = i * 16 + g, c0 = j * 8 + 2 * t;
      if (sel == 0) {
        *reinterpret_cast<float2*>(sw + r0 * LW + c0) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(sw + (r0 + 8) * LW + c0) = make_float2(c[2], c[3]);
      } else {
        float* o = (sel == 1 ? sQR : sKR) + r0 * RW + c0;
        *reinterpret_cast<float2*>(o) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(o + 8 * RW) = make_float2(c[2], c[3]);
      }
    }
    __syncthreads();

    // ---- combine the three similarities (each through its batch-norm affine), softmax over m, and write W / W' as split pair images
    const float s0 = p.sim_s[h], t0 = p.sim_t[h], s1 = p.sim_s[p.heads + h], t1 = p.sim_t[p.heads + h];
    const float s2 = p.sim_s[2 * p.heads + h], t2 = p.sim_t[2 * p.heads + h];
    for (int l = warp; l < L; l += KA_TC_THREADS / 32) {
      float* row = sw + l * LW;
      float x[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int m = lane + 32 * u;
        x[u] = m < L ? fmaf(row[m], s0, t0) + fmaf(sQR[l * RW + m - l + L - 1], s1, t1) + fmaf(sKR[m * RW + m - l + L - 1], s2, t2) : -INFINITY;
      }
      float mx = fmaxf(x[0], x[1]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
      const float y0 = lane < L ? expf(x[0] - mx) : 0.f, y1 = lane + 32 < L ? expf(x[1] - mx) : 0.f;
      float sum = y0 + y1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
      const float inv = 1.f / sum;
      __syncwarp();
      row[lane] = y0 * inv;                                        // lanes >= L write zeros: LW >= 64 is not guaranteed, ML + 8 is
      if (lane + 32 < LW) row[lane + 32] = y1 * inv;
      __syncwarp();
      auto wat = [&](int m) -> float { return (m >= 0 && m < L) ? row[m] : 0.f; };
      for (int jp = lane; jp < MP; jp += 32) {                     // W pairs (2 jp, 2 jp + 1)
        uint32_t hh, ll;
        ka_split(wat(2 * jp), wat(2 * jp + 1), hh, ll);
        wh[l * SM + jp] = hh; wh[w_lo + l * SM + jp] = ll;
      }
      for (int jp = lane; jp < RT / 2; jp += 32) {                 // W' pairs: r = 2 jp, m = r + l - (L - 1)
        const int m = 2 * jp + l - (L - 1);
        uint32_t hh, ll;
        ka_split(wat(m), wat(m + 1), hh, ll);
        wph[l * SR + jp] = hh; wph[wp_lo + l * SR + jp] = ll;
      }
    }
    __syncthreads();                                               // QR / KR are dead: their area becomes V^T

    // ---- V^T hi | lo from the registers: lanes = consecutive m pairs (conflict-free transposed stores)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = tid + u * KA_TC_THREADS;
      if (e < MP * (dv / 4)) {
        const int d4 = e / MP, mp = e - d4 * MP;
        const float4 a = va[u], b = vb[u];
        uint32_t hh, ll;
        ka_split(a.x, b.x, hh, ll); vth[(4 * d4 + 0) * SM + mp] = hh; vth[vt_lo + (4 * d4 + 0) * SM + mp] = ll;
        ka_split(a.y, b.y, hh, ll); vth[(4 * d4 + 1) * SM + mp] = hh; vth[vt_lo + (4 * d4 + 1) * SM + mp] = ll;
        ka_split(a.z, b.z, hh, ll); vth[(4 * d4 + 2) * SM + mp] = hh; vth[vt_lo + (4 * d4 + 2) * SM + mp] = ll;
        ka_split(a.w, b.w, hh, ll); vth[(4 * d4 + 3) * SM + mp] = hh; vth[vt_lo + (4 * d4 + 3) * SM + mp] = ll;
      }
    }
    if (item + (int)gridDim.x < p.n_items) prefetch_qk(item + gridDim.x);      // lands while this item's retrieval runs
    __syncthreads();

    // ---- retrieval: yc = W V and yr = W' RV, one 16 x 8 output tile per job
    const long long out0 = (long long)(s / p.seq_inner) * p.out_outer + (long long)(s % p.seq_inner) * p.out_inner;
    for (int job = warp; job < MT * NDt; job += KA_TC_THREADS / 32) {
      const int i = job / NDt, j = job - i * NDt;
      float yc[4] = {0.f, 0.f, 0.f, 0.f}, yr[4] = {0.f, 0.f, 0.f, 0.f};
      const int cb = h * dv + j * 8 + 2 * t;                       // this thread's two channels; their affines load under the MMAs
      const float2 osc = *reinterpret_cast<const float2*>(p.out_s + cb), otc = *reinterpret_cast<const float2*>(p.out_t + cb);
      const float2 osr = *reinterpret_cast<const float2*>(p.out_s + Vd + cb), otr = *reinterpret_cast<const float2*>(p.out_t + Vd + cb);
      ka_tile(yc, wh + (i * 16 + g) * SM + t, w_lo, vth + (j * 8 + g) * SM + t, vt_lo, SM, SM, ML / 16);
      ka_tile(yr, wph + (i * 16 + g) * SR + t, wp_lo, rvh + (j * 8 + g) * SR + t, rv_lo, SR, SR, RT / 16);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = i * 16 + g + (u >> 1) * 8, c = h * dv + j * 8 + 2 * t + (u & 1);
        if (l < L)
          p.out[out0 + (long long)c * p.out_chan + (long long)l * p.out_pos] =
              (u & 1) ? fmaf(yc[u], osc.y, otc.y) + fmaf(yr[u], osr.y, otr.y) : fmaf(yc[u], osc.x, otc.x) + fmaf(yr[u], osr.x, otr.x);
      }
    }
    __syncthreads();                                               // operands and weights are overwritten by the next item
  }
}

}  // namespace axvs
