// kMaX pixel-decoder axial attention core.
// Reference: AxialAttention.forward, Vk/kmax_deeplab/modeling/pixel_decoder/kmax_pixel_decoder.py:128-157 (eval mode):
//   logits[l, m] = BN0(q_l . k_m) + BN1(q_l . rq[l, m]) + BN2(k_m . rk[l, m])        rq / rk / rv[l, m] = embedding[m - l + 254]
//   w = softmax_m(logits)                                                              (fp32, :147-148)
//   y[c, l] = BNa(sum_m w[l, m] v_m[c]) + BNb(sum_m w[l, m] rv[l, m][c])
// The 1x1 qkv convolution with its folded batch norm is the tcgen05 GEMM (gemm.cuh, fp32 token rows [q | k | v] out); this kernel is
// everything after it, in fp32.  One CTA per (head, sequence): q, k, v of the head and the 2L-1 rows of each embedding table that a
// length-L axis can address live in shared memory (rows padded to an odd word count: conflict-free column walks), the L x L weights too.
// Sequences are addressed by strides so the same kernel serves the 1-D module ([N, C, L]) and both passes of AxialAttention2D
// (height axis: token rows at stride W, token-major output feeding the width-axis GEMM; width axis: NCHW output).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace axvs {

constexpr int KA_MAX_L = 64;
constexpr int KA_MAX_SPAN = 255;
constexpr int KA_THREADS = 256;

struct KmaxAxialParams {
  const float* qkv;            // fp32 token rows [rows, 2 * H * dk + H * dv]
  int ld;                      // row length
  int L, heads, dk, dv;
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
  long long out_outer, out_inner, out_chan, out_pos;
};

__host__ __device__ inline size_t kmax_axial_smem_bytes(int L, int dk, int dv) {
  const int pk = dk + 1, pv = dv + 1, R = 2 * L - 1;
  return sizeof(float) * ((size_t)2 * L * pk + (size_t)L * pv + (size_t)2 * R * pk + (size_t)R * pv + (size_t)L * (L + 1));
}

__global__ void __launch_bounds__(KA_THREADS) kmax_axial_attn_kernel(const KmaxAxialParams p) {
  extern __shared__ float ka_smem[];
  const int L = p.L, dk = p.dk, dv = p.dv, pk = dk + 1, pv = dv + 1, R = 2 * L - 1;
  float* sq = ka_smem;                    // [L][pk]
  float* sk = sq + L * pk;                // [L][pk]
  float* sv = sk + L * pk;                // [L][pv]
  float* rq = sv + L * pv;                // [R][pk]   row r = relative distance m - l + (L - 1); rq | rk double as the output staging
  float* rk = rq + R * pk;                // [R][pk]
  float* rv = rk + R * pk;                // [R][pv]
  float* sw = rv + R * pv;                // [L][L + 1] logits -> weights
  const int h = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const int Kd = p.heads * dk;
  const long long row0 = (long long)(s / p.seq_inner) * p.row_outer + (long long)(s % p.seq_inner) * p.row_inner;

  // ---- stage q, k, v of this head and the addressable embedding rows
  for (int e = tid; e < L * dk; e += KA_THREADS) {
    const int l = e / dk, d = e - l * dk;
    const float* row = p.qkv + (size_t)(row0 + l * p.row_pos) * p.ld;
    sq[l * pk + d] = __ldg(row + h * dk + d);
    sk[l * pk + d] = __ldg(row + Kd + h * dk + d);
  }
  for (int e = tid; e < L * dv; e += KA_THREADS) {
    const int l = e / dv, d = e - l * dv;
    sv[l * pv + d] = __ldg(p.qkv + (size_t)(row0 + l * p.row_pos) * p.ld + 2 * Kd + h * dv + d);
  }
  const int e0 = KA_MAX_SPAN - 1 - (L - 1);                        // embedding row of relative distance -(L - 1)
  for (int e = tid; e < R * dk; e += KA_THREADS) {
    const int r = e / dk, d = e - r * dk;
    rq[r * pk + d] = __ldg(p.emb_q + (size_t)(e0 + r) * dk + d);
    rk[r * pk + d] = __ldg(p.emb_k + (size_t)(e0 + r) * dk + d);
  }
  for (int e = tid; e < R * dv; e += KA_THREADS) {
    const int r = e / dv, d = e - r * dv;
    rv[r * pv + d] = __ldg(p.emb_v + (size_t)(e0 + r) * dv + d);
  }
  __syncthreads();

  // ---- similarity logits: three dot products per (l, m), each through its own batch-norm affine        (:137-145)
  const float s0 = p.sim_s[h], t0 = p.sim_t[h], s1 = p.sim_s[p.heads + h], t1 = p.sim_t[p.heads + h];
  const float s2 = p.sim_s[2 * p.heads + h], t2 = p.sim_t[2 * p.heads + h];
  for (int e = tid; e < L * L; e += KA_THREADS) {
    const int l = e / L, m = e - l * L;
    const float* q = sq + l * pk;
    const float* k = sk + m * pk;
    const float* a = rq + (m - l + L - 1) * pk;
    const float* b = rk + (m - l + L - 1) * pk;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll 8
    for (int d = 0; d < dk; ++d) {
      const float qd = q[d], kd = k[d];
      c0 = fmaf(qd, kd, c0);
      c1 = fmaf(qd, a[d], c1);
      c2 = fmaf(kd, b[d], c2);
    }
    sw[l * (L + 1) + m] = fmaf(c0, s0, t0) + fmaf(c1, s1, t1) + fmaf(c2, s2, t2);
  }
  __syncthreads();

  // ---- softmax over m, one warp per row                                                              (:147-148)
  for (int l = tid >> 5; l < L; l += KA_THREADS / 32) {
    float* row = sw + l * (L + 1);
    const int lane = tid & 31;
    const float x0 = lane < L ? row[lane] : -INFINITY, x1 = lane + 32 < L ? row[lane + 32] : -INFINITY;
    float mx = fmaxf(x0, x1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    const float y0 = lane < L ? expf(x0 - mx) : 0.f, y1 = lane + 32 < L ? expf(x1 - mx) : 0.f;
    float sum = y0 + y1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    const float inv = 1.f / sum;
    if (lane < L) row[lane] = y0 * inv;
    if (lane + 32 < L) row[lane + 32] = y1 * inv;
  }
  __syncthreads();

  // ---- retrieval: content and positional parts, their batch-norm affines, sum                        (:150-157)
  const float* os = p.out_s;
  const float* ot = p.out_t;
  const int Vd = p.heads * dv;
  const long long out0 = (long long)(s / p.seq_inner) * p.out_outer + (long long)(s % p.seq_inner) * p.out_inner;
  // channels-first outputs: stage [dv][L + 1] in the (now dead) rq | rk area and store along l
  const bool pos_fast = p.out_pos == 1 && dv * (L + 1) <= 2 * R * pk;
  float* stg = rq;
  for (int e = tid; e < L * dv; e += KA_THREADS) {
    const int l = e / dv, d = e - l * dv;
    const float* wrow = sw + l * (L + 1);
    float yc = 0.f, yr = 0.f;
    for (int m = 0; m < L; ++m) {
      const float wm = wrow[m];
      yc = fmaf(wm, sv[m * pv + d], yc);
      yr = fmaf(wm, rv[(m - l + L - 1) * pv + d], yr);
    }
    const int c = h * dv + d;
    const float y = fmaf(yc, os[c], ot[c]) + fmaf(yr, os[Vd + c], ot[Vd + c]);
    if (pos_fast) stg[d * (L + 1) + l] = y;
    else p.out[out0 + (long long)c * p.out_chan + (long long)l * p.out_pos] = y;
  }
  if (pos_fast) {
    __syncthreads();
    for (int e = tid; e < L * dv; e += KA_THREADS) {
      const int d = e / L, l = e - d * L;
      p.out[out0 + (long long)(h * dv + d) * p.out_chan + l] = stg[d * (L + 1) + l];
    }
  }
}

}  // namespace axvs
