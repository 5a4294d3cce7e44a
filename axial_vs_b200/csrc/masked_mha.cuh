// Masked multi-head attention of a FEW queries over MANY keys: the cross- / self-attention core of the Tube-Link mask decoder
// (mmcv MultiheadAttention = torch.nn.MultiheadAttention inside DetrTransformerDecoderLayer, TL/mmdet/models/utils/transformer.py:408-451,
// called at TL/models/video/tube_link_vis/mask2former_video_cc_head.py:883-894 with 100 queries, 8 heads of 32 channels, keys = T*h*w pixels
// of one pyramid level (1500 / 6000 / 24000 at T = 5, 480 x 640) and a boolean attn_mask [B*heads, 100, keys], True = blocked):
//   out[b, i, head*32 + :] = softmax_j( q[b, i, head] . k[b, j, head] + (mask[b*heads + head, i, j] ? -inf : 0) ) @ v[b, j, head]
// q arrives projected and scaled (in_proj + bias, * head_dim^-0.5 * log2 e), k / v projected.
//
// Flash-decoding split: grid (key splits, heads, batch x query blocks); thread = one query row (q, running max / sum and the 32 output
// channels in registers); the CTA's keys go through shared memory in stages of 64 (K and V rows of the head, fp32), the query's 64 mask
// bytes of a stage come straight from global memory (the mask is the largest operand: 100 x keys x 8 bytes per batch element).  Two
// passes per stage (scores + maximum, then probabilities and the weighted sum), so one rescale per 64 keys.  A second kernel combines the
// per-split (max, sum, o) triples.  fp32 SIMT: the op is bound by the mask / K / V streams, not by its 4*N*L*32 FLOPs per head.
#pragma once
#include "ptx.cuh"

namespace axvs {

constexpr int MM_D = 32;
constexpr int MM_KC = 64;
constexpr int MM_THREADS = 128;
constexpr int MM_REC = MM_D + 2;            // (max, sum, o[32]) per (split, batch, head, query)

struct MaskedMhaParams {
  const float* q;            // [B, Nq, H*32] or [Nq, B, H*32] (seq_first)
  const float* k;            // [B, L, H*32] / [L, B, H*32]
  const float* v;
  const uint8_t* mask;       // [B*H, Nq, L], non-zero = blocked; may be null
  float* partial;            // [splits][B][H][Nq][MM_REC]
  int B, H, Nq, L, keys_per_cta, qblocks, seq_first;
};

__global__ void __launch_bounds__(MM_THREADS) masked_mha_partial_kernel(const MaskedMhaParams p) {
  __shared__ __align__(16) float sk[MM_KC * MM_D];
  __shared__ __align__(16) float sv[MM_KC * MM_D];
  const int split = blockIdx.x, h = blockIdx.y;
  const int b = blockIdx.z / p.qblocks, qb = blockIdx.z - b * p.qblocks;
  const int tid = threadIdx.x;
  const int qi = qb * MM_THREADS + tid;
  const bool active = qi < p.Nq;
  const int C = p.H * MM_D;
  float q[MM_D];
  {
    const size_t row = p.seq_first ? (size_t)(active ? qi : 0) * p.B + b : (size_t)b * p.Nq + (active ? qi : 0);
    const float4* src = reinterpret_cast<const float4*>(p.q + row * C + h * MM_D);
#pragma unroll
    for (int i = 0; i < MM_D / 4; ++i) {
      const float4 t = __ldg(src + i);
      q[4 * i] = t.x; q[4 * i + 1] = t.y; q[4 * i + 2] = t.z; q[4 * i + 3] = t.w;
    }
  }
  float m = -INFINITY, l = 0.f, o[MM_D];
#pragma unroll
  for (int i = 0; i < MM_D; ++i) o[i] = 0.f;
  const int k_begin = split * p.keys_per_cta, k_end = min(p.L, k_begin + p.keys_per_cta);
  const uint8_t* mrow = p.mask ? p.mask + ((size_t)(b * p.H + h) * p.Nq + (active ? qi : 0)) * p.L : nullptr;
  const bool mvec = (p.L & 15) == 0;          // 16-byte mask loads need rows that start 16-byte aligned
  for (int k0 = k_begin; k0 < k_end; k0 += MM_KC) {
    const int nk = min(MM_KC, k_end - k0);
    __syncthreads();                             // the previous stage has been consumed
    for (int i = tid; i < MM_KC * MM_D / 4; i += MM_THREADS) {
      const int j = i >> 3, c4 = i & 7;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (j < nk) {
        const size_t row = p.seq_first ? (size_t)(k0 + j) * p.B + b : (size_t)b * p.L + k0 + j;
        kk = __ldg(reinterpret_cast<const float4*>(p.k + row * C + h * MM_D) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(p.v + row * C + h * MM_D) + c4);
      }
      reinterpret_cast<float4*>(sk)[i] = kk;
      reinterpret_cast<float4*>(sv)[i] = vv;
    }
    uint32_t mw[MM_KC / 4];                      // this query's mask bytes of the stage
    if (mrow && active) {
      if (mvec && nk == MM_KC) {
#pragma unroll
        for (int i = 0; i < MM_KC / 16; ++i) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(mrow + k0) + i);
          mw[4 * i] = u.x; mw[4 * i + 1] = u.y; mw[4 * i + 2] = u.z; mw[4 * i + 3] = u.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < MM_KC / 4; ++i) {
          uint32_t w = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * i + e;
            if (j < nk && __ldg(mrow + k0 + j)) w |= 0xFFu << (8 * e);
          }
          mw[i] = w;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < MM_KC / 4; ++i) mw[i] = 0u;
    }
    __syncthreads();
    // pass 1: scores of the stage (exp2 domain: q carries scale * log2 e) and their maximum
    float s[MM_KC];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MM_KC; ++j) {
      const float4* kr = reinterpret_cast<const float4*>(sk + j * MM_D);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < MM_D / 4; ++c) {
        const float4 t = kr[c];
        a0 = fmaf(q[4 * c], t.x, a0); a1 = fmaf(q[4 * c + 1], t.y, a1);
        a0 = fmaf(q[4 * c + 2], t.z, a0); a1 = fmaf(q[4 * c + 3], t.w, a1);
      }
      const bool blocked = j >= nk || ((mw[j >> 2] >> (8 * (j & 3))) & 0xFFu) != 0u;
      s[j] = blocked ? -INFINITY : a0 + a1;
      mx = fmaxf(mx, s[j]);
    }
    // pass 2: probabilities and the weighted sum; one rescale of the running state per stage
    if (mx != -INFINITY) {
      const float mn = fmaxf(m, mx);
      const float corr = exp2f(m - mn);          // m = -inf on the first unmasked stage: 2^-inf = 0, o and l are 0 anyway
      m = mn;
      l *= corr;
#pragma unroll
      for (int i = 0; i < MM_D; ++i) o[i] *= corr;
#pragma unroll
      for (int j = 0; j < MM_KC; ++j) {
        const float pj = exp2f(s[j] - mn);       // blocked: 2^-inf = 0
        l += pj;
        const float4* vr = reinterpret_cast<const float4*>(sv + j * MM_D);
#pragma unroll
        for (int c = 0; c < MM_D / 4; ++c) {
          const float4 t = vr[c];
          o[4 * c] = fmaf(pj, t.x, o[4 * c]); o[4 * c + 1] = fmaf(pj, t.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(pj, t.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pj, t.w, o[4 * c + 3]);
        }
      }
    }
  }
  if (active) {
    float* rec = p.partial + ((((size_t)split * p.B + b) * p.H + h) * p.Nq + qi) * MM_REC;
    rec[0] = m;
    rec[1] = l;
#pragma unroll
    for (int i = 0; i < MM_D; ++i) rec[2 + i] = o[i];
  }
}

// Per-split outputs instead of a combination: x[b, i, s, h*32 + c] = o_s[c] / l_s -- with the key splits placed on the FRAMES of a sequence
// (keys_per_cta = keys per frame) this is the per-frame-softmax spatial attention of TrajectoryAttention in fp32 (CC:104-110,
// WC/temporal_attention.py:47-60), used by the split-precision cross-clip path.  out [B, Nq, splits, H*32] (batch-first).
__global__ void __launch_bounds__(256) masked_mha_per_split_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits, int B, int H, int Nq) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)splits * B * H * Nq * MM_D;
  if (idx >= total) return;
  const int c = (int)(idx % MM_D);
  long long r = idx / MM_D;
  const int qi = (int)(r % Nq); r /= Nq;
  const int h = (int)(r % H); r /= H;
  const int b = (int)(r % B);
  const int s = (int)(r / B);
  const float* rec = partial + ((((size_t)s * B + b) * H + h) * Nq + qi) * MM_REC;
  out[(((size_t)b * Nq + qi) * splits + s) * (size_t)(H * MM_D) + h * MM_D + c] = __ldg(rec + 2 + c) / __ldg(rec + 1);
}

// out[b, i, h*32 + c] = sum_s o_s[c] 2^(m_s - M) / sum_s l_s 2^(m_s - M), M = max_s m_s; one thread per (b, h, i, c).  A row whose keys are
// ALL blocked gives 0 / 0 = NaN exactly like torch.nn.MultiheadAttention (the caller un-blocks such rows first, cc head :877-879).
__global__ void __launch_bounds__(256) masked_mha_combine_kernel(const float* __restrict__ partial, float* __restrict__ out32, __nv_bfloat16* __restrict__ out16,
                                                                 int splits, int B, int H, int Nq, int seq_first) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * H * Nq * MM_D;
  if (idx >= total) return;
  const int c = (int)(idx % MM_D);
  long long r = idx / MM_D;
  const int qi = (int)(r % Nq); r /= Nq;
  const int h = (int)(r % H);
  const int b = (int)(r / H);
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, __ldg(partial + ((((size_t)s * B + b) * H + h) * Nq + qi) * MM_REC));
  float num = 0.f, den = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* rec = partial + ((((size_t)s * B + b) * H + h) * Nq + qi) * MM_REC;
    const float ms = __ldg(rec);
    const float w = ms == -INFINITY ? 0.f : exp2f(ms - M);
    den = fmaf(__ldg(rec + 1), w, den);
    num = fmaf(__ldg(rec + 2 + c), w, num);
  }
  const float y = num / den;
  const size_t row = seq_first ? (size_t)qi * B + b : (size_t)b * Nq + qi;
  const size_t off = row * (size_t)(H * MM_D) + h * MM_D + c;
  if (out32) out32[off] = y;
  if (out16) out16[off] = __float2bfloat16(y);
}

}  // namespace axvs
