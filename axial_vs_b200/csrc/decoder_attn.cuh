// Clip-level kMaX decoder attention, query side:
//   * query_self_attn_kernel : AttentionOperation.forward, Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py:49-71
//                              (logits q.k, per-head BatchNorm on the logits, fp32 softmax, retrieved value, per-channel BN, GELU)
//   * kmeans_partial_kernel + kmeans_reduce_kernel : the k-means cross-attention update of kMaXTransformerLayer.forward, same file :196-208
//                              (argmax over the L cluster centres per pixel, one-hot, 'blm,bdm->bdl' sum of pixel values per cluster)
// Both are fp32 end to end, as the reference runs them (it disables autocast around the softmax and the update).
// The attention is tiny (L = 128 queries); the k-means update is HBM-bound: it reads L + D floats per pixel once.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "attn.cuh"

namespace axvs {

constexpr int QSA_DK = 16;   // key depth per head   (base_filters * key_expansion / heads = 128 / 8)
constexpr int QSA_DV = 32;   // value depth per head (256 / 8)

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// q, k fp32 [N, heads*16, L]; v fp32 [N, heads*32, L] (channel = head*depth + d, the reference's reshape at :215-217).
// sim_affine [2*heads] = (scale, shift) of the eval-mode BN on the logits, val_affine [2*heads*32] likewise per output channel.
// out fp32 [N, heads*32, L].  grid (heads, N), 128 threads, one query per thread; k and v of the head staged in shared memory.
__global__ void __launch_bounds__(128) query_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                              const float* __restrict__ sim_affine, const float* __restrict__ val_affine,
                                                              float* __restrict__ out, int heads, int L) {
  extern __shared__ __align__(16) float qsa_smem[];
  float* sk = qsa_smem;                  // [16][L]
  float* sv = qsa_smem + QSA_DK * L;     // [32][L]
  const int h = blockIdx.x, n = blockIdx.y;
  const float* qh = q + ((size_t)n * heads + h) * QSA_DK * L;
  const float* kh = k + ((size_t)n * heads + h) * QSA_DK * L;
  const float* vh = v + ((size_t)n * heads + h) * QSA_DV * L;
  for (int i = threadIdx.x; i < QSA_DK * L; i += blockDim.x) sk[i] = __ldg(kh + i);
  for (int i = threadIdx.x; i < QSA_DV * L; i += blockDim.x) sv[i] = __ldg(vh + i);
  __syncthreads();
  const float a = __ldg(sim_affine + 2 * h), c = __ldg(sim_affine + 2 * h + 1);
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float qr[QSA_DK];
#pragma unroll
    for (int d = 0; d < QSA_DK; ++d) qr[d] = __ldg(qh + d * L + l);
    float mx = -INFINITY;
    for (int m = 0; m < L; ++m) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < QSA_DK; ++d) s = fmaf(qr[d], sk[d * L + m], s);
      mx = fmaxf(mx, fmaf(s, a, c));
    }
    float acc[QSA_DV], den = 0.f;
#pragma unroll
    for (int d = 0; d < QSA_DV; ++d) acc[d] = 0.f;
    for (int m = 0; m < L; ++m) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < QSA_DK; ++d) s = fmaf(qr[d], sk[d * L + m], s);
      const float p = expf(fmaf(s, a, c) - mx);
      den += p;
#pragma unroll
      for (int d = 0; d < QSA_DV; ++d) acc[d] = fmaf(p, sv[d * L + m], acc[d]);
    }
    const float inv = 1.f / den;
    float* o = out + ((size_t)n * heads + h) * QSA_DV * L + l;
#pragma unroll
    for (int d = 0; d < QSA_DV; ++d) {
      const int ch = h * QSA_DV + d;
      o[(size_t)d * L] = gelu_erf_f(fmaf(acc[d] * inv, __ldg(val_affine + 2 * ch), __ldg(val_affine + 2 * ch + 1)));
    }
  }
}

// ---- k-means update --------------------------------------------------------------------------------------------------
constexpr int KM_D = 256;        // total value depth
constexpr int KM_PT = 64;        // pixels per tile
constexpr int KM_LMAX = 128;     // cluster centres (queries)
constexpr int KM_THREADS = 512;
constexpr int KM_ROWS = KM_THREADS / (KM_PT / 2);   // tile rows covered by one load sweep of the CTA (one pixel pair per thread)
constexpr int KM_SEGS = 4;                          // cluster-row segments of the argmax (threads 0..255)
constexpr int KM_LREG = KM_LMAX / KM_ROWS;          // staged logits pairs per thread
constexpr int KM_VREG = KM_D / KM_ROWS;             // staged value pairs per thread
constexpr int KM_SMEM_BYTES = (KM_LMAX * KM_D + KM_LMAX * KM_PT + KM_D * KM_PT) * 4 + (2 * KM_SEGS * KM_PT + KM_PT + KM_LMAX) * 4;
static_assert(KM_SMEM_BYTES <= 227 * 1024, "kmeans_partial_kernel shared memory");

template <bool VEC>
__device__ __forceinline__ float2 km_load_pair(const float* row, int m, int m_end) {
  if (VEC) {                                        // M even: every pixel pair of every row is 8-byte aligned
    return m < m_end ? __ldg(reinterpret_cast<const float2*>(row + m)) : make_float2(0.f, 0.f);
  } else {
    float2 r;
    r.x = m < m_end ? __ldg(row + m) : 0.f;
    r.y = m + 1 < m_end ? __ldg(row + m + 1) : 0.f;
    return r;
  }
}

// logits fp32 [N, L, M] (mask logits of the clip, M = T*H*W pixels), pv fp32 [N, 256, M].
// One CTA = one clip x one chunk of pixels, 64-pixel tiles.  The next tile ([L][64] logits, [256][64] values) is prefetched into
// registers (24 x 8-byte loads per thread, 96 KiB in flight per SM) while the current one is processed from shared memory:
//   argmax over the L logits of each pixel (first maximum wins, torch.max semantics), then thread (d, g) adds tile[d][p] into
//   acc[idx[p]][d] for the clusters with idx & 1 == g (private cells: no atomics, deterministic order).
// (4-byte cp.async was measured 3x slower here: LDGSTS.32 issue clogs the load/store unit in front of the shared-memory
// read-modify-writes.)  The chunk's sums go to partial [N, chunks, L, 256], the per-cluster pixel counts to
// counts [N, chunks, L], the assignment (optional) to assign [N, M].
template <bool VEC>
__global__ void __launch_bounds__(KM_THREADS) kmeans_partial_kernel(const float* __restrict__ logits, const float* __restrict__ pv,
                                                                    float* __restrict__ partial, int* __restrict__ counts,
                                                                    int* __restrict__ assign, int L, int M, int chunk_pixels) {
  extern __shared__ __align__(16) float km_smem[];
  float* acc = km_smem;                                           // [L][256]
  float* ltile = acc + KM_LMAX * KM_D;                            // [L][64]
  float* tile = ltile + KM_LMAX * KM_PT;                          // [256][32 pairs], pair index XOR (d & 31): conflict-free both ways
  float* bestv = tile + KM_D * KM_PT;                             // [segs][64]
  int* besti = reinterpret_cast<int*>(bestv + KM_SEGS * KM_PT);   // [segs][64]
  int* idx = besti + KM_SEGS * KM_PT;                             // [64]
  int* cnt = idx + KM_PT;                                         // [L]
  uint8_t* idx8 = reinterpret_cast<uint8_t*>(idx);                // cluster id per pixel of the tile, one byte each
  const int tid = threadIdx.x, n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int m_begin = chunk * chunk_pixels, m_end = min(M, m_begin + chunk_pixels);
  const float* lg = logits + (size_t)n * L * M;
  const float* pvn = pv + (size_t)n * KM_D * M;
  const int pr = tid & 31, row = tid >> 5;                        // load mapping: pixel pair, first row
  const int p = tid & (KM_PT - 1), seg = tid >> 6;                // argmax mapping
  const int d = tid & (KM_D - 1), g = tid >> 8;                   // accumulate mapping

  float2 rl[KM_LREG], rv[KM_VREG];
  auto prefetch = [&](int m0) {
    const int m = m0 + 2 * pr;
#pragma unroll
    for (int i = 0; i < KM_LREG; ++i) {
      const int l = row + i * KM_ROWS;
      rl[i] = l < L ? km_load_pair<VEC>(lg + (size_t)l * M, m, m_end) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < KM_VREG; ++i) rv[i] = km_load_pair<VEC>(pvn + (size_t)(row + i * KM_ROWS) * M, m, m_end);
  };
  prefetch(m_begin);
  for (int i = tid; i < L * KM_D; i += KM_THREADS) acc[i] = 0.f;
  for (int i = tid; i < L; i += KM_THREADS) cnt[i] = 0;

  const int lseg = (L + KM_SEGS - 1) / KM_SEGS, l0 = seg * lseg, l1 = min(L, l0 + lseg);
#ifdef AXVS_KM_PROFILE
  long long tp[5] = {0, 0, 0, 0, 0}, t0 = clock64(), t1;
#define KM_TICK(i) do { t1 = clock64(); tp[i] += t1 - t0; t0 = t1; } while (0)
#else
#define KM_TICK(i)
#endif
  for (int m0 = m_begin; m0 < m_end; m0 += KM_PT) {
    const int np = min(KM_PT, m_end - m0);
#pragma unroll
    for (int i = 0; i < KM_LREG; ++i) *reinterpret_cast<float2*>(ltile + (row + i * KM_ROWS) * KM_PT + 2 * pr) = rl[i];
#pragma unroll
    for (int i = 0; i < KM_VREG; ++i) {
      const int dd = row + i * KM_ROWS;
      *reinterpret_cast<float2*>(tile + dd * KM_PT + ((pr ^ (dd & 31)) << 1)) = rv[i];
    }
    __syncthreads();
    KM_TICK(0);
    if (m0 + KM_PT < m_end) prefetch(m0 + KM_PT);
    KM_TICK(1);
    if (seg < KM_SEGS) {
      float bv = -INFINITY;
      int bi = l0 < L ? l0 : 0;
      if (p < np) {
#pragma unroll 8
        for (int l = l0; l < l1; ++l) {
          const float x = ltile[l * KM_PT + p];
          if (x > bv) { bv = x; bi = l; }
        }
      }
      bestv[seg * KM_PT + p] = bv;
      besti[seg * KM_PT + p] = bi;
    }
    __syncthreads();
    if (tid < np) {
      float v0 = bestv[tid];
      int i0 = besti[tid];
#pragma unroll
      for (int s = 1; s < KM_SEGS; ++s) {
        const float vs = bestv[s * KM_PT + tid];
        if (vs > v0) { v0 = vs; i0 = besti[s * KM_PT + tid]; }
      }
      idx8[tid] = (uint8_t)i0;
      atomicAdd(&cnt[i0], 1);
      if (assign) assign[(size_t)n * M + m0 + tid] = i0;
    } else if (tid < KM_PT) {
      idx8[tid] = 0xFF;                                           // beyond the chunk: matches neither accumulate group
    }
    __syncthreads();
    KM_TICK(2);
    // cluster ids of the tile as 64 bytes in registers (0xFF = beyond the chunk); the unrolled loop then has no dependent
    // shared-memory address loads in front of each read-modify-write
    const int dsw = d & 31;
    uint32_t lw[KM_PT / 4];
#pragma unroll
    for (int i = 0; i < KM_PT / 16; ++i) {
      const uint4 u = reinterpret_cast<const uint4*>(idx8)[i];
      lw[4 * i] = u.x; lw[4 * i + 1] = u.y; lw[4 * i + 2] = u.z; lw[4 * i + 3] = u.w;
    }
#pragma unroll
    for (int j = 0; j < KM_PT / 2; ++j) {
      const uint32_t la = (lw[j >> 1] >> (16 * (j & 1))) & 0xFFu, lb = (lw[j >> 1] >> (16 * (j & 1) + 8)) & 0xFFu;
      const bool a = la < (uint32_t)KM_LMAX && (la & 1u) == (uint32_t)g, b = lb < (uint32_t)KM_LMAX && (lb & 1u) == (uint32_t)g;
      if (a || b) {
        const float2 v = *reinterpret_cast<const float2*>(tile + d * KM_PT + ((j ^ dsw) << 1));
        if (a && b && la == lb) {
          acc[la * KM_D + d] += v.x + v.y;
        } else {
          if (a) acc[la * KM_D + d] += v.x;
          if (b) acc[lb * KM_D + d] += v.y;
        }
      }
    }
    __syncthreads();
    KM_TICK(3);
  }
#ifdef AXVS_KM_PROFILE
  if (tid == 0 && blockIdx.x == 1 && blockIdx.y == 1)
    printf("kmeans phases (clk) M=%d: store+wait %lld prefetch_issue %lld argmax %lld accumulate %lld\n", M, tp[0], tp[1], tp[2], tp[3]);
#endif
  float* po = partial + ((size_t)n * chunks + chunk) * L * KM_D;
  for (int i = tid; i < L * KM_D; i += KM_THREADS) po[i] = acc[i];
  int* co = counts + ((size_t)n * chunks + chunk) * L;
  for (int i = tid; i < L; i += KM_THREADS) co[i] = cnt[i];
}

// out[n, d, l] = sum_chunk partial[n, chunk, l, d]  (/ max(count[n, l], 1) when advanced).  32 x 32 transposing tiles.
__global__ void __launch_bounds__(256) kmeans_reduce_kernel(const float* __restrict__ partial, const int* __restrict__ counts,
                                                            float* __restrict__ out, int chunks, int L, int advanced) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, l_base = blockIdx.y * 32, d_base = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int l = l_base + r;
    float s = 0.f;
    if (l < L) {
      for (int c = 0; c < chunks; ++c) s += __ldg(partial + (((size_t)n * chunks + c) * L + l) * KM_D + d_base + tx);
      if (advanced) {
        int cn = 0;
        for (int c = 0; c < chunks; ++c) cn += __ldg(counts + ((size_t)n * chunks + c) * L + l);
        s = s / fmaxf((float)cn, 1.f);
      }
    }
    t[r][tx] = s;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int d = d_base + r, l = l_base + tx;
    if (l < L) out[((size_t)n * KM_D + d) * L + l] = t[tx][r];
  }
}

}  // namespace axvs
