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
constexpr int KM_TILE_LD = KM_PT + 1;
constexpr int KM_SMEM_BYTES = (KM_LMAX * KM_D + KM_D * KM_TILE_LD + 4 * KM_PT) * 4 + (4 * KM_PT + KM_PT + KM_LMAX) * 4;
static_assert(KM_SMEM_BYTES <= 227 * 1024, "kmeans_partial_kernel shared memory");

// logits fp32 [N, L, M] (mask logits of the clip, M = T*H*W pixels), pv fp32 [N, 256, M].
// One CTA = one clip x one chunk of pixels.  Per tile of 64 pixels: (1) argmax over the L logits of each pixel (first maximum wins,
// torch.max semantics), (2) the 256 x 64 value tile is staged in shared memory with coalesced row reads, (3) thread d adds
// tile[d][p] into acc[idx[p]][d] (private column: no atomics, deterministic order).  The chunk's sums go to
// partial [N, chunks, L, 256], the per-cluster pixel counts to counts [N, chunks, L], the assignment (optional) to assign [N, M].
__global__ void __launch_bounds__(256) kmeans_partial_kernel(const float* __restrict__ logits, const float* __restrict__ pv,
                                                             float* __restrict__ partial, int* __restrict__ counts,
                                                             int* __restrict__ assign, int L, int M, int chunk_pixels) {
  extern __shared__ __align__(16) float km_smem[];
  float* acc = km_smem;                                   // [L][256]
  float* tile = acc + KM_LMAX * KM_D;                     // [256][65]
  float* bestv = tile + KM_D * KM_TILE_LD;                // [4][64]
  int* besti = reinterpret_cast<int*>(bestv + 4 * KM_PT); // [4][64]
  int* idx = besti + 4 * KM_PT;                           // [64]
  int* cnt = idx + KM_PT;                                 // [L]
  const int tid = threadIdx.x, n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  for (int i = tid; i < L * KM_D; i += 256) acc[i] = 0.f;
  for (int i = tid; i < L; i += 256) cnt[i] = 0;
  const int m_begin = chunk * chunk_pixels, m_end = min(M, m_begin + chunk_pixels);
  const float* lg = logits + (size_t)n * L * M;
  const float* pvn = pv + (size_t)n * KM_D * M;
  const int p = tid & (KM_PT - 1), seg = tid >> 6;
  const int lseg = (L + 3) >> 2, l0 = seg * lseg, l1 = min(L, l0 + lseg);
  __syncthreads();
  for (int m0 = m_begin; m0 < m_end; m0 += KM_PT) {
    const int np = min(KM_PT, m_end - m0);
    // (1) argmax, 4 segments of cluster rows per pixel
    float bv = -INFINITY;
    int bi = l0 < L ? l0 : 0;
    if (p < np) {
      for (int l = l0; l < l1; ++l) {
        const float x = __ldg(lg + (size_t)l * M + m0 + p);
        if (x > bv) { bv = x; bi = l; }
      }
    }
    bestv[seg * KM_PT + p] = bv;
    besti[seg * KM_PT + p] = bi;
    // (2) value tile
    for (int i = tid; i < KM_D * KM_PT; i += 256) {
      const int d = i >> 6, pp = i & (KM_PT - 1);
      tile[d * KM_TILE_LD + pp] = pp < np ? __ldg(pvn + (size_t)d * M + m0 + pp) : 0.f;
    }
    __syncthreads();
    if (tid < np) {
      float v0 = bestv[tid];
      int i0 = besti[tid];
#pragma unroll
      for (int s = 1; s < 4; ++s) {
        const float vs = bestv[s * KM_PT + tid];
        if (vs > v0) { v0 = vs; i0 = besti[s * KM_PT + tid]; }
      }
      idx[tid] = i0;
      atomicAdd(&cnt[i0], 1);
      if (assign) assign[(size_t)n * M + m0 + tid] = i0;
    }
    __syncthreads();
    // (3) accumulate: thread d owns column d of acc
    for (int pp = 0; pp < np; ++pp) acc[idx[pp] * KM_D + tid] += tile[tid * KM_TILE_LD + pp];
    __syncthreads();
  }
  float* po = partial + ((size_t)n * chunks + chunk) * L * KM_D;
  for (int i = tid; i < L * KM_D; i += 256) po[i] = acc[i];
  int* co = counts + ((size_t)n * chunks + chunk) * L;
  for (int i = tid; i < L; i += 256) co[i] = cnt[i];
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
