// Within-clip input / output projections: Conv2d 1x1 + GroupNorm(32) on either side of the
// transformer encoder -- WC/msdeformattn.py:355-375 (definition), :413-416 (input side), :432-434 (output side).
// The 1x1 convolutions run on the generic tcgen05 GEMM (gemm.cuh): the input side reads the NCHW backbone feature directly
// (transposing fp32 -> bf16 producers, a_diag = 3) and writes token-major rows [image, pixel, 256] -- the layout the
// temporal layers consume, so the reference's flatten(2).transpose(1,2) copy disappears; the output side reads token rows
// (a_diag = 4) and writes NCHW.  GroupNorm is applied in place by the kernels below (deterministic: fixed-order reductions).
#pragma once
#include "simt.cuh"

namespace axvs {

constexpr int GN_GROUPS = 32;
constexpr int GN_CHUNKS = 8;          // pixel chunks per image in the token-major statistics pass

// ---- token-major [images, HW, 256] (img_stride floats between images: HW * 256, or more when the level is a slice of the multi-level
// token tensor): group g = channels 8g..8g+7 of every pixel --------------------------------------------
// pass 1: partial (sum, sumsq) per (image, chunk, group); thread t: channel quad t & 63, pixel slot t >> 6 (4 pixels per sweep)
__global__ void __launch_bounds__(256) gn_tokens_stats_kernel(const float* __restrict__ x, float2* __restrict__ partial, int HW, long long img_stride) {
  __shared__ float2 red[4][64];
  const int img = blockIdx.y, chunk = blockIdx.x;
  const int p0 = (int)(((long long)HW * chunk) / GN_CHUNKS), p1 = (int)(((long long)HW * (chunk + 1)) / GN_CHUNKS);
  const int cq = threadIdx.x & 63, slot = threadIdx.x >> 6;
  const float* base = x + (size_t)img * (size_t)img_stride + cq * 4;
  float s = 0.f, q = 0.f;
  for (int p = p0 + slot; p < p1; p += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * C256));
    s += v.x + v.y + v.z + v.w;
    q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  red[slot][cq] = make_float2(s, q);
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {                  // group = channel quads 2g, 2g+1; slots in fixed order
    const int g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      ts += red[sl][2 * g].x + red[sl][2 * g + 1].x;
      tq += red[sl][2 * g].y + red[sl][2 * g + 1].y;
    }
    partial[((size_t)img * GN_CHUNKS + chunk) * GN_GROUPS + g] = make_float2(ts, tq);
  }
}
// pass 2: y = (x - mean_g) * rstd_g * gamma_c + beta_c in place
__global__ void __launch_bounds__(256) gn_tokens_apply_kernel(float* __restrict__ x, const float2* __restrict__ partial, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int HW, float eps, long long img_stride) {
  __shared__ float2 stat[GN_GROUPS];               // (mean, rstd)
  const int img = blockIdx.y, chunk = blockIdx.x;
  if (threadIdx.x < GN_GROUPS) {
    float ts = 0.f, tq = 0.f;
    for (int c = 0; c < GN_CHUNKS; ++c) {
      const float2 v = partial[((size_t)img * GN_CHUNKS + c) * GN_GROUPS + threadIdx.x];
      ts += v.x; tq += v.y;
    }
    const float cnt = 8.f * (float)HW;
    const float mean = ts / cnt;
    const float var = fmaxf(tq / cnt - mean * mean, 0.f);
    stat[threadIdx.x] = make_float2(mean, rsqrtf(var + eps));
  }
  __syncthreads();
  const int p0 = (int)(((long long)HW * chunk) / GN_CHUNKS), p1 = (int)(((long long)HW * (chunk + 1)) / GN_CHUNKS);
  const int cq = threadIdx.x & 63, slot = threadIdx.x >> 6;
  const float2 st = stat[cq >> 1];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + cq), be = __ldg(reinterpret_cast<const float4*>(beta) + cq);
  float* base = x + (size_t)img * (size_t)img_stride + cq * 4;
  for (int p = p0 + slot; p < p1; p += 4) {
    float4 v = *reinterpret_cast<float4*>(base + (size_t)p * C256);
    v.x = (v.x - st.x) * st.y * ga.x + be.x; v.y = (v.y - st.x) * st.y * ga.y + be.y;
    v.z = (v.z - st.x) * st.y * ga.z + be.z; v.w = (v.w - st.x) * st.y * ga.w + be.w;
    *reinterpret_cast<float4*>(base + (size_t)p * C256) = v;
  }
}

// ---- NCHW [images, C, HW]: group g = the contiguous block of (C/32) * HW floats; one CTA per (image, group), two passes --------
// (the second pass re-reads what the first just streamed: with ~8 CTAs per SM the block is usually still in L2).  HW is odd at every
// pyramid level, so channel rows are not 16-byte aligned, but the GROUP block is (C/32 is a multiple of 4): both passes walk it with
// float4 accesses, and the second one picks the (scale, shift) pair per element (a float4 can straddle two channels).
__global__ void __launch_bounds__(512) gn_nchw_kernel(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      int C, int HW, float eps) {
  __shared__ float2 red[16];
  __shared__ float2 stat;
  __shared__ float s_sc[64], s_sh[64];
  const int img = blockIdx.y, g = blockIdx.x, cpg = C / GN_GROUPS;
  const size_t n = (size_t)cpg * HW;
  float* base = x + ((size_t)img * C + (size_t)g * cpg) * HW;
  const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (n % 4 == 0) && cpg <= 64 && HW >= 4;
  float s = 0.f, q = 0.f;
  if (vec) {
    const float4* b4 = reinterpret_cast<const float4*>(base);
    for (size_t i = threadIdx.x; i < n / 4; i += 512) {
      const float4 v = b4[i];
      s += v.x + v.y + v.z + v.w;
      q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (size_t i = threadIdx.x; i < n; i += 512) {
      const float v = base[i];
      s += v; q += v * v;
    }
  }
  s = warp_sum(s); q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(s, q);
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int w = 0; w < 16; ++w) { ts += red[w].x; tq += red[w].y; }
    const float mean = ts / (float)n;
    stat = make_float2(mean, rsqrtf(fmaxf(tq / (float)n - mean * mean, 0.f) + eps));
  }
  __syncthreads();
  const float2 st = stat;
  if (vec) {
    if (threadIdx.x < cpg) {                          // per channel: one (scale, shift) pair
      const float sc = st.y * __ldg(gamma + g * cpg + threadIdx.x);
      s_sc[threadIdx.x] = sc;
      s_sh[threadIdx.x] = __ldg(beta + g * cpg + threadIdx.x) - st.x * sc;
    }
    __syncthreads();
    float4* b4 = reinterpret_cast<float4*>(base);
    for (size_t i = threadIdx.x; i < n / 4; i += 512) {
      const int e = (int)(4 * i), c = e / HW, left = (c + 1) * HW - e;      // elements of this float4 still in channel c (>= 1)
      const float sc0 = s_sc[c], sh0 = s_sh[c];
      const float sc1 = left < 4 ? s_sc[c + 1] : sc0, sh1 = left < 4 ? s_sh[c + 1] : sh0;
      float4 v = b4[i];
      v.x = v.x * sc0 + sh0;
      v.y = left > 1 ? v.y * sc0 + sh0 : v.y * sc1 + sh1;
      v.z = left > 2 ? v.z * sc0 + sh0 : v.z * sc1 + sh1;
      v.w = left > 3 ? v.w * sc0 + sh0 : v.w * sc1 + sh1;
      b4[i] = v;
    }
    return;
  }
  for (int c = 0; c < cpg; ++c) {
    const float sc = st.y * __ldg(gamma + g * cpg + c), sh = __ldg(beta + g * cpg + c) - st.x * sc;
    float* row = base + (size_t)c * HW;
    for (int i = threadIdx.x; i < HW; i += 512) row[i] = row[i] * sc + sh;
  }
}

}  // namespace axvs
