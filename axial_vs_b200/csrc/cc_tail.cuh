// Cross-clip tracking module, everything after the trajectory attention (SURVEY.md section 8 rows A9 tail and A10):
//   * aspp_tail_kernel   : channels-first LayerNorm (eps 1e-6) -> GELU -> + residual -> LayerNorm (eps 1e-5)      CC:195-201, 293-295
//   * cc_class_pool_kernel: class-activation head (256 -> 1), softmax over clips, weighted sum of class embeddings  CC:47-50
//   * mask_einsum_kernel : mask logits = pixel_feature^T . mask_kernel per clip + 1-channel BatchNorm affine,       CC:62-69
//                          written directly in the reference's final '(B T) C (V H) W -> B C (T V) H W' order.
// The three dilated temporal convolutions and the 1x1 projections run on the tcgen05 GEMM (gemm.cuh, gather mode 2).
#pragma once
#include "attn.cuh"
#include "simt.cuh"

namespace axvs {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// y [rows,256] fp32 = ASPP projection output (no bias); z = block input (residual).  One warp per row.
//   u = GELU(LN_cf(y));  out = LN(u + z)         (LN over the 256 channels of a row in both cases)
__global__ void aspp_tail_kernel(const float* __restrict__ y, const float* __restrict__ z, const float* __restrict__ g1, const float* __restrict__ b1,
                                 const float* __restrict__ g2, const float* __restrict__ b2, float* __restrict__ out32,
                                 __nv_bfloat16* __restrict__ out16, int rows, float eps1, float eps2) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    float v[8], zz[8];
    {
      const float4* y4 = reinterpret_cast<const float4*>(y + (size_t)r * C256) + lane * 2;
      const float4* z4 = reinterpret_cast<const float4*>(z + (size_t)r * C256) + lane * 2;
      const float4 a = __ldg(y4), c = __ldg(y4 + 1), d = __ldg(z4), e = __ldg(z4 + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
      zz[0] = d.x; zz[1] = d.y; zz[2] = d.z; zz[3] = d.w; zz[4] = e.x; zz[5] = e.y; zz[6] = e.z; zz[7] = e.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    float mu = warp_sum(s) * (1.f / C256), q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] -= mu; q += v[i] * v[i]; }
    float rstd = rsqrtf(warp_sum(q) * (1.f / C256) + eps1);
    s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane * 8 + i;
      v[i] = gelu_erf(v[i] * rstd * __ldg(g1 + c) + __ldg(b1 + c)) + zz[i];
      s += v[i];
    }
    mu = warp_sum(s) * (1.f / C256);
    q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] -= mu; q += v[i] * v[i]; }
    rstd = rsqrtf(warp_sum(q) * (1.f / C256) + eps2);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane * 8 + i;
      v[i] = v[i] * rstd * __ldg(g2 + c) + __ldg(b2 + c);
    }
    float4* o = reinterpret_cast<float4*>(out32 + (size_t)r * C256) + lane * 2;
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
    if (out16) {
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
      reinterpret_cast<uint4*>(out16 + (size_t)r * C256)[lane] = u;
    }
  }
}

// ce bf16 [T*Q, 256] rows (t, q).  logit[t,q] = w . ce[t,q,:] + b;  a = softmax_t(logit);  pooled[q,:] = sum_t a[t,q] ce[t,q,:]
// One warp per query; pooled written as bf16 [Q, 256] (A operand of the class head GEMM).
__global__ void cc_class_pool_kernel(const __nv_bfloat16* __restrict__ ce, const float* __restrict__ w, float b, __nv_bfloat16* __restrict__ pooled,
                                     int T, int Q) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = __ldg(w + lane * 8 + i);
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int t = 0; t < T; ++t) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(ce + ((size_t)t * Q + q) * C256) + lane);
    float x[8];
    float2 f;
    f = unpack_bf16x2(u.x); x[0] = f.x; x[1] = f.y;
    f = unpack_bf16x2(u.y); x[2] = f.x; x[3] = f.y;
    f = unpack_bf16x2(u.z); x[4] = f.x; x[5] = f.y;
    f = unpack_bf16x2(u.w); x[6] = f.x; x[7] = f.y;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] * wv[i];
    s = warp_sum(s) + b;
    const float mn = fmaxf(m, s), corr = __expf(m - mn), pe = __expf(s - mn);
    l = l * corr + pe;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = acc[i] * corr + pe * x[i];
    m = mn;
  }
  const float inv = 1.f / l;
  uint4 o;
  o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
  o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
  reinterpret_cast<uint4*>(pooled + (size_t)q * C256)[lane] = o;
}

// out[q, t, p] = bn_scale * sum_c pixel[t, c, p] * mk[(t*Q + q), c] + bn_shift        (c = 0..127, q < Q <= 128)
// pixel fp32 [T, 128, P] (channels first, as the reference holds it), mk bf16 rows (t, q) with leading dimension ld_mk,
// out fp32 [Q, T, P].  One CTA = one clip x 128 pixels: the pixel tile is converted to bf16 in shared memory (k = channel,
// n = pixel, read with ldmatrix.trans), 8 warps x 16 queries, mma.sync m16n8k16 with fp32 accumulation.  HBM-bound
// (64 KiB read + 64 KiB written per CTA).
constexpr int ME_PT = 128;
__global__ void __launch_bounds__(256) mask_einsum_kernel(const float* __restrict__ pixel, const __nv_bfloat16* __restrict__ mk, int ld_mk,
                                                          float* __restrict__ out, int T, int Q, int P, float bn_scale, float bn_shift) {
  extern __shared__ __align__(128) uint8_t me_smem[];
  // sB: [128 c][128 p] bf16, rows of 256 B = 16 chunks of 16 B, chunk index XOR (row & 7) (bank-conflict-free ldmatrix)
  // sA: [128 q][128 c] bf16, same layout
  uint8_t* sB = me_smem;
  uint8_t* sA = me_smem + 128 * 256;
  const int t = blockIdx.y;
  const int p0 = blockIdx.x * ME_PT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto off = [](int row, int chunk) { return row * 256 + ((chunk ^ (row & 7)) << 4); };
  // mask kernel tile (bf16 already): 128 rows x 16 chunks
  for (int c = tid; c < 128 * 16; c += 256) {
    const int q = c >> 4, ch = c & 15;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (q < Q) v = __ldg(reinterpret_cast<const uint4*>(mk + ((size_t)t * Q + q) * ld_mk) + ch);
    *reinterpret_cast<uint4*>(sA + off(q, ch)) = v;
  }
  // pixel tile fp32 -> bf16: 128 channels x 128 pixels; each thread converts 8 consecutive pixels per step
  const float* px = pixel + (size_t)t * 128 * P;
  const bool vec_ok = (P % 4 == 0);
  for (int c = tid; c < 128 * 16; c += 256) {
    const int ch_row = c >> 4, chunk = c & 15;
    const int p = p0 + chunk * 8;
    float x[8];
    const float* src = px + (size_t)ch_row * P + p;
    if (vec_ok && p + 8 <= P) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (p + i < P) ? __ldg(src + i) : 0.f;
    }
    uint4 u;
    u.x = pack_bf16x2(x[0], x[1]); u.y = pack_bf16x2(x[2], x[3]);
    u.z = pack_bf16x2(x[4], x[5]); u.w = pack_bf16x2(x[6], x[7]);
    *reinterpret_cast<uint4*>(sB + off(ch_row, chunk)) = u;
  }
  __syncthreads();

  float acc[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {                     // k = channel, 16 per step
    uint32_t a[4];
    {
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldmatrix_x4(a, sA + off(r, ks * 2 + (lane >> 4)));
    }
#pragma unroll
    for (int jn = 0; jn < 16; jn += 2) {               // n = pixel, 8 per tile
      uint32_t b[4];
      const int krow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldmatrix_x4_trans(b, sB + off(krow, jn + (lane >> 4)));
      mma_bf16_16816(acc[jn], a, b[0], b[1]);
      mma_bf16_16816(acc[jn + 1], a, b[2], b[3]);
    }
  }
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int q = warp * 16 + g + h * 8;
    if (q >= Q) continue;
    float* orow = out + ((size_t)q * T + t) * P + p0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int p = j * 8 + t4 * 2;
      const float v0 = acc[j][h * 2] * bn_scale + bn_shift, v1 = acc[j][h * 2 + 1] * bn_scale + bn_shift;
      if (p0 + p + 1 < P && ((P & 1) == 0)) *reinterpret_cast<float2*>(orow + p) = make_float2(v0, v1);
      else {
        if (p0 + p < P) orow[p] = v0;
        if (p0 + p + 1 < P) orow[p + 1] = v1;
      }
    }
  }
}


// Split-precision variant of mask_einsum_kernel: both operands arrive in fp32 and are split into bf16 hi / lo = bf16(x - hi) halves in
// shared memory; acc += A_hi B_hi + A_hi B_lo + A_lo B_hi reproduces the fp32 product to ~2^-17 on the bf16 tensor cores.  The kernel
// stays HBM-bound (the pixel tile is read once), so the extra MMAs are free -- and the per-pixel argmax over the queries, which decides
// the final panoptic label, no longer flips on near-ties because of operand rounding (tests/test_error_budget_cpu.py).
// `channels` (a multiple of 128) may exceed one shared-memory tile: the contraction then walks 128-channel slabs (Tube-Link mask features
// have 256 channels, the Video-kMaX ones 128).
__global__ void __launch_bounds__(256) mask_einsum_split_kernel(const float* __restrict__ pixel, const float* __restrict__ mk, int ld_mk,
                                                                float* __restrict__ out, int T, int Q, int P, float bn_scale, float bn_shift, int channels) {
  extern __shared__ __align__(128) uint8_t me_smem[];
  uint8_t* sBh = me_smem;                    // [128 c][128 p] bf16 hi
  uint8_t* sBl = me_smem + 128 * 256;        // lo
  uint8_t* sAh = me_smem + 2 * 128 * 256;    // [128 q][128 c] bf16 hi
  uint8_t* sAl = me_smem + 3 * 128 * 256;
  const int t = blockIdx.y;
  const int p0 = blockIdx.x * ME_PT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto off = [](int row, int chunk) { return row * 256 + ((chunk ^ (row & 7)) << 4); };
  auto split8 = [](const float (&x)[8], uint4& hi, uint4& lo) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = x[i] - __bfloat162float(__float2bfloat16_rn(x[i]));
    hi = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    lo = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
  };
  float acc[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  for (int k0 = 0; k0 < channels; k0 += 128) {
  if (k0) __syncthreads();                             // every warp is done with the previous slab's tiles
  for (int c = tid; c < 128 * 16; c += 256) {
    const int q = c >> 4, ch = c & 15;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 0.f;
    if (q < Q) {
      const float4* s = reinterpret_cast<const float4*>(mk + ((size_t)t * Q + q) * ld_mk + k0 + ch * 8);
      const float4 a = __ldg(s), b = __ldg(s + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(sAh + off(q, ch)) = hi;
    *reinterpret_cast<uint4*>(sAl + off(q, ch)) = lo;
  }
  const float* px = pixel + ((size_t)t * channels + k0) * P;
  const bool vec_ok = (P % 4 == 0);
  for (int c = tid; c < 128 * 16; c += 256) {
    const int ch_row = c >> 4, chunk = c & 15;
    const int p = p0 + chunk * 8;
    float x[8];
    const float* src = px + (size_t)ch_row * P + p;
    if (vec_ok && p + 8 <= P) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (p + i < P) ? __ldg(src + i) : 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    *reinterpret_cast<uint4*>(sBh + off(ch_row, chunk)) = hi;
    *reinterpret_cast<uint4*>(sBl + off(ch_row, chunk)) = lo;
  }
  __syncthreads();

#pragma unroll 2
  for (int ks = 0; ks < 8; ++ks) {                     // k = channel, 16 per step
    uint32_t ah[4], al[4];
    {
      const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldmatrix_x4(ah, sAh + off(r, ks * 2 + (lane >> 4)));
      ldmatrix_x4(al, sAl + off(r, ks * 2 + (lane >> 4)));
    }
#pragma unroll
    for (int jn = 0; jn < 16; jn += 2) {               // n = pixel, 8 per tile
      uint32_t bh[4], bl[4];
      const int krow = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      ldmatrix_x4_trans(bh, sBh + off(krow, jn + (lane >> 4)));
      ldmatrix_x4_trans(bl, sBl + off(krow, jn + (lane >> 4)));
      mma_bf16_16816(acc[jn], al, bh[0], bh[1]);       // small terms first
      mma_bf16_16816(acc[jn], ah, bl[0], bl[1]);
      mma_bf16_16816(acc[jn], ah, bh[0], bh[1]);
      mma_bf16_16816(acc[jn + 1], al, bh[2], bh[3]);
      mma_bf16_16816(acc[jn + 1], ah, bl[2], bl[3]);
      mma_bf16_16816(acc[jn + 1], ah, bh[2], bh[3]);
    }
  }
  }                                                    // 128-channel slabs
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int q = warp * 16 + g + h * 8;
    if (q >= Q) continue;
    float* orow = out + ((size_t)q * T + t) * P + p0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int p = j * 8 + t4 * 2;
      const float v0 = acc[j][h * 2] * bn_scale + bn_shift, v1 = acc[j][h * 2 + 1] * bn_scale + bn_shift;
      if (p0 + p + 1 < P && ((P & 1) == 0)) *reinterpret_cast<float2*>(orow + p) = make_float2(v0, v1);
      else {
        if (p0 + p < P) orow[p] = v0;
        if (p0 + p + 1 < P) orow[p + 1] = v1;
      }
    }
  }
}

// z[r, tap*256 + c] = x[row of (b, clamp(t + (tap - 1) * dilation, 0, T-1), q), c]: the A operand of one dilated temporal convolution
// (Conv1d k = 3, padding 'same', replicate; CC:180-182) as an fp32 matrix, for the split-precision GEMM.  One warp per row.
__global__ void aspp_gather_kernel(const float* __restrict__ x, float* __restrict__ z, int rows, int T, int Q, int dilation) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const int t = (r / Q) % T;
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
      int t2 = t + (tap - 1) * dilation;
      t2 = t2 < 0 ? 0 : (t2 >= T ? T - 1 : t2);
      const float4* s = reinterpret_cast<const float4*>(x + ((size_t)r + (ptrdiff_t)(t2 - t) * Q) * C256) + lane * 2;
      float4* d = reinterpret_cast<float4*>(z + (size_t)r * 768 + tap * C256) + lane * 2;
      d[0] = __ldg(s);
      d[1] = __ldg(s + 1);
    }
  }
}

}  // namespace axvs
