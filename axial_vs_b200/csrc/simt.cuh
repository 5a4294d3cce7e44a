// Memory-bound helper kernels: axis permute + positional add + cast, LayerNorm, temporal (over-frames) attention,
// 3-D sine positional table.  All are one-warp-per-row, 128-bit vectorised, warp-shuffle reductions.
#pragma once
#include "gemm.cuh"

namespace axvs {

constexpr int C256 = 256;   // channel count the path is specialised for (d_model of every shipped config)

// A1[p,:] = bf16(src[c(p),:] + pos[c(p),:]),  A2[p,:] = bf16(src[c(p),:])   (p = pass-order row, c = canonical token)
// Reference: the two rearranges + with_pos_embed, WC/temporal_attention.py:197-200 and :206-209.
__global__ void pack_kq_kernel(const float* __restrict__ src, const float* __restrict__ pos, __nv_bfloat16* __restrict__ a1,
                               __nv_bfloat16* __restrict__ a2, int rows, int map_mode, AxialDims d) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < rows; p += gridDim.x * wpb) {
    const size_t c = (size_t)pass_to_canonical(p, map_mode, d);
    const float4* s4 = reinterpret_cast<const float4*>(src + c * C256) + lane * 2;
    float4 s0 = __ldg(s4), s1 = __ldg(s4 + 1);
    if (a2) {
      uint4 u;
      u.x = pack_bf16x2(s0.x, s0.y); u.y = pack_bf16x2(s0.z, s0.w);
      u.z = pack_bf16x2(s1.x, s1.y); u.w = pack_bf16x2(s1.z, s1.w);
      reinterpret_cast<uint4*>(a2 + (size_t)p * C256)[lane] = u;
    }
    if (pos) {
      const float4* p4 = reinterpret_cast<const float4*>(pos + (size_t)pos_row((uint32_t)c, d) * C256) + lane * 2;
      float4 q0 = __ldg(p4), q1 = __ldg(p4 + 1);
      s0.x += q0.x; s0.y += q0.y; s0.z += q0.z; s0.w += q0.w;
      s1.x += q1.x; s1.y += q1.y; s1.z += q1.z; s1.w += q1.w;
    }
    uint4 u;
    u.x = pack_bf16x2(s0.x, s0.y); u.y = pack_bf16x2(s0.z, s0.w);
    u.z = pack_bf16x2(s1.x, s1.y); u.w = pack_bf16x2(s1.z, s1.w);
    reinterpret_cast<uint4*>(a1 + (size_t)p * C256)[lane] = u;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// y = LayerNorm(x) over 256 channels (biased variance, eps), optional fp32 and bf16 outputs.
// Reference: nn.LayerNorm norm1 / norm2, WC/temporal_attention.py:217-218,184.
__global__ void layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                    float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, int rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(g) + lane * 2 + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2), b1 = __ldg(reinterpret_cast<const float4*>(b) + lane * 2 + 1);
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float4* x4 = reinterpret_cast<const float4*>(x + (size_t)r * C256) + lane * 2;
    float4 a = __ldg(x4), c = __ldg(x4 + 1);
    float s = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
    const float mu = warp_sum(s) * (1.f / C256);
    a.x -= mu; a.y -= mu; a.z -= mu; a.w -= mu; c.x -= mu; c.y -= mu; c.z -= mu; c.w -= mu;
    float q = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C256) + eps);
    a.x = a.x * rstd * g0.x + b0.x; a.y = a.y * rstd * g0.y + b0.y; a.z = a.z * rstd * g0.z + b0.z; a.w = a.w * rstd * g0.w + b0.w;
    c.x = c.x * rstd * g1.x + b1.x; c.y = c.y * rstd * g1.y + b1.y; c.z = c.z * rstd * g1.z + b1.z; c.w = c.w * rstd * g1.w + b1.w;
    if (y32) {
      float4* o = reinterpret_cast<float4*>(y32 + (size_t)r * C256) + lane * 2;
      o[0] = a; o[1] = c;
    }
    if (y16) {
      uint4 u;
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
      u.z = pack_bf16x2(c.x, c.y); u.w = pack_bf16x2(c.z, c.w);
      reinterpret_cast<uint4*>(y16 + (size_t)r * C256)[lane] = u;
    }
  }
}

// Temporal attention over the F frames of every (token, head):  a = softmax_f(q2 . k2_f),  o = sum_f a_f v2_f.
// q2 [rows,256] (already scaled), kv2 [rows*F, 512] (k2 | v2), o [rows,256]; one thread per (row, head), d = 32.
// Reference: WC/temporal_attention.py:66-73.
__global__ void temporal_attn_kernel(const __nv_bfloat16* __restrict__ q2, const __nv_bfloat16* __restrict__ kv2,
                                     __nv_bfloat16* __restrict__ o, int rows, int F) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 8) return;
  const int row = idx >> 3, head = idx & 7;
  float q[32];
  {
    const uint4* q4 = reinterpret_cast<const uint4*>(q2 + (size_t)row * C256 + head * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(q4 + i);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      q[8 * i] = f0.x; q[8 * i + 1] = f0.y; q[8 * i + 2] = f1.x; q[8 * i + 3] = f1.y;
      q[8 * i + 4] = f2.x; q[8 * i + 5] = f2.y; q[8 * i + 6] = f3.x; q[8 * i + 7] = f3.y;
    }
  }
  float m = -INFINITY, l = 0.f, acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int f = 0; f < F; ++f) {
    const __nv_bfloat16* base = kv2 + ((size_t)row * F + f) * 512 + head * 32;
    const uint4* k4 = reinterpret_cast<const uint4*>(base);
    const uint4* v4 = reinterpret_cast<const uint4*>(base + 256);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(k4 + i);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      s += q[8 * i] * f0.x + q[8 * i + 1] * f0.y + q[8 * i + 2] * f1.x + q[8 * i + 3] * f1.y + q[8 * i + 4] * f2.x +
           q[8 * i + 5] * f2.y + q[8 * i + 6] * f3.x + q[8 * i + 7] * f3.y;
    }
    const float mn = fmaxf(m, s);
    const float corr = __expf(m - mn), pe = __expf(s - mn);
    l = l * corr + pe;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u = __ldg(v4 + i);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      acc[8 * i] = acc[8 * i] * corr + pe * f0.x;         acc[8 * i + 1] = acc[8 * i + 1] * corr + pe * f0.y;
      acc[8 * i + 2] = acc[8 * i + 2] * corr + pe * f1.x; acc[8 * i + 3] = acc[8 * i + 3] * corr + pe * f1.y;
      acc[8 * i + 4] = acc[8 * i + 4] * corr + pe * f2.x; acc[8 * i + 5] = acc[8 * i + 5] * corr + pe * f2.y;
      acc[8 * i + 6] = acc[8 * i + 6] * corr + pe * f3.x; acc[8 * i + 7] = acc[8 * i + 7] * corr + pe * f3.y;
    }
    m = mn;
  }
  const float inv = 1.f / l;
  uint4* o4 = reinterpret_cast<uint4*>(o + (size_t)row * C256 + head * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(acc[8 * i] * inv, acc[8 * i + 1] * inv);
    u.y = pack_bf16x2(acc[8 * i + 2] * inv, acc[8 * i + 3] * inv);
    u.z = pack_bf16x2(acc[8 * i + 4] * inv, acc[8 * i + 5] * inv);
    u.w = pack_bf16x2(acc[8 * i + 6] * inv, acc[8 * i + 7] * inv);
    o4[i] = u;
  }
}

// Slow path for the attention visualiser (Vk/maxtron_deeplab/maxtron_wc_model.py:598-611): materialise the per-frame
// softmax maps  maps[(seq*8 + head), q, f, i] = softmax_i(scale * Q[q] . K[f*n + i])  in fp32 (the reference's
// `space_attn`, WC/temporal_attention.py:54,76).  One warp per (seq, head, q, f); qkv head-major bf16.
__global__ void attn_maps_kernel(const __nv_bfloat16* __restrict__ qkv, size_t rows_total, float* __restrict__ maps, int num_seq, int N, int n,
                                 int F, float scale) {
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)num_seq * 8 * N * F;
  for (size_t w = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); w < total; w += (size_t)gridDim.x * (blockDim.x >> 5)) {
    const int f = (int)(w % F);
    size_t r = w / F;
    const int q = (int)(r % N); r /= N;
    const int head = (int)(r % 8);
    const int seq = (int)(r / 8);
    const __nv_bfloat16* qp = qkv + ((size_t)head * rows_total + (size_t)seq * N + q) * 32;
    const __nv_bfloat16* kp = qkv + ((size_t)(8 + head) * rows_total + (size_t)seq * N + (size_t)f * n) * 32;
    const float qv = __bfloat162float(qp[lane]);
    float* out = maps + w * n;
    float mx = -INFINITY;
    for (int i = 0; i < n; ++i) {                     // logits: lane = head-dim element, reduce over the 32 lanes
      float s = warp_sum(qv * __bfloat162float(kp[(size_t)i * 32 + lane])) * scale;
      if (lane == 0) out[i] = s;
      mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.f;
    for (int i = lane; i < n; i += 32) { const float e = __expf(out[i] - mx); out[i] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int i = lane; i < n; i += 32) out[i] *= inv;
  }
}

// pos[b,t,h,w,c] = cat(sine(y), sine(x))[c] + sine(z)[c] + level_embed[c]      (channels-last, fp32)
// Reference: PositionEmbeddingSine3D.forward (normalize=True, scale=2pi, temperature 1e4), WC/pos_embeddings.py:86-130,
// plus the level embedding add, WC/msdeformattn.py:112-115.  One thread per (t,h,w, channel pair).
__global__ void pos3d_kernel(float* __restrict__ out, const float* __restrict__ level_embed, int B, int T, int H, int W) {
  const int pairs = C256 / 2;
  const size_t total = (size_t)T * H * W * pairs;
  const float two_pi = 6.283185307179586f;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int cp = (int)(idx % pairs);
    size_t r = idx / pairs;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int t = (int)(r / H);
    const int c = cp * 2;                       // even channel: sin, odd channel: cos
    // z term: 256 channels, exponent 2*floor(c/2)/256
    const float zt = (float)(t + 1) / ((float)T + 1e-6f) * two_pi;
    const float dz = powf(10000.f, (float)c / 256.f);
    float zs, zc;
    sincosf(zt / dz, &zs, &zc);
    // y term for c < 128, x term for c >= 128: exponent 2*floor(c'/2)/128
    const int cc = c & 127;
    const float coord = (c < 128) ? (float)(h + 1) / ((float)H + 1e-6f) * two_pi : (float)(w + 1) / ((float)W + 1e-6f) * two_pi;
    const float dxy = powf(10000.f, (float)cc / 128.f);
    float s, co;
    sincosf(coord / dxy, &s, &co);
    float v0 = s + zs, v1 = co + zc;
    if (level_embed) { v0 += level_embed[c]; v1 += level_embed[c + 1]; }
    for (int b = 0; b < B; ++b) {
      float2* o = reinterpret_cast<float2*>(out + ((((size_t)b * T + t) * H + h) * W + w) * C256 + c);
      *o = make_float2(v0, v1);
    }
  }
}

}  // namespace axvs
