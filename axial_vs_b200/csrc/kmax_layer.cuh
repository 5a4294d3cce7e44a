// Helper kernels of the kMaX transformer decoder layer's pixel side (row A11, Video-kMaX half):
// Vk/maxtron_deeplab/modeling/transformer_decoder/maxtron_transformer_decoder.py:75-124 (kMaXPredictor) and :184-232 (kMaXTransformerLayer.forward).
// The 1x1 convolutions with their eval-mode batch norms folded in run on the library GEMMs over TOKEN ROWS [clip * pixel, channel]
// (channels-last); the reference's tensors are channel-major [N, C, TH, W].  These kernels are the layout / elementwise glue, each one
// HBM-bound pass:
//   cm_to_rows_kernel     [N, C, M] -> rows [N*M, C], optional GELU on the way (`F.gelu(pixel_feature)`, :186)        32x32 smem transposes
//   rows_to_cm_kernel     rows [N*M, ld] -> [N, C, M] for the first C channels, optional L2 normalisation over C (F.normalize, :102)
//   dwconv5_kernel        depthwise 5x5 convolution (padding 2, over the [TH, W] plane, frames stacked along H exactly as the reference
//                         stacks them) + folded batch norm + GELU on channels-last rows (:78-79, :98)
//   add_act_kernel        y = act(a + b) (the `query + update` -> GELU steps, :212-213, :219-220)
#pragma once
#include "cc_tail.cuh"

namespace axvs {


// grid (ceil(M / 32), ceil(C / 32), N), 256 threads
__global__ void __launch_bounds__(256) cm_to_rows_kernel(const float* __restrict__ x, float* __restrict__ rows, int C, int M, int act) {
  __shared__ float t[32][33];
  const int n = blockIdx.z, m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, m = m0 + tx;
    float v = 0.f;
    if (c < C && m < M) v = __ldg(x + ((size_t)n * C + c) * M + m);
    t[r][tx] = act == 2 ? gelu_erf(v) : v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, c = c0 + tx;
    if (m < M && c < C) rows[((size_t)n * M + m) * C + c] = t[tx][r];
  }
}

// grid (ceil(M / 32), N), 256 threads; C <= 256 (one warp pass per 32 channels); normalize: divide every pixel's C-vector by max(||.||_2, 1e-12)
__global__ void __launch_bounds__(256) rows_to_cm_kernel(const float* __restrict__ rows, int ld, float* __restrict__ out, int C, int M, int normalize) {
  __shared__ float t[256][33];
  __shared__ float inv[32];
  const int n = blockIdx.y, m0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {                       // pixel r of the tile: its C channels, 32 at a time (coalesced row reads)
    const int m = m0 + r;
    float ss = 0.f;
    for (int c = tx; c < C; c += 32) {
      const float v = m < M ? __ldg(rows + ((size_t)n * M + m) * ld + c) : 0.f;
      t[c][r] = v;
      ss = fmaf(v, v, ss);
    }
    if (normalize) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (tx == 0) inv[r] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
  }
  __syncthreads();
  for (int c = ty; c < C; c += 8) {
    const int m = m0 + tx;
    if (m < M) out[((size_t)n * C + c) * M + m] = normalize ? t[c][tx] * inv[tx] : t[c][tx];
  }
}

// x, y: channels-last rows [N, H, W, C]; w: [C][25] (row-major taps), affine: [C][2] = folded batch norm (scale, shift).  One thread per
// (pixel, 4 channels); C % 4 == 0.
__global__ void __launch_bounds__(256) dwconv5_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ affine,
                                                      float* __restrict__ y, int N, int H, int W, int C, int act) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int c4 = C >> 2;
  const long long total = (long long)N * H * W * c4;
  if (idx >= total) return;
  const int c = (int)(idx % c4) * 4;
  long long r = idx / c4;
  const int xw = (int)(r % W); r /= W;
  const int yh = (int)(r % H);
  const int n = (int)(r / H);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
    const int yy = yh + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int xx = xw + dx;
      if (xx < 0 || xx >= W) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + yy) * W + xx) * C + c));
      const int tap = (dy + 2) * 5 + dx + 2;
      acc.x = fmaf(v.x, __ldg(w + (c + 0) * 25 + tap), acc.x);
      acc.y = fmaf(v.y, __ldg(w + (c + 1) * 25 + tap), acc.y);
      acc.z = fmaf(v.z, __ldg(w + (c + 2) * 25 + tap), acc.z);
      acc.w = fmaf(v.w, __ldg(w + (c + 3) * 25 + tap), acc.w);
    }
  }
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(affine + 2 * c)), a1 = __ldg(reinterpret_cast<const float4*>(affine + 2 * c + 4));
  float4 o = make_float4(fmaf(acc.x, a0.x, a0.y), fmaf(acc.y, a0.z, a0.w), fmaf(acc.z, a1.x, a1.y), fmaf(acc.w, a1.z, a1.w));
  if (act == 2) o = make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w));
  *reinterpret_cast<float4*>(y + (((size_t)n * H + yh) * W + xw) * C + c) = o;
}

__global__ void __launch_bounds__(256) add_act_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n, int act) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float v = a[i] + (b ? b[i] : 0.f);
  y[i] = act == 2 ? gelu_erf(v) : (act == 1 ? fmaxf(v, 0.f) : v);
}

}  // namespace axvs
