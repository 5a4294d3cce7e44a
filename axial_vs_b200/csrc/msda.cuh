// Multi-scale deformable attention sampling: the gather at the heart of MSDeformAttn,
// WC/ops/modules/ms_deform_attn.py:102-121 with the bilinear rule of WC/ops/functions/ms_deform_attn_func.py:51-72
// (grid_sample, align_corners = False, zero padding) == WC/ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304.
// The four Linear layers around it run on the tcgen05 GEMM; this kernel fuses the softmax over the L*P logits of a head,
// the sampling-location arithmetic and the weighted bilinear gather.
#pragma once
#include "simt.cuh"

namespace axvs {

constexpr int MSDA_MAX_LEVELS = 4;
constexpr int MSDA_MAX_LP = 16;      // levels * points per head

struct MsdaDims {
  int L, P, len;                     // levels, points, tokens per image (sum of H*W)
  int H[MSDA_MAX_LEVELS], W[MSDA_MAX_LEVELS], start[MSDA_MAX_LEVELS];
};

// value bf16 [images*len, 256] (value_proj output, channel = head*32 + j); oa fp32 [images*len, ld_oa]:
// columns [0, 8*L*P*2) = sampling offsets ordered (head, level, point, xy), then 8*L*P attention logits (head, level, point);
// ref fp32 [images*len, L, 2] normalised (x, y) reference points (ref_rows > 0: only that many rows, shared by every image);
// out bf16 [images*len, 256].
// 256 threads = 8 tokens x (8 heads x 4 eight-channel groups): one warp per token, each thread gathers 8 channels with 16-byte loads.
__global__ void __launch_bounds__(256) msda_sample_kernel(const __nv_bfloat16* __restrict__ value, const float* __restrict__ oa, int ld_oa,
                                                          const float* __restrict__ ref, int ref_rows, __nv_bfloat16* __restrict__ out, int rows, MsdaDims d) {
  const int token = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (token >= rows) return;
  const int lane = threadIdx.x & 31, head = lane >> 2, cg = lane & 3;
  const int LP = d.L * d.P;
  const float* o_row = oa + (size_t)token * ld_oa;
  const float* lg = o_row + 8 * LP * 2 + head * LP;
  float aw[MSDA_MAX_LP];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MSDA_MAX_LP; ++i) {
    aw[i] = i < LP ? __ldg(lg + i) : -INFINITY;
    mx = fmaxf(mx, aw[i]);
  }
  float den = 0.f;
#pragma unroll
  for (int i = 0; i < MSDA_MAX_LP; ++i) {
    aw[i] = i < LP ? expf(aw[i] - mx) : 0.f;
    den += aw[i];
  }
  const float inv = 1.f / den;
  const size_t img_row0 = (size_t)(token / d.len) * d.len;
  const __nv_bfloat16* vbase = value + img_row0 * C256 + head * 32 + cg * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int l = 0; l < MSDA_MAX_LEVELS; ++l) {
    if (l >= d.L) break;
    const int H = d.H[l], W = d.W[l];
    const size_t rrow = ref_rows > 0 ? (size_t)(token % ref_rows) : (size_t)token;
    const float rx = __ldg(ref + (rrow * d.L + l) * 2), ry = __ldg(ref + (rrow * d.L + l) * 2 + 1);
    const __nv_bfloat16* vl = vbase + (size_t)d.start[l] * C256;
    for (int pt = 0; pt < d.P; ++pt) {
      const float2 off = __ldg(reinterpret_cast<const float2*>(o_row + ((head * d.L + l) * d.P + pt) * 2));
      // location = ref + off / (W, H) in [0,1]; pixel coordinate = location * size - 0.5 (align_corners = False)
      const float x = (rx + off.x / (float)W) * (float)W - 0.5f, y = (ry + off.y / (float)H) * (float)H - 0.5f;
      const float xf = floorf(x), yf = floorf(y);
      const int x0 = (int)xf, y0 = (int)yf;
      const float fx = x - xf, fy = y - yf;
      const float a = aw[l * d.P + pt] * inv;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int xi = x0 + (t & 1), yi = y0 + (t >> 1);
        if (xi >= 0 && xi < W && yi >= 0 && yi < H) {                      // zero padding outside the map
          const float wgt = a * ((t & 1) ? fx : 1.f - fx) * ((t >> 1) ? fy : 1.f - fy);
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(vl + (size_t)(yi * W + xi) * C256));
          const float2 v0 = unpack_bf16x2(u.x), v1 = unpack_bf16x2(u.y), v2 = unpack_bf16x2(u.z), v3 = unpack_bf16x2(u.w);
          acc[0] = fmaf(wgt, v0.x, acc[0]); acc[1] = fmaf(wgt, v0.y, acc[1]);
          acc[2] = fmaf(wgt, v1.x, acc[2]); acc[3] = fmaf(wgt, v1.y, acc[3]);
          acc[4] = fmaf(wgt, v2.x, acc[4]); acc[5] = fmaf(wgt, v2.y, acc[5]);
          acc[6] = fmaf(wgt, v3.x, acc[6]); acc[7] = fmaf(wgt, v3.y, acc[7]);
        }
      }
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(out + (size_t)token * C256 + head * 32 + cg * 8) = o;
}

}  // namespace axvs
