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

// Specialised form for compile-time (L, P) -- the shipped configs use 3 or 4 levels x 4 points.  An `ncu --set full` capture of the generic
// kernel above showed it ISSUE-bound (issue slots 75 % busy, L1 49 %, L2 31 %, DRAM 8 %): every one of the four lanes of a head repeated the
// head's softmax and the location / weight arithmetic of all L*P samples.  Here the four lanes SHARE that work: lane cg owns samples
// cg, cg + 4, ... (logit -> exp, location, validity, the four bilinear weights, one packed tap index), the softmax statistics are combined
// with two xor-shuffles, and each sample's five values are broadcast from its owner.  Out-of-range taps keep a clamped (safe) address and
// a zero weight, so the sixteen tap loads of a level are issued without branches; the channel pairs accumulate with packed fma.rn.f32x2.
// VALUE_HM: value is head-major [image][head][len][32] (written by the fused front kernel) instead of token-major [image*len][256].
template <int L, int P, bool VALUE_HM>
__global__ void __launch_bounds__(256) msda_sample_lp_kernel(const __nv_bfloat16* __restrict__ value, const float* __restrict__ oa, int ld_oa,
                                                             const float* __restrict__ ref, int ref_rows, __nv_bfloat16* __restrict__ out, int rows,
                                                             MsdaDims d) {
  constexpr int LP = L * P;
  constexpr int OWN = (LP + 3) / 4;                       // samples owned per lane
#ifdef AXVS_MSDA_PB
  constexpr int PB = AXVS_MSDA_PB;
#else
  constexpr int PB = (P % 2 == 0) ? 2 : 1;
#endif
  const int token = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (token >= rows) return;                              // whole warps leave together
  const int lane = threadIdx.x & 31, head = lane >> 2, cg = lane & 3;
  const float* o_row = oa + (size_t)token * ld_oa;
  const int img = token / d.len;
  const size_t rrow = ref_rows > 0 ? (size_t)(token % ref_rows) : (size_t)token;
  // ---- owner work: samples s = cg + 4 j
  float e[OWN], w00[OWN], w01[OWN], w10[OWN], w11[OWN];
  int pk[OWN];                                            // tap (y0c, x0c) index inside the level | dx << 30 | dy << 31
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < OWN; ++j) {
    const int s_ = cg + 4 * j;
    e[j] = s_ < LP ? __ldg(o_row + 8 * LP * 2 + head * LP + s_) : -INFINITY;
    mx = fmaxf(mx, e[j]);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float den = 0.f;
#pragma unroll
  for (int j = 0; j < OWN; ++j) {
    e[j] = cg + 4 * j < LP ? expf(e[j] - mx) : 0.f;
    den += e[j];
  }
  den += __shfl_xor_sync(0xffffffffu, den, 1);
  den += __shfl_xor_sync(0xffffffffu, den, 2);
  const float inv = 1.f / den;
#pragma unroll
  for (int j = 0; j < OWN; ++j) {
    const int s_ = cg + 4 * j;
    w00[j] = w01[j] = w10[j] = w11[j] = 0.f;
    pk[j] = 0;
    if (s_ < LP) {
      const int l = (P == 4) ? j : s_ / P;                  // P == 4: owner slot j holds a sample of level j (compile-time level constants)
      const int H = d.H[l], W = d.W[l];
      const float rx = __ldg(ref + (rrow * L + l) * 2), ry = __ldg(ref + (rrow * L + l) * 2 + 1);
      const float2 off = __ldg(reinterpret_cast<const float2*>(o_row + (head * LP + s_) * 2));
      // location = ref + off / (W, H) in [0,1]; pixel coordinate = location * size - 0.5 (align_corners = False)
      const float x = (rx + off.x / (float)W) * (float)W - 0.5f, y = (ry + off.y / (float)H) * (float)H - 0.5f;
      const float xf = floorf(x), yf = floorf(y);
      const float fx = x - xf, fy = y - yf;
      // clamp in float first: far-away locations must not overflow the integer conversion
      const int x0 = (int)fminf(fmaxf(xf, -2.f), (float)W), y0 = (int)fminf(fmaxf(yf, -2.f), (float)H);
      const float a = e[j] * inv;
      const float ax0 = (x0 >= 0 && x0 < W) ? 1.f - fx : 0.f, ax1 = (x0 + 1 >= 0 && x0 + 1 < W) ? fx : 0.f;      // zero padding outside the map
      const float ay0 = (y0 >= 0 && y0 < H) ? a * (1.f - fy) : 0.f, ay1 = (y0 + 1 >= 0 && y0 + 1 < H) ? a * fy : 0.f;
      w00[j] = ay0 * ax0; w01[j] = ay0 * ax1; w10[j] = ay1 * ax0; w11[j] = ay1 * ax1;
      const int x0c = min(max(x0, 0), W - 1), x1c = min(max(x0 + 1, 0), W - 1);
      const int y0c = min(max(y0, 0), H - 1), y1c = min(max(y0 + 1, 0), H - 1);
      pk[j] = (int)((unsigned)(y0c * W + x0c) | ((unsigned)(x1c - x0c) << 30) | ((unsigned)(y1c - y0c) << 31));
    }
  }
  // ---- gather: every lane needs every sample of its head
  const size_t tok_stride = VALUE_HM ? 32 : C256;         // elements between consecutive tokens of one head
  const __nv_bfloat16* vbase = VALUE_HM ? value + ((size_t)(img * 8 + head) * d.len) * 32 + cg * 8
                                        : value + (size_t)img * d.len * C256 + head * 32 + cg * 8;
  float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  const int grp = lane & ~3;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int W = d.W[l];
    const __nv_bfloat16* vl = vbase + (size_t)d.start[l] * tok_stride;
#pragma unroll
    for (int p0 = 0; p0 < P; p0 += PB) {                  // PB points = 4 PB independent 16-byte tap loads in flight per lane
      uint4 u[PB][4];
      float tw[PB][4];
#pragma unroll
      for (int q = 0; q < PB; ++q) {
        const int s_ = l * P + p0 + q, owner = grp | (s_ & 3), j = s_ >> 2;
        const int k = __shfl_sync(0xffffffffu, pk[j], owner);
        tw[q][0] = __shfl_sync(0xffffffffu, w00[j], owner);
        tw[q][1] = __shfl_sync(0xffffffffu, w01[j], owner);
        tw[q][2] = __shfl_sync(0xffffffffu, w10[j], owner);
        tw[q][3] = __shfl_sync(0xffffffffu, w11[j], owner);
        const int i00 = k & 0x3fffffff, dx = (k >> 30) & 1, dy = ((unsigned)k >> 31) ? W : 0;
        u[q][0] = __ldg(reinterpret_cast<const uint4*>(vl + (size_t)i00 * tok_stride));
        u[q][1] = __ldg(reinterpret_cast<const uint4*>(vl + (size_t)(i00 + dx) * tok_stride));
        u[q][2] = __ldg(reinterpret_cast<const uint4*>(vl + (size_t)(i00 + dy) * tok_stride));
        u[q][3] = __ldg(reinterpret_cast<const uint4*>(vl + (size_t)(i00 + dy + dx) * tok_stride));
      }
#pragma unroll
      for (int q = 0; q < PB; ++q) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 w2 = make_float2(tw[q][t], tw[q][t]);
          acc[0] = fma_f32x2(w2, unpack_bf16x2(u[q][t].x), acc[0]);
          acc[1] = fma_f32x2(w2, unpack_bf16x2(u[q][t].y), acc[1]);
          acc[2] = fma_f32x2(w2, unpack_bf16x2(u[q][t].z), acc[2]);
          acc[3] = fma_f32x2(w2, unpack_bf16x2(u[q][t].w), acc[3]);
        }
      }
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0].x, acc[0].y);
  o.y = pack_bf16x2(acc[1].x, acc[1].y);
  o.z = pack_bf16x2(acc[2].x, acc[2].y);
  o.w = pack_bf16x2(acc[3].x, acc[3].y);
  *reinterpret_cast<uint4*>(out + (size_t)token * C256 + head * 32 + cg * 8) = o;
}

// value_hm != 0: head-major value (see msda_sample_lp_kernel); only the specialised (L, P) pairs support it.
inline bool launch_msda_sample(const __nv_bfloat16* value, int value_hm, const float* oa, int ld_oa, const float* ref, int ref_rows,
                               __nv_bfloat16* out, int rows, const MsdaDims& d, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (d.L == 3 && d.P == 4) {
    if (value_hm) msda_sample_lp_kernel<3, 4, true><<<grid, 256, 0, st>>>(value, oa, ld_oa, ref, ref_rows, out, rows, d);
    else msda_sample_lp_kernel<3, 4, false><<<grid, 256, 0, st>>>(value, oa, ld_oa, ref, ref_rows, out, rows, d);
    return true;
  }
  if (d.L == 4 && d.P == 4) {
    if (value_hm) msda_sample_lp_kernel<4, 4, true><<<grid, 256, 0, st>>>(value, oa, ld_oa, ref, ref_rows, out, rows, d);
    else msda_sample_lp_kernel<4, 4, false><<<grid, 256, 0, st>>>(value, oa, ld_oa, ref, ref_rows, out, rows, d);
    return true;
  }
  if (value_hm) return false;
  msda_sample_kernel<<<grid, 256, 0, st>>>(value, oa, ld_oa, ref, ref_rows, out, rows, d);
  return true;
}

}  // namespace axvs
