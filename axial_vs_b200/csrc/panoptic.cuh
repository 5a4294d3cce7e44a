// Mask-wise panoptic post-processing, GPU-resident.
// Reference: MaXTronWCDeepLab.panoptic_mask_inference, Vk/maxtron_deeplab/maxtron_wc_model.py:439-553 (copy: maxtron_cc_model.py:460-574):
// a 128-iteration Python loop with three `.item()` host syncs and four full-frame boolean kernels per iteration.  Here:
//   1. pano_pixel_kernel  : per pixel softmax over the N slots in ONE pass (running max / rescaled sum, four largest logits); every
//                           slot above the pixel threshold becomes a candidate of the pixel (at most 4: the scores of one pixel sum
//                           to 1 and the threshold is >= 0.2).  Per-slot pixel counts and score sums (64-bit fixed point: order-
//                           independent, hence reproducible).  Pixels with ONE candidate are counted per slot; pixels with several
//                           are counted per candidate SET in a hash table (a few thousand distinct sets: the borders between masks).
//   2. pano_rank_kernel   : one CTA: class softmax / label / confidence per slot, reorder score, descending rank
//   3. pano_greedy_kernel : one CTA walks the slots in rank order over the candidate-set table instead of the pixels: the number of
//                           still-unassigned pixels of a slot = its single-candidate pixels + the counts of the live sets that
//                           contain it; accepting a slot kills those sets.  One block reduction per slot, no pixel traffic.
//   4. pano_paint_kernel  : per pixel, the accepted candidate of lowest rank gives the id (fully parallel).
// Integer / byte work, HBM-bound: kernel 1 streams the N x P logits once (4 N bytes per pixel); kernel 4 moves 8 bytes per pixel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace axvs {

constexpr int PANO_MAX_SLOTS = 255;          // slot ids are bytes; 0xFF = no candidate
constexpr int PANO_CAND = 4;
constexpr float PANO_FIX = 4294967296.0f;    // 2^32 fixed-point scale of the score sums

struct PanoParams {
  const float* mask_cls;     // [N, C1]  (C1 = classes + void)
  const float* mask_pred;    // [N, P]
  int N, C1;
  long long P;
  const int* cat_ids;        // [C1 - 1] label -> category id
  const int* is_thing;       // [C1 - 1]
  int label_divisor;
  float pixel_thr, thing_thr, stuff_thr, overlap_thr, w_cls, w_mask;
  int* out;                  // [P] final ids, -1 = unassigned
  int* segments;             // [1 + 4 N]: count, then (slot, label, is_thing, final id) per opened segment
  // workspace
  uint32_t* cand;            // [P] the pixel's candidate slots, one byte each in ascending slot order (0xFF = none)
  uint32_t* set_key;         // [set_cap] hash table of candidate sets with two or more members: key = the cand word (0xFFFFFFFF = empty)
  uint32_t* set_cnt;         // [set_cap] pixels per set
  uint32_t* set_pos;         // [set_cap] table positions of the occupied entries, in claim order
  uint32_t* set_dkey;        // [set_cap] dense copies (claim order) made by the greedy kernel
  uint32_t* set_dcnt;
  uint32_t set_cap;          // power of two >= 2 P
  int* n_sets;               // [1]
  int* count;                // [N] pixels above the threshold (original_pixel_number)
  int* single;               // [N] ... of which the slot is the only candidate
  unsigned long long* sum;   // [N]
  int* order;                // [N] slot of rank r
  int* rank;                 // [N]
  int* label;                // [N]
  int* confident;            // [N]
  int* final_id;             // [N] id painted by the slot, -2 = rejected
};

__global__ void pano_zero_kernel(PanoParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.N) { p.count[i] = 0; p.single[i] = 0; p.sum[i] = 0ull; }
  if (i == 0) *p.n_sets = 0;
}

__device__ __forceinline__ void pano_set_insert(const PanoParams& p, uint32_t key) {
  uint32_t h = (key * 0x9E3779B1u) & (p.set_cap - 1);
  while (true) {
    const uint32_t old = atomicCAS(p.set_key + h, 0xFFFFFFFFu, key);
    if (old == 0xFFFFFFFFu) p.set_pos[atomicAdd(p.n_sets, 1)] = h;     // claimed a fresh entry
    if (old == 0xFFFFFFFFu || old == key) { atomicAdd(p.set_cnt + h, 1u); return; }
    h = (h + 1) & (p.set_cap - 1);                                       // capacity >= 2 P: a free entry always exists
  }
}

__global__ void __launch_bounds__(256) pano_pixel_kernel(PanoParams p) {
  __shared__ int s_cnt[PANO_MAX_SLOTS + 1], s_single[PANO_MAX_SLOTS + 1];
  __shared__ unsigned long long s_sum[PANO_MAX_SLOTS + 1];
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) { s_cnt[i] = 0; s_single[i] = 0; s_sum[i] = 0ull; }
  __syncthreads();
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px < p.P) {
    // ONE pass over the pixel's N logits: running maximum m and sum s = sum exp(x - m) (rescaled whenever a batch of 16 raises the
    // maximum) and the four largest logits.  Only those can pass the threshold: scores sum to 1 and the threshold is >= 0.2.
    const float* x = p.mask_pred + px;
    float m = -INFINITY, s = 0.f;
    float tv[PANO_CAND] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int ti[PANO_CAND] = {0xFF, 0xFF, 0xFF, 0xFF};
    for (int n0 = 0; n0 < p.N; n0 += 16) {                     // 16 independent loads in flight per thread
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = n0 + j < p.N ? __ldg(x + (size_t)(n0 + j) * p.P) : -INFINITY;
      float bm = v[0];
#pragma unroll
      for (int j = 1; j < 16; ++j) bm = fmaxf(bm, v[j]);
      if (bm > m) { s *= expf(m - bm); m = bm; }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s += expf(v[j] - m);
        if (v[j] > tv[3]) {
          tv[3] = v[j]; ti[3] = n0 + j;
#pragma unroll
          for (int k = 3; k > 0; --k)
            if (tv[k] > tv[k - 1]) {
              const float a = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = a;
              const int b = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = b;
            }
        }
      }
    }
    int k = 0;
    int cs[PANO_CAND] = {0xFF, 0xFF, 0xFF, 0xFF};
#pragma unroll
    for (int q = 0; q < PANO_CAND; ++q) {
      const float sc = expf(tv[q] - m) / s;
      if (ti[q] != 0xFF && sc > p.pixel_thr) {
        cs[q] = ti[q];
        ++k;
        atomicAdd(&s_cnt[ti[q]], 1);
        atomicAdd(&s_sum[ti[q]], (unsigned long long)(sc * PANO_FIX));
      }
    }
    // canonical form: ascending slot order, empty bytes (0xFF) last -- a 4-element sorting network
#define PANO_CSWAP(a, b) { const int lo = min(cs[a], cs[b]), hi = max(cs[a], cs[b]); cs[a] = lo; cs[b] = hi; }
    PANO_CSWAP(0, 1) PANO_CSWAP(2, 3) PANO_CSWAP(0, 2) PANO_CSWAP(1, 3) PANO_CSWAP(1, 2)
#undef PANO_CSWAP
    const uint32_t c = (uint32_t)cs[0] | ((uint32_t)cs[1] << 8) | ((uint32_t)cs[2] << 16) | ((uint32_t)cs[3] << 24);
    if (k == 1) atomicAdd(&s_single[cs[0]], 1);
    else if (k > 1) pano_set_insert(p, c);
    p.cand[px] = c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.N; i += blockDim.x)
    if (s_cnt[i]) {
      atomicAdd(&p.count[i], s_cnt[i]);
      atomicAdd(&p.sum[i], s_sum[i]);
      if (s_single[i]) atomicAdd(&p.single[i], s_single[i]);
    }
}

// one CTA of 1024 threads: a warp per slot for the class softmax, then thread n = slot n for the ranking
__global__ void __launch_bounds__(1024) pano_rank_kernel(PanoParams p) {
  __shared__ float s_score[PANO_MAX_SLOTS + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.C1 <= 128) {                                           // touch every slot's logits first: all the loads of a warp are in flight together
    float t = 0.f;
    for (int n = warp; n < p.N; n += 32) {
      const float* c = p.mask_cls + (size_t)n * p.C1;
#pragma unroll
      for (int j = 0; j < 4; ++j) t += lane + 32 * j < p.C1 ? c[lane + 32 * j] : 0.f;
    }
    if (t == 12345.678f) s_score[0] = t;                       // keeps the loads alive; the passes below hit L1
  }
  for (int n = warp; n < p.N; n += 32) {
    const float* c = p.mask_cls + (size_t)n * p.C1;
    float m = -INFINITY;
    for (int j = lane; j < p.C1; j += 32) m = fmaxf(m, c[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    float s = 0.f;
    for (int j = lane; j < p.C1; j += 32) s += expf(c[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    float best = -1.f;
    int lab = 0x7FFFFFFF;
    for (int j = lane; j < p.C1 - 1; j += 32) {                // the void class (last) is dropped
      const float v = expf(c[j] - m) / s;
      if (v > best) { best = v; lab = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                         // first maximum wins
      const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
      const int ol = __shfl_xor_sync(0xFFFFFFFFu, lab, o);
      if (ob > best || (ob == best && ol < lab)) { best = ob; lab = ol; }
    }
    if (lane == 0) {
      const int cnt = p.count[n];
      const float ms = (__ull2float_rn(p.sum[n]) * (1.0f / 4294967296.0f)) / fmaxf((float)cnt, 1.0f);
      const float a = p.w_cls == 1.0f ? best : powf(best, p.w_cls);
      const float b = p.w_mask == 1.0f ? ms : powf(ms, p.w_mask);
      s_score[n] = a * b;
      p.label[n] = lab;
      p.confident[n] = best > (p.is_thing[lab] ? p.thing_thr : p.stuff_thr);
    }
  }
  __syncthreads();
  const int n = threadIdx.x;
  if (n < p.N) {
    const float score = s_score[n];
    int rank = 0;
    for (int j = 0; j < p.N; ++j) rank += (s_score[j] > score) || (s_score[j] == score && j < n);      // descending, ties by slot index
    p.order[rank] = n;
    p.rank[n] = rank;
  }
}

// one CTA of PANO_GREEDY_THREADS threads (few warps: the per-slot barrier is the critical path); thread t owns the candidate sets
// t, t + 128, ...: the first KEEP of them in registers, the rest in a dense copy it alone updates (a dead set has count 0)
constexpr int PANO_GREEDY_THREADS = 128;
__global__ void __launch_bounds__(PANO_GREEDY_THREADS) pano_greedy_kernel(PanoParams p) {
  constexpr int NT = PANO_GREEDY_THREADS;
  __shared__ int s_acc[3];
  __shared__ int s_rep[PANO_MAX_SLOTS + 1];            // first slot with the same label
  __shared__ int s_lab_cnt[PANO_MAX_SLOTS + 1];        // accepted slots per label, indexed by the label's first slot
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_sets = *p.n_sets;
  constexpr int KEEP = 16;                             // sets per thread held in registers (2048 in all)
  uint32_t key[KEEP], cnt[KEEP];
#pragma unroll
  for (int j = 0; j < KEEP; ++j) {
    const int q = tid + j * NT;
    const uint32_t pos = q < n_sets ? p.set_pos[q] : 0u;
    key[j] = q < n_sets ? p.set_key[pos] : 0xFFFFFFFFu;
    cnt[j] = q < n_sets ? p.set_cnt[pos] : 0u;
  }
  for (int q = tid + KEEP * NT; q < n_sets; q += NT) {  // dense copy of the remaining sets (entry q is private to this thread)
    const uint32_t pos = p.set_pos[q];
    p.set_dkey[q] = p.set_key[pos];
    p.set_dcnt[q] = p.set_cnt[pos];
  }
  __shared__ int s_order[PANO_MAX_SLOTS + 1], s_total[PANO_MAX_SLOTS + 1], s_single[PANO_MAX_SLOTS + 1], s_label[PANO_MAX_SLOTS + 1];
  __shared__ int s_thing[PANO_MAX_SLOTS + 1], s_cat[PANO_MAX_SLOTS + 1];
  if (tid < 3) s_acc[tid] = 0;
  for (int n = tid; n < p.N; n += NT) {                // per-slot tables: no dependent global loads inside the serial loop
    p.final_id[n] = -2;
    s_order[n] = p.order[n];
    const int lab = p.label[n];
    s_total[n] = p.confident[n] ? p.count[n] : 0;                // 0 = the slot can never be accepted
    s_single[n] = p.single[n];
    s_label[n] = lab;
    s_thing[n] = p.is_thing[lab];
    s_cat[n] = p.cat_ids[lab];
    s_lab_cnt[n] = 0;
  }
  __syncthreads();
  for (int n = tid; n < p.N; n += NT) {
    int rep = n;
    for (int j = n - 1; j >= 0; --j) rep = s_label[j] == s_label[n] ? j : rep;
    s_rep[n] = rep;
  }
  __syncthreads();
  int nseg = 0, it = 0;                                // nseg: thread 0 only; it: kept identically by every thread
  for (int r = 0; r < p.N; ++r) {
    const int cur = s_order[r];
    const int total = s_total[cur];
    if (total == 0) continue;                          // uniform across the block; such a slot paints nothing (:493-501)
    const uint32_t pat = (uint32_t)cur * 0x01010101u;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < KEEP; ++j) mine += __vcmpeq4(key[j], pat) ? (int)cnt[j] : 0;      // dead sets have cnt = 0
    for (int q = tid + KEEP * NT; q < n_sets; q += NT) mine += __vcmpeq4(p.set_dkey[q], pat) ? (int)p.set_dcnt[q] : 0;
    mine = __reduce_add_sync(0xFFFFFFFFu, mine);
    if (lane == 0 && mine) atomicAdd(&s_acc[it % 3], mine);
    if (tid == 0) s_acc[(it + 1) % 3] = 0;
    __syncthreads();
    const int fresh = s_acc[it % 3] + s_single[cur];
    ++it;
    if (!((float)fresh > __fmul_rn((float)total, p.overlap_thr))) continue;   // new_pixel_number > original_pixel_number * overlap_threshold (:499)
    // the slot takes every live set it belongs to
#pragma unroll
    for (int j = 0; j < KEEP; ++j)
      if (__vcmpeq4(key[j], pat)) cnt[j] = 0;
    for (int q = tid + KEEP * NT; q < n_sets; q += NT)
      if (__vcmpeq4(p.set_dkey[q], pat)) p.set_dcnt[q] = 0;
    if (tid == 0) {                                    // segment bookkeeping: thread 0 only, O(1) per slot
      const int rep = s_rep[cur];
      const int same = s_lab_cnt[rep];                 // earlier accepted slots of the same label (label <-> category id is 1:1)
      s_lab_cnt[rep] = same + 1;
      const int thing = s_thing[cur];
      const int id = thing ? s_cat[cur] * p.label_divisor + same : s_cat[cur];
      p.final_id[cur] = id;
      if (thing || same == 0) {                        // a merged stuff region opens no new segment (:506-508)
        int* sg = p.segments + 1 + 4 * nseg;
        sg[0] = cur; sg[1] = s_label[cur]; sg[2] = thing; sg[3] = id;
        ++nseg;
      }
    }
  }
  if (tid == 0) p.segments[0] = nseg;
}

__global__ void __launch_bounds__(256) pano_paint_kernel(PanoParams p) {
  __shared__ int s_rank[PANO_MAX_SLOTS + 1], s_id[PANO_MAX_SLOTS + 1];
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) { s_rank[i] = p.rank[i]; s_id[i] = p.final_id[i]; }
  __syncthreads();
  const long long px = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (px >= p.P) return;
  const uint32_t c = p.cand[px];
  int best = 0x7FFFFFFF, id = -1;
#pragma unroll
  for (int k = 0; k < PANO_CAND; ++k) {
    const uint32_t n = (c >> (8 * k)) & 0xFFu;
    if (n != 0xFFu && s_id[n] != -2 && s_rank[n] < best) { best = s_rank[n]; id = s_id[n]; }
  }
  p.out[px] = id;
}

}  // namespace axvs
