// Per-frame-softmax spatial attention of TrajectoryAttention (Appendix A steps 2-5):
//   x[s, q, f, head*32 + :] = softmax_i( scale * Q[s,q,head,:] . K[s, f*n+i, head, :] ) @ V[s, f*n+i, head, :]
// Reference: WC/temporal_attention.py:47-60 (the softmax is taken independently inside every key frame).
//
// v1 implementation: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate), flash-style online softmax
// over 64-key tiles so that any tokens-per-frame n works (n = H or W for the axial layers, H*W for the
// non-axial "trajectory" layer, Q for the cross-clip module).  One CTA = 64 queries of one (sequence, head);
// K/V tiles are double-buffered with cp.async; the score matrix never leaves registers.
#pragma once
#include "ptx.cuh"

namespace axvs {

constexpr int ATT_QT = 64;   // queries per CTA (4 warps x 16)
constexpr int ATT_KT = 64;   // keys per tile
constexpr int ATT_D = 32;    // head dim

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = smem_u32(smem_dst);
  const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 2^x for x <= 0 (and -inf -> 0): the hardware approximation without exp2f's range handling
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// smem tiles are [rows][32] bf16 = 64 B rows = 4 chunks of 16 B, chunk index XOR-swizzled with (row>>1)&3
__device__ __forceinline__ int att_off(int row, int chunk) { return row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4); }

// qkv: [num_seq*N, ld] bf16 with q at column q_col, k at k_col, v at v_col (+ head*32);  x: [num_seq*N, F, 256] bf16
// operand (which, head) of row r starts at qkv + which * which_stride + head * head_stride + r * ld  (elements):
//   row-major [rows, 768]:           ld = 768, which_stride = 256,           head_stride = 32
//   head-major [3][8][rows][32]:     ld = 32,  which_stride = 8 * rows * 32, head_stride = rows * 32
__global__ void __launch_bounds__(128) spatial_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, size_t which_stride,
                                                           size_t head_stride, __nv_bfloat16* __restrict__ x, int N, int n, int F,
                                                           float scale_log2e) {
  __shared__ __align__(128) uint8_t sQ[ATT_QT * 64];
  __shared__ __align__(128) uint8_t sK[2][ATT_KT * 64];
  __shared__ __align__(128) uint8_t sV[2][ATT_KT * 64];

  const int q0 = blockIdx.x * ATT_QT;
  const int seq = blockIdx.y;
  const int head = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t seq_row0 = (size_t)seq * N;

  // ---- Q tile -> smem (64 rows x 4 chunks = 256 chunks, 2 per thread)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = i * 128 + tid;
    const int r = c >> 2, ch = c & 3;
    const bool ok = (q0 + r) < N;
    const __nv_bfloat16* src = qkv + head * head_stride + (seq_row0 + (ok ? q0 + r : 0)) * ld + ch * 8;
    cp_async16(sQ + att_off(r, ch), src, ok);
  }
  const int tiles_per_frame = (n + ATT_KT - 1) / ATT_KT;
  const int total_tiles = F * tiles_per_frame;

  auto load_kv = [&](int it, int buf) {
    const int f = it / tiles_per_frame, kt = it % tiles_per_frame;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = i * 128 + tid;
      const int r = c >> 2, ch = c & 3;
      const int key = kt * ATT_KT + r;
      const bool ok = key < n;
      const __nv_bfloat16* base = qkv + head * head_stride + (seq_row0 + (size_t)f * n + (ok ? key : 0)) * ld + ch * 8;
      cp_async16(sK[buf] + att_off(r, ch), base + which_stride, ok);
      cp_async16(sV[buf] + att_off(r, ch), base + 2 * which_stride, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  // wait for Q (+ first K/V tile), build the Q fragments once
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qa[2][4];
  {
    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(qa[ks], sQ + att_off(r, ks * 2 + (lane >> 4)));
  }

  const int g = lane >> 2, t4 = lane & 3;
  float m_run[2], l_run[2], acc[4][4];

  for (int it = 0; it < total_tiles; ++it) {
    const int buf = it & 1;
    const int f = it / tiles_per_frame, kt = it % tiles_per_frame;
    if (kt == 0) {
      m_run[0] = m_run[1] = -INFINITY;
      l_run[0] = l_run[1] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    }
    if (it + 1 < total_tiles) load_kv(it + 1, buf ^ 1);
    cp_async_commit();
    if (it > 0) {
      cp_async_wait<1>();   // tile `it` has landed (tile it+1 may still be in flight)
      __syncthreads();
    }

    // ---- S = Q K^T for 16 queries x 64 keys
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t kb[4];
      ldmatrix_x4(kb, sK[buf] + att_off(j * 8 + (lane & 7), lane >> 3));
      mma_bf16_16816(s[j], qa[0], kb[0], kb[1]);
      mma_bf16_16816(s[j], qa[1], kb[2], kb[3]);
    }
    // ---- mask keys beyond the frame, online softmax (log2 domain)
    const int key_base = kt * ATT_KT;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = key_base + j * 8 + t4 * 2 + (e & 1);
        const float v = (key < n) ? s[j][e] * scale_log2e : -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float mn = fmaxf(m_run[h], mx[h]);   // finite: every tile has >= 1 valid key
      corr[h] = exp2f(m_run[h] - mn);
      m_run[h] = mn;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pa[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_run[0]), p1 = exp2f(s[j][1] - m_run[0]);
      const float p2 = exp2f(s[j][2] - m_run[1]), p3 = exp2f(s[j][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + rs[h];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j][0] *= corr[0]; acc[j][1] *= corr[0];
      acc[j][2] *= corr[1]; acc[j][3] *= corr[1];
    }
    // ---- acc += P V   (16 x 64) x (64 x 32)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jn = 0; jn < 4; jn += 2) {
        uint32_t vb[4];
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(vb, sV[buf] + att_off(key, jn + (lane >> 4)));
        mma_bf16_16816(acc[jn], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(acc[jn + 1], pa[ks], vb[2], vb[3]);
      }
    }
    // ---- end of frame: normalise and write x[q, f, head*32 : head*32+32]
    if (kt == tiles_per_frame - 1) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float l = l_run[h];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float inv = 1.f / l;
        const int q = q0 + warp * 16 + g + h * 8;
        if (q < N) {
          __nv_bfloat16* dst = x + ((seq_row0 + q) * F + f) * 256 + head * ATT_D + t4 * 2;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint32_t*>(dst + j * 8) = pack_bf16x2(acc[j][h * 2] * inv, acc[j][h * 2 + 1] * inv);
        }
      }
    }
    __syncthreads();   // everyone is done with buf before it is refilled two iterations later
  }
}

}  // namespace axvs

namespace axvs {

// ------------------------------------------------------------------------------------------------------------------
// v2: one CTA per (sequence, head); the whole Q of the sequence is staged in shared memory, K_f / V_f are double-
// buffered per frame, and the per-frame softmax is ONE-SHOT (n <= 16 * NT16 <= 176 keys: the score row lives in
// registers), so there is no online rescaling and key tiles are padded to 16 instead of 64.  Inputs are head-major
// (qkv[which][head][row][32], written by qkv_fused_kernel); outputs are written straight into the SWIZZLE_128B tile
// images that traj_fused_kernel loads by TMA (x_f per frame + the own-frame copy x_diag).
// ------------------------------------------------------------------------------------------------------------------
constexpr int ATT2_KB = 16384;   // bytes of one 128-row x 64-column image K-block

// One (16-query block, key frame) work item of a (sequence, head): S = Q K_f^T, one-shot softmax over the frame's n keys, P V_f,
// then the x_f (+ x_diag) rows written as 16-byte pieces of the SWIZZLE_128B tile images traj_*_kernel consumes.
// sQ / sK / sV: [rows][64 B] tiles (att_off swizzle); stg: this warp's 1 KiB staging; kb / ch0: K-block and first 16-byte chunk
// of the head's 32 channels inside an image row.
template <int NT16>
__device__ __forceinline__ void attn_item(const uint8_t* sQ, const uint8_t* sK, const uint8_t* sV, uint8_t* stg, int mb, int f, int N, int n,
                                          float scale_log2e, size_t seq_row0, int kb, int ch0, int tiles, uint8_t* __restrict__ x_img,
                                          uint8_t* __restrict__ xd_img, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
    uint32_t qa[2][4];
    {
      const int r = mb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) ldmatrix_x4(qa[ks], sQ + att_off(r, ks * 2 + (lane >> 4)));
    }
    float s[2 * NT16][4];
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      uint32_t kf[4];
      ldmatrix_x4(kf, sK + att_off(j * 8 + (lane & 7), lane >> 3));
      mma_bf16_16816(s[j], qa[0], kf[0], kf[1]);
      mma_bf16_16816(s[j], qa[1], kf[2], kf[3]);
    }
    // row maxima over the raw scores (keys >= n masked; only 8-key tiles straddling n need the per-element test), then
    // p = 2^(s * scale - max * scale) with the scale folded into one FMA per element
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      if (j * 8 + 8 > n) {                               // warp-uniform
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j * 8 + t4 * 2 + (e & 1) >= n) s[j][e] = -INFINITY;
      }
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      mx[h] *= -scale_log2e;                             // every row has at least one valid key: finite
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pa[NT16][4];
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      const float p0 = ex2_approx(fmaf(s[j][0], scale_log2e, mx[0])), p1 = ex2_approx(fmaf(s[j][1], scale_log2e, mx[0]));
      const float p2 = ex2_approx(fmaf(s[j][2], scale_log2e, mx[1])), p3 = ex2_approx(fmaf(s[j][3], scale_log2e, mx[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NT16; ++ks) {
#pragma unroll
      for (int jn = 0; jn < 4; jn += 2) {
        uint32_t vb[4];
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4_trans(vb, sV + att_off(key, jn + (lane >> 4)));
        mma_bf16_16816(acc[jn], pa[ks], vb[0], vb[1]);
        mma_bf16_16816(acc[jn + 1], pa[ks], vb[2], vb[3]);
      }
    }
    // normalise -> per-warp staging (16 rows x 64 B) -> 16-byte stores into the tile images
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float l = rs[h];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = __frcp_rn(l);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint32_t*>(stg + att_off(g + h * 8, j) + t4 * 4) = pack_bf16x2(acc[j][h * 2] * inv, acc[j][h * 2 + 1] * inv);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = lane + 32 * i;
      const int row = idx >> 2, chunk = idx & 3;
      const int qi = mb * 16 + row;                     // query index inside the sequence
      if (qi < N) {
        const uint4 v = *reinterpret_cast<const uint4*>(stg + att_off(row, chunk));
        const size_t r = seq_row0 + qi;
        const size_t off = ((r >> 7) * 4 + kb) * ATT2_KB + sw128_offset((uint32_t)(r & 127), ch0 + chunk);
        *reinterpret_cast<uint4*>(x_img + (size_t)f * tiles * 4 * ATT2_KB + off) = v;
        if ((unsigned)(qi - f * n) < (unsigned)n) *reinterpret_cast<uint4*>(xd_img + off) = v;   // qi / n == f without the division
      }
    }
    __syncwarp();
}


template <int NT16>
__global__ void __launch_bounds__(128) spatial_attn_v2_kernel(const __nv_bfloat16* __restrict__ qkv, size_t rows_total, uint8_t* __restrict__ x_img,
                                                              uint8_t* __restrict__ xd_img, int tiles, int N, int n, int F, float scale_log2e,
                                                              int all_frames) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  constexpr int NP = 16 * NT16;
  const int n_mblk = (N + 15) >> 4;
  const int kv_bufs = all_frames ? F : 2;
  uint8_t* sQ = att_smem;                                  // [n_mblk*16][64 B]
  uint8_t* sKV = sQ + (size_t)n_mblk * 16 * 64;            // [kv_bufs][K | V][NP][64 B]
  uint8_t* sStage = sKV + (size_t)kv_bufs * 2 * NP * 64;   // [4 warps][16 rows][64 B]

  const int seq = blockIdx.x >> 3, head = blockIdx.x & 7;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t seq_row0 = (size_t)seq * N;
  const __nv_bfloat16* gq = qkv + ((size_t)(0 * 8 + head) * rows_total + seq_row0) * 32;
  const __nv_bfloat16* gk = qkv + ((size_t)(1 * 8 + head) * rows_total + seq_row0) * 32;
  const __nv_bfloat16* gv = qkv + ((size_t)(2 * 8 + head) * rows_total + seq_row0) * 32;

  for (int c = tid; c < n_mblk * 16 * 4; c += 128) {
    const int r = c >> 2, ch = c & 3;
    const bool ok = r < N;
    cp_async16(sQ + att_off(r, ch), gq + (size_t)(ok ? r : 0) * 32 + ch * 8, ok);
  }
  auto load_kv = [&](int f, int buf) {
    uint8_t* sK = sKV + (size_t)buf * 2 * NP * 64;
    uint8_t* sV = sK + NP * 64;
    for (int c = tid; c < NP * 4; c += 128) {
      const int r = c >> 2, ch = c & 3;
      const bool ok = r < n;
      const size_t off = (size_t)(f * n + (ok ? r : 0)) * 32 + ch * 8;
      cp_async16(sK + att_off(r, ch), gk + off, ok);
      cp_async16(sV + att_off(r, ch), gv + off, ok);
    }
  };

  uint8_t* stg = sStage + warp * 1024;
  const int kb = head >> 1, ch0 = (head & 1) * 4;

  auto process = [&](int mb, int f, const uint8_t* sK, const uint8_t* sV) {
    attn_item<NT16>(sQ, sK, sV, stg, mb, f, N, n, scale_log2e, seq_row0, kb, ch0, tiles, x_img, xd_img, lane);
  };

  if (all_frames) {
    // every frame's K/V resident: the (query block, frame) items are spread evenly over the 4 warps
    for (int f = 0; f < F; ++f) load_kv(f, f);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    int mb = warp / F, f = warp - mb * F;                 // item = mb * F + f, advanced by 4 per iteration without divisions
    for (int item = warp; item < n_mblk * F; item += 4) {
      const uint8_t* sK = sKV + (size_t)f * 2 * NP * 64;
      process(mb, f, sK, sK + NP * 64);
      f += 4;
      while (f >= F) { f -= F; ++mb; }
    }
  } else {
    load_kv(0, 0);
    cp_async_commit();
    for (int f = 0; f < F; ++f) {
      const int buf = f & 1;
      if (f + 1 < F) load_kv(f + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const uint8_t* sK = sKV + (size_t)buf * 2 * NP * 64;
      for (int mb = warp; mb < n_mblk; mb += 4) process(mb, f, sK, sK + NP * 64);
      __syncthreads();   // all warps done with this frame's K/V buffer before it is refilled
    }
  }
}

// Persistent, double-buffered variant for sequences whose q | k | v of one head fit twice in shared memory (the axial passes):
// each CTA walks (sequence, head) work items with a grid stride and loads the NEXT item's operands with cp.async while the four
// warps compute the current one, so the ~2 us load round trip that made up 30 % of the one-shot kernel's samples is hidden.
template <int NT16>
__global__ void __launch_bounds__(128) spatial_attn_v3_kernel(const __nv_bfloat16* __restrict__ qkv, size_t rows_total, uint8_t* __restrict__ x_img,
                                                              uint8_t* __restrict__ xd_img, int tiles, int N, int n, int F, float scale_log2e,
                                                              int num_work) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  constexpr int NP = 16 * NT16;
  const int n_mblk = (N + 15) >> 4;
  const size_t q_bytes = (size_t)n_mblk * 16 * 64, kv_bytes = (size_t)F * 2 * NP * 64, buf_bytes = q_bytes + kv_bytes;
  uint8_t* sStage = att_smem + 2 * buf_bytes;              // [4 warps][16 rows][64 B]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* stg = sStage + warp * 1024;

  auto load_item = [&](int w, int buf) {
    const int seq = w >> 3, head = w & 7;
    const size_t seq_row0 = (size_t)seq * N;
    const __nv_bfloat16* gq = qkv + ((size_t)(0 * 8 + head) * rows_total + seq_row0) * 32;
    const __nv_bfloat16* gk = qkv + ((size_t)(1 * 8 + head) * rows_total + seq_row0) * 32;
    const __nv_bfloat16* gv = qkv + ((size_t)(2 * 8 + head) * rows_total + seq_row0) * 32;
    uint8_t* sQ = att_smem + (size_t)buf * buf_bytes;
    uint8_t* sKV = sQ + q_bytes;
    for (int c = tid; c < n_mblk * 16 * 4; c += 128) {
      const int r = c >> 2, ch = c & 3;
      const bool ok = r < N;
      cp_async16(sQ + att_off(r, ch), gq + (size_t)(ok ? r : 0) * 32 + ch * 8, ok);
    }
    for (int f = 0; f < F; ++f) {
      uint8_t* sK = sKV + (size_t)f * 2 * NP * 64;
      uint8_t* sV = sK + NP * 64;
      for (int c = tid; c < NP * 4; c += 128) {
        const int r = c >> 2, ch = c & 3;
        const bool ok = r < n;
        const size_t off = (size_t)(f * n + (ok ? r : 0)) * 32 + ch * 8;
        cp_async16(sK + att_off(r, ch), gk + off, ok);
        cp_async16(sV + att_off(r, ch), gv + off, ok);
      }
    }
  };

  int w = blockIdx.x;
  if (w < num_work) load_item(w, 0);
  cp_async_commit();
  const int mb0 = warp / F, f0 = warp - mb0 * F;
  for (int it = 0; w < num_work; w += gridDim.x, ++it) {
    const int buf = it & 1;
    if (w + (int)gridDim.x < num_work) load_item(w + gridDim.x, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();                                    // the current item's operands have landed
    __syncthreads();
    const int seq = w >> 3, head = w & 7;
    const uint8_t* sQ = att_smem + (size_t)buf * buf_bytes;
    const uint8_t* sKV = sQ + q_bytes;
    int mb = mb0, f = f0;                                  // item = mb * F + f, advanced by 4 per iteration without divisions
    for (int item = warp; item < n_mblk * F; item += 4) {
      const uint8_t* sK = sKV + (size_t)f * 2 * NP * 64;
      attn_item<NT16>(sQ, sK, sK + NP * 64, stg, mb, f, N, n, scale_log2e, (size_t)seq * N, head >> 1, (head & 1) * 4, tiles, x_img, xd_img, lane);
      f += 4;
      while (f >= F) { f -= F; ++mb; }
    }
    __syncthreads();                                       // every warp is done with this buffer before it is refilled
  }
}

}  // namespace axvs
