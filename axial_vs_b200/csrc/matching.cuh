// Clip-to-clip query matching, GPU-resident (second half).
//
// Reference: MaXTronWCDeepLab.match_from_embds (Vk/maxtron_deeplab/maxtron_wc_model.py:391-400; identical copy in maxtron_cc_model.py) and
// the chains that call it (maxtron_wc_model.py:342-346, maxtron_cc_model.py:292-295):
//     cur, tgt L2-normalised rows;  C[cur, tgt] = 1 - cur . tgt;  indices = scipy.optimize.linear_sum_assignment(C^T)[1]
// and clip i is matched against the ALREADY PERMUTED clip i-1.  The reference moves every cost matrix to the host for scipy; here one CTA
// per video walks the whole chain on the device: normalise, cost matrix, exact assignment, permute, next clip -- no host hop.
//
// The assignment is scipy's own algorithm (scipy/optimize/rectangular_lsap/rectangular_lsap.cpp: Crouse's shortest-augmenting-path variant
// of Jonker-Volgenant), restated with the SAME arithmetic (float64 duals, the same left-to-right sums) and the same tie rules (the
// `remaining` list filled in reverse and shrunk by swap-with-last; among equal shortest-path costs the scan keeps the first column unless a
// later one is unassigned), so the permutation is identical to scipy's on the same cost matrix, ties included (tests/test_matching_*.py).
// The scan over the remaining columns is what the CTA parallelises: one thread per column, two block-wide reductions per step that
// reproduce the sequential scan's result exactly.
#pragma once
#include "ptx.cuh"

namespace axvs {

constexpr int LS_MAX_N = 256;
constexpr int LS_THREADS = 256;

struct LsapShared {
  double u[LS_MAX_N], v[LS_MAX_N], spc[LS_MAX_N];
  int path[LS_MAX_N], col4row[LS_MAX_N], row4col[LS_MAX_N], remaining[LS_MAX_N];
  unsigned char SR[LS_MAX_N], SC[LS_MAX_N];
  int sink;
#ifdef AXVS_MATCH_PROF
  long long steps;
#endif
};

// Solve min sum_i cost[i, col4row[i]] for a square n x n matrix (row-major, row = target, column = current); the result is left in
// s.col4row[0..n) (s.sink < 0: infeasible).  Runs in WARP 0 of the CTA; the other warps wait at the closing barrier.
//
// Every relaxation step is one dependent chain (pick a column -> its row -> relax that row's costs -> pick ...), ~1000 steps per 128 x 128
// problem, so the step latency is everything: block-wide barriers (256 threads, four per step) gave 1.8 ms per problem, a single warp
// working on shared-memory arrays 2.2 ms (the compiler cannot overlap the loads of one column with the stores of the previous one).  Here
// lane l OWNS the columns l, l + 32, ... and keeps their duals, shortest-path costs, assignments and scan positions in registers; per
// step it reads one cost per owned column and two warp reductions pick the column.  The sequential scan's choice -- the first minimum
// unless a later equal one is unassigned -- equals "the last unassigned minimum in scan order if there is one, else the first minimum".
template <int CPL>
__device__ void lsap_solve_t(const float* __restrict__ cost, int n, LsapShared& s) {
  const int lane = threadIdx.x & 31;
  double v[CPL], spc[CPL];
  int r4c[CPL], pos[CPL], pth[CPL];
#pragma unroll
  for (int m = 0; m < CPL; ++m) { v[m] = 0.0; r4c[m] = -1; pth[m] = -1; }
  for (int k = lane; k < n; k += 32) { s.u[k] = 0.0; s.col4row[k] = -1; s.row4col[k] = -1; }
  if (lane == 0) s.sink = 0;
  __syncwarp();
  bool feasible = true;
  for (int curRow = 0; curRow < n && feasible; ++curRow) {
    uint32_t rem = 0, sc = 0;                                  // bit m: owned column lane + 32 m is still to be scanned / was scanned (SC)
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const int j = lane + 32 * m;
      spc[m] = INFINITY;
      pos[m] = n - 1 - j;                                      // remaining[it] = n - it - 1
      if (j < n) rem |= 1u << m;
    }
    for (int k = lane; k < n; k += 32) { s.remaining[k] = n - k - 1; s.SR[k] = 0; }
    __syncwarp();
    int i = curRow, nrem = n, sink = -1;
    double minVal = 0.0;
    while (sink == -1) {
      if (lane == 0) s.SR[i] = 1;
      const double ui = s.u[i];
      const float* crow = cost + (size_t)i * n;
      double best = INFINITY;
#pragma unroll
      for (int m = 0; m < CPL; ++m) {
        if (rem & (1u << m)) {
          const double r = minVal + (double)crow[lane + 32 * m] - ui - v[m];
          if (r < spc[m]) { pth[m] = i; spc[m] = r; }
          best = fmin(best, spc[m]);
        }
      }
      double lowest = best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lowest = fmin(lowest, __shfl_xor_sync(0xffffffffu, lowest, o));
      if (!(lowest < INFINITY)) { feasible = false; break; }
      int code = -1;
#pragma unroll
      for (int m = 0; m < CPL; ++m)
        if ((rem & (1u << m)) && spc[m] == lowest) code = max(code, (r4c[m] == -1) ? (0x10000 + pos[m]) : (0xFFFF - pos[m]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) code = max(code, __shfl_xor_sync(0xffffffffu, code, o));
      const int index = code >= 0x10000 ? code - 0x10000 : 0xFFFF - code;
      const int j = s.remaining[index];                        // the picked column, and the column that takes its scan position
      const int jl = s.remaining[nrem - 1];
      int own = -1;
#pragma unroll
      for (int m = 0; m < CPL; ++m) if ((j >> 5) == m) own = r4c[m];
      own = __shfl_sync(0xffffffffu, own, j & 31);
      __syncwarp();
      if (lane == (j & 31)) { rem &= ~(1u << (j >> 5)); sc |= 1u << (j >> 5); }
      if (lane == (jl & 31)) {
#pragma unroll
        for (int m = 0; m < CPL; ++m) if ((jl >> 5) == m) pos[m] = index;
      }
      if (lane == 0) s.remaining[index] = jl;
      --nrem;
      minVal = lowest;
      if (own == -1) sink = j; else i = own;
      __syncwarp();
    }
    if (!feasible) break;
    // ---- dual update (rows from shared memory, columns in registers), then the augmentation walk by lane 0
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const int j = lane + 32 * m;
      if (j < n) { s.spc[j] = spc[m]; s.path[j] = pth[m]; }
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      if (k == curRow) s.u[k] += minVal;
      else if (s.SR[k]) s.u[k] += minVal - s.spc[s.col4row[k]];
    }
#pragma unroll
    for (int m = 0; m < CPL; ++m) if (sc & (1u << m)) v[m] -= minVal - spc[m];
    __syncwarp();
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = s.path[j];
        s.row4col[j] = r;
        const int t = s.col4row[r];
        s.col4row[r] = j;
        j = t;
        if (r == curRow) break;
      }
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < CPL; ++m) {
      const int j = lane + 32 * m;
      if (j < n) r4c[m] = s.row4col[j];
    }
  }
  if (lane == 0 && !feasible) s.sink = -2;
}

__device__ void lsap_solve(const float* __restrict__ cost, int n, LsapShared& s) {
  if (threadIdx.x < 32) {
    if (n <= 32) lsap_solve_t<1>(cost, n, s);
    else if (n <= 64) lsap_solve_t<2>(cost, n, s);
    else if (n <= 128) lsap_solve_t<4>(cost, n, s);
    else lsap_solve_t<8>(cost, n, s);
  }
  __syncthreads();
}

// cost [batch, n, n] fp32 (row = target, column = current) -> col4row [batch, n]; one CTA per matrix.  Infeasible -> every entry -1.
// (the solver reads one cost row per relaxation step, a dependent access: the matrix is staged in shared memory when it fits -- n <= 208 --
// which took the 128 x 128 problem from 1.8 ms to well under scipy's 0.35 ms)
__global__ void __launch_bounds__(LS_THREADS) lsap_kernel(const float* __restrict__ cost, int n, int* __restrict__ col4row, int cost_in_smem) {
  __shared__ LsapShared s;
  extern __shared__ float ls_cost[];
  const float* c = cost + (size_t)blockIdx.x * n * n;
  if (cost_in_smem) {
    for (int k = threadIdx.x; k < n * n; k += LS_THREADS) ls_cost[k] = c[k];
    __syncthreads();
    c = ls_cost;
  }
  lsap_solve(c, n, s);
  for (int k = threadIdx.x; k < n; k += LS_THREADS) col4row[(size_t)blockIdx.x * n + k] = s.sink < 0 ? -1 : s.col4row[k];
}

// emb [videos, clips, n, e] fp32 -> indices [videos, clips, n]: indices[v, 0] = identity; indices[v, i] aligns clip i to the already
// aligned clip i - 1 (match_from_embds chained as maxtron_wc_model.py:342-346).  ws per video: 2 * n * e (normalised rows of the target
// and of the current clip) + n * n (cost) floats.  One CTA per video.
__global__ void __launch_bounds__(LS_THREADS) match_chain_kernel(const float* __restrict__ emb, int clips, int n, int e, int* __restrict__ indices,
                                                                 float* __restrict__ ws_all, int cost_in_smem) {
  __shared__ LsapShared s;
  extern __shared__ float ls_cost[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ev = emb + (size_t)blockIdx.x * clips * n * e;
  int* iv = indices + (size_t)blockIdx.x * clips * n;
  float* tgt = ws_all + (size_t)blockIdx.x * ((2 * (size_t)n * e + (size_t)n * n + 3) & ~(size_t)3);   // 16-byte aligned per video
  float* cur = tgt + (size_t)n * e;
  float* cost = cost_in_smem ? ls_cost : cur + (size_t)n * e;
  // x / x.norm(dim=1)[:, None]: one warp per row, fp32
  auto normalise = [&](const float* src, float* dst) {
    for (int r = warp; r < n; r += LS_THREADS / 32) {
      float ss = 0.f;
      for (int c = lane; c < e; c += 32) { const float x = src[(size_t)r * e + c]; ss += x * x; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float nrm = sqrtf(ss);
      for (int c = lane; c < e; c += 32) dst[(size_t)r * e + c] = src[(size_t)r * e + c] / nrm;
    }
  };
  normalise(ev, tgt);
  for (int k = tid; k < n; k += LS_THREADS) iv[k] = k;
  __syncthreads();
#ifdef AXVS_MATCH_PROF
  long long t_norm = 0, t_cost = 0, t_solve = 0, t_perm = 0;
  if (tid == 0) s.steps = 0;
#define MP(x) x
#else
#define MP(x)
#endif
  for (int i = 1; i < clips; ++i) {
    MP(long long c0 = clock64();)
    normalise(ev + (size_t)i * n * e, cur);
    __syncthreads();
    MP(long long c1 = clock64(); t_norm += c1 - c0;)
    // cost[t, c] = 1 - cur[c] . tgt[t]: a warp takes one target row at a time, its lanes the current rows c = lane, lane + 32, ...;
    // the 32 partial dot products of a lane's 4-element K slices are coalesced 16-byte loads (a transposed read of cur would be
    // stride-e: 2.2 M clk per 128 x 128 pair, this form: ~0.1 M)
    for (int t = warp; t < n; t += LS_THREADS / 32) {
      const float* b = tgt + (size_t)t * e;
      for (int c0 = 0; c0 < n; c0 += 32) {
        // lanes cooperate on 32 rows c0..c0+31: lane l accumulates row c0 + l, reading float4 pieces; consecutive lanes read
        // different rows, so stage through registers: each lane reads ITS row with 16-byte loads (e % 4 == 0 is checked by the host)
        const int c = c0 + lane;
        float acc = 0.f;
        if (c < n) {
          const float4* a4 = reinterpret_cast<const float4*>(cur + (size_t)c * e);
          const float4* b4 = reinterpret_cast<const float4*>(b);
          for (int k = 0; k < e / 4; ++k) {
            const float4 x = a4[k], y = b4[k];
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
          }
          cost[(size_t)t * n + c] = 1.f - acc;
        }
      }
    }
    __syncthreads();
    MP(long long c2 = clock64(); t_cost += c2 - c1;)
    lsap_solve(cost, n, s);
    __syncthreads();
    MP(long long c3 = clock64(); t_solve += c3 - c2;)
    for (int k = tid; k < n; k += LS_THREADS) iv[(size_t)i * n + k] = s.sink < 0 ? -1 : s.col4row[k];
    // the aligned current clip becomes the next target: tgt[t] = cur[col4row[t]]
    if (s.sink >= 0) {
      for (int idx = tid; idx < n * e; idx += LS_THREADS) {
        const int t = idx / e, c = idx - t * e;
        tgt[idx] = cur[(size_t)s.col4row[t] * e + c];
      }
    }
    __syncthreads();
    MP(t_perm += clock64() - c3;)
  }
  MP(if (tid == 0 && blockIdx.x == 0) printf("[match prof] norm %lld cost %lld solve %lld perm %lld clk, %lld steps over %d pairs\n", t_norm, t_cost, t_solve, t_perm, s.steps, clips - 1);)
}

}  // namespace axvs
