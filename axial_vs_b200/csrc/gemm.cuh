// Persistent warp-specialised tcgen05 GEMM:  out = epilogue( A[M,K] (bf16, row-major) * W[n_out,K]^T + bias ).
//
//  * A K-blocks (128 rows x 64 bf16) are gathered by producer warps (arbitrary row map) into SWIZZLE_128B
//    K-major shared-memory tiles; W K-blocks (256 rows x 64 bf16) arrive by 1-D TMA bulk copies from a
//    pre-swizzled packed image (see pack_weight_kernel), so no tensor map is needed.
//  * One elected thread issues tcgen05.mma (M=128, N=256, K=16) into a double-buffered TMEM accumulator
//    (2 x 256 fp32 columns); four epilogue warps drain it with tcgen05.ld (one row per thread).
//  * mbarrier rings: full/empty per smem stage, full/empty per TMEM stage.  Tiles are assigned
//    round-robin to a grid of min(#tiles, #SMs) CTAs.
#pragma once
#include "ptx.cuh"

namespace axvs {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KiB
constexpr int GEMM_W_BYTES = GEMM_BN * GEMM_BK * 2;   // 32 KiB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_W_BYTES;
constexpr int GEMM_EPI_WARPS = 8;   // two groups of four (TMEM lane quarter = warp & 3), each draining 128 of the 256 accumulator columns
constexpr int GEMM_APROD_WARPS = 8;
constexpr int GEMM_THREADS = (GEMM_EPI_WARPS + 2 + GEMM_APROD_WARPS) * 32;   // 320
constexpr int GEMM_STG_BYTES = GEMM_EPI_WARPS * 4096;   // per-warp 32 rows x 128 B transpose staging of the epilogue
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + GEMM_STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum RowMap : int { MAP_NONE = 0, MAP_HPASS = 1, MAP_WPASS = 2 };

struct AxialDims {
  int B, T, H, W;
  uint32_t pos_mod;     // != 0: the positional table covers ONE clip ([T, H, W, 256], pos_mod = T*H*W rows) and is shared by all B clips
  uint32_t pos_magic;   // floor(2^32 / pos_mod)
};

__host__ __forceinline__ AxialDims make_dims(int B, int T, int H, int W, bool shared_pos = false) {
  AxialDims d{B, T, H, W, 0u, 0u};
  if (shared_pos) {
    d.pos_mod = (uint32_t)T * H * W;
    d.pos_magic = (uint32_t)(0x100000000ull / d.pos_mod);
  }
  return d;
}

// canonical token -> row of the positional table (c mod pos_mod by a reciprocal multiply: the estimate is low by at most one)
__device__ __forceinline__ uint32_t pos_row(uint32_t c, const AxialDims& d) {
  if (!d.pos_mod) return c;
  const uint32_t r = c - __umulhi(c, d.pos_magic) * d.pos_mod;
  return r >= d.pos_mod ? r - d.pos_mod : r;
}

// pass-order row index -> canonical token index ((b*T + t)*H + h)*W + w
__device__ __forceinline__ int pass_to_canonical(int p, int mode, const AxialDims& d) {
  if (mode == MAP_HPASS) {   // p = ((b*W + w)*T + t)*H + h
    int h = p % d.H;
    int r = p / d.H;
    int t = r % d.T;
    r /= d.T;
    int w = r % d.W;
    int b = r / d.W;
    return ((b * d.T + t) * d.H + h) * d.W + w;
  } else if (mode == MAP_WPASS) {   // p = ((b*H + h)*T + t)*W + w
    int w = p % d.W;
    int r = p / d.W;
    int t = r % d.T;
    r /= d.T;
    int h = r % d.H;
    int b = r / d.H;
    return ((b * d.T + t) * d.H + h) * d.W + w;
  }
  return p;
}

struct GemmParams {
  // A operand
  const __nv_bfloat16* A;
  int lda;      // elements
  int M;        // logical rows
  int K;        // multiple of 64
  int a_diag;   // 1: logical row r reads A row r*F + (r % N) / n   (own-frame rows of x[rows, F, C])
                // 2: temporal 3-tap gather for the cross-clip ASPP convs (K = 3*256): rows are (b, t, q) with a_n = Q rows
                //    per time step, a_N = T steps, a_F = dilation; K-block kb reads channels (kb&3)*64.. of the row at
                //    time clamp(t + (kb/4 - 1) * dilation, 0, T-1)  (Conv1d k=3, padding 'same', replicate; CC:180-182)
                // 3: A32 is an fp32 NCHW feature map [images, K channels, a_n pixels] (Conv2d 1x1 input, WC/msdeformattn.py:355-358):
                //    logical row r = image r / a_n, pixel r % a_n; the producers transpose while converting (thread = pixel row,
                //    eight 4-byte loads at a stride of a_n floats -> one 16-byte chunk of the K-major image)
                // 4: A32 is an fp32 row-major matrix [M, lda] (token rows), converted to bf16 on the fly
  int a_N, a_n, a_F;
  int a_split;  // 1 = split-precision product (modes 3 / 4: fp32 source split by the producers; mode 0: A = bf16 [hi | lo], see the producer).  K counts 3 x the source channels: K-blocks [0, K/3) and [2K/3, K) carry
                //   hi = bf16(a), K-blocks [K/3, 2K/3) carry lo = bf16(a - hi); with weights packed as [hi | hi | lo] the accumulator
                //   holds a_hi w_hi + a_lo w_hi + a_hi w_lo, i.e. the fp32 product to ~2^-17 (kmax_axial: a softmax sits behind this GEMM)
  const float* A32;
  const float* A32b;    // mode 4 only: optional second fp32 matrix added element-wise (e.g. src + pos), same layout as A32
  int a32b_rows;        // > 0: A32b has only this many rows and is broadcast (row r reads A32b row r % a32b_rows)
  // W operand (packed) and bias
  const uint8_t* Wp;
  int w_rows_total;   // rows of the packed image (K-block stride = w_rows_total * 128 B)
  int w_row0;         // first weight row used by this GEMM (multiple of 8)
  int n_out;    // multiple of 256
  int n_valid;  // > 0: only columns < n_valid (a multiple of 32) are stored; the rest of n_out is zero-weight padding
  const float* bias;
  // epilogue
  float scale;  // applied after bias
  int relu;     // activation: 0 none, 1 ReLU, 2 GELU (erf form, nn.GELU default)
  void* out;    // bf16 or fp32
  int ldo;      // elements
  int out_col0; // column offset inside out rows
  int out_bf16;
  const float* resid;   // optional fp32 residual, same row map / ld as out (fp32 path only)
  int map_mode;         // RowMap applied to output (and residual) rows
  AxialDims dims;
  int out_ch;           // NCHW outputs: channels of the output tensor (0 = n_out); smaller than n_out when n_out is padded to 256 (n_valid)
  int a_img_rows;       // mode 4, > 0: logical row r lives at A32 + (r / a_img_rows) * a_img_stride + (r % a_img_rows) * lda (one pyramid level
  long long a_img_stride;   //   inside the multi-level token tensor [images, sum(H_l*W_l), 256]: a_img_rows = H_l*W_l, a_img_stride = len * 256)
  int out_img_rows;     // token-major fp32 outputs, > 0: the same per-image addressing for the output (and residual) rows
  long long out_img_stride;
  int out_nchw;         // > 0: fp32 output is NCHW [images, out_ch, out_nchw pixels] (Conv2d 1x1 output): row r = image r / out_nchw,
                        //      pixel r % out_nchw; every column is one coalesced 4-byte store per lane (lanes = consecutive pixels)
};

// v holds the accumulator for columns col..col+31 of logical row `row` (= row0 + lane; rows >= M are not stored).
// Token-major outputs go through a per-warp 4 KiB shared-memory transpose so that one store instruction writes whole rows pieces:
// 4 rows x 128 contiguous bytes (fp32) or 8 rows x 64 bytes (bf16) instead of 32 scattered 16-byte pieces; the residual is read the
// same way.  NCHW outputs are already coalesced (lanes = consecutive pixels of one channel).
__device__ __forceinline__ void gemm_epilogue_store(const GemmParams& p, int row, int row0, int col, float (&v)[32], const float4 (&b)[8],
                                                    uint8_t* stg, int lane) {
  // bias (loaded by the caller before it waits for the accumulator), scale, activation: one uniform branch per group, so the GELU
  // body (erff, 32 x unrolled) is not fetched by the projections that do not use it
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[4 * i] += b[i].x; v[4 * i + 1] += b[i].y; v[4 * i + 2] += b[i].z; v[4 * i + 3] += b[i].w; }
  if (p.scale != 1.f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= p.scale;
  }
  if (p.relu == 1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (p.relu == 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
  }
  if (p.out_nchw > 0) {
    if (row >= p.M) return;
    const int img = row / p.out_nchw, pix = row - img * p.out_nchw;
    float* o = reinterpret_cast<float*>(p.out) + ((size_t)img * (p.out_ch > 0 ? p.out_ch : p.n_out) + col) * p.out_nchw + pix;
#pragma unroll
    for (int i = 0; i < 32; ++i) o[(size_t)i * p.out_nchw] = v[i];
    return;
  }
  if (p.out_bf16) {
    // staging row = 64 B (4 pieces of 16 B), piece index XOR-swizzled with (row >> 1) & 3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      *reinterpret_cast<uint4*>(stg + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = u;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rl = 8 * i + (lane >> 2), piece = lane & 3;
      const uint4 u = *reinterpret_cast<const uint4*>(stg + rl * 64 + ((piece ^ ((rl >> 1) & 3)) << 4));
      if (row0 + rl < p.M) {
        const int orow = pass_to_canonical(row0 + rl, p.map_mode, p.dims);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)orow * p.ldo + p.out_col0 + col + piece * 8) = u;
      }
    }
    __syncwarp();
    return;
  }
  // fp32: staging row = 128 B (8 pieces), piece index XOR-swizzled with row & 7
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(stg + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  __syncwarp();
  float4 r[8];
  auto out_off = [&](int orow) -> size_t {
    if (p.out_img_rows > 0) {
      const int img = orow / p.out_img_rows;
      return (size_t)img * (size_t)p.out_img_stride + (size_t)(orow - img * p.out_img_rows) * p.ldo;
    }
    return (size_t)orow * p.ldo;
  };
  if (p.resid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + (lane >> 3), piece = lane & 7;
      r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + rl < p.M) {
        const int orow = pass_to_canonical(row0 + rl, p.map_mode, p.dims);
        r[i] = __ldg(reinterpret_cast<const float4*>(p.resid + out_off(orow) + p.out_col0 + col) + piece);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + (lane >> 3), piece = lane & 7;
    float4 u = *reinterpret_cast<const float4*>(stg + rl * 128 + ((piece ^ (rl & 7)) << 4));
    if (p.resid) { u.x += r[i].x; u.y += r[i].y; u.z += r[i].z; u.w += r[i].w; }
    if (row0 + rl < p.M) {
      const int orow = pass_to_canonical(row0 + rl, p.map_mode, p.dims);
      *(reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_off(orow) + p.out_col0 + col) + piece) = u;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_kernel(const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + GEMM_STAGES * GEMM_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + GEMM_STG_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * GEMM_STAGES;    // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_chunks = p.n_out / GEMM_BN;
  const int num_tiles = m_tiles * n_chunks;
  const int num_kb = p.K / GEMM_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1 + GEMM_APROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == GEMM_EPI_WARPS + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < GEMM_EPI_WARPS) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    uint32_t acc = 0, acc_phase = 0;
    uint8_t* stg = stg_base + warp * 4096;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / n_chunks, nc = tile % n_chunks;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = mt * GEMM_BM + (warp & 3) * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * GEMM_BN;
      const int c_begin = (warp >> 2) * (GEMM_BN / 2);
#pragma unroll 1
      for (int c = c_begin; c < c_begin + GEMM_BN / 2; c += 32) {
        if (p.n_valid > 0 && nc * GEMM_BN + c >= p.n_valid) break;      // padding columns
        float v[32];
        float4 b[8];
        tmem_ld32(taddr + c, v);
        if (p.bias) {                              // eight independent 16-byte loads, in flight while the accumulator is read
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc * GEMM_BN + c);
#pragma unroll
          for (int i = 0; i < 8; ++i) b[i] = __ldg(b4 + i);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        tmem_ld_wait();
        gemm_epilogue_store(p, row, mt * GEMM_BM + (warp & 3) * 32, nc * GEMM_BN + c, v, b, stg, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp == GEMM_EPI_WARPS) {
    // ===================== W producer: TMA bulk copies of packed weight K-blocks =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nc = tile % n_chunks;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* dst = smem + stage * GEMM_STAGE_BYTES + GEMM_A_BYTES;
          const uint8_t* src = p.Wp + ((size_t)kb * p.w_rows_total + (size_t)p.w_row0 + (size_t)nc * GEMM_BN) * 128;
          mbar_arrive_expect_tx(&full_bar[stage], GEMM_W_BYTES);
          tma_bulk_g2s(dst, src, GEMM_W_BYTES, &full_bar[stage]);
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == GEMM_EPI_WARPS + 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(GEMM_BM, GEMM_BN);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint32_t w_addr = a_addr + GEMM_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            umma_bf16(tmem_base + acc * GEMM_BN, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_addr + k * 32),
                      idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                      // frees the smem stage when the MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== A producers: gather rows -> swizzled smem =====================
    const int ptid = threadIdx.x - (GEMM_EPI_WARPS + 2) * 32;   // 0..127
    constexpr int PT = GEMM_APROD_WARPS * 32;
    constexpr int ITERS = GEMM_BM * 8 / PT;                     // 16-byte chunks per thread per K-block
    uint32_t stage = 0, phase = 0;
    if (p.a_diag >= 3) {
      // fp32 sources, loads software-pipelined over two register sets ACROSS K-blocks and tiles: work item it = (this CTA's tile
      // it / num_kb, K-block it % num_kb); item it+1 is in flight while item it is converted and stored, so the short K loops of
      // the 256-channel projections never drain the pipeline at a tile boundary.
      // Mode 3 (NCHW): q = i * PT + ptid -> pixel row q & 127, 16-byte chunk (8 channels) q >> 7: a warp's 32 lanes read 32
      // consecutive pixels of one channel per instruction.  Mode 4: row q >> 3, chunk q & 7 as for bf16 sources.
      // Plain bursts: a thread requests its ITERS x 8 (mode 3) / ITERS x 2 (mode 4) loads of a K-block, then converts and stores them; the other
      // seven producer warps cover the gap.  (The first version software-pipelined two register sets across K-blocks: the consumer of one set
      // then waits on a scoreboard shared with the other set's loads issued just before it, so only one set was ever in flight per warp --
      // the same effect as in the q|k|v producers, tools/microbench/ldg_rows.cu vs ldg_burst.cu.)  The row -> (image, pixel) division is
      // done once per tile, not once per K-block.
      float fa[ITERS][8];
      const float* rowp[ITERS];                                   // first element of the thread's row / pixel for K-block 0 (null: past the end)
      size_t row2[ITERS];                                         // mode 4: element offset into A32b
      auto prepare = [&](int tile) {
        const int mt = tile / n_chunks;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int q = i * PT + ptid;
          rowp[i] = nullptr;
          row2[i] = 0;
          if (p.a_diag == 3) {
            const int r = mt * GEMM_BM + (q & 127);
            if (r < p.M) {
              const int img = r / p.a_n;
              rowp[i] = p.A32 + ((size_t)img * (p.a_split ? p.K / 3 : p.K) + (q >> 7) * 8) * p.a_n + (r - img * p.a_n);
            }
          } else {
            const int r = mt * GEMM_BM + (q >> 3);
            if (r < p.M) {
              size_t o = (size_t)r * p.lda;
              if (p.a_img_rows > 0) {
                const int img = r / p.a_img_rows;
                o = (size_t)img * (size_t)p.a_img_stride + (size_t)(r - img * p.a_img_rows) * p.lda;
              }
              rowp[i] = p.A32 + o + (q & 7) * 8;
              if (p.A32b) row2[i] = p.a32b_rows > 0 ? (size_t)(r % p.a32b_rows) * p.lda + (q & 7) * 8 : o + (q & 7) * 8;
            }
          }
        }
      };
      auto load = [&](int kbi, float (&f)[ITERS][8]) {
        const int kb = p.a_split ? kbi % (num_kb / 3) : kbi;      // source K-block
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          if (rowp[i] == nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[i][j] = 0.f;
          } else if (p.a_diag == 3) {
            const float* s = rowp[i] + (size_t)kb * GEMM_BK * p.a_n;
#pragma unroll
            for (int j = 0; j < 8; ++j) f[i][j] = __ldg(s + (size_t)j * p.a_n);
          } else {
            const float4* s = reinterpret_cast<const float4*>(rowp[i] + kb * GEMM_BK);
            float4 a = __ldg(s), b = __ldg(s + 1);
            if (p.A32b) {
              const float4* s2 = reinterpret_cast<const float4*>(p.A32b + row2[i] + kb * GEMM_BK);
              const float4 a2 = __ldg(s2), b2 = __ldg(s2 + 1);
              a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w; b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
            }
            f[i][0] = a.x; f[i][1] = a.y; f[i][2] = a.z; f[i][3] = a.w; f[i][4] = b.x; f[i][5] = b.y; f[i][6] = b.z; f[i][7] = b.w;
          }
        }
      };
      auto store = [&](const float (&f)[ITERS][8], int it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = smem + stage * GEMM_STAGE_BYTES;
        const bool lo_part = p.a_split && ((it % num_kb) / (num_kb / 3)) == 1;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int q = i * PT + ptid;
          const uint32_t off = (p.a_diag == 3) ? sw128_offset(q & 127, q >> 7) : sw128_offset(q >> 3, q & 7);
          float g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = lo_part ? f[i][j] - __bfloat162float(__float2bfloat16_rn(f[i][j])) : f[i][j];
          *reinterpret_cast<uint4*>(dst + off) = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]),
                                                            pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        prepare(tile);
#pragma unroll 1
        for (int kb = 0; kb < num_kb; ++kb) {
          load(kb, fa);
          store(fa, kb);
        }
      }
    } else
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / n_chunks;
      // source row pointers are K-block independent
      const __nv_bfloat16* rowp[ITERS];
      int tstep[ITERS];
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        const int q = i * PT + ptid;
        const int r = mt * GEMM_BM + (q >> 3);
        tstep[i] = 0;
        if (r < p.M) {
          size_t ar = (p.a_diag == 1) ? (size_t)r * p.a_F + (size_t)((r % p.a_N) / p.a_n) : (size_t)r;
          rowp[i] = p.A + ar * p.lda + (q & 7) * 8;
          if (p.a_diag == 2) tstep[i] = (r / p.a_n) % p.a_N;
        } else {
          rowp[i] = nullptr;
        }
      }
      // two register sets: the loads of K-block kb + 1 are in flight while K-block kb waits for its stage and is stored
      auto loadk = [&](int kb, uint4 (&v)[ITERS]) {
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          if (p.a_diag == 2) {
            int t2 = tstep[i] + ((kb >> 2) - 1) * p.a_F;
            t2 = t2 < 0 ? 0 : (t2 >= p.a_N ? p.a_N - 1 : t2);
            v[i] = rowp[i] ? ldg_nc_v4(rowp[i] + (ptrdiff_t)(t2 - tstep[i]) * p.a_n * p.lda + (kb & 3) * GEMM_BK) : make_uint4(0, 0, 0, 0);
          } else {
            // a_split with a bf16 source: A holds [hi | lo] (2 x K/3 columns); K-blocks of the middle third read the lo half
            const int nk3 = num_kb / 3;
            const int kcol = p.a_split ? ((kb / nk3 == 1 ? nk3 : 0) + kb % nk3) * GEMM_BK : kb * GEMM_BK;
            v[i] = rowp[i] ? ldg_nc_v4(rowp[i] + kcol) : make_uint4(0, 0, 0, 0);
          }
        }
      };
      auto storek = [&](const uint4 (&v)[ITERS]) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = smem + stage * GEMM_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
          const int q = i * PT + ptid;
          *reinterpret_cast<uint4*>(dst + sw128_offset(q >> 3, q & 7)) = v[i];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
      };
      uint4 va[ITERS], vb[ITERS];
      loadk(0, va);
      for (int kb = 0; kb < num_kb; kb += 2) {
        if (kb + 1 < num_kb) loadk(kb + 1, vb);
        storek(va);
        if (kb + 1 < num_kb) {
          if (kb + 2 < num_kb) loadk(kb + 2, va);
          storek(vb);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == GEMM_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 [n_out, K] row-major -> bf16 packed image: [K/64][n_out/8][8 rows x 128 B, 16-B chunks XOR-swizzled by row&7]
__global__ void pack_weight_kernel(const float* __restrict__ w, int n_out, int K, uint8_t* __restrict__ packed) {
  const int total = n_out * (K / 8);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / (K / 8);
    const int c8 = idx % (K / 8);     // 8-element chunk along K
    const int kb = c8 >> 3, ch = c8 & 7;
    const float4* s = reinterpret_cast<const float4*>(w + (size_t)r * K + c8 * 8);
    float4 a = s[0], b = s[1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y);
    u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(packed + (size_t)kb * n_out * 128 + sw128_offset(r, ch)) = u;
  }
}

// fp32 [n_out, K] -> bf16 "unit" image for the fused kernels: 32 KiB units = [2 K-blocks][128 rows x 128 B, SWIZZLE_128B],
// i.e. one TMA bulk copy brings the operand of 8 UMMAs (N = 128, K = 128).  Unit order: k_major == 0 -> unit(rt, kg) =
// rt * (K/128) + kg (all K of a 128-row tile contiguous); k_major == 1 -> unit = kg * (n_out/128) + rt;
// k_major == 2 -> units of 256 rows x ONE K-block (the B operand of an N = 256 UMMA), unit = (r / 256) * (K / 64) + kb.
__global__ void pack_weight_units_kernel(const float* __restrict__ w, int n_out, int K, int k_major, uint8_t* __restrict__ packed) {
  const int total = n_out * (K / 8);
  const int KG = K / 128, RT = n_out / 128;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / (K / 8);
    const int c8 = idx % (K / 8);
    const int kb = c8 >> 3, ch = c8 & 7;
    const int rt = r >> 7, rl = r & 127, kg = kb >> 1, kb2 = kb & 1;
    const size_t unit = k_major ? (size_t)kg * RT + rt : (size_t)rt * KG + kg;
    const float4* s = reinterpret_cast<const float4*>(w + (size_t)r * K + c8 * 8);
    float4 a = s[0], b = s[1];
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y);
    u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y);
    u.w = pack_bf16x2(b.z, b.w);
    if (k_major == 2) {   // N = 256 units: [256 rows x 128 B] of ONE K-block, unit = (row tile of 256) * (K / 64) + K-block
      *reinterpret_cast<uint4*>(packed + ((size_t)(r >> 8) * (K / 64) + kb) * 32768 + sw128_offset(r & 255, ch)) = u;
      continue;
    }
    *reinterpret_cast<uint4*>(packed + unit * 32768 + (size_t)kb2 * 16384 + sw128_offset(rl, ch)) = u;
  }
}

}  // namespace axvs
