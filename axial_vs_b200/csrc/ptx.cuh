// sm_100a PTX wrappers: mbarrier, TMA bulk copy, tcgen05 (UMMA / TMEM), proxy fences.
// Hand-written for this project; bit layouts follow the PTX ISA tcgen05 descriptor tables.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace axvs {

// Optional wait-time profile (debug builds with -DAXVS_WAIT_PROFILE): cycles spent in each class of mbarrier wait,
// accumulated per role and summed over CTAs into g_wait_prof[]; read back with axvs_debug_read_waits().
#ifdef AXVS_WAIT_PROFILE
__device__ unsigned long long g_wait_prof[64];
#define AXVS_PROF_DECL(n) long long prof_acc_[n] = {}; const long long prof_t0_ = clock64();
#define AXVS_PROF_WAIT(i, stmt) { const long long t_ = clock64(); stmt; prof_acc_[i] += clock64() - t_; }
__device__ unsigned long long g_trace[512];
// timeline trace of ONE tile iteration of CTA 0 (clock64 is an SM-wide counter, so the roles' stamps are comparable)
#define AXVS_TRACE(cond, slot) if ((cond) && blockIdx.x == 0) g_trace[slot] = (unsigned long long)clock64();
#define AXVS_PROF_MARK(v) const long long v = clock64();
#define AXVS_PROF_SPAN(i, since) prof_acc_[i] += clock64() - (since);
#define AXVS_PROF_FLUSH(base, n, cond) if (cond) { for (int i_ = 0; i_ < (n); ++i_) atomicAdd(&g_wait_prof[(base) + i_], (unsigned long long)prof_acc_[i_]); \
                                                   atomicAdd(&g_wait_prof[(base) + (n)], (unsigned long long)(clock64() - prof_t0_)); }
#else
#define AXVS_PROF_DECL(n)
#define AXVS_PROF_WAIT(i, stmt) stmt;
#define AXVS_PROF_MARK(v)
#define AXVS_TRACE(cond, slot)
#define AXVS_PROF_SPAN(i, since)
#define AXVS_PROF_FLUSH(base, n, cond)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (context error, visible to the host) instead of hanging the GPU.
// (No printf on the failure path: a call site inside the epilogue loops would force the accumulator registers
// to be spilled around it.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifndef AXVS_NO_DEADLOCK_TRAP
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
#ifdef AXVS_TRAP_DEBUG
    if (++spins > (1u << 22)) {   // debug builds: say who is stuck on which barrier (byte offset from the dynamic smem base)
      extern __shared__ uint8_t axvs_dbg_smem_base[];
      printf("[axvs deadlock] block %d thread %d barrier +%u parity %u\n", blockIdx.x, threadIdx.x,
             smem_u32(bar) - smem_u32(axvs_dbg_smem_base), parity);
      break;                      // carry on with garbage so the kernel ends and the printf buffer is flushed
    }
#else
    if (++spins > (1u << 24)) __trap();
#endif
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ------------------------------------------------------------------ fences
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ TMA bulk copy (1-D, no tensor map)
// global -> shared, completion signalled on an mbarrier as transaction bytes. 16 B aligned, size % 16 == 0.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// shared -> global (bulk async group of the issuing thread).  The generic-proxy writes of the source must be fenced
// (fence.proxy.async) before the issue; the source may be overwritten after wait_group.read.
__device__ __forceinline__ void tma_bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ TMEM allocation
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, rows of 128 B, 8-row groups 1024 B apart.
//   [0,14) start address >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;             // LBO (ignored for swizzled K-major), 16 B
  d |= (uint64_t)(1024 >> 4) << 32;   // SBO = 1024 B
  d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;             // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: A = B = bf16, D = fp32, both operands K-major, dense.
//   [4,6) D fmt (1 = f32) | [7,10) A fmt (1 = bf16) | [10,13) B fmt | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued UMMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------ TMEM -> registers
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i), columns c..c+31.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ misc
// Byte offset of the 16-byte chunk `chunk` (0..7) of row `row` inside a SW128 K-major tile (128 B rows).
__device__ __host__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// bf16x2(max(lo, 0), max(hi, 0)): the ReLU is a modifier of the conversion instruction
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// packed fp32 pair add (one instruction for two lanes of a float2 held in a 64-bit register pair)
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {   // a * b + c, both lanes in one instruction
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 mul_f32x2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace axvs

namespace axvs {
// registers -> TMEM, 32 lanes x 32 columns (inverse of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// warp-group register re-allocation (all 4 warps of an aligned warpgroup must execute the same instruction)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// ---- warp-converged issue path (measured in tools/microbench/umma_issue.cu: a lone diverged thread pays ~570 clk per
// K-block for descriptor build + R2UR traffic; a converged warp with an elected lane reaches the tensor-pipe rate)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One K-block (4 UMMAs) + up to two commits, executed by the elected lane of a converged warp.
__device__ __forceinline__ void umma_kblock_elect(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool accumulate,
                                                  uint64_t* commit_a, uint64_t* commit_b) {
  const uint32_t a_lo = umma_desc_lo(a_addr), w_lo = umma_desc_lo(w_addr);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo(tmem_d, a_lo + 2 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
    if (commit_a) umma_commit(commit_a);
    if (commit_b) umma_commit(commit_b);
  }
  __syncwarp();
}
// One 32 KiB weight unit = two K-blocks (8 UMMAs, N = 128, K = 128) + up to four commits, elected lane of a converged warp.
// a0 / a1: shared-memory addresses of the two A K-blocks; w: address of the unit (its K-blocks are 16 KiB apart).
__device__ __forceinline__ void umma_unit_elect(uint32_t tmem_d, uint32_t a0, uint32_t a1, uint32_t w, uint32_t idesc, bool accumulate,
                                                uint64_t* c0, uint64_t* c1, uint64_t* c2, uint64_t* c3) {
  const uint32_t a0_lo = umma_desc_lo(a0), a1_lo = umma_desc_lo(a1), w_lo = umma_desc_lo(w);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo(tmem_d, a0_lo + 2 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo(tmem_d, a1_lo + 2 * k, w_lo + (16384 >> 4) + 2 * k, idesc, 1u);
    if (c0) umma_commit(c0);
    if (c1) umma_commit(c1);
    if (c2) umma_commit(c2);
    if (c3) umma_commit(c3);
  }
  __syncwarp();
}
// ---- A operand in TENSOR MEMORY (validated in tools/microbench/umma_tmem_a.cu): row m of A in lane m, 32-bit column c of the
// A region holds the bf16 elements k = 2c (low half) and k = 2c + 1 (high half); one K = 16 UMMA consumes 8 columns.
// Measured: N = 128 UMMAs run at 64 clk (the full tensor rate) in this form against ~90 clk with A in shared memory, where
// the 4 KiB (A) + 4 KiB (B) operand reads per instruction saturate the 128 B/clk shared-memory port.
__device__ __forceinline__ void umma_bf16_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One 32 KiB weight unit (8 UMMAs, N = 128, K = 128) with A in tensor memory: ta0 / ta1 = TMEM addresses (lane 0) of the
// 32 packed columns holding K = 0..63 / 64..127.
__device__ __forceinline__ void umma_unit_elect_ts(uint32_t tmem_d, uint32_t ta0, uint32_t ta1, uint32_t w, uint32_t idesc, bool accumulate,
                                                   uint64_t* c0, uint64_t* c1, uint64_t* c2) {
  const uint32_t w_lo = umma_desc_lo(w);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ts_lo(tmem_d, ta0 + 8 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_ts_lo(tmem_d, ta1 + 8 * k, w_lo + (16384 >> 4) + 2 * k, idesc, 1u);
    if (c0) umma_commit(c0);
    if (c1) umma_commit(c1);
    if (c2) umma_commit(c2);
  }
  __syncwarp();
}
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// shared memory -> TMEM: 128 rows x 32 bytes (one K = 16 slice of a SWIZZLE_128B K-major bf16 tile) -> 128 lanes x 8 columns,
// i.e. exactly the tensor-memory A operand of one UMMA (validated in tools/microbench/umma_tmem_a.cu).  Executes in issue order
// with the tcgen05.mma of the same thread; completion is tracked by tcgen05.commit like an MMA.
__device__ __forceinline__ void tmem_cp_128x256b_lo(uint32_t taddr, uint32_t desc_lo) {
  asm volatile(
      "{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\t"
      "tcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}" ::"r"(taddr), "r"(desc_lo), "r"(UMMA_DESC_HI)
      : "memory");
}
// One K-block image (128 rows x 64 bf16) -> 32 TMEM columns.
__device__ __forceinline__ void tmem_cp_kblock(uint32_t taddr, uint32_t smem_addr) {
  const uint32_t lo = umma_desc_lo(smem_addr);
#pragma unroll
  for (int k = 0; k < 4; ++k) tmem_cp_128x256b_lo(taddr + 8 * k, lo + 2 * k);
}
// registers -> TMEM, 32 lanes x 32 columns of raw 32-bit words (packed bf16 pairs)
__device__ __forceinline__ void tmem_st32u(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
  __syncwarp();
}

// Issue the 4 UMMAs (K = 64 = 4 x 16) of one K-block: D[128 x N] (+)= A_kblock[128 x 64] * W_kblock[N x 64]^T.
// a_addr / w_addr: shared-memory byte addresses of SWIZZLE_128B K-major tiles (1024 B aligned).
__device__ __forceinline__ void umma_kblock(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    umma_bf16(tmem_d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(w_addr + k * 32), idesc, (accumulate || k) ? 1u : 0u);
}
}  // namespace axvs

namespace axvs {
// ------------------------------------------------------------------ CTA pairs (cta_group::2), validated in tools/microbench/umma_2cta.cu
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// same, default (CTA-scope release) semantics: enough when the hand-off carries no generic-proxy data (e.g. "TMEM stage
// drained", ordered by tcgen05.fence) -- a cluster-scope release makes the warp wait for all its earlier writes
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait with acquire at cluster scope (arrivals may come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
#ifndef AXVS_NO_DEADLOCK_TRAP
    if (++spins > (1u << 24)) __trap();
#endif
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, both CTAs] (+)= [A_cta0; A_cta1] (256 x 16) * [B_cta0; B_cta1]^T (N x 16); issued by the leader CTA only
__device__ __forceinline__ void umma_bf16_lo_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// One weight unit for a CTA pair: each CTA holds HALF of the unit's rows as [2 K-blocks][64 rows x 128 B] (K-blocks 8 KiB
// apart); 8 UMMAs with M = 256 (128 rows per CTA), N = 128.  Elected lane of a converged warp in the leader CTA.
__device__ __forceinline__ void umma_unit_elect_pair(uint32_t tmem_d, uint32_t a0, uint32_t a1, uint32_t w, uint32_t idesc, bool accumulate,
                                                     uint64_t* c0, uint64_t* c1, uint64_t* c2, uint64_t* c3) {
  const uint32_t a0_lo = umma_desc_lo(a0), a1_lo = umma_desc_lo(a1), w_lo = umma_desc_lo(w);
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo_pair(tmem_d, a0_lo + 2 * k, w_lo + 2 * k, idesc, (accumulate || k) ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16_lo_pair(tmem_d, a1_lo + 2 * k, w_lo + (8192 >> 4) + 2 * k, idesc, 1u);
    if (c0) umma_commit_pair(c0);
    if (c1) umma_commit_pair(c1);
    if (c2) umma_commit_pair(c2);
    if (c3) umma_commit_pair(c3);
  }
  __syncwarp();
}
}  // namespace axvs
