// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a (inline PTX) and the one
// shared-memory operand format every tensor-core kernel in this library uses.
//
// Operand tiles ("chunked" format, no swizzle)
// -------------------------------------------
// A bf16 matrix T[ROWS][COLS] is stored in shared memory as 16-byte chunks of 8
// consecutive columns, chunks of one column-group contiguous over rows:
//
//     byte_offset(row, col) = row * 16 + (col / 8) * (ROWS * 16) + (col % 8) * 2
//
// This is the canonical SWIZZLE_NONE ("interleaved") UMMA layout for BOTH majors:
//   * rows = M/N index, cols = K index  -> K-major operand:  SBO = 128 B (next 8-row group),
//                                          LBO = ROWS*16 B (next 8-column K group)
//   * rows = K index, cols = M/N index  -> MN-major operand: SBO = ROWS*16 B (next 8 M/N
//                                          elements), LBO = 128 B (next 8 K rows)
// (core matrix = 8 rows x 16 bytes = 128 contiguous bytes in either reading).  A warp that
// writes one chunk per lane with consecutive rows stores 512 contiguous bytes: no bank
// conflicts, and the tile written for one GEMM (e.g. H1 as the K-major A of the forward
// pass) is read by another as its transpose (H1^T as the MN-major A of the weight-gradient
// GEMM) without being rewritten.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace rl8 {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// try_wait suspends the thread for a hardware-bounded time per attempt; the attempt cap turns
// a protocol bug (a barrier that never completes) into a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}
// The same with a suspend-time hint (ns): the thread sleeps in hardware until the phase completes instead of
// returning after a few dozen cycles.  For kernels whose many waiting warps would otherwise burn issue slots
// polling (measured on the split update kernels: 40 % of all executed instructions were wait-loop overhead);
// the latency-critical single-issuer waits of the bf16 kernels keep the short form.
constexpr uint32_t kMbarSuspendHintNs = 0x989680u;
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendHintNs)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_sleep(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_sleep(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s of SM clock: a protocol bug, not a wait
  }
}

// ---- bulk async copy global -> shared (TMA, non-tensor form; SASS: UBLKCP) ----------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bulk async copy shared -> global (bulk-group completion; SASS: UBLKCP) -------------------
// pull `bytes` of global memory into L2 ahead of the bulk copy that will stage it (no smem needed)
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// ... and have completed entirely (writes performed)
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// 128-bit vector reduction to global memory (sm_90+)
__device__ __forceinline__ void red_add_v4(float* gptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM -----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp gets row
// (lane field of taddr) + lane, columns [col, col + 32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 columns (small accumulators)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// raw 32-bit forms (packed bf16 pairs parked in TMEM next to the accumulators)
__device__ __forceinline__ void tmem_ld32_raw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32_raw(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
      "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_raw(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// issue-only forms: several loads in flight, one tmem_wait_ld() before the registers are used
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
        "=f"(v[7]), "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]),
        "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]),
        "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
        "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bits per instruction, x4 = 32 columns: thread (g = lane / 4, t = lane % 4) of the warp receives, for
// column block k = 0..3, rows {g, g + 8} of the 16 addressed lanes and columns {8 k + 2 t, 8 k + 2 t + 1}:
// v[4 k + 0..1] = row g, v[4 k + 2..3] = row g + 8 (the warp-level MMA accumulator layout; pinned by
// rl8_tc_selftest_tmem_16x256b).  Issue-only: tmem_wait_ld() before the registers are used.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void reg_fence16f(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]),
                    "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]),
                    "+f"(v[13]), "+f"(v[14]), "+f"(v[15])::"memory");
}
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Compiler-level fence on registers written by an issue-only tcgen05.ld: ordered after the
// (volatile) wait, it "redefines" the values so no use can be scheduled above the wait.
__device__ __forceinline__ void reg_fence16(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                    "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),
                    "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void reg_fence32(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]),
                    "+f"(v[7]), "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]),
                    "+f"(v[13]), "+f"(v[14]), "+f"(v[15])::"memory");
  asm volatile("" : "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]),
                    "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                    "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])::"memory");
}
__device__ __forceinline__ void tmem_wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// One lane of a CONVERGED warp (the lowest active one).  Guard single-thread tcgen05 / bulk-copy issue
// with `warp == k && elect_one()` rather than `tid == 32k`: behind a thread-index test the compiler
// wraps every tcgen05.mma in its own ELECT / BRA.U.ANY serialisation loop (~75 cycles per
// instruction, measured with rl8_tc_bench_mma); behind elect.sync it knows one lane is active.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// The CTA's single issuing thread: one elected lane of warp 0 (call it from converged code).
__device__ __forceinline__ bool cta_issuer() { return threadIdx.x < 32 && elect_one(); }

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor (SWIZZLE_NONE, sm_100 version field = 1).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::f16: bf16 x bf16 -> f32, dense.
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ |
         ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// the same with fp16 operands (format 0)
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return instr_desc(M, N, a_mn_major, b_mn_major) & ~((1u << 7) | (1u << 10));
}

// kind::tf32: fp32 containers in shared memory (10-bit mantissa used), K = 8 per instruction,
// both operands K-major.  A 16-byte chunk holds 4 elements; the core matrix is still 8 rows x 16 B.
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// round-to-nearest (ties away) onto the tf32 grid: what the kernels store, so the tensor core
// never sees mantissa bits it would drop
__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues on behalf of the CTA.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in tensor memory (128 lanes = rows, two bf16 K elements per 32-bit column,
// 8 columns per K = 16 instruction): D[tmem] (+)= A[tmem] * B[smem].  No shared-memory traffic for A.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- chunked operand tiles ------------------------------------------------------------------------
template <int ROWS>
__device__ __forceinline__ uint32_t chunk_offset(int row, int col8 /* col / 8 */) {
  return (uint32_t)row * 16u + (uint32_t)col8 * (uint32_t)(ROWS * 16);
}
// SFU-based gate non-linearities: ex2.approx.ftz / rcp.approx.ftz issued directly (2 ulp each) -- 4 instructions per
// sigmoid, 5 per tanh.  (__expf / __fdividef wrap every MUFU in denormal and range handling: the epilogue measured
// ~240 instructions per hidden unit and was issue-bound.)  This is the bf16 path, whose gate pre-activations already
// carry bf16 operand rounding; the fp32 path's cell kernel keeps the exact functions.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1 / (1 + 2^(-x log2 e)): ex2 -> 0 or +inf at the extremes, rcp(inf) = 0
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f)); }
// 2 sigmoid(2 x) - 1 (absolute error ~2e-7; saturates to +-1 without a clamp)
__device__ __forceinline__ float tanh_fast(float x) {
  return fmaf(2.0f, rcp_ftz(1.0f + ex2_ftz(x * -2.8853900817779268f)), -1.0f);
}

// T128 layout of a bf16 matrix X[R][C] in GLOBAL memory (R padded to a multiple of 128): 16-byte chunks of 8
// consecutive columns, [row tile of 128][column group][row in tile][8] -- every 128-row x 8k-column block is stored
// exactly as the chunked shared-memory operand tile it will become, whichever GEMM reads it (rows as M / N with the
// columns as K, or rows as K with the columns as M / N), so staging an operand is one bulk copy.
__host__ __device__ __forceinline__ int64_t t128_offset(int64_t r, int col8, int ncol8) {
  return (((r >> 7) * ncol8 + col8) << 11) + ((r & 127) << 4);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// relu + round-to-nearest pack of two floats (lo -> bits 0..15)
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void store_chunk_relu(uint8_t* tile, uint32_t off, const float* v) {
  uint4 q;
  q.x = pack_relu_bf16x2(v[0], v[1]);
  q.y = pack_relu_bf16x2(v[2], v[3]);
  q.z = pack_relu_bf16x2(v[4], v[5]);
  q.w = pack_relu_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + off) = q;
}
// 0xFFFF in each half whose bf16 value is > 0
__device__ __forceinline__ uint32_t gt0_mask_bf16x2(uint32_t h) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&h);
  return __hgt2_mask(v, __floats2bfloat162_rn(0.0f, 0.0f));
}
// store 8 consecutive columns of one row
__device__ __forceinline__ void store_chunk(uint8_t* tile, uint32_t off, const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + off) = q;
}

// Issue the K loop of one GEMM whose operands are chunked tiles.
//   a_rows_is_k / b_rows_is_k: operand tile rows index K (MN-major) instead of M/N (K-major)
//   A tile has A_ROWS rows, B tile B_ROWS rows; k_total = contraction length (multiple of 16)
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_saddr, int A_ROWS,
                                           bool a_rows_is_k, uint32_t b_saddr, int B_ROWS,
                                           bool b_rows_is_k, int M, int N, int k_total,
                                           bool accumulate_first) {
  const uint32_t idesc = instr_desc(M, N, a_rows_is_k ? 1 : 0, b_rows_is_k ? 1 : 0);
  // K-major: LBO = ROWS*16, SBO = 128, 16 K elements = 2 column groups -> advance 2*ROWS*16
  // MN-major: LBO = 128, SBO = ROWS*16, 16 K elements = 16 rows -> advance 256
  const uint32_t a_lbo = a_rows_is_k ? 128u : (uint32_t)A_ROWS * 16u;
  const uint32_t a_sbo = a_rows_is_k ? (uint32_t)A_ROWS * 16u : 128u;
  const uint32_t b_lbo = b_rows_is_k ? 128u : (uint32_t)B_ROWS * 16u;
  const uint32_t b_sbo = b_rows_is_k ? (uint32_t)B_ROWS * 16u : 128u;
  const uint32_t a_step = a_rows_is_k ? 256u : 2u * A_ROWS * 16u;
  const uint32_t b_step = b_rows_is_k ? 256u : 2u * B_ROWS * 16u;
  const int steps = k_total / 16;
  // The start-address field is the low 14 bits of the descriptor (bytes >> 4): stepping K is one
  // 64-bit add per operand, so the issuing thread spends its cycles on tcgen05.mma, not on
  // rebuilding descriptors (a tile never crosses the 256 KB the field spans).
  uint64_t ad = smem_desc(a_saddr, a_lbo, a_sbo);
  uint64_t bd = smem_desc(b_saddr, b_lbo, b_sbo);
  const uint64_t a_inc = a_step >> 4, b_inc = b_step >> 4;
#pragma unroll 4
  for (int s = 0; s < steps; ++s) {
    mma_bf16(d_tmem, ad, bd, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    ad += a_inc, bd += b_inc;
  }
}

}  // namespace tc
}  // namespace rl8
