// fp32-accurate GEMMs on the half-precision tensor pipe (kind::f16).  BF16 PIECES: every fp32 operand x is split into
// bf16 pieces
//
//     x = p0 + p1 + p2 (+ 2^-27 |x|),   p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1)
//
// (each subtraction is exact in fp32) and a product a*b is accumulated in fp32 TMEM as the sum of the
// piece products that matter:
//
//     6 terms  a0b0 + a0b1 + a1b0 + a1b1 + a0b2 + a2b0     dropped terms <= 2^-25 |ab|   ("x3")
//     3 terms  a0b0 + a0b1 + a1b0  (two pieces per operand)  dropped terms <= 2^-16 |ab|   ("x2")
//
// The x3 form reproduces an fp32 nn.Linear within fp32 rounding (the reference's arithmetic,
// src/rl8/models/_feedforward.py:263-289, 336-362); the x2 form is used where the consumer is a gradient
// (tests bound gradients at 1e-4).
//
// FP16 PIECES ("h2", the form the update kernels use).  An fp16 piece carries 11 significant bits against bf16's 8, and
// kind::f16 multiplies fp16 operands at the same rate, so TWO pieces per operand and THREE piece products
//
//     x = h0 + h1 (+ 2^-22 |x|),   h0 = f16(x), h1 = f16(x - h0);      a0b0 + a0b1 + a1b0     dropped a1b1 <= 2^-22 |ab|
//
// carry 22 bits per operand: 8e-8 of the result at K = 256 (tests/test_cpu_split_arith.py) against 6e-9 for the six bf16
// products and 5e-7 for an fp32 dot product accumulated in fp32 (the reference's own nn.Linear) -- i.e. still below the
// rounding the reference itself carries, at half the tensor-pipe work and two thirds of the splitting work.  The
// price is fp16's exponent range: every operand tensor is multiplied by a power of two s (exact) chosen from a bound of
// its magnitude so that |x s| <= 2^14 (pow2_scale_for); elements down to 2^-17 of the bound keep the full 2^-22
// relative accuracy, smaller ones an absolute error of 2^-39 of the bound (fp16 subnormals).  The consumer multiplies
// the accumulator by 1 / (s_a s_b), also exact.
//
// The MMAs are tcgen05.mma.cta_group::2: a pair of CTAs (one cluster) works on one 256-row tile, each CTA
// stages ITS 128 rows of A and ITS 128-row half of B (the N index), so per CTA an N = 256 instruction reads
// 4 KB + 4 KB of shared memory per 128 cycles instead of 4 KB + 8 KB -- with three pieces per operand the
// single-CTA form would sit at the shared-memory bandwidth limit.  Accumulators: 128 lanes x 256 columns
// in each CTA's tensor memory (rows of the leader = tile rows 0..127, of the peer = 128..255).
#pragma once
#include <cuda_fp16.h>

#include "tc.cuh"

namespace rl8 {
namespace tc {

// ---- cluster / pair primitives --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Arrive on the mbarrier at the same offset in CTA `rank`.  No cluster-scope release / acquire: nothing written
// through the generic proxy crosses CTAs behind these barriers -- operand tiles are read by each CTA's tensor core
// from its own shared memory (made visible by the writer's fence.proxy.async), accumulators are ordered by the
// tcgen05 fences -- and a cluster-scope acquire compiles to an L1 invalidate (CCTL.IVALL) on every poll.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  const uint32_t addr = map_to_cta(smem_u32(bar), rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// waits on a barrier whose arrivals may come from the peer CTA or from a multicast commit
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// ... sleeping form (suspend-time hint): measured better where all 16 worker warps wait at once and nothing else
// needs their issue slots (the weight-gradient kernel's ring), worse on the latency-critical waits
__device__ __forceinline__ void mbar_wait_cluster_sleep(uint64_t* bar, uint32_t parity) { mbar_wait_sleep(bar, parity); }

// ---- tensor memory, pair form (one warp of EACH CTA of the pair calls these) -------------------------------
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem of both CTAs] (+)= A * B with M = 256 (128 rows per CTA), issued by one thread of the LEADER CTA.
// The descriptors address the leader's shared memory; the peer's tensor core reads the same offsets of its own.
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier at this offset in BOTH CTAs receives one arrival when every MMA issued so far has completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// ---- splitting ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf16lo_f32(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi_f32(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// packed fp32 pairs (FFMA2 / FADD2 on sm_100: two independent round-to-nearest operations per instruction)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float2 f16x2_f32(uint32_t p) { return __half22float2(*reinterpret_cast<const __half2*>(&p)); }

// The largest power of two s with bound * s <= 2^14 (fp16 pieces: 2^14 leaves a factor 4 below the largest finite
// fp16 value for the rounding of the bound itself); 1 when the bound is zero or not finite.  Any finite bound gets a
// normal s (k >= -113); tiny bounds stop at 2^60 (tensors below 2^-46 lose piece precision, not range) so that the
// product of two inverse scales never underflows.
__host__ __device__ __forceinline__ float pow2_scale_for(float bound) {
  if (!(bound > 0.0f) || !(bound < 1.0e38f)) return 1.0f;
  int e;
  frexpf(bound, &e);  // bound = m 2^e, 0.5 <= m < 1
  int k = 14 - e;
  k = k > 60 ? 60 : k;
  return ldexpf(1.0f, k);
}

// two fp32 values -> NP packed 16-bit pairs (piece k of both values in q[k]; x0 in the low half); F16: fp16 pieces
template <int NP, bool F16 = false>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t* q) {
  if constexpr (F16) {
    static_assert(NP <= 2, "fp16 pieces: two carry 22 bits");
    q[0] = cvt_f16x2(x0, x1);
    if constexpr (NP > 1) {
      const float2 r = fsub2(make_float2(x0, x1), f16x2_f32(q[0]));  // exact: the residual of a rounding
      q[1] = cvt_f16x2(r.x, r.y);
    }
  } else {
    float2 x = make_float2(x0, x1);
    q[0] = cvt_bf16x2(x.x, x.y);
    if constexpr (NP > 1) {
      x = fsub2(x, make_float2(bf16lo_f32(q[0]), bf16hi_f32(q[0])));  // exact: the residual of a rounding
      q[1] = cvt_bf16x2(x.x, x.y);
    }
    if constexpr (NP > 2) {
      x = fsub2(x, make_float2(bf16lo_f32(q[1]), bf16hi_f32(q[1])));
      q[2] = cvt_bf16x2(x.x, x.y);
    }
  }
}
// eight consecutive K (or MN) elements -> one 16-byte chunk per piece at tile_k + off
template <int NP, bool F16 = false>
__device__ __forceinline__ void store_split_chunk(uint8_t* const* piece_tiles, uint32_t off, const float* v) {
  uint32_t q[4][NP];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair<NP, F16>(v[2 * i], v[2 * i + 1], q[i]);
#pragma unroll
  for (int k = 0; k < NP; ++k)
    *reinterpret_cast<uint4*>(piece_tiles[k] + off) = make_uint4(q[0][k], q[1][k], q[2][k], q[3][k]);
}

// piece pairs (a piece, b piece) in issue order: small products first within a K step
//   x3: a2b0 a0b2 a1b1 a1b0 a0b1 a0b0     x2: a1b0 a0b1 a0b0
template <int NP>
struct Terms;
template <>
struct Terms<3> {
  static constexpr int n = 6;
  __host__ __device__ static constexpr int a(int i) { return i == 0 ? 2 : i == 1 ? 0 : i == 2 ? 1 : i == 3 ? 1 : 0; }
  __host__ __device__ static constexpr int b(int i) { return i == 0 ? 0 : i == 1 ? 2 : i == 2 ? 1 : i == 3 ? 0 : i == 4 ? 1 : 0; }
};
template <>
struct Terms<2> {
  static constexpr int n = 3;
  __host__ __device__ static constexpr int a(int i) { return i == 0 ? 1 : 0; }
  __host__ __device__ static constexpr int b(int i) { return i == 1 ? 1 : 0; }
};

}  // namespace tc
}  // namespace rl8
