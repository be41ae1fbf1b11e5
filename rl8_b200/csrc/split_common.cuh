// Shared pieces of the split-operand pair kernels (split_tc.cu: forward / collect; update_x3.cu: PPO update).
#pragma once
#include "mlp_tc.cuh"
#include "split_tc.cuh"

namespace rl8 {

using namespace tc;

constexpr int kXKc = 32;                              // contraction values per ring stage (two K = 16 instructions)
constexpr int kXPieceBytes = TILE * (kXKc / 8) * 16;  // 8192: one piece of one operand stage, [128 rows][4 column groups]
constexpr int kXThreads = 544;                        // 16 worker warps + the TMEM / MMA warp
constexpr int kXImgBytes = (H / kXKc) * 2 * 3 * kXPieceBytes;  // 393216: W2 piece image (room for three pieces)

// Piece format of the update kernels (update_x3.cu): two fp16 pieces + power-of-two operand scales (three piece
// products per fp32 product), or -- make X3_BF16=1, the first form of this path -- three bf16 pieces (six).
#ifndef RL8_X3_F16
#define RL8_X3_F16 1
#endif
constexpr bool kUF16 = RL8_X3_F16 != 0;
constexpr int kUNP = kUF16 ? 2 : 3;  // pieces per operand of the forward contraction
__host__ __device__ constexpr uint32_t upd_idesc(int a_mn_major, int b_mn_major) {
  return kUF16 ? instr_desc_f16(256, H, a_mn_major, b_mn_major) : instr_desc(256, H, a_mn_major, b_mn_major);
}

// Magnitudes behind the operand scales of the fp16 form; device memory, rewritten by every rl8_ppo_minibatch call.
struct X3Scales {
  float w2f[2];      // power-of-two scale of the forward W2 piece image, per network (pack_w2_pieces_kernel)
  float w2b[2];      // ... of the transposed image (value network: with W3 folded in)
  uint32_t dmax[2];  // bits of max |dOut| over the rows seen so far, per network (forward kernel, atomicMax)
  uint32_t omax;     // bits of max |obs| over the T slabs of the batch (absmax_bits_kernel)
  uint32_t pad;
};

// One ring stage of a K-major operand pair: piece tiles [128 rows][4 column groups of 8 K values],
// off(r, c) = r * 16 + c * 2048  (LBO 2048 = next K group, SBO 128 = next 8 rows).
//
// K ORDER.  Stage kc holds the global K groups {8 g + kc : g = 0..3} as its local column groups g, so the
// worker thread (row, g) that fills local group g of every stage covers the 64 consecutive K values
// [64 g, 64 g + 64) of its row over the 8 stages of a tile (its ReLU-mask bits are two whole words).
// Any order works for a contraction as long as both operands use the same one.
template <int NP>
struct StageX {
  uint8_t a[NP][kXPieceBytes];
  uint8_t b[NP][kXPieceBytes];
};
__host__ __device__ constexpr int stage_kgroup(int kc, int g) { return g * 8 + kc; }  // global group of 8 K values

// every piece product of one ring stage, M = 256 (pair), N = 256, both operands K-major
template <int NP>
__device__ __forceinline__ void issue_stage(uint32_t acc_tmem, const StageX<NP>& stg, uint32_t idesc,
                                            bool accumulate_first) {
  using T = Terms<NP>;
#pragma unroll
  for (int ks = 0; ks < kXKc / 16; ++ks) {
#pragma unroll
    for (int i = 0; i < T::n; ++i) {
      const uint64_t ad = smem_desc(smem_u32(stg.a[T::a(i)]) + ks * 4096, 2048, 128);
      const uint64_t bd = smem_desc(smem_u32(stg.b[T::b(i)]) + ks * 4096, 2048, 128);
      mma_bf16_pair(acc_tmem, ad, bd, idesc, (accumulate_first || ks > 0 || i > 0) ? 1u : 0u);
    }
  }
}

// The same with the leading products a0 b0 and the correction products in separate accumulators (the tensor pipe
// truncates the accumulator after every instruction: the leading accumulator then sees 2 truncations per stage
// instead of 2 * n, and the corrections are 2^-8 of the result).  The consumer adds the two in fp32.
template <int NP>
__device__ __forceinline__ void issue_stage_split(uint32_t lead_tmem, uint32_t corr_tmem, const StageX<NP>& stg,
                                                  uint32_t idesc, bool accumulate_first) {
  using T = Terms<NP>;
#pragma unroll
  for (int ks = 0; ks < kXKc / 16; ++ks) {
#pragma unroll
    for (int i = 0; i < T::n; ++i) {
      const uint64_t ad = smem_desc(smem_u32(stg.a[T::a(i)]) + ks * 4096, 2048, 128);
      const uint64_t bd = smem_desc(smem_u32(stg.b[T::b(i)]) + ks * 4096, 2048, 128);
      const bool lead = T::a(i) == 0 && T::b(i) == 0;
      const bool first = !accumulate_first && ks == 0 && (lead || i == 0);
      mma_bf16_pair(lead ? lead_tmem : corr_tmem, ad, bd, idesc, first ? 0u : 1u);
    }
  }
}

// One operand is a 0 / 1 mask (a single exact piece, tile b[0]); the other has NP pieces: NP products per K step,
// a0 b0 into the leading accumulator, the rest into the corrections.
template <int NP>
__device__ __forceinline__ void issue_stage_mask_b(uint32_t lead_tmem, uint32_t corr_tmem, const StageX<NP>& stg,
                                                   uint32_t idesc, bool accumulate_first) {
#pragma unroll
  for (int ks = 0; ks < kXKc / 16; ++ks) {
    const uint64_t bd = smem_desc(smem_u32(stg.b[0]) + ks * 4096, 2048, 128);
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {  // small products first
      const uint64_t ad = smem_desc(smem_u32(stg.a[i]) + ks * 4096, 2048, 128);
      // lead_tmem == corr_tmem: one accumulator, only the first product issued (i = NP - 1) overwrites it
      const bool first = !accumulate_first && ks == 0 && (i == NP - 1 || (i == 0 && lead_tmem != corr_tmem));
      mma_bf16_pair(i == 0 ? lead_tmem : corr_tmem, ad, bd, idesc, first ? 0u : 1u);
    }
  }
}

// [W1 | b1] in fp32, input-major: w1t[d][i] = W1[i][d] (d < D <= 7; rows D..6 unused), w1t[7][i] = b1[i];
// scale: a power of two (exact) -- h1_chunk then yields scale * H1, bit for bit.  A NEGATIVE scale stages
// -|scale| [W1 | b1] for h1_chunk<true> (no entry is -0: a bias of zero must not turn an exact zero sum into -0).
__device__ __forceinline__ void stage_w1t(float (*w1t)[H], const NetParams& np, float scale = 1.0f) {
  for (int e = threadIdx.x; e < H * 8; e += blockDim.x) {
    const int i = e & (H - 1), d = e >> 8;
    w1t[d][i] = __fadd_rn(scale * (d < np.D ? np.w1[i * np.D + d] : (d == 7 ? np.b1[i] : 0.0f)), 0.0f);
  }
}
// max over the block of a non-negative value (all threads call; red: >= 32 floats of shared memory)
__device__ __forceinline__ float block_max_nonneg(float v, float* red) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = 0.0f;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmaxf(m, red[w]);
  return m;
}
// Bound of H1 = relu(W1 obs + b1) over every row whose |obs| <= omax:  max_i (|b1[i]| + omax sum_d |W1[i][d]|)
__device__ __forceinline__ float h1_bound(const NetParams& np, float omax, float* red) {
  float b = 0.0f;
  if (threadIdx.x < H) {
    const int i = threadIdx.x;
    float sum = 0.0f;
    for (int d = 0; d < np.D; ++d) sum += fabsf(np.w1[i * np.D + d]);
    b = fabsf(np.b1[i]) + omax * sum;
  }
  return block_max_nonneg(b, red);
}
// max_j sum_p |W3[p][j]|: dZ2 = [H2 > 0] .* (dOut W3) is bounded by max |dOut| times this
__device__ __forceinline__ float w3_colsum_bound(const NetParams& np, float* red) {
  float b = 0.0f;
  if (threadIdx.x < H)
    for (int p = 0; p < np.P; ++p) b += fabsf(np.w3[p * H + threadIdx.x]);
  return block_max_nonneg(b, red);
}
// the row's observations duplicated into register pairs (operands of the packed FMAs)
struct ObsPairs {
  float2 v[7];
};
__device__ __forceinline__ ObsPairs obs_pairs(const float* ob) {
  ObsPairs o;
#pragma unroll
  for (int d = 0; d < 7; ++d) o.v[d] = make_float2(ob[d], ob[d]);
  return o;
}
// Z1[row][c0 .. c0 + 8) = b1[c] + sum_d obs[d] W1[c][d] in fp32: bias first, then d ascending -- one FMA per term
// (two columns per FFMA2); H1 = relu(Z1) -> v; returns the 8 ReLU-mask bits (bit e: column c0 + e is positive).
// NEG: w1t holds -s [W1 | b1] (stage_w1t with a negative scale), so the sums are -s Z1: v = min(-s Z1, 0) = -s H1 and
// the mask bit is the SIGN bit of the sum, gathered with one funnel shift per column instead of compare / select / or
// (the consumers fold the sign into the factor that removes the operand scales).
template <bool NEG = false>
__device__ __forceinline__ uint32_t h1_chunk(const float (*w1t)[H], const ObsPairs& ob, int D, int c0, float* v) {
  float2 z[4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(&w1t[7][c0]);
    const float4 b1 = *reinterpret_cast<const float4*>(&w1t[7][c0 + 4]);
    z[0] = make_float2(b0.x, b0.y), z[1] = make_float2(b0.z, b0.w);
    z[2] = make_float2(b1.x, b1.y), z[3] = make_float2(b1.z, b1.w);
  }
#pragma unroll
  for (int d = 0; d < 7; ++d) {
    if (d < D) {  // uniform
      const float4 w0 = *reinterpret_cast<const float4*>(&w1t[d][c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&w1t[d][c0 + 4]);
      z[0] = ffma2(ob.v[d], make_float2(w0.x, w0.y), z[0]);
      z[1] = ffma2(ob.v[d], make_float2(w0.z, w0.w), z[1]);
      z[2] = ffma2(ob.v[d], make_float2(w1.x, w1.y), z[2]);
      z[3] = ffma2(ob.v[d], make_float2(w1.z, w1.w), z[3]);
    }
  }
  uint32_t bits = 0u;
  if constexpr (NEG) {
#pragma unroll
    for (int j = 3; j >= 0; --j) {  // bits = (bits << 1) | sign: the last column shifted in lands in bit 0
      bits = __funnelshift_l(__float_as_uint(z[j].y), bits, 1);
      bits = __funnelshift_l(__float_as_uint(z[j].x), bits, 1);
      v[2 * j] = fminf(z[j].x, 0.0f), v[2 * j + 1] = fminf(z[j].y, 0.0f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bits |= (z[j].x > 0.0f ? 1u : 0u) << (2 * j) | (z[j].y > 0.0f ? 1u : 0u) << (2 * j + 1);
      v[2 * j] = fmaxf(z[j].x, 0.0f), v[2 * j + 1] = fmaxf(z[j].y, 0.0f);
    }
  }
  return bits;
}

// Two rows at once (the same 8 columns): every [W1 | b1] value is read from shared memory once for both rows -- the
// producers of the split kernels are bound by the number of shared-memory load instructions, not by their bytes.
template <bool NEG>
__device__ __forceinline__ void h1_chunk2(const float (*w1t)[H], const ObsPairs& oa, const ObsPairs& ob, int D, int c0,
                                          float* va, float* vb, uint32_t& bits_a, uint32_t& bits_b) {
  float2 za[4], zb[4];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(&w1t[7][c0]);
    const float4 b1 = *reinterpret_cast<const float4*>(&w1t[7][c0 + 4]);
    za[0] = zb[0] = make_float2(b0.x, b0.y), za[1] = zb[1] = make_float2(b0.z, b0.w);
    za[2] = zb[2] = make_float2(b1.x, b1.y), za[3] = zb[3] = make_float2(b1.z, b1.w);
  }
#pragma unroll
  for (int d = 0; d < 7; ++d) {
    if (d < D) {  // uniform
      const float4 w0 = *reinterpret_cast<const float4*>(&w1t[d][c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&w1t[d][c0 + 4]);
      const float2 p0 = make_float2(w0.x, w0.y), p1 = make_float2(w0.z, w0.w);
      const float2 p2 = make_float2(w1.x, w1.y), p3 = make_float2(w1.z, w1.w);
      za[0] = ffma2(oa.v[d], p0, za[0]), zb[0] = ffma2(ob.v[d], p0, zb[0]);
      za[1] = ffma2(oa.v[d], p1, za[1]), zb[1] = ffma2(ob.v[d], p1, zb[1]);
      za[2] = ffma2(oa.v[d], p2, za[2]), zb[2] = ffma2(ob.v[d], p2, zb[2]);
      za[3] = ffma2(oa.v[d], p3, za[3]), zb[3] = ffma2(ob.v[d], p3, zb[3]);
    }
  }
  bits_a = bits_b = 0u;
  static_assert(NEG, "h1_chunk2: the sign-bit form only");
#pragma unroll
  for (int j = 3; j >= 0; --j) {
    bits_a = __funnelshift_l(__float_as_uint(za[j].y), bits_a, 1);
    bits_a = __funnelshift_l(__float_as_uint(za[j].x), bits_a, 1);
    bits_b = __funnelshift_l(__float_as_uint(zb[j].y), bits_b, 1);
    bits_b = __funnelshift_l(__float_as_uint(zb[j].x), bits_b, 1);
    va[2 * j] = fminf(za[j].x, 0.0f), va[2 * j + 1] = fminf(za[j].y, 0.0f);
    vb[2 * j] = fminf(zb[j].x, 0.0f), vb[2 * j + 1] = fminf(zb[j].y, 0.0f);
  }
}

// eight mask bits -> eight 16-bit values {0, 1} (one 16-byte operand chunk; a mask is exact in ONE piece)
template <bool F16 = false>
__device__ __forceinline__ uint4 mask_byte_to_bf16x8(uint32_t byte) {
  constexpr uint32_t one = F16 ? 0x3c00u : 0x3f80u;
  uint32_t q[4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    q[e] = ((byte >> (2 * e)) & 1u) * one | ((byte >> (2 * e + 1)) & 1u) * (one << 16);
  return make_uint4(q[0], q[1], q[2], q[3]);
}

// the same through a 16-entry table in shared memory (entry n: the two words of mask bits n & 15): two 8-byte loads
// per chunk, no bank conflicts (the 16 entries cover the 32 banks once, equal entries broadcast)
template <bool F16>
__device__ __forceinline__ void fill_mask_lut(uint2* lut) {
  if (threadIdx.x < 16) {
    const uint4 q = mask_byte_to_bf16x8<F16>(threadIdx.x);
    lut[threadIdx.x] = make_uint2(q.x, q.y);
  }
}
__device__ __forceinline__ uint4 mask_byte_lut(const uint2* lut, uint32_t byte) {
  const uint2 lo = lut[byte & 15u], hi = lut[(byte >> 4) & 15u];
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ void worker_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// the four worker warps {q, q + 4, q + 8, q + 12} that share the accumulator lanes 32 q .. 32 q + 31 (named barriers 2..5)
__device__ __forceinline__ void quarter_bar_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 2) : "memory"); }
// the four worker warps {4 cq .. 4 cq + 3} that share the accumulator columns 64 cq .. 64 cq + 63 (named barriers 6..9)
__device__ __forceinline__ void colq_bar_sync(int cq) { asm volatile("bar.sync %0, 128;" ::"r"(cq + 6) : "memory"); }

// lane l of the warp returns sum over the 32 lanes m of x_m[l]  (x is destroyed): 31 shuffles
__device__ __forceinline__ float transpose_reduce32(float* x, int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? x[i] : x[i + o];
      const float keep = up ? x[i + o] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return x[0];
}

// Buffer coordinates (slab t, env n) of minibatch row `rw` (update kernels); false past the minibatch.
// A needs: M, small, slab_nenv, slab_env0, rows, row_begin, T.
template <class A>
__device__ __forceinline__ bool minibatch_row_to_tn(const A& a, int64_t rw, int64_t& t, int64_t& n) {
  if (rw >= a.M) return false;
  if (a.small) {
    if (a.slab_nenv > 0) {
      const uint32_t ne = (uint32_t)a.slab_nenv, rr = (uint32_t)rw;
      const uint32_t tt = rr / ne;
      t = tt, n = a.slab_env0 + (rr - tt * ne);
    } else {
      const uint32_t g = a.rows ? (uint32_t)a.rows[rw] : (uint32_t)(a.row_begin + rw);
      const uint32_t TT = (uint32_t)a.T, nn = g / TT;
      t = g - nn * TT, n = nn;
    }
    return true;
  }
  if (a.slab_nenv > 0) {
    t = rw / a.slab_nenv, n = a.slab_env0 + (rw - t * a.slab_nenv);
  } else {
    const int64_t g = a.rows ? a.rows[rw] : a.row_begin + rw;
    n = g / a.T, t = g - n * a.T;
  }
  return true;
}

// split_tc.cu
// one piece image: X(n, k) = W2[n][k] (transpose = 0) or W2[k][n] * kscale[k] (transpose = 1, kscale may be null)
struct PackJob {
  const float* w2;
  uint8_t* img;
  const float* kscale;
  float* scale_out;
  int transpose;
};
struct PackJobs {
  PackJob j[4];
};
// pieces = 3 / 2: bf16 pieces;  pieces = -2: two fp16 pieces of W2 * s, the power of two s written to *scale_out
int launch_pack_w2_jobs(const PackJobs& jobs, int njobs, int pieces, cudaStream_t st);  // one launch, njobs <= 4
int launch_pack_w2_pieces(const float* w2, uint8_t* img, int transpose, int pieces, cudaStream_t st,
                          const float* kscale = nullptr, float* scale_out = nullptr);
// *out_bits = max(*out_bits, bits of max |x[i]|)  (NaNs are skipped)
int launch_absmax_bits(const float* x, int64_t n, uint32_t* out_bits, cudaStream_t st);

}  // namespace rl8
