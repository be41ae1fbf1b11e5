// fp32-accurate tensor-core path (RL8_PREC_FP32_TC): split operands (two fp16 or three bf16 pieces per fp32 value),
// tcgen05.mma.cta_group::2 (see split_tc.cuh).
//
//   tc3_selftest_kernel   D[256][256] = A[256][K] * B[256][K]^T with a selectable set of piece products: pins the
//                         pair plumbing (cluster launch, pair TMEM allocation, N-split B operand, multicast commit)
//                         and measures what each term set costs in accuracy.
#include "split_common.cuh"

namespace rl8 {

using namespace tc;

constexpr int kXStages = kUF16 ? 6 : 4;  // ring depth of the forward kernel (32 KB / 48 KB stages)

// ---- pair selftest ---------------------------------------------------------------------------------------------
// One cluster of two CTAs, 256 threads each, synchronous K loop in chunks of 32: stage -> cluster barrier ->
// leader issues -> multicast commit -> both CTAs wait.
struct SmemSelf {
  uint8_t a[3][128 * 4 * 16];  // piece tiles [128 rows][4 column groups of 8 K]: off(r, c) = r*16 + c*2048
  uint8_t b[3][128 * 4 * 16];
  uint64_t bar;
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
tc3_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K,
                    int terms) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemSelf& s = *reinterpret_cast<SmemSelf*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&s.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(&s.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  uint8_t* const a_tiles[3] = {s.a[0], s.a[1], s.a[2]};
  uint8_t* const b_tiles[3] = {s.b[0], s.b[1], s.b[2]};
  const bool f16 = (terms & 64) != 0;  // two fp16 pieces per operand (the caller scales the operands into range)
  const uint32_t idesc = f16 ? instr_desc_f16(256, 256, 0, 0) : instr_desc(256, 256, 0, 0);
  uint32_t parity = 0;
  bool first = true;
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = tid & 127, cg = (tid >> 7) + 2 * i;
      const int64_t g = (int64_t)(128 * rank + row) * K + k0 + cg * 8;
      float v[8];
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(A + g);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(A + g + 4);
      if (f16) store_split_chunk<2, true>(a_tiles, (uint32_t)(row * 16 + cg * 2048), v);
      else store_split_chunk<3>(a_tiles, (uint32_t)(row * 16 + cg * 2048), v);
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(B + g);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(B + g + 4);
      if (f16) store_split_chunk<2, true>(b_tiles, (uint32_t)(row * 16 + cg * 2048), v);
      else store_split_chunk<3>(b_tiles, (uint32_t)(row * 16 + cg * 2048), v);
    }
    fence_async_smem();
    fence_before_sync();
    cluster_sync_all();
    if (rank == 0 && warp == 0 && elect_one()) {
      fence_after_sync();
      // (a piece, b piece) of term bit i
      const int ta[6] = {0, 0, 1, 1, 0, 2}, tb[6] = {0, 1, 0, 1, 2, 0};
      for (int ks = 0; ks < 2; ++ks) {
        for (int i = 5; i >= 0; --i) {
          if (!((terms >> i) & 1)) continue;
          const uint64_t ad = smem_desc(smem_u32(s.a[ta[i]]) + ks * 4096, 2048, 128);
          const uint64_t bd = smem_desc(smem_u32(s.b[tb[i]]) + ks * 4096, 2048, 128);
          mma_bf16_pair(tmem, ad, bd, idesc, first ? 0u : 1u);
          first = false;
        }
      }
      mma_commit_pair(&s.bar);
    }
    mbar_wait_cluster(&s.bar, parity);
    parity ^= 1u;
    fence_after_sync();
  }
  {
    const int q = warp & 3, half = warp >> 2;
    const int64_t row = 128 * rank + q * 32 + lane;
    for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) D[row * 256 + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 512);
}


// ---- pair MMA pace -----------------------------------------------------------------------------------------------
// Every pair issues `reps` ring stages' worth of piece products (terms per K step, two K steps per stage, operands at
// the ring's offsets; the data is whatever the shared memory holds) back to back with one commit at the end.
// out[pair] = {SM cycles, nanoseconds} around the issue + wait of the pair's leader.
struct SmemPace {
  StageX<3> ring[4];
  uint64_t bar;
  uint32_t tmem_base;
};
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
tc3_pace_kernel(long long* __restrict__ out, int reps, int terms, int n_cols, int mn_major) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemPace& s = *reinterpret_cast<SmemPace*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < (int)(sizeof(s.ring) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(s.ring)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(&s.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(&s.tmem_base, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  // mn_major: both operands as [K rows][128 M / N columns] tiles (the weight-gradient kernel's form: LBO 128, SBO 512,
  // 16 K rows = 256 bytes) instead of K-major [128 rows][K] tiles (LBO 2048, SBO 128, 16 K values = 4096 bytes)
  const uint32_t idesc = instr_desc(256, n_cols, mn_major, mn_major);
  const uint32_t lbo = mn_major ? 128u : 2048u, sbo = mn_major ? 512u : 128u, kstep = mn_major ? 256u : 4096u;
  long long c0 = 0, t0 = 0;
  if (rank == 0 && warp == 0) {
    c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
        const StageX<3>& stg = s.ring[r & 3];
        for (int ks = 0; ks < 2; ++ks)
          for (int i = 0; i < terms; ++i) {
            const uint64_t ad = smem_desc(smem_u32(stg.a[i % 3]) + ks * kstep, lbo, sbo);
            const uint64_t bd = smem_desc(smem_u32(stg.b[(i / 3) % 3]) + ks * kstep, lbo, sbo);
            mma_bf16_pair(tmem, ad, bd, idesc, (r | ks | i) ? 1u : 0u);
          }
      }
      mma_commit_pair(&s.bar);
    }
    __syncwarp();
  }
  mbar_wait_cluster(&s.bar, 0);
  fence_after_sync();
  if (rank == 0 && tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[2 * (blockIdx.x >> 1)] = clock64() - c0;
    out[2 * (blockIdx.x >> 1) + 1] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 512);
}


// ---- W2 piece images ------------------------------------------------------------------------------------------------
// The weight operand of a 256x256 contraction, split once per call into NP pieces (bf16, or fp16 pieces of the weights
// times a power of two derived from max |X|, written to *scale_out) and laid out so that the half
// a CTA of the pair needs for one ring stage (its 128 rows, the stage's four K groups, every piece) is one contiguous
// run of NP x 8 KB:
//     img[kc][half = n / 128][piece][ (n % 128) * 16 + g * 2048 + (k % 8) * 2 ]     with  k / 8 = 8 g + kc
// (the K order of StageX).  transpose = 0: X(n, k) = W2[n][k]  (forward Z2 = H1 W2^T: n = unit j, k = input i);
// transpose = 1: X(n, k) = W2[k][n]  (backward dH1^T = W2^T dZ2^T: n = input i, k = unit j).
// kscale (transpose = 1 only): X(n, k) = W2[k][n] * kscale[k] in fp32 -- the value network's backward operand with
// its one-row head folded in (update_x3.cu).
// jobs.j[blockIdx.y] is the image this block works on (up to four per launch: the update packs both networks' forward
// and transposed images at once)
template <int NP, bool F16>
__global__ void __launch_bounds__(256) pack_w2_pieces_kernel(PackJobs jobs) {
  const PackJob& job = jobs.j[blockIdx.y];
  const float* __restrict__ w2 = job.w2;
  uint8_t* __restrict__ img = job.img;
  const int transpose = job.transpose;
  const float* __restrict__ kscale = job.kscale;
  float* __restrict__ scale_out = job.scale_out;
  float s = 1.0f;
  if constexpr (F16) {
    // fp16 pieces: every block derives the same power-of-two scale from max |X| (64 K values, L2 hits)
    __shared__ float red[32];
    float m = 0.0f;
    for (int i = threadIdx.x; i < H * H / 4; i += blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(w2)[i];
      const float ks = kscale ? kscale[(i * 4) / H] : 1.0f;  // (kscale: X(n, k) = W2[k][n] kscale[k], k = the row of W2)
      m = fmaxf(m, fmaxf(fmaxf(fabsf(__fmul_rn(v.x, ks)), fabsf(__fmul_rn(v.y, ks))),
                         fmaxf(fabsf(__fmul_rn(v.z, ks)), fabsf(__fmul_rn(v.w, ks)))));
    }
    s = pow2_scale_for(block_max_nonneg(m, red));
    if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out) *scale_out = s;
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (n, 8 consecutive k)
  if (idx >= H * (H / 8)) return;
  const int n = idx & (H - 1), k8 = idx >> 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = transpose ? w2[(k8 * 8 + e) * H + n] : w2[n * H + k8 * 8 + e];
  if (kscale) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __fmul_rn(v[e], kscale[k8 * 8 + e]);
  }
  if constexpr (F16) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= s;
  }
  const int g = k8 >> 3, kc = k8 & 7, half = n >> 7, r = n & 127;
  uint8_t* base = img + (size_t)((kc * 2 + half) * NP) * kXPieceBytes;
  uint8_t* tiles[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) tiles[p] = base + (size_t)p * kXPieceBytes;
  store_split_chunk<NP, F16>(tiles, (uint32_t)(r * 16 + g * 2048), v);
}

int launch_pack_w2_jobs(const PackJobs& jobs, int njobs, int pieces, cudaStream_t st) {
  if (njobs < 1 || njobs > 4) return RL8_ERR_ARG;
  for (int i = 0; i < njobs; ++i) {
    if (jobs.j[i].kscale && !jobs.j[i].transpose) return RL8_ERR_ARG;
    if (pieces == -2 && !jobs.j[i].scale_out) return RL8_ERR_ARG;
  }
  const dim3 grid(H * (H / 8) / 256, njobs);
  if (pieces == 3) pack_w2_pieces_kernel<3, false><<<grid, 256, 0, st>>>(jobs);
  else if (pieces == 2) pack_w2_pieces_kernel<2, false><<<grid, 256, 0, st>>>(jobs);
  else if (pieces == -2) pack_w2_pieces_kernel<2, true><<<grid, 256, 0, st>>>(jobs);
  else return RL8_ERR_ARG;
  return check_launch("pack_w2_pieces");
}
int launch_pack_w2_pieces(const float* w2, uint8_t* img, int transpose, int pieces, cudaStream_t st,
                          const float* kscale, float* scale_out) {
  PackJobs jobs{};
  jobs.j[0] = PackJob{w2, img, kscale, scale_out, transpose};
  return launch_pack_w2_jobs(jobs, 1, pieces, st);
}

// max |x| as float bits (non-negative floats order like their bit patterns), NaNs skipped
__global__ void __launch_bounds__(256) absmax_bits_kernel(const float* __restrict__ x, int64_t n,
                                                          uint32_t* __restrict__ out_bits) {
  float m = 0.0f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int64_t n4 = n >> 2;
    for (int64_t i = i0; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (int64_t i = (n4 << 2) + i0; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
  } else {
    for (int64_t i = i0; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out_bits, __float_as_uint(m));
}
int launch_absmax_bits(const float* x, int64_t n, uint32_t* out_bits, cudaStream_t st) {
  if (n <= 0) return RL8_OK;
  int64_t blocks = ceil_div(n, (int64_t)256 * 16);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  absmax_bits_kernel<<<(int)blocks, 256, 0, st>>>(x, n, out_bits);
  return check_launch("absmax_bits");
}

// ---- forward kernel ---------------------------------------------------------------------------------------------------
// out[rows][P] of one network, fp32-accurate:  H1 = relu([obs] W1^T + b1) is computed on CUDA cores (K = D <= 7, exact
// fp32 FMAs) by the 16 worker warps straight into split A stages;  Z2 = H1 W2^T is the split pair GEMM (two fp16
// pieces per operand and 3 piece products, or -- X3_BF16 build -- three bf16 pieces and 6; B stages bulk-copied from
// the piece image);  H2 = relu(Z2 + b2) and the P-wide head are the epilogue.  fp16 pieces: the A operand is
// -s_h H1 with s_h from the bound h1_bound(max |obs| of the rows), the image carries s_w, and the epilogue folds
// 1 / (s_h s_w) into its bias add (all powers of two: the results are those of the un-scaled sums, bit for bit).
// A cluster of two CTAs owns 256-row tiles; accumulators are double-buffered in tensor memory (2 x 256 columns), so
// while the workers run the epilogue of tile j the tensor pipe drains the ring stages they produced for tile j + 1.
//   worker warps 0..15 : produce(tile j + 1) -> epilogue(tile j)
//   warp 16            : pair TMEM allocation; in the leader CTA one elected lane issues every tcgen05.mma
// Barriers (same offsets in both CTAs): full[s] (leader's, 32 worker-warp arrivals: A written, + 1 from the peer's warp
// 16: the peer's B half landed), bfull[s] (local, the bulk copy of the CTA's B half), empty[s] / acc_full[b] (multicast commits),
// acc_empty[b] (leader's, 32 worker-warp arrivals: accumulator b has been read).
struct SmemX3F {
  StageX<kUNP> ring[kXStages];  // 196608
  float w1t[8][H];              //   8192  -s_h [W1^T (D <= 7 rows) | b1 in row 7]
  float b2[H];                  //   1024  -b2
  float red[32];
  float w3[kMaxPT][H];          //   4096
  float part[4][TILE][kMaxPT];  //   8192  head partial sums per column quarter
  uint64_t full[kXStages], bfull[kXStages], empty[kXStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemX3F) <= 227 * 1024, "SmemX3F exceeds the 227 KB CTA limit");

template <int P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXThreads, 1)
tc3_forward_kernel(NetParams np, RowMap map, int64_t rows, float* __restrict__ out, int tanh_col1,
                   const float* __restrict__ w2_scale, const uint32_t* __restrict__ omax_bits) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemX3F& s = *reinterpret_cast<SmemX3F*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kXStages; ++i) {
      mbar_init(&s.full[i], 33);  // 32 worker warps of the pair + the peer's bulk copy (forwarded by its warp 16)
      mbar_init(&s.bfull[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.acc_full[0], 1), mbar_init(&s.acc_full[1], 1);
    mbar_init(&s.acc_empty[0], 32), mbar_init(&s.acc_empty[1], 32);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc_pair(&s.tmem_base, 512);
  float s_h = 1.0f, inv_scale = 1.0f;
  if constexpr (kUF16) {
    s_h = pow2_scale_for(h1_bound(np, __uint_as_float(*omax_bits), s.red));
    inv_scale = (1.0f / s_h) * (1.0f / *w2_scale);
  }
  stage_w1t(s.w1t, np, -s_h);                                           // h1_chunk<true>: A = -s_h H1
  for (int i = tid; i < H; i += blockDim.x) s.b2[i] = 0.0f - np.b2[i];  // -b2, never -0
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const int64_t ntiles = (rows + 255) / 256;
  const int64_t pr = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;

  if (warp < 16) {
    const int rloc = tid & 127, g = tid >> 7;
    const int64_t ds = map.dstride();
    const int D = np.D;
    uint32_t kcount = 0;
    auto produce = [&](int64_t tile) {
      float obf[7];
      const int64_t row = tile * 256 + rank * 128 + rloc;
#pragma unroll
      for (int d = 0; d < 7; ++d) obf[d] = 0.0f;
      if (row < rows) {
        const float* base = map.obs + map.offset(row);
#pragma unroll
        for (int d = 0; d < 7; ++d)
          if (d < D) obf[d] = __ldg(base + (int64_t)d * ds);
      }
      const ObsPairs ob = obs_pairs(obf);
      // two ring stages per generic -> async proxy fence; no worker warp waits for a bulk copy (the leader's MMA warp
      // waits for its own half of a W2 stage, the peer's warp 16 forwards the peer's) -- as in x3_update_f_kernel
      static_assert(kXStages % 2 == 0 && (H / kXKc) % 2 == 0, "stage pairs");
      for (int kc = 0; kc < H / kXKc; kc += 2, kcount += 2) {
        const int st = (int)(kcount % kXStages);
        const uint32_t use = kcount / kXStages;
        if (use > 0) {
          mbar_wait_cluster(&s.empty[st], (use - 1) & 1);
          mbar_wait_cluster(&s.empty[st + 1], (use - 1) & 1);
        }
        __syncwarp();
        if (warp == 0 && elect_one()) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint8_t* src = np.w2_img + (size_t)(((kc + h) * 2 + rank) * kUNP) * kXPieceBytes;
            mbar_expect_tx(&s.bfull[st + h], kUNP * kXPieceBytes);
#pragma unroll
            for (int p = 0; p < kUNP; ++p)
              bulk_g2s(s.ring[st + h].b[p], src + (size_t)p * kXPieceBytes, kXPieceBytes, &s.bfull[st + h]);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[8];
          h1_chunk<true>(s.w1t, ob, D, stage_kgroup(kc + h, g) * 8, v);  // -s_h H1
          uint8_t* tiles[kUNP];
#pragma unroll
          for (int p = 0; p < kUNP; ++p) tiles[p] = s.ring[st + h].a[p];
          store_split_chunk<kUNP, kUF16>(tiles, (uint32_t)(rloc * 16 + g * 2048), v);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(&s.full[st], 0);
          mbar_arrive_cluster(&s.full[st + 1], 0);
        }
      }
    };
    auto epilogue = [&](int64_t tile, int64_t j) {
      const int buf = (int)(j & 1);
      mbar_wait_cluster(&s.acc_full[buf], (uint32_t)((j >> 1) & 1));
      fence_after_sync();
      const int q = warp & 3, cq = warp >> 2, r = q * 32 + lane;
      const uint32_t acc = tmem + (uint32_t)(buf * H) + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
      float v0[32], v1[32];
      tmem_ld32_nowait(acc, v0);
      tmem_ld32_nowait(acc + 32, v1);
      tmem_wait_ld();
      reg_fence32(v0);
      reg_fence32(v1);
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&s.acc_empty[buf], 0);  // the accumulator may be overwritten
      float dot[P];
#pragma unroll
      for (int p = 0; p < P; ++p) dot[p] = 0.0f;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const int col0 = cq * 64 + c2 * 32;
        const float* v = c2 ? v1 : v0;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          // the accumulator holds -s Z2 and s.b2 holds -b2: fma(v, 1 / s, -b2) = -(Z2 + b2), rounded once like the
          // un-scaled sum; h = min(., 0) = -H2 enters the head sums negated
          const float4 b = *reinterpret_cast<const float4*>(&s.b2[col0 + jj]);
          const float h0 = fminf(fmaf(v[jj], inv_scale, b.x), 0.0f), h1 = fminf(fmaf(v[jj + 1], inv_scale, b.y), 0.0f);
          const float h2 = fminf(fmaf(v[jj + 2], inv_scale, b.z), 0.0f), h3 = fminf(fmaf(v[jj + 3], inv_scale, b.w), 0.0f);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const float4 w = *reinterpret_cast<const float4*>(&s.w3[p][col0 + jj]);
            dot[p] = fmaf(-h0, w.x, dot[p]);
            dot[p] = fmaf(-h1, w.y, dot[p]);
            dot[p] = fmaf(-h2, w.z, dot[p]);
            dot[p] = fmaf(-h3, w.w, dot[p]);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < P; ++p) s.part[cq][r][p] = dot[p];
      worker_bar_sync();
      if (tid < TILE) {
        const int64_t row = tile * 256 + rank * 128 + tid;
        if (row < rows) {
#pragma unroll
          for (int p = 0; p < P; ++p) {
            float v = ((s.part[0][tid][p] + s.part[1][tid][p]) + (s.part[2][tid][p] + s.part[3][tid][p])) + np.b3[p];
            if (tanh_col1 && p == 1) v = tanhf(v);
            out[row * P + p] = v;
          }
        }
      }
      worker_bar_sync();  // part[] may be rewritten
    };
    if (n_my > 0) produce(pr);
    for (int64_t j = 0; j < n_my; ++j) {
      if (j + 1 < n_my) produce(pr + (j + 1) * npairs);
      epilogue(pr + j * npairs, j);
    }
  } else if (rank == 0) {
    // MMA issue: the whole warp follows the barriers, one elected lane issues
    const uint32_t idesc = upd_idesc(0, 0);
    uint32_t kcount = 0;
    for (int64_t j = 0; j < n_my; ++j) {
      const int buf = (int)(j & 1);
      if (j >= 2) mbar_wait_cluster(&s.acc_empty[buf], (uint32_t)(((j >> 1) - 1) & 1));
      for (int kc = 0; kc < H / kXKc; ++kc, ++kcount) {
        const int st = (int)(kcount % kXStages);
        mbar_wait_cluster(&s.full[st], (kcount / kXStages) & 1);
        mbar_wait(&s.bfull[st], (kcount / kXStages) & 1);  // this CTA's half of the W2 stage has landed
        fence_after_sync();
        if (elect_one()) {
          issue_stage<kUNP>(tmem + (uint32_t)(buf * H), s.ring[st], idesc, kc > 0);
          mma_commit_pair(&s.empty[st]);
          if (kc == H / kXKc - 1) mma_commit_pair(&s.acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // peer CTA, warp 16: tells the leader's full[] barrier when THIS CTA's half of a W2 stage has landed
    const uint32_t total = (uint32_t)(n_my * (H / kXKc));
    for (uint32_t kcount = 0; kcount < total; ++kcount) {
      const int st = (int)(kcount % kXStages);
      mbar_wait(&s.bfull[st], (kcount / kXStages) & 1);
      if ((tid & 31) == 0) mbar_arrive_cluster(&s.full[st], 0);
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its tensor memory until here
  if (warp == 16) tmem_dealloc_pair(tmem, 512);
}

// max |obs| over the rows of a map (NaNs skipped) -> *out_bits (atomicMax of the float's bit pattern)
__global__ void __launch_bounds__(256) absmax_rows_kernel(RowMap map, int64_t rows, int D, uint32_t* __restrict__ out_bits) {
  float m = 0.0f;
  const int64_t ds = map.dstride();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float* base = map.obs + map.offset(r);
    for (int d = 0; d < D; ++d) m = fmaxf(m, fabsf(__ldg(base + (int64_t)d * ds)));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out_bits, __float_as_uint(m));
}
static int launch_absmax_rows(const RowMap& map, int64_t rows, int D, uint32_t* out_bits, cudaStream_t st) {
  if (rows <= 0) return RL8_OK;
  int64_t blocks = ceil_div(rows, (int64_t)256 * 4);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  absmax_rows_kernel<<<(int)blocks, 256, 0, st>>>(map, rows, D, out_bits);
  return check_launch("absmax_rows");
}

// w2_scale / omax_bits (device): the scale of the W2 piece image and the bits of max |obs| over the rows (fp16 pieces)
static int launch_forward_x3(const NetParams& np, const RowMap& map, int64_t rows, float* out, int tanh_col1,
                             const float* w2_scale, const uint32_t* omax_bits, cudaStream_t st) {
  const int64_t ntiles = ceil_div(rows, 256);
  const int grid = 2 * (int)(ntiles < kNumSMs / 2 ? ntiles : kNumSMs / 2);
  int rc;
#define RL8_FWDX(PV)                                                                                   \
  case PV:                                                                                             \
    if ((rc = set_smem((const void*)tc3_forward_kernel<PV>, sizeof(SmemX3F)))) return rc;               \
    tc3_forward_kernel<PV><<<grid, kXThreads, sizeof(SmemX3F), st>>>(np, map, rows, out, tanh_col1,     \
                                                                     w2_scale, omax_bits);             \
    break;
  switch (np.P) {
    RL8_FWDX(1) RL8_FWDX(2) RL8_FWDX(3) RL8_FWDX(4)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_FWDX
  return check_launch("tc3_forward");
}

// workspace: the piece image, then 256 bytes of operand magnitudes {float w2 scale; uint32 max |obs| bits}
int64_t forward_x3_workspace() { return kXImgBytes + 256; }

static int zero_words(void* p, size_t bytes, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, st);
  if (e != cudaSuccess) {
    set_last_error("cudaMemsetAsync", e);
    return RL8_ERR_CUDA;
  }
  return RL8_OK;
}

int mlp_forward_x3(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out, int tanh_col1,
                   void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (m->H != H || m->D > 7 || m->P > kMaxPT) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < forward_x3_workspace()) return RL8_ERR_WORKSPACE;
  uint8_t* img = (uint8_t*)workspace;
  float* w2_scale = (float*)(img + kXImgBytes);
  uint32_t* omax = (uint32_t*)(w2_scale + 1);
  int rc;
  if (kUF16) {
    if ((rc = zero_words(w2_scale, 8, st))) return rc;
    if ((rc = launch_absmax_rows(map, rows, m->D, omax, st))) return rc;
  }
  if ((rc = launch_pack_w2_pieces(which ? m->vf_w2 : m->pi_w2, img, 0, kUF16 ? -2 : 3, st, nullptr, w2_scale))) return rc;
  return launch_forward_x3(net_params(m, which, img), map, rows, out, tanh_col1, w2_scale, omax, st);
}

// collect() in RL8_PREC_FP32_TC: per step the split policy forward + the fused sample / env-step / buffer-write kernel
// of the fp32 path (collect.cu), then one split value pass over all T + 1 observation slabs.
int collect_tail(const rl8_rollout* ro, int t, const float* feat, cudaStream_t st, uint32_t* omax_next = nullptr,
                 uint32_t* omax_all = nullptr);

// workspace: two piece images, the operand magnitudes {w2 scale pi, w2 scale vf, max |obs| bits of all slabs, of slab
// 0 .. T}, the head outputs of a step
static int64_t collect_x3_scale_bytes(int32_t T) { return round_up((int64_t)(T + 4) * 4, 256); }
int64_t collect_x3_workspace(const rl8_model*, int64_t N, int32_t T) {
  return 2 * (int64_t)kXImgBytes + collect_x3_scale_bytes(T) + N * kMaxPT * 4;
}

int collect_x3(const rl8_model* model, const rl8_rollout* ro, void* workspace, int64_t workspace_bytes,
               cudaStream_t st) {
  if (model->H != H || model->P > kMaxPT || model->D > 7) return RL8_ERR_UNSUPPORTED;
  const int64_t N = ro->N;
  if (!workspace || workspace_bytes < collect_x3_workspace(model, N, ro->T)) return RL8_ERR_WORKSPACE;
  uint8_t* img_pi = (uint8_t*)workspace;
  uint8_t* img_vf = img_pi + kXImgBytes;
  float* w2_scale = (float*)(img_vf + kXImgBytes);           // [2]
  uint32_t* omax_all = (uint32_t*)(w2_scale + 2);            // all T + 1 slabs (value pass)
  uint32_t* omax_t = omax_all + 1;                           // [T + 1]: slab t (policy forward of step t)
  float* feat = (float*)((uint8_t*)w2_scale + collect_x3_scale_bytes(ro->T));
  int rc;
  if (kUF16 && (rc = zero_words(w2_scale, (size_t)collect_x3_scale_bytes(ro->T), st))) return rc;
  {
    PackJobs jobs{};
    jobs.j[0] = PackJob{model->pi_w2, img_pi, nullptr, w2_scale, 0};
    jobs.j[1] = PackJob{model->vf_w2, img_vf, nullptr, w2_scale + 1, 0};
    if ((rc = launch_pack_w2_jobs(jobs, 2, kUF16 ? -2 : 3, st))) return rc;
  }
  const NetParams np_pi = net_params(model, 0, img_pi), np_vf = net_params(model, 1, img_vf);
  const int continuous = ro->dist_kind != RL8_DIST_CATEGORICAL;
  RowMap map{};
  map.mode = 0, map.stride_r = 1, map.stride_d = N, map.D = model->D;
  // max |obs| per slab: slab 0 (written by the reset / the previous rollout) by its own kernel, slab t + 1 by the tail
  // kernel of step t that writes it; omax_all collects the maximum over all T + 1 slabs for the value pass
  if (kUF16) {
    if ((rc = launch_absmax_bits(ro->obs, (int64_t)model->D * N, omax_t, st))) return rc;
    if ((rc = launch_absmax_bits(ro->obs, (int64_t)model->D * N, omax_all, st))) return rc;
  }
  for (int t = 0; t < ro->T; ++t) {
    map.obs = ro->obs + (int64_t)t * model->D * N;
    if ((rc = launch_forward_x3(np_pi, map, N, feat, continuous, w2_scale, omax_t + t, st))) return rc;
    if ((rc = collect_tail(ro, t, feat, st, kUF16 ? omax_t + t + 1 : nullptr, kUF16 ? omax_all : nullptr))) return rc;
  }
  RowMap vmap{};
  vmap.obs = ro->obs, vmap.mode = 2, vmap.D = model->D, vmap.N = N, vmap.T = ro->T;
  return launch_forward_x3(np_vf, vmap, (int64_t)(ro->T + 1) * N, ro->values, 0, w2_scale + 1, omax_all, st);
}

}  // namespace rl8

using namespace rl8;

// Test hook: the pair MMA with split operands (tests/test_gpu_split.py).  terms: bit 0 a0b0, 1 a0b1, 2 a1b0,
// 3 a1b1, 4 a0b2, 5 a2b0; bit 6: the pieces are fp16 (two per operand, bits 0..3 only).
extern "C" int rl8_tc3_selftest(const float* A, const float* B, float* D, int32_t K, int32_t terms,
                                rl8_stream_t stream) {
  if (!A || !B || !D || K < 32 || (K % 32) || terms < 1 || terms > 127 || (terms & 63) == 0 ||
      ((terms & 64) && (terms & 48)))
    return RL8_ERR_ARG;
  cudaError_t e = cudaFuncSetAttribute(tc3_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SmemSelf));
  if (e != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute", e);
    return RL8_ERR_CUDA;
  }
  tc3_selftest_kernel<<<2, 256, sizeof(SmemSelf), (cudaStream_t)stream>>>(A, B, D, K, terms);
  return check_launch("tc3_selftest");
}

// Development hook: pace of back-to-back pair MMAs on `pairs` clusters at once (tools/bench_pair_mma.py).
// terms + 16: both operands MN-major (the weight-gradient kernel's tiles).
extern "C" int rl8_tc3_bench_pace(long long* out, int32_t pairs, int32_t reps, int32_t terms, int32_t n_cols,
                                  rl8_stream_t stream) {
  const int mn_major = terms >= 16 ? 1 : 0;
  terms -= 16 * mn_major;
  if (!out || pairs < 1 || pairs > kNumSMs / 2 || reps < 1 || terms < 1 || terms > 9 ||
      (n_cols != 64 && n_cols != 128 && n_cols != 256))
    return RL8_ERR_ARG;
  cudaError_t e = cudaFuncSetAttribute(tc3_pace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SmemPace));
  if (e != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute", e);
    return RL8_ERR_CUDA;
  }
  tc3_pace_kernel<<<2 * pairs, 128, sizeof(SmemPace), (cudaStream_t)stream>>>(out, reps, terms, n_cols, mn_major);
  return check_launch("tc3_pace");
}
