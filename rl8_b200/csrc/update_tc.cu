// PPO update on tensor cores (RL8_PREC_BF16): forward + clipped losses + hand-derived backward
// of both default networks for one minibatch, as two persistent kernels.
//
//   tc_update_h_kernel  ("activation" kernel; CTAs alternate between the policy and the value
//     network, each with ITS W2 resident in smem).  Per 128-row tile:
//       H1 (CUDA cores) -> MMA1  Z2 = H1 * W2^T            (K-major A, K-major B)
//       heads + per-row PPO loss -> dOut                    (CUDA cores, fp32)
//       H2 tile          -> MMA-G3  gW3^T += H2^T * dOut    (MN-major A, MN-major B, N = 8)
//       dZ2 tile         -> MMA2  dH1 = dZ2 * W2            (K-major A, MN-major B: the SAME
//                                                            W2 image read transposed)
//                           MMA-Gb2 [.,gb2] += dZ2^T * [obs,1]
//       dZ1 tile         -> MMA-G1  [gW1,gb1] += dZ1^T * [obs,1]
//     The thin gradients (gW1, gb1, gb2, gW3) accumulate in TMEM over all tiles of the CTA and
//     are flushed once.  The ReLU mask of layer 2 (256 bits / row) and dOut go to a small
//     scratch (80 B / row / network) for the weight-gradient kernel.
//   tc_update_w_kernel  ("weight" kernel).  gW2 needs a 256x256 fp32 accumulator = ALL 512 TMEM
//     columns of an SM, so it cannot share an SM with the kernel above.  Per 128-row tile it
//     recomputes H1 (K = D <= 8, cheap), rebuilds dZ2 from dOut and the mask bits, and issues
//     gW2 += dZ2^T * H1 (both operands MN-major); accumulators stay in TMEM for the whole
//     kernel and are flushed with one red.global.add per element.
//
// GEMM work per row and network: 3 x 256 x 256 MACs -- the minimum (forward, dH1, gW2); nothing
// is recomputed on the tensor cores and no activation tile ever reaches HBM.
#include "mlp_tc.cuh"
#include "ppo_loss_math.cuh"

namespace rl8 {

using namespace tc;

struct UpdArgs {
  const float* obs;      // [T+1][D][N]
  const void* actions;   // [T+1][N]
  const float* logp;     // [T+1][N]
  const float* adv;      // [T+1][N]
  const float* ret;      // [T+1][N]
  const int64_t* rows;   // minibatch row indices (n*T + t) or null
  int64_t row_begin, M, N;
  int T, dist_kind;
  rl8_ppo_hparams hp;
  float inv_denom;
  uint32_t* mask[2];     // [M][8]   ReLU mask of layer 2, bit (j % 32) of word j / 32
  float* dout[2];        // [M][4]   d(loss)/d(head outputs)
  float *gw1[2], *gb1[2], *gw2[2], *gb2[2], *gw3[2], *gb3[2];
  double* sums;          // [5]
};

struct SmemH {
  uint8_t w2[kW2Bytes];
  uint8_t a_tile[kTileBytes];
  float w1t[8][H];
  float b1[H], b2[H];
  float w3[kMaxPT][H];
  float obs[8][TILE];
  float part[2][TILE][kMaxPT];
  float dout[TILE][kMaxPT];
  uint8_t dout_tile[TILE * 16];  // bf16 [r][8]: MN-major B operand (N = 8)
  uint8_t aug_tile[TILE * 16];   // bf16 [r][8] = [obs_0..obs_{D-1}, 1, 0..]
  int64_t row_idx[TILE];         // t*N + n of the tile's rows (-1: past the minibatch)
  float gb3[kMaxPT];
  uint64_t bar_w, bar_mma[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemH) + 1024 <= 227 * 1024, "SmemH exceeds the 227 KB CTA limit");

// TMEM columns of kernel H
constexpr uint32_t kColMain = 0, kColG3 = 256, kColGb2 = 320, kColG1 = 384;  // +32 per 128-block

// Stage the tile's row indices and observations (gathered through the minibatch row list).
template <class S>
__device__ __forceinline__ void stage_rows(S& s, const UpdArgs& a, int64_t tile, int D) {
  const int tid = threadIdx.x;
  if (tid < TILE) {
    const int64_t rw = tile * TILE + tid;
    int64_t idx = -1;
    if (rw < a.M) {
      const int64_t g = a.rows ? a.rows[rw] : a.row_begin + rw;
      const int64_t n = g / a.T, t = g - n * a.T;
      idx = t * a.N + n;
    }
    s.row_idx[tid] = idx;
  }
  __syncthreads();
  for (int i = tid; i < 8 * TILE; i += blockDim.x) {
    const int d = i / TILE, rr = i - d * TILE;
    const int64_t idx = s.row_idx[rr];
    float v = 0.0f;
    if (d < D && idx >= 0) {
      const int64_t t = idx / a.N, n = idx - t * a.N;
      v = a.obs[(t * D + d) * a.N + n];
    }
    s.obs[d][rr] = v;
  }
  __syncthreads();
}

// dZ2 chunks of this thread's 128 columns from the layer-2 ReLU mask words and dOut:
// dz2[r][j] = mask ? sum_p dout[r][p] * w3[p][j] : 0  -> bf16 chunks in `tile`.
template <int PN, class S>
__device__ __forceinline__ void dz2_to_tile(S& s, uint8_t* tile, const uint32_t* mask_words, int r,
                                            int half) {
  float dr[PN];
#pragma unroll
  for (int p = 0; p < PN; ++p) dr[p] = s.dout[r][p];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t word = mask_words[w];
    const int col0 = half * 128 + w * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int p = 0; p < PN; ++p) {
        const float4 wa = *reinterpret_cast<const float4*>(&s.w3[p][col0 + 8 * k]);
        const float4 wb = *reinterpret_cast<const float4*>(&s.w3[p][col0 + 8 * k + 4]);
        v[0] = fmaf(dr[p], wa.x, v[0]), v[1] = fmaf(dr[p], wa.y, v[1]);
        v[2] = fmaf(dr[p], wa.z, v[2]), v[3] = fmaf(dr[p], wa.w, v[3]);
        v[4] = fmaf(dr[p], wb.x, v[4]), v[5] = fmaf(dr[p], wb.y, v[5]);
        v[6] = fmaf(dr[p], wb.z, v[6]), v[7] = fmaf(dr[p], wb.w, v[7]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = ((word >> (8 * k + e)) & 1u) ? v[e] : 0.0f;
      store_chunk(tile, chunk_offset<TILE>(r, col0 / 8 + k), v);
    }
  }
}

template <int PN, bool POLICY>
__device__ __forceinline__ void update_h_body(SmemH& s, const NetParams& np, const UpdArgs& a,
                                              int net) {
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x;
  const int r = tid & (TILE - 1), half = tid >> 7, q = (tid >> 5) & 3;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const int D = np.D;
  const int64_t ntiles = (a.M + TILE - 1) / TILE;
  const int nctas = gridDim.x >> 1;
  float b3[PN];
#pragma unroll
  for (int p = 0; p < PN; ++p) b3[p] = np.b3[p];
  if (tid < kMaxPT) s.gb3[tid] = 0.0f;
  double s_ent = 0, s_pol = 0, s_vf = 0, s_kl = 0;
  uint32_t ph0 = 0, ph1 = 0;
  int it = 0;
  const bool continuous = POLICY && a.dist_kind != RL8_DIST_CATEGORICAL;

  for (int64_t tile = blockIdx.x >> 1; tile < ntiles; tile += nctas, ++it) {
    const int64_t row = tile * TILE + r;
    const bool valid = row < a.M;
    // ---- 1. rows, observations, [obs, 1] operand tile ------------------------------------------
    stage_rows(s, a, tile, D);
    if (tid < TILE) {
      float v[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) v[d] = d < D ? s.obs[d][tid] : (d == D ? 1.0f : 0.0f);
      store_chunk(s.aug_tile, (uint32_t)tid * 16u, v);
    }
    // ---- 2. H1 (keeps the layer-1 ReLU mask in registers) -> MMA1 ----------------------------------
    uint32_t mask1[4], mask2[4];
    layer1_to_tile(s, D, mask1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + kColMain, smem_u32(s.a_tile), TILE, false, smem_u32(s.w2), H, false, TILE, H,
                 H, false);
      mma_commit(&s.bar_mma[0]);
    }
    mbar_wait(&s.bar_mma[0], ph0);
    ph0 ^= 1;
    fence_after_sync();
    // ---- 3. one pass over Z2: head partial sums, H2 tile (bf16), layer-2 mask ----------------------
    {
      float dot[PN];
#pragma unroll
      for (int p = 0; p < PN; ++p) dot[p] = 0.0f;
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {
        const int col0 = half * 128 + c4 * 32;
        float v[32];
        tmem_ld32(tmem + kColMain + lane_base + (uint32_t)col0, v);
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&s.b2[col0 + j]);
          v[j] = fmaxf(v[j] + b.x, 0.0f), v[j + 1] = fmaxf(v[j + 1] + b.y, 0.0f);
          v[j + 2] = fmaxf(v[j + 2] + b.z, 0.0f), v[j + 3] = fmaxf(v[j + 3] + b.w, 0.0f);
#pragma unroll
          for (int p = 0; p < PN; ++p) {
            const float4 w = *reinterpret_cast<const float4*>(&s.w3[p][col0 + j]);
            dot[p] = fmaf(v[j], w.x, dot[p]);
            dot[p] = fmaf(v[j + 1], w.y, dot[p]);
            dot[p] = fmaf(v[j + 2], w.z, dot[p]);
            dot[p] = fmaf(v[j + 3], w.w, dot[p]);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) word |= (v[j + e] > 0.0f ? 1u : 0u) << (j + e);
        }
        mask2[c4] = word;
#pragma unroll
        for (int k = 0; k < 4; ++k) store_chunk(s.a_tile, chunk_offset<TILE>(r, col0 / 8 + k), v + 8 * k);
      }
#pragma unroll
      for (int p = 0; p < PN; ++p) s.part[half][r][p] = dot[p];
      if (valid)
        *reinterpret_cast<uint4*>(a.mask[net] + row * 8 + half * 4) =
            make_uint4(mask2[0], mask2[1], mask2[2], mask2[3]);
    }
    __syncthreads();
    // ---- 4. per-row loss -> dOut ---------------------------------------------------------------------------
    if (tid < TILE) {
      float o[PN], d_o[PN];
#pragma unroll
      for (int p = 0; p < PN; ++p) o[p] = s.part[0][tid][p] + s.part[1][tid][p] + b3[p];
#pragma unroll
      for (int p = 0; p < PN; ++p) d_o[p] = 0.0f;
      if (valid) {
        const int64_t idx = s.row_idx[tid];
        RowLoss L;
        if constexpr (POLICY) {
          if (continuous) o[1] = tanhf(o[1]);
          const float act = a.dist_kind == RL8_DIST_CATEGORICAL
                                ? (float)((const long long*)a.actions)[idx]
                                : ((const float*)a.actions)[idx];
          ppo_policy_row<PN>(a.dist_kind, o, act, a.logp[idx], a.adv[idx], a.hp, a.inv_denom, d_o, L);
          s_ent += L.entropy, s_pol += L.policy, s_kl += L.kl;
        } else {
          ppo_value_row(o[0], a.ret[idx], a.hp, a.inv_denom, d_o, L);
          s_vf += L.vf;
        }
      }
      float v8[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) v8[p] = p < PN ? d_o[p] : 0.0f;
      store_chunk(s.dout_tile, (uint32_t)tid * 16u, v8);
      *reinterpret_cast<float4*>(&s.dout[tid][0]) = make_float4(v8[0], v8[1], v8[2], v8[3]);
      if (valid) *reinterpret_cast<float4*>(a.dout[net] + row * 4) = make_float4(v8[0], v8[1], v8[2], v8[3]);
      // gb3 += sum_r dOut: one shared atomic per warp
#pragma unroll
      for (int p = 0; p < PN; ++p) {
        const float w = warp_sum(d_o[p]);
        if ((tid & 31) == 0) atomicAdd(&s.gb3[p], w);
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- 5. gW3^T += H2^T * dOut ----------------------------------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
        issue_gemm(tmem + kColG3 + 32 * jb, smem_u32(s.a_tile) + jb * 32768, TILE, true,
                   smem_u32(s.dout_tile), TILE, true, TILE, 8, TILE, it > 0);
      mma_commit(&s.bar_mma[1]);
    }
    mbar_wait(&s.bar_mma[1], ph1);
    ph1 ^= 1;
    fence_after_sync();
    // ---- 6. dZ2 tile (from the mask words; no TMEM traffic) -----------------------------------------------------
    dz2_to_tile<PN>(s, s.a_tile, mask2, r, half);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- 7. dH1 = dZ2 * W2 ; [., gb2] += dZ2^T * [obs, 1] -----------------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
      issue_gemm(tmem + kColMain, smem_u32(s.a_tile), TILE, false, smem_u32(s.w2), H, true, TILE, H, H,
                 false);
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
        issue_gemm(tmem + kColGb2 + 32 * jb, smem_u32(s.a_tile) + jb * 32768, TILE, true,
                   smem_u32(s.aug_tile), TILE, true, TILE, 8, TILE, it > 0);
      mma_commit(&s.bar_mma[0]);
    }
    mbar_wait(&s.bar_mma[0], ph0);
    ph0 ^= 1;
    fence_after_sync();
    // ---- 8. dZ1 tile = dH1 masked by the layer-1 ReLU mask ------------------------------------------------------------
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {
      const int col0 = half * 128 + c4 * 32;
      float v[32];
      tmem_ld32(tmem + kColMain + lane_base + (uint32_t)col0, v);
      uint32_t word = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) word = (c4 == w) ? mask1[w] : word;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = ((word >> j) & 1u) ? v[j] : 0.0f;
#pragma unroll
      for (int k = 0; k < 4; ++k) store_chunk(s.a_tile, chunk_offset<TILE>(r, col0 / 8 + k), v + 8 * k);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- 9. [gW1, gb1] += dZ1^T * [obs, 1] --------------------------------------------------------------------------------
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int ib = 0; ib < 2; ++ib)
        issue_gemm(tmem + kColG1 + 32 * ib, smem_u32(s.a_tile) + ib * 32768, TILE, true,
                   smem_u32(s.aug_tile), TILE, true, TILE, 8, TILE, it > 0);
      mma_commit(&s.bar_mma[1]);
    }
    mbar_wait(&s.bar_mma[1], ph1);
    ph1 ^= 1;
    fence_after_sync();
  }

  // ---- flush the thin gradients (threads: warps 0-3 -> block 0, warps 4-7 -> block 1) ------------------------
  if (it > 0) {
    const int blk = half;
    const int c = blk * 128 + q * 32 + (tid & 31);  // hidden unit of this thread
    float v[8];
    tmem_ld8(tmem + kColG3 + 32 * blk + lane_base, v);
#pragma unroll
    for (int p = 0; p < PN; ++p) atomicAdd(a.gw3[net] + p * H + c, v[p]);
    tmem_ld8(tmem + kColGb2 + 32 * blk + lane_base, v);
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d == D) atomicAdd(a.gb2[net] + c, v[d]);
    tmem_ld8(tmem + kColG1 + 32 * blk + lane_base, v);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (d < D) atomicAdd(a.gw1[net] + c * D + d, v[d]);
      if (d == D) atomicAdd(a.gb1[net] + c, v[d]);
    }
    __syncthreads();
    if (tid < PN) atomicAdd(a.gb3[net] + tid, s.gb3[tid]);
  }
  __shared__ double red[32];
  double sv[4] = {s_ent, s_pol, s_vf, s_kl};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double t = block_sum(sv[i], red);
    if (tid == 0 && t != 0.0) atomicAdd(a.sums + i, t);
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(a.sums + 4, (double)a.M);
}

template <int P>
__global__ void __launch_bounds__(256, 1)
tc_update_h_kernel(NetParams np_pi, NetParams np_vf, UpdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemH& s = *reinterpret_cast<SmemH*>(smem_raw);
  const int net = blockIdx.x & 1;
  const NetParams np = net ? np_vf : np_pi;
  cta_setup(s, np, 512);
  if (net == 0) update_h_body<P, true>(s, np, a, 0);
  else update_h_body<1, false>(s, np, a, 1);
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(s.tmem_base, 512);
}

// ---- weight-gradient kernel -----------------------------------------------------------------------------------
struct SmemW {
  uint8_t a_tile[kTileBytes];   // H1 tile   (MN-major B: rows = K = r, cols = i)
  uint8_t dz_tile[kTileBytes];  // dZ2 tile  (MN-major A: rows = K = r, cols = j)
  float w1t[8][H];
  float b1[H];
  float w3[kMaxPT][H];
  float obs[8][TILE];
  float dout[TILE][kMaxPT];
  int64_t row_idx[TILE];
  uint64_t bar_mma;
  uint32_t tmem_base;
};
static_assert(sizeof(SmemW) <= 227 * 1024, "SmemW exceeds the 227 KB CTA limit");

template <int PN>
__device__ __forceinline__ void update_w_body(SmemW& s, const NetParams& np, const UpdArgs& a, int net) {
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x;
  const int r = tid & (TILE - 1), half = tid >> 7, q = (tid >> 5) & 3;
  const int D = np.D;
  const int64_t ntiles = (a.M + TILE - 1) / TILE;
  const int nctas = gridDim.x >> 1;
  uint32_t ph = 0;
  int it = 0;
  for (int64_t tile = blockIdx.x >> 1; tile < ntiles; tile += nctas, ++it) {
    const int64_t row = tile * TILE + r;
    // mask words and dOut of this thread's row straight from the scratch (coalesced 16 B each)
    uint4 mw = make_uint4(0u, 0u, 0u, 0u);
    if (row < a.M) mw = *reinterpret_cast<const uint4*>(a.mask[net] + row * 8 + half * 4);
    if (tid < TILE) {
      float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < a.M) d4 = *reinterpret_cast<const float4*>(a.dout[net] + row * 4);
      *reinterpret_cast<float4*>(&s.dout[tid][0]) = d4;
    }
    stage_rows(s, a, tile, D);  // ends with __syncthreads: s.dout visible too
    layer1_to_tile(s, D);
    const uint32_t mask2[4] = {mw.x, mw.y, mw.z, mw.w};
    dz2_to_tile<PN>(s, s.dz_tile, mask2, r, half);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
        issue_gemm(tmem + 256 * jb, smem_u32(s.dz_tile) + jb * 32768, TILE, true, smem_u32(s.a_tile),
                   TILE, true, TILE, H, TILE, it > 0);
      mma_commit(&s.bar_mma);
    }
    mbar_wait(&s.bar_mma, ph);
    ph ^= 1;
    fence_after_sync();
  }
  if (it > 0) {
    const int jb = half;
    const int j = jb * 128 + q * 32 + (tid & 31);
    float* dst = a.gw2[net] + (int64_t)j * H;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      float v[32];
      tmem_ld32(tmem + 256 * jb + ((uint32_t)(q * 32) << 16) + 32 * c, v);
#pragma unroll
      for (int e = 0; e < 32; ++e) atomicAdd(dst + 32 * c + e, v[e]);
    }
  }
}

template <int P>
__global__ void __launch_bounds__(256, 1)
tc_update_w_kernel(NetParams np_pi, NetParams np_vf, UpdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemW& s = *reinterpret_cast<SmemW*>(smem_raw);
  const int net = blockIdx.x & 1;
  const NetParams np = net ? np_vf : np_pi;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s.bar_mma, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, 512);
  for (int i = tid; i < 8 * H; i += blockDim.x) {
    const int d = i / H, c = i - d * H;
    s.w1t[d][c] = d < np.D ? np.w1[c * np.D + d] : 0.0f;
  }
  for (int i = tid; i < H; i += blockDim.x) s.b1[i] = np.b1[i];
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (net == 0) update_w_body<P>(s, np, a, 0);
  else update_w_body<1>(s, np, a, 1);
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(s.tmem_base, 512);
}

// ---- host -----------------------------------------------------------------------------------------------------
int64_t ppo_tc_workspace(const rl8_model*, int64_t max_rows) {
  return 2 * (int64_t)kW2Bytes + 2 * max_rows * (8 * 4 + 4 * 4) + 256;
}

int ppo_minibatch_tc(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch,
                     const int64_t* rows, int64_t row_begin, int64_t M, double mean_denominator,
                     const rl8_ppo_hparams* hp, double* loss_sums, void* workspace,
                     int64_t workspace_bytes, cudaStream_t st) {
  if (model->H != H || model->P > kMaxPT || model->D > 7) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < ppo_tc_workspace(model, M)) return RL8_ERR_WORKSPACE;
  uint8_t* img_pi = (uint8_t*)workspace;
  uint8_t* img_vf = img_pi + kW2Bytes;
  uint32_t* mask0 = (uint32_t*)(img_vf + kW2Bytes);
  uint32_t* mask1 = mask0 + M * 8;
  float* dout0 = (float*)(mask1 + M * 8);
  float* dout1 = dout0 + M * 4;
  int rc;
  if ((rc = launch_pack_w2(model->pi_w2, img_pi, st))) return rc;
  if ((rc = launch_pack_w2(model->vf_w2, img_vf, st))) return rc;
  const NetParams np_pi = net_params(model, 0, img_pi), np_vf = net_params(model, 1, img_vf);
  UpdArgs a;
  a.obs = batch->obs, a.actions = batch->actions, a.logp = batch->logp;
  a.adv = batch->advantages, a.ret = batch->returns;
  a.rows = rows, a.row_begin = row_begin, a.M = M, a.N = batch->N, a.T = batch->T;
  a.dist_kind = batch->dist_kind, a.hp = *hp;
  a.inv_denom = (float)((double)hp->loss_scale / mean_denominator);
  a.mask[0] = mask0, a.mask[1] = mask1, a.dout[0] = dout0, a.dout[1] = dout1;
  a.gw1[0] = (float*)grads->pi_w1, a.gb1[0] = (float*)grads->pi_b1, a.gw2[0] = (float*)grads->pi_w2;
  a.gb2[0] = (float*)grads->pi_b2, a.gw3[0] = (float*)grads->pi_w3, a.gb3[0] = (float*)grads->pi_b3;
  a.gw1[1] = (float*)grads->vf_w1, a.gb1[1] = (float*)grads->vf_b1, a.gw2[1] = (float*)grads->vf_w2;
  a.gb2[1] = (float*)grads->vf_b2, a.gw3[1] = (float*)grads->vf_w3, a.gb3[1] = (float*)grads->vf_b3;
  a.sums = loss_sums;
  const int64_t ntiles = ceil_div(M, TILE);
  int grid = (int)(2 * ntiles < kNumSMs ? 2 * ntiles : kNumSMs);
  grid &= ~1;
#define RL8_UPD(PV)                                                                            \
  case PV:                                                                                     \
    if ((rc = set_smem((const void*)tc_update_h_kernel<PV>, sizeof(SmemH)))) return rc;         \
    tc_update_h_kernel<PV><<<grid, 256, sizeof(SmemH), st>>>(np_pi, np_vf, a);                  \
    if ((rc = check_launch("tc_update_h"))) return rc;                                          \
    if ((rc = set_smem((const void*)tc_update_w_kernel<PV>, sizeof(SmemW)))) return rc;         \
    tc_update_w_kernel<PV><<<grid, 256, sizeof(SmemW), st>>>(np_pi, np_vf, a);                  \
    break;
  switch (model->P) {
    RL8_UPD(2) RL8_UPD(3) RL8_UPD(4)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_UPD
  return check_launch("tc_update_w");
}

}  // namespace rl8
