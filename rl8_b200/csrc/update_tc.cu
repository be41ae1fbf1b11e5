// PPO update on tensor cores (RL8_PREC_BF16): forward + clipped losses + hand-derived backward
// of both default networks for one minibatch, as two persistent kernels per row chunk.
//
//   tc_update_h_kernel  ("activation" kernel, 16 warps; CTAs alternate between the policy and the
//     value network, each with ITS W2 resident in smem).  Per 128-row tile, every contraction is
//     a tcgen05.mma and the CUDA cores only convert, mask and evaluate the per-row loss:
//       Z1  = [obs,1] * [W1,b1]^T          kind::tf32, K = 8 (one instruction)
//       H1  = relu(Z1) -> packed bf16 in 128 TMEM columns: the A operand of Z2 (read from tensor memory,
//                                          no shared-memory traffic) and later the layer-1 ReLU mask
//       Z2  = H1 * W2^T                     kind::f16 (bf16), K = 256, two N = 128 halves
//       H2  = relu(Z2 + b2) -> tile;  head dot products, per-row PPO loss -> dOut
//       gW3^T += H2^T * dOut                thin GEMM (N = 16)
//       dZ2 = [H2 > 0] .* (dOut * W3) -> tile -> one 64 KB bulk store to the dZ2 scratch
//       dH1 = dZ2 * W2  (the SAME W2 image read MN-major, one N = 256 GEMM);  [., gb2] += dZ2^T * [obs,1]
//       dZ1 = [H1 > 0] .* dH1 -> tile;  [gW1, gb1] += dZ1^T * [obs,1]
//     The thin gradients accumulate in TMEM over all tiles of the CTA and are flushed once.
//   tc_update_w_kernel  ("weight" kernel).  gW2 = dZ2^T * H1 needs a 256x256 fp32 accumulator =
//     all 512 TMEM columns, so each CTA owns HALF of the hidden units j of one network (256
//     columns) and uses the other 256 columns to recompute Z1 -> H1 with the same tf32
//     instruction (bit-identical to the activation kernel).  dZ2 half-tiles arrive by bulk
//     async copy (double-buffered), so the loop is: TMA -> [tf32 MMA -> relu/pack epilogue]
//     overlapped with the previous tile's gW2 MMAs; a 17th warp issues every bulk copy and MMA.
//
// GEMM work per row and network: 3 x 256 x 256 MACs -- the minimum (forward, dH1, gW2).
// Activations never reach HBM except the bf16 dZ2 tile (512 B / row / network, written once and
// read once).
#include "mlp_tc.cuh"
#include <stdlib.h>

#include "ppo_loss_math.cuh"

namespace rl8 {

using namespace tc;

constexpr int kUpdThreads = 512;
constexpr int64_t kChunkRows = 1 << 21;  // rows per kernel pair: bounds the dZ2 scratch at 2 GiB (of 180 GB)

struct UpdArgs {
  const float* obs;      // [T+1][D][N]
  const void* actions;   // [T+1][N]
  const float* logp;     // [T+1][N]
  const float* adv;      // [T+1][N]
  const float* ret;      // [T+1][N]
  const int64_t* rows;   // minibatch row indices (n*T + t) or null
  int64_t row_begin;     // first flattened row when rows == null
  int64_t M;             // rows of the minibatch
  int64_t row_off, Mc;   // this chunk: minibatch rows [row_off, row_off + Mc)
  int64_t N;
  int64_t slab_env0, slab_nenv;  // nenv > 0: order-free traversal t-major over envs [env0, env0+nenv)
  int T, dist_kind;
  int small;             // every row count fits 31 bits: 32-bit index arithmetic
  int n_pi;              // activation kernel: CTAs [0, n_pi) run the policy network, the rest the value network
  rl8_ppo_hparams hp;
  float inv_denom;
  uint8_t* dz[2];        // [tiles of the chunk][64 KB] bf16 dZ2 tile images per network
  float *gw1[2], *gb1[2], *gw2[2], *gb2[2], *gw3[2], *gb3[2];
  double* sums;          // [5]
  unsigned long long* phase;  // debug: [16] cycle sums of the activation kernel's phases, or null
};

// Phase timing (debug hook rl8_tc_phase_buffer): one observer thread stamps the clock at every phase
// boundary it passes together with the rest of the CTA.
struct PhaseClock {
  unsigned long long* dst;
  long long last;
  __device__ __forceinline__ PhaseClock(unsigned long long* d) : dst(d), last(0) {
    if (dst) last = clock64();
  }
  __device__ __forceinline__ void mark(int i) {
    if (dst) {
      const long long now = clock64();
      atomicAdd(dst + i, (unsigned long long)(now - last));
      last = now;
    }
  }
};

// Buffer coordinates (slab t, env n) of minibatch row `rw`; false past the minibatch.
__device__ __forceinline__ bool row_to_tn(const UpdArgs& a, int64_t rw, int64_t& t, int64_t& n) {
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

// ---- shared pieces of both kernels -------------------------------------------------------------------
// [W1 | b1] rounded to tf32, chunked fp32 [256 rows i][8]: off(i, d) = i*16 + (d/4)*4096 + (d%4)*4
__device__ __forceinline__ void stage_w1aug(uint8_t* w1aug, const NetParams& np) {
  for (int e = threadIdx.x; e < H * 8; e += blockDim.x) {
    const int i = e >> 3, d = e & 7;
    float v = 0.0f;
    if (d < np.D) v = np.w1[i * np.D + d];
    else if (d == np.D) v = np.b1[i];
    *reinterpret_cast<float*>(w1aug + i * 16 + (d >> 2) * (H * 16) + (d & 3) * 4) = tf32_round(v);
  }
}

// The two [obs, 1] slots (d = dsel, dsel + 4) of row rr this thread stages; loaded a tile ahead.
struct ObsRegs {
  float v0, v1;
};
__device__ __forceinline__ ObsRegs load_obs(const UpdArgs& a, bool valid, int64_t t, int64_t n, int D) {
  const int d0 = (threadIdx.x >> 7) & 3;  // 0..3
  ObsRegs o;
  o.v0 = d0 == D ? 1.0f : 0.0f;
  o.v1 = d0 + 4 == D ? 1.0f : 0.0f;
  if (valid) {
    const float* base = a.obs + t * (int64_t)D * a.N + n;
    if (d0 < D) o.v0 = __ldg(base + (int64_t)d0 * a.N);
    if (d0 + 4 < D) o.v1 = __ldg(base + (int64_t)(d0 + 4) * a.N);
  }
  return o;
}
// aug32: fp32 (tf32-rounded) chunked [128 rows][8]: off(r, d) = r*16 + (d/4)*2048 + (d%4)*4
__device__ __forceinline__ void store_aug32(uint8_t* aug32, const ObsRegs& o) {
  const int rr = threadIdx.x & (TILE - 1), d0 = (threadIdx.x >> 7) & 3;
  *reinterpret_cast<float*>(aug32 + rr * 16 + d0 * 4) = tf32_round(o.v0);
  *reinterpret_cast<float*>(aug32 + rr * 16 + TILE * 16 + d0 * 4) = tf32_round(o.v1);
}

__device__ __forceinline__ void issue_z1(uint32_t d_tmem, const uint8_t* aug32, const uint8_t* w1aug) {
  mma_tf32(d_tmem, smem_desc(smem_u32(aug32), TILE * 16, 128), smem_desc(smem_u32(w1aug), H * 16, 128),
           instr_desc_tf32(TILE, H), 0u);
}

// Column ownership of the epilogues: thread (q = warp % 4, cq = warp / 4) owns row 32q + lane and the
// two 32-column groups [32cq, 32cq + 32) and [128 + 32cq, 128 + 32cq + 32): one group per half of an
// accumulator, so the epilogue of half 0 overlaps the MMAs of half 1.
__device__ __forceinline__ int group_col0(int cq, int h) { return 128 * h + 32 * cq; }

// relu + pack 32 fp32 columns -> 16 bf16 pairs
__device__ __forceinline__ void relu_pack32(const float* v, uint32_t* hp) {
#pragma unroll
  for (int i = 0; i < 16; ++i) hp[i] = pack_relu_bf16x2(v[2 * i], v[2 * i + 1]);
}
__device__ __forceinline__ void store_group(uint8_t* tile, int r, int col0, const uint32_t* hp) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    *reinterpret_cast<uint4*>(tile + chunk_offset<TILE>(r, col0 / 8 + k)) =
        make_uint4(hp[4 * k], hp[4 * k + 1], hp[4 * k + 2], hp[4 * k + 3]);
}

// ---- activation kernel ---------------------------------------------------------------------------------
struct SmemH {
  uint8_t w2[kW2Bytes];        // 131072
  uint8_t a_tile[kTileBytes];  //  65536  H1 -> H2 -> dZ2 -> dZ1
  uint8_t w1aug[H * 32];       //   8192
  uint8_t w3img[16 * H * 2];   //   8192  bf16 chunked [16 rows p][256 cols j]: off(p, c8) = p*16 + c8*256
  union {
    uint8_t aug32[TILE * 32];     //  tf32 [obs, 1] (A of the layer-1 MMA): dead once Z1 is in TMEM
    float part[3][TILE][kMaxPT];  //  head partial sums of column quarters 1..3 (phases D, E)
  } u;                         //   6144
  uint8_t thin[3][TILE * 16];  //   6144  bf16 [r][8]: [0] = [obs, 1, 0..], [1] = [dOut, 0..], [2] = zeros (the
                               //         second 8-column group of both N = 16 operands)
  float b2[H];                 //   1024
  float w3[kMaxPT][H];         //   4096
  uint64_t bar_w, bar[7];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemH) + 1024 <= 227 * 1024, "SmemH exceeds the 227 KB CTA limit");

// TMEM columns of the activation kernel
constexpr uint32_t kColMain = 0;     // 256: Z1 / Z2 / G / dH1
constexpr uint32_t kColH1 = 256;     // 128: packed bf16 H1 (for the layer-1 ReLU mask)
constexpr uint32_t kColThin = 384;   // 6 x 16: gW3, gb2, [gW1 gb1], two 128-unit blocks each
constexpr int kThinN = 16;
// mbarriers: each completes exactly once per tile, so one phase bit (tile parity) serves all
enum { kBZ1 = 0, kBZ2A, kBZ2B, kBT1, kBDA, kBDB, kBT3 };  // kBZ1: Z1 of a tile + phase J of the tile before it

// D[128 units][16] (+)= X[:, 128-unit block]^T * Y with X the activation tile (MN-major A) and Y a
// [r][16] operand whose second 8-column group is the shared zero block at b_saddr + b_sbo.
__device__ __forceinline__ void issue_thin(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr,
                                           uint32_t b_sbo, bool accumulate) {
  const uint32_t idesc = instr_desc(TILE, kThinN, 1, 1);
  uint64_t ad = smem_desc(a_saddr, 128, TILE * 16), bd = smem_desc(b_saddr, 128, b_sbo);
#pragma unroll
  for (int k = 0; k < TILE / 16; ++k) {
    mma_bf16(d_tmem, ad, bd, idesc, (k > 0 || accumulate) ? 1u : 0u);
    ad += 256 >> 4, bd += 256 >> 4;  // 16 rows of K per instruction
  }
}

template <int PN, bool POLICY>
__device__ __forceinline__ void update_h_body(SmemH& s, const NetParams& np, const UpdArgs& a, int net,
                                              int cta, int nctas) {
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cq = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const int D = np.D;
  const int64_t ntiles = (a.Mc + TILE - 1) / TILE;
  float b3[PN], gb3_acc[PN];
#pragma unroll
  for (int p = 0; p < PN; ++p) b3[p] = np.b3[p], gb3_acc[p] = 0.0f;
  float s_ent = 0, s_pol = 0, s_vf = 0, s_kl = 0;  // per-thread partial sums (<= a few hundred rows)
  int it = 0;
  const bool continuous = POLICY && a.dist_kind != RL8_DIST_CATEGORICAL;

  // [obs, 1] of row (tid & 127) in bf16 -> thin[0] (B operand of the gb2 / gW1 thin GEMMs)
  auto store_thin0 = [&](const ObsRegs& o) {
    const int rr = tid & (TILE - 1), d0 = tid >> 7;
    __nv_bfloat16* row16 = reinterpret_cast<__nv_bfloat16*>(s.thin[0] + rr * 16);
    row16[d0] = __float2bfloat16(o.v0);
    row16[d0 + 4] = __float2bfloat16(o.v1);
  };
  // prefetch of the first tile; its layer-1 MMA is issued here, every later one together with the
  // previous tile's last thin GEMM (phase J), so a tile starts with Z1 already in flight
  int64_t tile = cta;
  int64_t tn = 0, nn = 0;
  bool valid_next = tile < ntiles && row_to_tn(a, a.row_off + tile * TILE + (tid & (TILE - 1)), tn, nn);
  ObsRegs obs_next = load_obs(a, valid_next, tn, nn, D);
  int64_t idx_next = valid_next ? tn * a.N + nn : -1;
  if (tile < ntiles) {
    store_aug32(s.u.aug32, obs_next);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (cta_issuer()) {
      fence_after_sync();
      issue_z1(tmem + kColMain, s.u.aug32, s.w1aug);
      mma_commit(&s.bar[kBZ1]);
    }
  }

  PhaseClock pc(a.phase && tid == kUpdThreads - 1 && cta == 0 ? a.phase + 8 * net : nullptr);
  for (; tile < ntiles; tile += nctas, ++it) {
    const uint32_t ph = (uint32_t)(it & 1);
    // ---- A. this tile's per-row loss inputs (threads 0..127), in flight until phase E ---------------------
    const int64_t idx = idx_next;  // of row (tid & 127)
    const ObsRegs obs_cur = obs_next;
    // (the action stays a raw 32-bit word until phase E: converting here would wait for the load)
    uint32_t in_act_raw = 0u;
    float in_logp = 0.0f, in_tgt = 0.0f;
    if (tid < TILE && idx >= 0) {
      if constexpr (POLICY) {
        in_act_raw = a.dist_kind == RL8_DIST_CATEGORICAL ? ((const uint32_t*)a.actions)[2 * idx]
                                                         : ((const uint32_t*)a.actions)[idx];
        in_logp = a.logp[idx];
        in_tgt = a.adv[idx];
      } else {
        in_tgt = a.ret[idx];
      }
    }
    // ---- B. prefetch the next tile's observations (staged in phase G) ---------------------------------------
    const bool has_next = tile + nctas < ntiles;
    {
      const int64_t nt = tile + nctas;
      valid_next = has_next && row_to_tn(a, a.row_off + nt * TILE + (tid & (TILE - 1)), tn, nn);
      obs_next = load_obs(a, valid_next, tn, nn, D);
      idx_next = valid_next ? tn * a.N + nn : -1;
    }
    // Z1 of this tile is in TMEM and the previous tile's last thin GEMM has read the tile and thin[0]
    mbar_wait(&s.bar[kBZ1], ph);
    fence_after_sync();
    store_thin0(obs_cur);  // read by the MMAs of phases H and J, several barriers from here
    pc.mark(0);  // A + B: loads, Z1 / previous J round trip
    // ---- C. H1 = relu(Z1) packed to bf16 in TMEM only: it is the A operand of Z2 = H1 * W2^T (read by the tensor
    //         core straight from tensor memory: no shared-memory traffic, the activation tile stays free) and,
    //         in phase I, the layer-1 ReLU mask --------------------------------------------------------------------
    {
      float v0[32], v1[32];
      tmem_ld32_nowait(tmem + kColMain + lane_base + (uint32_t)group_col0(cq, 0), v0);
      tmem_ld32_nowait(tmem + kColMain + lane_base + (uint32_t)group_col0(cq, 1), v1);
      tmem_wait_ld();
      reg_fence32(v0);
      reg_fence32(v1);
      uint32_t hp[16];
      relu_pack32(v0, hp);
      tmem_st16_raw(tmem + kColH1 + lane_base + (uint32_t)(group_col0(cq, 0) / 2), hp);
      relu_pack32(v1, hp);
      tmem_st16_raw(tmem + kColH1 + lane_base + (uint32_t)(group_col0(cq, 1) / 2), hp);
      tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (cta_issuer()) {
      fence_after_sync();
      const uint32_t idesc = instr_desc(TILE, 128, 0, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // B = W2 rows [128 h, 128 h + 128) K-major: LBO = H*16 (next 8 K), SBO = 128, 16 K = 2 * H * 16 bytes
        uint64_t bd = smem_desc(smem_u32(s.w2) + h * 2048, H * 16, 128);
#pragma unroll
        for (int ks = 0; ks < H / 16; ++ks) {
          mma_bf16_ts(tmem + kColMain + 128 * h, tmem + kColH1 + 8 * ks, bd, idesc, ks > 0 ? 1u : 0u);
          bd += (2 * H * 16) >> 4;
        }
        mma_commit(&s.bar[kBZ2A + h]);
      }
    }
    pc.mark(1);  // C: H1 epilogue
    // ---- D. H2 = relu(Z2 + b2) -> tile, head partial sums; half 0 is processed under the MMAs of half 1 -------
    float dot[PN];
#pragma unroll
    for (int p = 0; p < PN; ++p) dot[p] = 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col0 = group_col0(cq, h);
      mbar_wait(&s.bar[kBZ2A + h], ph);
      fence_after_sync();
      float v[32];
      tmem_ld32(tmem + kColMain + lane_base + (uint32_t)col0, v);
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
      }
      uint32_t hp[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) hp[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      store_group(s.a_tile, r, col0, hp);  // nothing reads the tile during this phase
    }
    if (cq > 0) {
#pragma unroll
      for (int p = 0; p < PN; ++p) s.u.part[cq - 1][r][p] = dot[p];
    }
    __syncthreads();
    pc.mark(2);  // D: Z2 MMAs + H2 epilogue
    // ---- E. per-row loss -> dOut (threads 0..127 own row tid and column quarter 0) -------------------------------------
    if (tid < TILE) {
      float o[PN], d_o[PN];
#pragma unroll
      for (int p = 0; p < PN; ++p)
        o[p] = ((dot[p] + s.u.part[0][tid][p]) + s.u.part[1][tid][p]) + s.u.part[2][tid][p] + b3[p];
#pragma unroll
      for (int p = 0; p < PN; ++p) d_o[p] = 0.0f;
      if (idx >= 0) {
        RowLoss L;
        if constexpr (POLICY) {
          if (continuous) o[1] = tanhf(o[1]);
          const float in_act = a.dist_kind == RL8_DIST_CATEGORICAL ? (float)(int)in_act_raw
                                                                   : __uint_as_float(in_act_raw);
          ppo_policy_row<PN>(a.dist_kind, o, in_act, in_logp, in_tgt, a.hp, a.inv_denom, d_o, L);
          s_ent += L.entropy, s_pol += L.policy, s_kl += L.kl;
        } else {
          ppo_value_row(o[0], in_tgt, a.hp, a.inv_denom, d_o, L);
          s_vf += L.vf;
        }
      }
      float v8[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) v8[p] = p < PN ? d_o[p] : 0.0f;
      store_chunk(s.thin[1], (uint32_t)tid * 16u, v8);
#pragma unroll
      for (int p = 0; p < PN; ++p) gb3_acc[p] += d_o[p];
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    pc.mark(3);  // E: row loss
    // ---- F. G = dOut * W3 (K = 16: one instruction), awaited;  gW3^T += H2^T * dOut completes under phase G ---------
    if (cta_issuer()) {
      fence_after_sync();
      mma_bf16(tmem + kColMain, smem_desc(smem_u32(s.thin[1]), TILE * 16, 128),
               smem_desc(smem_u32(s.w3img), 128, 16 * 16), instr_desc(TILE, H, 0, 1), 0u);
      mma_commit(&s.bar[kBT1]);
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
        issue_thin(tmem + kColThin + kThinN * jb, smem_u32(s.a_tile) + jb * 32768, smem_u32(s.thin[1]),
                   TILE * 16, it > 0);
      mma_commit(&s.bar[kBT3]);
    }
    mbar_wait(&s.bar[kBT1], ph);
    fence_after_sync();
    pc.mark(4);  // F: G MMA round trip
    // ---- G. dZ2 = [H2 > 0] .* G over H2: the first column group is computed while the thin GEMM still reads H2 ---------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[32];
      tmem_ld32(tmem + kColMain + lane_base + (uint32_t)group_col0(cq, h), v);
      uint32_t o[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 h2 =
            *reinterpret_cast<const uint4*>(s.a_tile + chunk_offset<TILE>(r, group_col0(cq, h) / 8 + k));
        o[4 * k + 0] = pack_bf16x2(v[8 * k + 0], v[8 * k + 1]) & gt0_mask_bf16x2(h2.x);
        o[4 * k + 1] = pack_bf16x2(v[8 * k + 2], v[8 * k + 3]) & gt0_mask_bf16x2(h2.y);
        o[4 * k + 2] = pack_bf16x2(v[8 * k + 4], v[8 * k + 5]) & gt0_mask_bf16x2(h2.z);
        o[4 * k + 3] = pack_bf16x2(v[8 * k + 6], v[8 * k + 7]) & gt0_mask_bf16x2(h2.w);
      }
      if (h == 0) {  // gW3's MMAs have read H2: the tile may be overwritten
        mbar_wait(&s.bar[kBT3], ph);
        fence_after_sync();
      }
      store_group(s.a_tile, r, group_col0(cq, h), o);
    }
    if (has_next) store_aug32(s.u.aug32, obs_next);  // `part` (same bytes) was last read in phase E
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    pc.mark(5);  // G: dZ2 epilogue
    // ---- H. dZ2 tile -> scratch;  dH1 = dZ2 * W2 in two column halves;  [., gb2] += dZ2^T * [obs, 1] --------------------------
    if (cta_issuer()) {
      fence_after_sync();
      uint8_t* dst = a.dz[net] + tile * (int64_t)kTileBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) bulk_s2g(dst + i * 16384, s.a_tile + i * 16384, 16384);
      bulk_commit();
      // one N = 256 GEMM: the phase is bound by shared-memory bandwidth (operand fetch + the bulk store + the
      // dZ1 stores), and N = 256 fetches the A operand once per K step instead of once per column half
      issue_gemm(tmem + kColMain, smem_u32(s.a_tile), TILE, false, smem_u32(s.w2), H, true, TILE, H, H, false);
      mma_commit(&s.bar[kBDA]);
#pragma unroll
      for (int jb = 0; jb < 2; ++jb)
        issue_thin(tmem + kColThin + kThinN * (2 + jb), smem_u32(s.a_tile) + jb * 32768, smem_u32(s.thin[0]),
                   2 * TILE * 16, it > 0);
      mma_commit(&s.bar[kBDB]);
      bulk_wait_read();  // the store engine has read the tile
    }
    // ---- I. dZ1 = [H1 > 0] .* dH1 in registers while the thin gb2 GEMM still reads dZ2, then stored over it ----------------
    {
      uint32_t held[16], hp[16];
      mbar_wait(&s.bar[kBDA], ph);
      fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int col0 = group_col0(cq, h);
        float v[32];
        uint32_t* out = h == 0 ? held : hp;
        tmem_ld16_raw(tmem + kColH1 + lane_base + (uint32_t)(col0 / 2), out);
        tmem_ld32_nowait(tmem + kColMain + lane_base + (uint32_t)col0, v);
        tmem_wait_ld();
        reg_fence16(out);
        reg_fence32(v);
#pragma unroll
        for (int i = 0; i < 16; ++i) out[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]) & gt0_mask_bf16x2(out[i]);
      }
      mbar_wait(&s.bar[kBDB], ph);  // every MMA that reads dZ2 is done
      fence_after_sync();
      __syncthreads();  // ... and the issuing lane's bulk_wait_read precedes every overwrite of the tile
      store_group(s.a_tile, r, group_col0(cq, 0), held);
      store_group(s.a_tile, r, group_col0(cq, 1), hp);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    pc.mark(6);  // H + I: dH1 MMAs + dZ1 epilogue
    // ---- J. [gW1, gb1] += dZ1^T * [obs, 1];  Z1 of the next tile -- one commit, awaited at the top of the next tile -----
    // (Z1 first on its own barrier was measured: the thin GEMM then delays the next tile's Z2 on the in-order pipe
    // by more than the shorter wait saves)
    if (cta_issuer()) {
      fence_after_sync();
#pragma unroll
      for (int ib = 0; ib < 2; ++ib)
        issue_thin(tmem + kColThin + kThinN * (4 + ib), smem_u32(s.a_tile) + ib * 32768, smem_u32(s.thin[0]),
                   2 * TILE * 16, it > 0);
      if (has_next) issue_z1(tmem + kColMain, s.u.aug32, s.w1aug);
      mma_commit(&s.bar[kBZ1]);
    }
    pc.mark(7);  // J: issue only
  }
  if (it > 0) {  // the last tile's thin GEMM: accumulators complete, tile free
    mbar_wait(&s.bar[kBZ1], (uint32_t)(it & 1));
    fence_after_sync();
  }
  if (cta_issuer()) bulk_wait_all();  // dZ2 stores have landed before the kernel ends

  // ---- flush the thin gradients (warps 0-3 -> units 0..127, warps 4-7 -> units 128..255) ---------------------------------
  if (it > 0 && warp < 8) {
    const int blk = warp >> 2;
    const int c = blk * 128 + q * 32 + lane;  // hidden unit of this thread
    float v[8];
    tmem_ld8(tmem + kColThin + kThinN * blk + lane_base, v);
#pragma unroll
    for (int p = 0; p < PN; ++p) atomicAdd(a.gw3[net] + p * H + c, v[p]);
    tmem_ld8(tmem + kColThin + kThinN * (2 + blk) + lane_base, v);
#pragma unroll
    for (int d = 0; d < 8; ++d)
      if (d == D) atomicAdd(a.gb2[net] + c, v[d]);
    tmem_ld8(tmem + kColThin + kThinN * (4 + blk) + lane_base, v);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      if (d < D) atomicAdd(a.gw1[net] + c * D + d, v[d]);
      if (d == D) atomicAdd(a.gb1[net] + c, v[d]);
    }
  }
  if (it > 0 && tid < TILE) {
#pragma unroll
    for (int p = 0; p < PN; ++p) {
      const float w = warp_sum(gb3_acc[p]);
      if (lane == 0) atomicAdd(a.gb3[net] + p, w);
    }
  }
  __shared__ double red[32];
  double sv[4] = {(double)s_ent, (double)s_pol, (double)s_vf, (double)s_kl};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double t = block_sum(sv[i], red);
    if (tid == 0 && t != 0.0) atomicAdd(a.sums + i, t);
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(a.sums + 4, (double)a.Mc);
}

template <int P>
__global__ void __launch_bounds__(kUpdThreads, 1)
tc_update_h_kernel(NetParams np_pi, NetParams np_vf, UpdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemH& s = *reinterpret_cast<SmemH*>(smem_raw);
  const int net = (int)blockIdx.x < a.n_pi ? 0 : 1;
  const int cta = net ? (int)blockIdx.x - a.n_pi : (int)blockIdx.x;
  const int nctas = net ? (int)gridDim.x - a.n_pi : a.n_pi;
  const NetParams np = net ? np_vf : np_pi;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s.bar_w, 1);
#pragma unroll
    for (int i = 0; i < 7; ++i) mbar_init(&s.bar[i], 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_expect_tx(&s.bar_w, kW2Bytes);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      bulk_g2s(s.w2 + i * (kW2Bytes / 8), np.w2_img + i * (kW2Bytes / 8), kW2Bytes / 8, &s.bar_w);
  }
  stage_w1aug(s.w1aug, np);
  for (int i = tid; i < H; i += blockDim.x) s.b2[i] = np.b2[i];
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  for (int i = tid; i < 16 * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    *reinterpret_cast<__nv_bfloat16*>(s.w3img + p * 16 + (c >> 3) * 256 + (c & 7) * 2) =
        __float2bfloat16(p < np.P ? np.w3[p * H + c] : 0.0f);
  }
  for (int i = tid; i < 3 * TILE * 16 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s.thin)[i] = 0u;
  __syncthreads();
  mbar_wait(&s.bar_w, 0);
  if (net == 0) update_h_body<P, true>(s, np, a, 0, cta, nctas);
  else update_h_body<1, false>(s, np, a, 1, cta, nctas);
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(s.tmem_base, 512);
}

// ---- weight-gradient kernel -----------------------------------------------------------------------------------
struct SmemW {
  uint8_t h1_tile[2][kTileBytes];      // 131072  H1 tiles (MN-major B: rows = K = r, cols = i)
  uint8_t dz_tile[2][kTileBytes / 2];  //  65536  dZ2 half tiles (MN-major A: rows = K = r, cols = 128 units j)
  uint8_t w1aug[H * 32];               //   8192
  uint8_t aug32[2][TILE * 32];         //   8192
  uint64_t bar_tma[2], bar_z, bar_g[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemW) + 512 <= 227 * 1024, "SmemW exceeds the 227 KB CTA limit");

// grid: blockIdx & 1 = network, (blockIdx >> 1) & 1 = half of the hidden units j, blockIdx >> 2 = CTA of the role
// block: 16 epilogue warps (Z1 -> H1 tiles) + ONE issuing warp (warp 16: bulk copies and every tcgen05.mma),
// so the issue / wait chain of the tensor pipe runs beside the epilogue instead of in series with it.
constexpr int kUpdWThreads = kUpdThreads + 32;

__global__ void __launch_bounds__(kUpdWThreads, 1)
tc_update_w_kernel(NetParams np_pi, NetParams np_vf, UpdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemW& s = *reinterpret_cast<SmemW*>(smem_raw);
  const int net = blockIdx.x & 1, jh = (blockIdx.x >> 1) & 1;
  const NetParams np = net ? np_vf : np_pi;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool epi = tid < kUpdThreads;  // epilogue thread (the rest is the issuing warp)
  const int q = warp & 3, cq = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;
  const int D = np.D;
  if (tid == 0) {
    mbar_init(&s.bar_tma[0], 1);
    mbar_init(&s.bar_tma[1], 1);
    mbar_init(&s.bar_z, 1);
    mbar_init(&s.bar_g[0], 1);
    mbar_init(&s.bar_g[1], 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, 512);
  stage_w1aug(s.w1aug, np);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  constexpr uint32_t kColG = 0, kColZ = 256;

  const int64_t ntiles = (a.Mc + TILE - 1) / TILE;
  const int nctas = gridDim.x >> 2;
  const int64_t tile0 = blockIdx.x >> 2;
  const int64_t n_my = tile0 < ntiles ? (ntiles - tile0 + nctas - 1) / nctas : 0;
  const uint8_t* dz_src = a.dz[net] + (int64_t)jh * (kTileBytes / 2);

  auto tma_load = [&](int64_t k) {  // issuing lane
    const int st = (int)(k & 1);
    const uint8_t* src = dz_src + (tile0 + k * nctas) * (int64_t)kTileBytes;
    mbar_expect_tx(&s.bar_tma[st], kTileBytes / 2);
    bulk_g2s(s.dz_tile[st], src, 16384, &s.bar_tma[st]);
    bulk_g2s(s.dz_tile[st] + 16384, src + 16384, 16384, &s.bar_tma[st]);
  };

  if (n_my > 0) {
    if (epi) {
      int64_t t0 = 0, n0 = 0;
      const bool valid = row_to_tn(a, a.row_off + tile0 * TILE + (tid & (TILE - 1)), t0, n0);
      store_aug32(s.aug32[0], load_obs(a, valid, t0, n0, D));
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (!epi && elect_one()) {
      fence_after_sync();
      tma_load(0);
      issue_z1(tmem + kColZ, s.aug32[0], s.w1aug);
      mma_commit(&s.bar_z);
    }
    // debug stamps (CTA 0): [16] wait for Z1 / stage, [17] H1 epilogue + barrier, and on the issuing
    // lane [18] wait for the dZ2 bulk copy, [19] wait for the previous gW2, [20] issue
    PhaseClock pc(a.phase && tid == kUpdThreads - 1 && blockIdx.x == 0 ? a.phase + 16 : nullptr);
    for (int64_t k = 0; k < n_my; ++k) {
      const int st = (int)(k & 1);
      const bool more = k + 1 < n_my;
      if (epi) {
        ObsRegs nxt;
        if (more) {
          int64_t t1 = 0, n1 = 0;
          const bool valid =
              row_to_tn(a, a.row_off + (tile0 + (k + 1) * nctas) * TILE + (tid & (TILE - 1)), t1, n1);
          nxt = load_obs(a, valid, t1, n1, D);
        }
        mbar_wait(&s.bar_z, (uint32_t)(k & 1));  // Z1(k) is in TMEM
        if (k >= 2) mbar_wait(&s.bar_g[st], (uint32_t)(((k - 2) >> 1) & 1));  // gW2(k-2) has read stage st
        fence_after_sync();
        pc.mark(0);
        {
          float v0[32], v1[32];
          tmem_ld32_nowait(tmem + kColZ + lane_base + (uint32_t)group_col0(cq, 0), v0);
          tmem_ld32_nowait(tmem + kColZ + lane_base + (uint32_t)group_col0(cq, 1), v1);
          tmem_wait_ld();
          reg_fence32(v0);
          reg_fence32(v1);
          uint32_t hp[16];
          relu_pack32(v0, hp);
          store_group(s.h1_tile[st], r, group_col0(cq, 0), hp);
          relu_pack32(v1, hp);
          store_group(s.h1_tile[st], r, group_col0(cq, 1), hp);
        }
        if (more) store_aug32(s.aug32[st ^ 1], nxt);
        fence_async_smem();
      }
      fence_before_sync();
      __syncthreads();
      pc.mark(1);
      if (!epi && elect_one()) {
        PhaseClock pi(a.phase && blockIdx.x == 0 ? a.phase + 18 : nullptr);
        fence_after_sync();
        if (more) {
          issue_z1(tmem + kColZ, s.aug32[st ^ 1], s.w1aug);
          mma_commit(&s.bar_z);
          // refill the other dZ2 stage as early as possible: gW2(k-1), issued one iteration ago, was its reader
          if (k >= 1) mbar_wait(&s.bar_g[st ^ 1], (uint32_t)(((k - 1) >> 1) & 1));
          pi.mark(1);
          tma_load(k + 1);
        }
        mbar_wait(&s.bar_tma[st], (uint32_t)((k >> 1) & 1));
        pi.mark(0);
        // gW2[j half][i] += dZ2[:, j half]^T * H1: A = dz half tile (MN-major), B = H1 tile (MN-major)
        issue_gemm(tmem + kColG, smem_u32(s.dz_tile[st]), TILE, true, smem_u32(s.h1_tile[st]), TILE, true,
                   TILE, H, TILE, k > 0);
        mma_commit(&s.bar_g[st]);
        pi.mark(2);
      }
    }
    // all MMAs done: the last commit covers every earlier one
    mbar_wait(&s.bar_g[(n_my - 1) & 1], (uint32_t)(((n_my - 1) >> 1) & 1));
    fence_after_sync();
    // flush: lane = unit j of this half, 256 columns i; warp (q, cq) -> lanes 32q.., columns 64cq..
    if (epi) {
      const int j = jh * 128 + r;
      float* dst = a.gw2[net] + (int64_t)j * H + cq * 64;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(tmem + kColG + lane_base + (uint32_t)(cq * 64 + h * 32), v);
#pragma unroll
        for (int e = 0; e < 32; e += 4) red_add_v4(dst + h * 32 + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(s.tmem_base, 512);
}

// ---- host -----------------------------------------------------------------------------------------------------
static unsigned long long* g_phase_buffer = nullptr;

// Debug hook: device buffer of 16 counters (8 phases x {policy, value} CTA 0) the activation kernel adds
// its per-phase cycle counts to; null switches the stamping off.
extern "C" int rl8_tc_phase_buffer(unsigned long long* device_counters) {
  g_phase_buffer = device_counters;
  return RL8_OK;
}

static int policy_ctas() {
  static int n = 0;
  if (!n) {
    const char* e = getenv("RL8_H_POLICY_CTAS");  // tuning knob; default measured on B200
    n = e ? atoi(e) : 80;
    if (n < 1 || n > kNumSMs - 1) n = kNumSMs / 2;
  }
  return n;
}

int64_t ppo_tc_workspace(const rl8_model*, int64_t max_rows) {
  const int64_t chunk = max_rows < kChunkRows ? max_rows : kChunkRows;
  return 2 * (int64_t)kW2Bytes + 2 * ceil_div(chunk, TILE) * (int64_t)kTileBytes + 256;
}

int ppo_minibatch_tc(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch,
                     const int64_t* rows, int64_t row_begin, int64_t M, double mean_denominator,
                     const rl8_ppo_hparams* hp, double* loss_sums, void* workspace,
                     int64_t workspace_bytes, cudaStream_t st) {
  if (model->H != H || model->P > kMaxPT || model->D > 7) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < ppo_tc_workspace(model, M)) return RL8_ERR_WORKSPACE;
  const int64_t chunk = M < kChunkRows ? M : kChunkRows;
  const int64_t chunk_tiles = ceil_div(chunk, TILE);
  uint8_t* img_pi = (uint8_t*)workspace;
  uint8_t* img_vf = img_pi + kW2Bytes;
  uint8_t* dz0 = img_vf + kW2Bytes;
  uint8_t* dz1 = dz0 + chunk_tiles * (int64_t)kTileBytes;
  int rc;
  if ((rc = launch_pack_w2(model->pi_w2, img_pi, st))) return rc;
  if ((rc = launch_pack_w2(model->vf_w2, img_vf, st))) return rc;
  const NetParams np_pi = net_params(model, 0, img_pi), np_vf = net_params(model, 1, img_vf);
  UpdArgs a;
  a.obs = batch->obs, a.actions = batch->actions, a.logp = batch->logp;
  a.adv = batch->advantages, a.ret = batch->returns;
  a.rows = rows, a.row_begin = row_begin, a.M = M, a.N = batch->N, a.T = batch->T;
  // A contiguous run of whole environments may be visited in any order (the minibatch is a sum):
  // walk it slab by slab so every load is coalesced in the horizon-major buffer.
  a.slab_env0 = 0, a.slab_nenv = 0;
  if (!rows && batch->T > 0 && row_begin % batch->T == 0 && M % batch->T == 0) {
    a.slab_env0 = row_begin / batch->T;
    a.slab_nenv = M / batch->T;
  }
  a.small = (M < (1ll << 31) && (int64_t)(batch->T + 1) * batch->N < (1ll << 31)) ? 1 : 0;
  a.dist_kind = batch->dist_kind, a.hp = *hp;
  a.inv_denom = (float)((double)hp->loss_scale / mean_denominator);
  a.dz[0] = dz0, a.dz[1] = dz1;
  a.gw1[0] = (float*)grads->pi_w1, a.gb1[0] = (float*)grads->pi_b1, a.gw2[0] = (float*)grads->pi_w2;
  a.gb2[0] = (float*)grads->pi_b2, a.gw3[0] = (float*)grads->pi_w3, a.gb3[0] = (float*)grads->pi_b3;
  a.gw1[1] = (float*)grads->vf_w1, a.gb1[1] = (float*)grads->vf_b1, a.gw2[1] = (float*)grads->vf_w2;
  a.gb2[1] = (float*)grads->vf_b2, a.gw3[1] = (float*)grads->vf_w3, a.gb3[1] = (float*)grads->vf_b3;
  a.sums = loss_sums;
  a.phase = g_phase_buffer;
  for (int64_t off = 0; off < M; off += chunk) {
    a.row_off = off;
    a.Mc = M - off < chunk ? M - off : chunk;
    const int64_t ntiles = ceil_div(a.Mc, TILE);
    int grid_h = (int)(2 * ntiles < kNumSMs ? 2 * ntiles : kNumSMs);
    grid_h &= ~1;
    // a policy tile costs more CUDA-core work than a value tile (wider head, heavier row loss)
    a.n_pi = grid_h == kNumSMs ? policy_ctas() : grid_h / 2;
    int grid_w = (int)(4 * ntiles < kNumSMs ? 4 * ntiles : kNumSMs);
    grid_w &= ~3;
#define RL8_UPD(PV)                                                                              \
  case PV:                                                                                       \
    if ((rc = set_smem((const void*)tc_update_h_kernel<PV>, sizeof(SmemH)))) return rc;           \
    tc_update_h_kernel<PV><<<grid_h, kUpdThreads, sizeof(SmemH), st>>>(np_pi, np_vf, a);          \
    break;
    switch (model->P) {
      RL8_UPD(2) RL8_UPD(3) RL8_UPD(4)
      default: return RL8_ERR_UNSUPPORTED;
    }
#undef RL8_UPD
    if ((rc = check_launch("tc_update_h"))) return rc;
    if ((rc = set_smem((const void*)tc_update_w_kernel, sizeof(SmemW)))) return rc;
    tc_update_w_kernel<<<grid_w, kUpdWThreads, sizeof(SmemW), st>>>(np_pi, np_vf, a);
    if ((rc = check_launch("tc_update_w"))) return rc;
  }
  return RL8_OK;
}

}  // namespace rl8
