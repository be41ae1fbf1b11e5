// PPO update in RL8_PREC_FP32_TC: forward + clipped losses + hand-derived backward of both default networks for one
// minibatch on split-operand pair MMAs (split_tc.cuh), three persistent kernels per row chunk.  No activation ever
// reaches HBM: every A / B operand that is an activation is RECOMPUTED on CUDA cores straight into split ring
// stages from 80 bytes per row and network of scratch (dOut [4] fp32 and the two 256-bit ReLU masks).
//
// Pieces: two fp16 pieces per fp32 operand and three piece products per fp32 product (a0b0 + a0b1 + a1b0), every
// operand tensor scaled by a power of two from a measured bound of its magnitude (X3Scales: max |obs| -> H1, max |W2|,
// max |dOut| -> dZ2) and the scales divided out of the accumulators -- exact, so the results are those of the
// un-scaled sums.  `make X3_BF16=1` builds the first form of this path instead: three bf16 pieces, six products.
// Sign convention: H1 and H2 are produced NEGATED (h1_chunk<true>, -b2 in shared memory) so that a ReLU-mask bit is
// the sign bit of the sum; the sign travels in the factor that removes the scales.
//
//   x3_update_f_kernel  rows x units   Z2 = H1 W2^T            (three piece products)
//       A = H1 = relu([obs] W1^T + b1) computed per stage;  B = W2 piece image (bulk copies)
//       epilogue: H2 = relu(Z2 + b2), head, per-row PPO loss -> dOut, loss sums, gb3; ReLU masks of H1 / H2 and
//       dOut -> scratch;  gW3 += H2^T dOut by warp-transposing reductions (the accumulator is read twice)
//   x3_update_b_kernel  inputs x rows  dH1^T = W2^T dZ2^T      (three piece products; value network: two, the B operand
//                                                               is the ReLU mask itself, exact in one piece)
//       A = W2^T piece image;  B = dZ2 = [H2 > 0] .* (dOut W3) computed per stage from the scratch
//       epilogue (lane = input unit i, columns = rows): dZ1 = [H1 > 0] .* dH1;  gW1[i][:] += dZ1 obs, gb1[i] += dZ1
//       as plain per-thread FMAs over the rows (no cross-lane work: that is why this GEMM is transposed)
//   x3_update_w_kernel  units x inputs gW2 = dZ2^T H1          (three / two piece products), contraction over rows
//       A = dZ2^T, B = H1, both computed per stage (MN-major tiles) from inputs that two loader warps fetch once per
//       CTA; gb2 += column sums of dZ2 in the producer; the 128 x 256 accumulator of each CTA lives in tensor memory
//
// GEMM work per row: policy network 3 + 3 + 3, value network 3 + 2 + 2 half-precision-rate 256x256 products (the bf16
// path does 3 per network).
#include <stdlib.h>

#include "ppo_loss_math.cuh"
#include "split_common.cuh"

// Development build (make ABLATE=1): RL8_X3_ABL switches parts of the worker code off and rl8_x3_debug_buffer
// collects the MMA warp's wait cycles (tools/ablate_x3.py, profiles/r02_x3_ablation.md).  In the production build
// both compile to nothing.
#ifdef RL8_X3_ABLATE
#define X3_ABL(a) ((a).abl)
#define X3_CLOCK() clock64()
#else
#define X3_ABL(a) 0
#define X3_CLOCK() 0ll
#endif

namespace rl8 {

using namespace tc;

constexpr int64_t kXChunkRows = 1 << 21;  // rows per kernel triple (scratch: 80 B per row and network)

struct UpdXArgs {
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
  int n_pi;              // CTA pairs [0, n_pi) run the policy network, the rest the value network (forward kernel)
  int n_pi_b, n_pi_w;    // the same for the input-gradient / weight-gradient kernels (their value pairs do half the MMAs)
  int abl;               // development: ablation bits (RL8_X3_ABL), 0 in production
  unsigned long long* dbg;  // development: wait-cycle counters of pair 0 of each network (rl8_x3_debug_buffer), or null
  rl8_ppo_hparams hp;
  float inv_denom;
  // scratch per network, indexed by the chunk-local row
  uint32_t* mask1[2];    // [Mc][8]  bit c of the row: H1[c] > 0
  uint32_t* mask2[2];    // [Mc][8]  bit c of the row: H2[c] > 0
  float* dout[2];        // [Mc][4]  d(loss) / d(head output)
  float *gw1[2], *gb1[2], *gw2[2], *gb2[2], *gw3[2], *gb3[2];
  double* sums;          // [5]
  X3Scales* sc;          // operand magnitudes behind the power-of-two scales of the fp16 piece form
};

// observations of minibatch row `rw` (zeros past the chunk / minibatch): slots 0..D-1, rest zero
__device__ __forceinline__ bool load_row_obs(const UpdXArgs& a, int64_t rowl, int D, float* ob, int64_t* idx) {
#pragma unroll
  for (int d = 0; d < 7; ++d) ob[d] = 0.0f;
  int64_t t = 0, n = 0;
  const bool valid = rowl < a.Mc && minibatch_row_to_tn(a, a.row_off + rowl, t, n);
  if (valid) {
    const float* base = a.obs + t * (int64_t)D * a.N + n;
#pragma unroll
    for (int d = 0; d < 7; ++d)
      if (d < D) ob[d] = __ldg(base + (int64_t)d * a.N);
  }
  if (idx) *idx = valid ? t * a.N + n : -1;
  return valid;
}

// ---- forward + loss kernel ---------------------------------------------------------------------------------------------
constexpr int kFStages = kUF16 ? 6 : 4;
struct SmemXF {
  StageX<kUNP> ring[kFStages];  // 196608
  float w1t[8][H];              //   8192
  float b2[H];                  //   1024  -b2
  float w3[kMaxPT][H];          //   4096
  float part[4][TILE][kMaxPT];  //   8192  head partial sums per column quarter
  float dsm[TILE][kMaxPT];      //   2048  dOut of the tile's rows (pass 2)
  // running sums of the 128 loss threads (thread tid owns slot tid): kept here, not in registers -- every one of the
  // 512 worker threads would carry them through the whole kernel at 96 registers per thread
  double lsum[TILE][4];         //   4096  entropy, policy, vf, kl
  float gb3s[TILE][kMaxPT];     //   2048  sum of dOut (head bias gradient)
  float dmx[TILE];              //    512  max |dOut| seen by loss thread tid
  float red[32];
  uint64_t full[kFStages], bfull[kFStages], empty[kFStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemXF) <= 227 * 1024, "SmemXF exceeds the 227 KB CTA limit");

template <int PN, bool POLICY>
__device__ __forceinline__ void update_f_workers(SmemXF& s, const NetParams& np, const UpdXArgs& a, int net,
                                                 int64_t pr, int64_t npairs, uint32_t rank, float inv_scale,
                                                 double* sv) {
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rloc = tid & 127, g = tid >> 7;
  const int q = warp & 3, cq = warp >> 2, r = q * 32 + lane;
  const int D = np.D;
  const int64_t ntiles = (a.Mc + 255) / 256;
  const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
  const bool continuous = POLICY && a.dist_kind != RL8_DIST_CATEGORICAL;
  float gw3_acc[2][PN];
#pragma unroll
  for (int p = 0; p < PN; ++p) gw3_acc[0][p] = gw3_acc[1][p] = 0.0f;
  if (tid < TILE) {  // (doubles: a thread sums hundreds of O(1) terms that cancel)
#pragma unroll
    for (int i = 0; i < 4; ++i) s.lsum[tid][i] = 0.0, s.gb3s[tid][i] = 0.0f;
    s.dmx[tid] = 0.0f;
  }
  uint32_t kcount = 0;

  auto produce = [&](int64_t tile) {
    float obf[7];
    const int64_t rowl = tile * 256 + rank * 128 + rloc;
    const bool valid = load_row_obs(a, rowl, D, obf, nullptr);
    const ObsPairs ob = obs_pairs(obf);
    uint32_t m0 = 0u, m1 = 0u;
    // two ring stages per generic -> async proxy fence (the fence and the warp barrier behind it were an eighth of the
    // kernel's stall samples with one per stage); kFStages is even, so a pair never wraps around the ring
    static_assert(kFStages % 2 == 0 && (H / kXKc) % 2 == 0, "stage pairs");
    for (int kc = 0; kc < H / kXKc; kc += 2, kcount += 2) {
      const int st = (int)(kcount % kFStages);
      const uint32_t use = kcount / kFStages;
      if (use > 0) {
        mbar_wait_cluster(&s.empty[st], (use - 1) & 1);
        mbar_wait_cluster(&s.empty[st + 1], (use - 1) & 1);
      }
      __syncwarp();
      if (warp == 0 && elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint8_t* src = np.w2_img + (size_t)(((kc + h) * 2 + rank) * kUNP) * kXPieceBytes;
          if (X3_ABL(a) & 64) {
            mbar_arrive(&s.bfull[st + h]);
          } else {
            mbar_expect_tx(&s.bfull[st + h], kUNP * kXPieceBytes);
#pragma unroll
            for (int p = 0; p < kUNP; ++p)
              bulk_g2s(s.ring[st + h].b[p], src + (size_t)p * kXPieceBytes, kXPieceBytes, &s.bfull[st + h]);
          }
        }
      }
      if (!(X3_ABL(a) & 1)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[8];
          const uint32_t bits = h1_chunk<true>(s.w1t, ob, D, stage_kgroup(kc + h, g) * 8, v);  // v = -s_h H1
          if (kc < 4) m0 |= bits << (8 * (kc + h));
          else m1 |= bits << (8 * (kc + h - 4));
          uint8_t* tiles[kUNP];
#pragma unroll
          for (int p = 0; p < kUNP; ++p) tiles[p] = s.ring[st + h].a[p];
          store_split_chunk<kUNP, kUF16>(tiles, (uint32_t)(rloc * 16 + g * 2048), v);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {  // (the bulk copies report through warp 16, see the kernel)
        mbar_arrive_cluster(&s.full[st], 0);
        mbar_arrive_cluster(&s.full[st + 1], 0);
      }
    }
    // H1 mask of columns [64 g, 64 g + 64) of this row
    if (valid) *reinterpret_cast<uint2*>(a.mask1[net] + rowl * 8 + 2 * g) = make_uint2(m0, m1);
  };

  long long ph[5] = {0, 0, 0, 0, 0};  // development: cycles in produce / pass 1 / barrier 1 / loss + barrier 2 / pass 2
  auto epilogue = [&](int64_t tile, int64_t j) {
    const int buf = (int)(j & 1);
    long long pc = X3_CLOCK();
    mbar_wait_cluster(&s.acc_full[buf], (uint32_t)((j >> 1) & 1));
    fence_after_sync();
    const uint32_t acc = tmem + (uint32_t)(buf * H) + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    const int64_t rowl = tile * 256 + rank * 128 + r;  // chunk-local row of this thread's lane
    // loss inputs of row tid (threads 0..127), requested now so that the loads land under pass 1
    float in_act = 0.0f, in_logp = 0.0f, in_tgt = 0.0f;  // action, old log-probability, advantage (policy) / return (value)
    bool in_valid = false;
    if (tid < TILE) {
      const int64_t rl = tile * 256 + rank * 128 + tid;
      int64_t t = 0, n = 0;
      in_valid = rl < a.Mc && minibatch_row_to_tn(a, a.row_off + rl, t, n);
      if (in_valid) {
        const int64_t idx = t * a.N + n;
        if constexpr (POLICY) {
          in_act = a.dist_kind == RL8_DIST_CATEGORICAL ? (float)((const long long*)a.actions)[idx]
                                                       : ((const float*)a.actions)[idx];
          in_logp = a.logp[idx], in_tgt = a.adv[idx];
        } else {
          in_tgt = a.ret[idx];
        }
      }
    }
    // ---- pass 1: H2, its mask, head partial sums
    float dot[PN];
#pragma unroll
    for (int p = 0; p < PN; ++p) dot[p] = 0.0f;
    uint32_t mk[2];
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      const int col0 = cq * 64 + c2 * 32;
      float v[32];
      tmem_ld32(acc + (uint32_t)(c2 * 32), v);
      uint32_t bits = 0u;
      if (X3_ABL(a) & 2) { dot[0] += v[0] + v[31]; mk[c2] = 0; continue; }
#pragma unroll
      for (int jj = 0; jj < 32; jj += 4) {
        // The accumulator holds -s Z2 (the A operand is -s_h H1) and s.b2 holds -b2, so z = -(Z2 + b2): inv_scale is a
        // power of two, the fma rounds once like the un-scaled sum; the ReLU-mask bit is the sign bit of z (one funnel
        // shift per column, first column ends up in bit 31), h = min(z, 0) = -H2 enters the head sums negated.
        const float4 b = *reinterpret_cast<const float4*>(&s.b2[col0 + jj]);
        const float z0 = fmaf(v[jj], inv_scale, b.x), z1 = fmaf(v[jj + 1], inv_scale, b.y);
        const float z2 = fmaf(v[jj + 2], inv_scale, b.z), z3 = fmaf(v[jj + 3], inv_scale, b.w);
        bits = __funnelshift_l(__float_as_uint(z0), bits, 1), bits = __funnelshift_l(__float_as_uint(z1), bits, 1);
        bits = __funnelshift_l(__float_as_uint(z2), bits, 1), bits = __funnelshift_l(__float_as_uint(z3), bits, 1);
        const float h0 = fminf(z0, 0.0f), h1 = fminf(z1, 0.0f), h2 = fminf(z2, 0.0f), h3 = fminf(z3, 0.0f);
#pragma unroll
        for (int p = 0; p < PN; ++p) {
          const float4 w = *reinterpret_cast<const float4*>(&s.w3[p][col0 + jj]);
          dot[p] = fmaf(-h0, w.x, dot[p]);
          dot[p] = fmaf(-h1, w.y, dot[p]);
          dot[p] = fmaf(-h2, w.z, dot[p]);
          dot[p] = fmaf(-h3, w.w, dot[p]);
        }
      }
      mk[c2] = __brev(bits);
    }
    if (rowl < a.Mc) *reinterpret_cast<uint2*>(a.mask2[net] + rowl * 8 + 2 * cq) = make_uint2(mk[0], mk[1]);
#pragma unroll
    for (int p = 0; p < PN; ++p) s.part[cq][r][p] = dot[p];
    ph[1] += X3_CLOCK() - pc, pc = X3_CLOCK();
    // Both barriers of the epilogue only connect the four warps of one row quarter q (column quarters cq = 0..3): the
    // loss thread of row r = 32 q + lane is lane `lane` of warp q, and pass 2 of these warps reads that warp's dOut.
    quarter_bar_sync(q);
    ph[2] += X3_CLOCK() - pc, pc = X3_CLOCK();
    // ---- per-row loss -> dOut (threads 0..127 own row tid)
    if (tid < TILE && !(X3_ABL(a) & 128)) {
      const int64_t rl = tile * 256 + rank * 128 + tid;
      float d_o[kMaxPT] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (in_valid) {
        float o[PN];
#pragma unroll
        for (int p = 0; p < PN; ++p)
          o[p] = ((s.part[0][tid][p] + s.part[1][tid][p]) + (s.part[2][tid][p] + s.part[3][tid][p])) + __ldg(np.b3 + p);
        RowLoss L;
        if constexpr (POLICY) {
          if (continuous) o[1] = tanhf(o[1]);
          ppo_policy_row<PN>(a.dist_kind, o, in_act, in_logp, in_tgt, a.hp, a.inv_denom, d_o, L);
          s.lsum[tid][0] += L.entropy, s.lsum[tid][1] += L.policy, s.lsum[tid][3] += L.kl;
        } else {
          ppo_value_row(o[0], in_tgt, a.hp, a.inv_denom, d_o, L);
          s.lsum[tid][2] += L.vf;
        }
#pragma unroll
        for (int p = 0; p < PN; ++p) s.gb3s[tid][p] += d_o[p];
        if constexpr (kUF16)
          s.dmx[tid] = fmaxf(s.dmx[tid], fmaxf(fmaxf(fabsf(d_o[0]), fabsf(d_o[1])), fmaxf(fabsf(d_o[2]), fabsf(d_o[3]))));
        *reinterpret_cast<float4*>(a.dout[net] + rl * 4) = make_float4(d_o[0], d_o[1], d_o[2], d_o[3]);
      }
      *reinterpret_cast<float4*>(s.dsm[tid]) = make_float4(d_o[0], d_o[1], d_o[2], d_o[3]);
    }
    // ---- pass 2: gW3[p][c] += sum_rows H2[row][c] dOut[row][p].  The accumulator is read a second time, now with
    // the 16x256b shape: thread (g = lane / 4, t = lane % 4) receives rows {g, g + 8, g + 16, g + 24} of the warp's lane
    // quarter and the 8 columns {8 k + 2 t, 8 k + 2 t + 1}, so four of the 32 rows are summed in registers and only
    // the 8 lanes that share t are left to reduce: 7 shuffles per 8 columns and head output (the 32x32b shape, one row
    // per thread, needs 31 per 32).  Lane (g, t) ends up with column col0 + 8 (g / 2) + 2 t + (g & 1).
    const int g4 = lane >> 2, t4 = lane & 3;
    // H2 of the thread's 4 rows x 8 columns of column block c2: hv[i][2 k + e] = row g + 8 i, column col0 + 8 k + 2 t + e
    auto load_h2 = [&](int c2, float (*hv)[8]) {
      const int col0 = cq * 64 + c2 * 32;
      uint32_t zr[2][16];
      tmem_ld_16x256b_x4(acc + (uint32_t)(c2 * 32), zr[0]);
      tmem_ld_16x256b_x4(acc + (16u << 16) + (uint32_t)(c2 * 32), zr[1]);
      tmem_wait_ld();
      if (c2 == 1) {  // last read of the accumulator: the tensor pipe may overwrite it
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&s.acc_empty[buf], 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 b = *reinterpret_cast<const float2*>(&s.b2[col0 + 8 * k + 2 * t4]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          hv[2 * h][2 * k] = fminf(fmaf(__uint_as_float(zr[h][4 * k]), inv_scale, b.x), 0.0f);  // -H2
          hv[2 * h][2 * k + 1] = fminf(fmaf(__uint_as_float(zr[h][4 * k + 1]), inv_scale, b.y), 0.0f);
          hv[2 * h + 1][2 * k] = fminf(fmaf(__uint_as_float(zr[h][4 * k + 2]), inv_scale, b.x), 0.0f);
          hv[2 * h + 1][2 * k + 1] = fminf(fmaf(__uint_as_float(zr[h][4 * k + 3]), inv_scale, b.y), 0.0f);
        }
      }
    };
    float hv[4][8];
    load_h2(0, hv);        // does not need dOut: the 12 warps without loss rows do it under the loss phase
    quarter_bar_sync(q);   // dsm (dOut of the quarter's rows) is complete
    ph[3] += X3_CLOCK() - pc, pc = X3_CLOCK();
    float dr[4][PN];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 dv = *reinterpret_cast<const float4*>(s.dsm[q * 32 + g4 + 8 * i]);
      const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int p = 0; p < PN; ++p) dr[i][p] = d4[p];
    }
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2) {
      if (c2 == 1) load_h2(1, hv);
      if (X3_ABL(a) & 4) { gw3_acc[c2][0] += hv[0][lane & 7]; continue; }
#pragma unroll
      for (int p = 0; p < PN; ++p) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          x[j] = fmaf(hv[3][j], dr[3][p], fmaf(hv[2][j], dr[2][p], fmaf(hv[1][j], dr[1][p], hv[0][j] * dr[0][p])));
        // the 8 lanes that share t (lane bits 2..4) exchange halves: lane (g, t) keeps entry j = g
#pragma unroll
        for (int o = 16, n = 4; o >= 4; o >>= 1, n >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send = up ? x[i] : x[i + n];
            const float keep = up ? x[i + n] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        gw3_acc[c2][p] -= x[0];  // (hv holds -H2)
      }
    }
    ph[4] += X3_CLOCK() - pc;
  };

  // (measured: producing the second half of tile j + 1 between the passes of the epilogue changes nothing)
  if (n_my > 0) produce(pr);
  for (int64_t j = 0; j < n_my; ++j) {
    const long long pc = X3_CLOCK();
    if (j + 1 < n_my) produce(pr + (j + 1) * npairs);
    ph[0] += X3_CLOCK() - pc;
    epilogue(pr + j * npairs, j);
  }
  if (X3_CLOCK() != 0 && a.dbg && pr == 0 && rank == 0 && (warp == 5 || warp == 1) && lane == 0) {
    unsigned long long* d = a.dbg + 16 * net + (warp == 5 ? 11 : 32 + 11);  // warp 5: no loss rows; warp 1: loss rows
#pragma unroll
    for (int i = 0; i < 5; ++i) atomicAdd(d + i, (unsigned long long)ph[i]);
  }
  // ---- flush
  if (n_my > 0) {
#pragma unroll
    for (int c2 = 0; c2 < 2; ++c2)
#pragma unroll
      for (int p = 0; p < PN; ++p)  // lane (g, t) of pass 2 owns column 8 (g / 2) + 2 t + (g & 1) of its 32
        atomicAdd(a.gw3[net] + p * H + cq * 64 + c2 * 32 + 8 * (lane >> 3) + 2 * (lane & 3) + ((lane >> 2) & 1),
                  gw3_acc[c2][p]);
    if (tid < TILE) {
#pragma unroll
      for (int p = 0; p < PN; ++p) {
        const float w = warp_sum(s.gb3s[tid][p]);
        if (lane == 0) atomicAdd(a.gb3[net] + p, w);
      }
      if constexpr (kUF16) {
        float m = s.dmx[tid];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0.0f) atomicMax(&a.sc->dmax[net], __float_as_uint(m));
      }
    }
  }
  if (tid < TILE) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sv[i] = s.lsum[tid][i];
  }
}

template <int P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXThreads, 1)
x3_update_f_kernel(NetParams np_pi, NetParams np_vf, UpdXArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemXF& s = *reinterpret_cast<SmemXF*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, pairs_all = gridDim.x >> 1;
  const int net = pair < a.n_pi ? 0 : 1;
  const int64_t pr = net ? pair - a.n_pi : pair, npairs = net ? pairs_all - a.n_pi : a.n_pi;
  const NetParams np = net ? np_vf : np_pi;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kFStages; ++i) {
      mbar_init(&s.full[i], 33);  // 32 worker warps of the pair + the peer's bulk copy (forwarded by its warp 16)
      mbar_init(&s.bfull[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.acc_full[0], 1), mbar_init(&s.acc_full[1], 1);
    mbar_init(&s.acc_empty[0], 32), mbar_init(&s.acc_empty[1], 32);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc_pair(&s.tmem_base, 512);
  // fp16 pieces: H1 leaves h1_chunk multiplied by s_h (folded into [W1 | b1]), the W2 image carries s_w, and the
  // epilogue multiplies the accumulator by 1 / (s_h s_w) inside its bias add
  float s_h = 1.0f, inv_scale = 1.0f;
  if constexpr (kUF16) {
    s_h = pow2_scale_for(h1_bound(np, __uint_as_float(a.sc->omax), s.red));
    inv_scale = (1.0f / s_h) * (1.0f / a.sc->w2f[net]);
  }
  stage_w1t(s.w1t, np, -s_h);                                              // h1_chunk<true>: A = -s_h H1
  for (int i = tid; i < H; i += blockDim.x) s.b2[i] = 0.0f - np.b2[i];     // -b2, never -0
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  double sv[4] = {0.0, 0.0, 0.0, 0.0};
  if (warp < 16) {
    if (net == 0) update_f_workers<P, true>(s, np, a, 0, pr, npairs, rank, inv_scale, sv);
    else update_f_workers<1, false>(s, np, a, 1, pr, npairs, rank, inv_scale, sv);
  } else if (rank == 0) {
    const int64_t ntiles = (a.Mc + 255) / 256;
    const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
    const uint32_t idesc = upd_idesc(0, 0);
    uint32_t kcount = 0;
    long long w_full[8] = {0, 0, 0, 0, 0, 0, 0, 0}, w_acc = 0;
    const long long t_begin = X3_CLOCK();
    for (int64_t j = 0; j < n_my; ++j) {
      const int buf = (int)(j & 1);
      long long c0 = X3_CLOCK();
      if (j >= 2) mbar_wait_cluster(&s.acc_empty[buf], (uint32_t)(((j >> 1) - 1) & 1));
      w_acc += X3_CLOCK() - c0;
#pragma unroll
      for (int kc = 0; kc < H / kXKc; ++kc, ++kcount) {
        const int st = (int)(kcount % kFStages);
        c0 = X3_CLOCK();
        mbar_wait_cluster(&s.full[st], (kcount / kFStages) & 1);
        mbar_wait(&s.bfull[st], (kcount / kFStages) & 1);  // this CTA's half of the W2 stage has landed
        w_full[kc] += X3_CLOCK() - c0;
        fence_after_sync();
        if (elect_one()) {
          issue_stage<kUNP>(tmem + (uint32_t)(buf * H), s.ring[st], idesc, kc > 0);
          mma_commit_pair(&s.empty[st]);
          if (kc == H / kXKc - 1) mma_commit_pair(&s.acc_full[buf]);
        }
        __syncwarp();
      }
    }
    if (X3_ABL(a) >= 0 && X3_CLOCK() != 0 && a.dbg && pr == 0 && (tid & 31) == 0) {
      unsigned long long* d = a.dbg + 16 * net;
      for (int i = 0; i < 8; ++i) atomicAdd(d + i, (unsigned long long)w_full[i]);
      atomicAdd(d + 8, (unsigned long long)w_acc);
      atomicAdd(d + 9, (unsigned long long)(X3_CLOCK() - t_begin));
      atomicAdd(d + 10, (unsigned long long)n_my);
    }
  }
  else {
    // peer CTA, warp 16: tells the leader's full[] barrier when THIS CTA's half of a W2 stage has landed, so that no
    // worker warp waits for a bulk copy
    const int64_t ntiles = (a.Mc + 255) / 256;
    const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
    const uint32_t total = (uint32_t)(n_my * (H / kXKc));
    for (uint32_t kcount = 0; kcount < total; ++kcount) {
      const int st = (int)(kcount % kFStages);
      mbar_wait(&s.bfull[st], (kcount / kFStages) & 1);
      if ((tid & 31) == 0) mbar_arrive_cluster(&s.full[st], 0);
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 16) tmem_dealloc_pair(tmem, 512);
  __shared__ double red[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double t = block_sum(sv[i], red);
    if (tid == 0 && t != 0.0) atomicAdd(a.sums + i, t);
  }
  if (blockIdx.x == 0 && tid == 0) atomicAdd(a.sums + 4, (double)a.Mc);
}

// dZ2[row][j0 .. j0 + 8) = [H2 > 0] .* (dOut W3) from the scratch values of the row
template <int PN>
__device__ __forceinline__ void dz2_chunk(const float (*w3)[H], const float* d, uint32_t mask_byte, int j0, float* v) {
  float2 g[4];
#pragma unroll
  for (int p = 0; p < PN; ++p) {
    const float4 w0 = *reinterpret_cast<const float4*>(&w3[p][j0]);
    const float4 w1 = *reinterpret_cast<const float4*>(&w3[p][j0 + 4]);
    const float2 dp = make_float2(d[p], d[p]);
    if (p == 0) {  // fma(d, w, 0) == d * w
      g[0] = fmul2(dp, make_float2(w0.x, w0.y)), g[1] = fmul2(dp, make_float2(w0.z, w0.w));
      g[2] = fmul2(dp, make_float2(w1.x, w1.y)), g[3] = fmul2(dp, make_float2(w1.z, w1.w));
    } else {
      g[0] = ffma2(dp, make_float2(w0.x, w0.y), g[0]), g[1] = ffma2(dp, make_float2(w0.z, w0.w), g[1]);
      g[2] = ffma2(dp, make_float2(w1.x, w1.y), g[2]), g[3] = ffma2(dp, make_float2(w1.z, w1.w), g[3]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = (mask_byte >> (2 * e)) & 1u ? g[e].x : 0.0f;
    v[2 * e + 1] = (mask_byte >> (2 * e + 1)) & 1u ? g[e].y : 0.0f;
  }
}

// ---- input-gradient kernel ---------------------------------------------------------------------------------------------
// dH1^T [input unit i][row] = sum_j W2[j][i] dZ2[row][j]:  M = i (the pair's 256: CTA c owns the half i / 128 = c),
// N = the 256 rows of a tile (CTA c stages ITS 128 rows as the B half), K = j in 8 stages of 32.
//   A stage: W2^T piece image (bulk copy, two pieces);  B stage: dZ2 chunks computed by the workers from dOut / mask2
//   epilogue: lane = input unit i, columns = tile rows: dZ1 = [H1 > 0] .* dH1 (mask1 bit i of the row),
//             gb1[i] += dZ1, gW1[i][d] += dZ1 obs[row][d] -- per-thread FMAs, the row's obs / mask word are broadcasts
template <int NPB>
struct SmemXB {
  static constexpr int kStages = NPB == 2 ? 6 : 4;
  StageX<NPB> ring[kStages];  // 196608
  float w3[kMaxPT][H];       //   4096
  float os[2][256][8];       //  16384  observations of the tile's rows (zero padded), per accumulator buffer
  uint32_t m1s[2][256][4];   //   8192  mask1 words of this CTA's input half, per tile row
  float red[32];
  uint2 lut[16];             //    128  mask nibble -> four 16-bit {0, 1}
  uint64_t full[kStages], bfull[kStages], empty[kStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemXB<2>) <= 227 * 1024 && sizeof(SmemXB<3>) <= 227 * 1024, "SmemXB exceeds the 227 KB CTA limit");
constexpr int kNPBMax = kUF16 ? 2 : 3;  // (fp16 pieces come in twos)

// VNET (the value network, one head output): dZ2[row][j] = dOut[row] * (mask2[row][j] W3[j]), so
//     dH1[row][i] = dOut[row] * sum_j (W3[j] W2[j][i]) mask2[row][j]
// -- the A image is W2^T with W3 folded in (pack_w2_pieces_kernel's kscale), the B operand is the MASK itself, exact
// in one bf16 piece (three piece products instead of six, and nothing to split), and dOut[row] multiplies the
// accumulator column in the epilogue (it travels in slot 7 of the row's staged observations).
// fp16 pieces: s_d scales dOut (policy network: |dZ2| <= max |dOut| max_j sum_p |W3[p][j]|), the W2^T image carries its
// own scale, the mask is 0 / 1; the gradient sums stay scaled until inv_scale multiplies them in the final atomics.
template <int PN, int NPB, bool VNET>
__device__ __forceinline__ void update_b_workers(SmemXB<NPB>& s, const NetParams& np, const UpdXArgs& a, int net,
                                                 int64_t pr, int64_t npairs, uint32_t rank, float s_d, float inv_scale) {
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rloc = tid & 127, g = tid >> 7;
  const int q = warp & 3, cq = warp >> 2;
  const int D = np.D;
  const int64_t ntiles = (a.Mc + 255) / 256;
  const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
  // gradient sums of input unit i as packed pairs: (gW1[i][0], [1]), ([2], [3]), ([4], [5]), ([6], gb1[i])
  float2 gacc[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) gacc[d] = make_float2(0.0f, 0.0f);
  uint32_t kcount = 0;
  constexpr int kStages = SmemXB<NPB>::kStages;
  constexpr int kTileStages = H / kXKc;
  // inputs of this thread's dZ2 chunks of the tile being produced (its row: dOut, the mask2 words of columns [64 g, +64))
  float d4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  uint32_t m0 = 0u, m1 = 0u;

  // The inputs of a tile are requested one tile ahead (fetch) and handed over when the tile starts (begin_tile):
  // the global-load latency hides under the production and epilogue of the tile in between.  fetch() only issues
  // loads -- nothing in it reads a loaded value.
  float pf_o[4] = {0.0f, 0.0f, 0.0f, 0.0f}, pf_dr = 0.0f;  // obs slots 4 half .. 4 half + 3 of row rt; its dOut (VNET)
  float4 pf_d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  uint4 pf_m1 = make_uint4(0u, 0u, 0u, 0u);
  uint2 pf_m2 = make_uint2(0u, 0u);
  auto fetch = [&](int64_t tile) {
    {
      // row of the tile and slots 4 half .. 4 half + 3: the four warps that share cq stage exactly the 64 rows their
      // epilogues read, so the hand-over needs a barrier of those four warps only
      const int rt = cq * 64 + q * 16 + (lane & 15), half = lane >> 4;
      const int64_t rl = tile * 256 + rt;
      int64_t t = 0, n = 0;
      const bool valid = rl < a.Mc && minibatch_row_to_tn(a, a.row_off + rl, t, n);
#pragma unroll
      for (int k = 0; k < 4; ++k) pf_o[k] = 0.0f;
      pf_dr = 0.0f, pf_m1 = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        const float* base = a.obs + t * (int64_t)D * a.N + n;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * half + k < D) pf_o[k] = __ldg(base + (int64_t)(4 * half + k) * a.N);
        if constexpr (VNET) pf_dr = a.dout[net][rl * 4];
        if (half == 0) pf_m1 = *reinterpret_cast<const uint4*>(a.mask1[net] + rl * 8 + 4 * rank);
      }
    }
    const int64_t rowl = tile * 256 + rank * 128 + rloc;
    pf_d = make_float4(0.0f, 0.0f, 0.0f, 0.0f), pf_m2 = make_uint2(0u, 0u);
    if (rowl < a.Mc) {
      if (!VNET) pf_d = *reinterpret_cast<const float4*>(a.dout[net] + rowl * 4);
      pf_m2 = *reinterpret_cast<const uint2*>(a.mask2[net] + rowl * 8 + 2 * g);
    }
  };
  auto begin_tile = [&](int64_t j) {
    const int buf = (int)(j & 1);
    // epilogue inputs of the tile: [obs | bias factor] and this CTA's mask1 words of all 256 rows.  Slot 7 multiplies
    // the bias gradient: 1, or -- value network -- dOut of the row, which then scales the observations too
    // (dZ1 = dOut * [H1 > 0] .* accumulator: the factor moves into the row's [obs | 1]).
    const int rt = cq * 64 + q * 16 + (lane & 15), half = lane >> 4;
    float4 o4 = make_float4(pf_o[0], pf_o[1], pf_o[2], pf_o[3]);
    if constexpr (VNET) o4.x *= pf_dr, o4.y *= pf_dr, o4.z *= pf_dr, o4.w *= pf_dr;
    if (half) o4.w = VNET ? pf_dr : 1.0f;
    *reinterpret_cast<float4*>(&s.os[buf][rt][4 * half]) = o4;
    if (half == 0) *reinterpret_cast<uint4*>(s.m1s[buf][rt]) = pf_m1;
    d4[0] = pf_d.x * s_d, d4[1] = pf_d.y * s_d, d4[2] = pf_d.z * s_d, d4[3] = pf_d.w * s_d;
    m0 = pf_m2.x, m1 = pf_m2.y;
  };
  // the 8 ring stages of a tile, two per generic -> async proxy fence (kStages is even: a pair never wraps around)
  static_assert(kStages % 2 == 0 && kTileStages % 2 == 0, "stage pairs");
  auto produce = [&]() {
    for (int kc = 0; kc < kTileStages; kc += 2, kcount += 2) {
      const int st = (int)(kcount % kStages);
      const uint32_t use = kcount / kStages;
      if (use > 0) {
        mbar_wait_cluster(&s.empty[st], (use - 1) & 1);
        mbar_wait_cluster(&s.empty[st + 1], (use - 1) & 1);
      }
      __syncwarp();
      if (warp == 0 && elect_one()) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint8_t* src = np.w2_img + (size_t)(((kc + h) * 2 + rank) * NPB) * kXPieceBytes;
          mbar_expect_tx(&s.bfull[st + h], NPB * kXPieceBytes);
#pragma unroll
          for (int p = 0; p < NPB; ++p)
            bulk_g2s(s.ring[st + h].a[p], src + (size_t)p * kXPieceBytes, kXPieceBytes, &s.bfull[st + h]);
        }
      }
      if (!(X3_ABL(a) & 8)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t byte = (kc < 4 ? m0 >> (8 * (kc + h)) : m1 >> (8 * (kc + h - 4))) & 0xffu;
          if constexpr (VNET) {
            *reinterpret_cast<uint4*>(s.ring[st + h].b[0] + rloc * 16 + g * 2048) = mask_byte_lut(s.lut, byte);
          } else {
            float v[8];
            dz2_chunk<PN>(s.w3, d4, byte, stage_kgroup(kc + h, g) * 8, v);
            uint8_t* tiles[NPB];
#pragma unroll
            for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st + h].b[p];
            store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(rloc * 16 + g * 2048), v);
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {  // (the bulk copies report through warp 16, see the kernel)
        mbar_arrive_cluster(&s.full[st], 0);
        mbar_arrive_cluster(&s.full[st + 1], 0);
      }
    }
  };

  auto epilogue = [&](int64_t j) {
    const int buf = (int)(j & 1);
    colq_bar_sync(cq);  // os / m1s of the 64 rows of this column quarter are complete
    mbar_wait_cluster(&s.acc_full[buf], (uint32_t)((j >> 1) & 1));
    fence_after_sync();
    const uint32_t acc = tmem + (uint32_t)(buf * H) + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    // mask1 bits of THIS thread's input unit for the warp's 64 rows, as two words (bit r: row 64 cq + 32 h + r): lane l
    // loads the word of row l and the warp transposes the 32 x 32 bit matrix with five shuffles -- instead of one
    // broadcast shared-memory load per row and thread (the epilogue is bound by its shared-memory instructions)
    uint32_t mbits[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t x = s.m1s[buf][cq * 64 + 32 * h + lane][q];
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        const uint32_t keep = sft == 16 ? 0x0000ffffu : sft == 8 ? 0x00ff00ffu : sft == 4 ? 0x0f0f0f0fu
                              : sft == 2 ? 0x33333333u : 0x55555555u;  // bit positions whose index has bit `sft` clear
        const uint32_t other = __shfl_xor_sync(0xffffffffu, x, sft);
        x = (lane & sft) ? (((other >> sft) & keep) | (x & ~keep)) : ((x & keep) | ((other & keep) << sft));
      }
      mbits[h] = x;
    }
#pragma unroll 1
    for (int c4 = 0; c4 < 4; ++c4) {
      float v[16];
      tmem_ld16_nowait(acc + (uint32_t)(c4 * 16), v);
      tmem_wait_ld();
      reg_fence16f(v);
      const uint32_t hw = mbits[c4 >> 1] >> (16 * (c4 & 1));  // bit e: row 64 cq + 16 c4 + e
      if (c4 == 3) {  // the accumulator has been read: the tensor pipe may reuse it for tile j + 2
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&s.acc_empty[buf], 0);
      }
      if (X3_ABL(a) & 16) { gacc[3].y += v[0] + v[15]; continue; }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int rt = cq * 64 + c4 * 16 + e;
        const float4 oa = *reinterpret_cast<const float4*>(&s.os[buf][rt][0]);
        const float4 ob = *reinterpret_cast<const float4*>(&s.os[buf][rt][4]);  // (obs 4..6, bias factor)
        const float dz1 = (hw >> e) & 1u ? v[e] : 0.0f;
        const float2 dz = make_float2(dz1, dz1);
        gacc[0] = ffma2(dz, make_float2(oa.x, oa.y), gacc[0]), gacc[1] = ffma2(dz, make_float2(oa.z, oa.w), gacc[1]);
        gacc[2] = ffma2(dz, make_float2(ob.x, ob.y), gacc[2]), gacc[3] = ffma2(dz, make_float2(ob.z, ob.w), gacc[3]);
      }
    }
    colq_bar_sync(cq);  // os / m1s of this buffer may be rewritten
  };

  // Two 256-column accumulators (one per tile parity, like the forward kernel: K = 256 is 96 instructions into one
  // accumulator, whose per-instruction truncation stays below 1e-5 of a gradient): the tensor pipe works on tile
  // j + 1 while the workers run the epilogue of tile j.
  if (n_my > 0) {
    fetch(pr);
    begin_tile(0);
    if (n_my > 1) fetch(pr + npairs);
    produce();
  }
  for (int64_t j = 0; j < n_my; ++j) {
    if (j + 1 < n_my) {
      begin_tile(j + 1);
      if (j + 2 < n_my) fetch(pr + (j + 2) * npairs);
      produce();
    }
    epilogue(j);
  }
  if (n_my > 0) {
    const int i = 128 * (int)rank + q * 32 + lane;
    const float gw1_acc[8] = {gacc[0].x, gacc[0].y, gacc[1].x, gacc[1].y, gacc[2].x, gacc[2].y, gacc[3].x, gacc[3].y};
#pragma unroll
    for (int d = 0; d < 7; ++d)
      if (d < D) atomicAdd(a.gw1[net] + i * D + d, gw1_acc[d] * inv_scale);
    atomicAdd(a.gb1[net] + i, gw1_acc[7] * inv_scale);
  }
}

template <int P, int NPB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXThreads, 1)
x3_update_b_kernel(NetParams np_pi, NetParams np_vf, UpdXArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemXB<NPB>& s = *reinterpret_cast<SmemXB<NPB>*>(smem_raw);
  constexpr int kStages = SmemXB<NPB>::kStages;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, pairs_all = gridDim.x >> 1;
  const int n_pi = a.n_pi_b;
  const int net = pair < n_pi ? 0 : 1;
  const int64_t pr = net ? pair - n_pi : pair, npairs = net ? pairs_all - n_pi : n_pi;
  const NetParams np = net ? np_vf : np_pi;  // w2_img: the TRANSPOSED NPB-piece image (value network: W3 folded in)
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&s.full[i], 33);  // 32 worker warps of the pair + the peer's bulk copy (forwarded by its warp 16)
      mbar_init(&s.bfull[i], 1);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.acc_full[0], 1), mbar_init(&s.acc_full[1], 1);
    mbar_init(&s.acc_empty[0], 32), mbar_init(&s.acc_empty[1], 32);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc_pair(&s.tmem_base, 512);
  float s_d = 1.0f, inv_scale = 1.0f;
  if constexpr (kUF16) {
    if (net == 0) s_d = pow2_scale_for(__uint_as_float(a.sc->dmax[0]) * w3_colsum_bound(np, s.red));
    inv_scale = (1.0f / s_d) * (1.0f / a.sc->w2b[net]);
  }
  fill_mask_lut<kUF16>(s.lut);
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  if (warp < 16) {
    if (net == 0) update_b_workers<P, NPB, false>(s, np, a, 0, pr, npairs, rank, s_d, inv_scale);
    else update_b_workers<1, NPB, true>(s, np, a, 1, pr, npairs, rank, s_d, inv_scale);
  } else if (rank == 0) {
    const int64_t ntiles = (a.Mc + 255) / 256;
    const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
    const uint32_t idesc = upd_idesc(0, 0);
    uint32_t kcount = 0;
    for (int64_t j = 0; j < n_my; ++j) {
      const int buf = (int)(j & 1);
      if (j >= 2) mbar_wait_cluster(&s.acc_empty[buf], (uint32_t)(((j >> 1) - 1) & 1));
      for (int kc = 0; kc < H / kXKc; ++kc, ++kcount) {
        const int st = (int)(kcount % kStages);
        mbar_wait_cluster(&s.full[st], (kcount / kStages) & 1);
        mbar_wait(&s.bfull[st], (kcount / kStages) & 1);  // this CTA's half of the W2^T stage has landed
        fence_after_sync();
        if (elect_one()) {
          const uint32_t acc = tmem + (uint32_t)(buf * H);
          if (net == 0) issue_stage<NPB>(acc, s.ring[st], idesc, kc > 0);
          else issue_stage_mask_b<NPB>(acc, acc, s.ring[st], idesc, kc > 0);
          mma_commit_pair(&s.empty[st]);
          if (kc == H / kXKc - 1) mma_commit_pair(&s.acc_full[buf]);
        }
        __syncwarp();
      }
    }
  }
  else {
    // peer CTA, warp 16: forwards the landing of this CTA's W2^T stages to the leader's full[] barriers
    const int64_t ntiles = (a.Mc + 255) / 256;
    const int64_t n_my = pr < ntiles ? (ntiles - pr + npairs - 1) / npairs : 0;
    const uint32_t total = (uint32_t)(n_my * (H / kXKc));
    for (uint32_t kcount = 0; kcount < total; ++kcount) {
      const int st = (int)(kcount % kStages);
      mbar_wait(&s.bfull[st], (kcount / kStages) & 1);
      if ((tid & 31) == 0) mbar_arrive_cluster(&s.full[st], 0);
      __syncwarp();
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 16) tmem_dealloc_pair(tmem, 512);
}

// ---- weight-gradient kernel ------------------------------------------------------------------------------------------
// gW2[j][i] += sum_rows dZ2[row][j] H1[row][i].  The contraction runs over rows: a ring stage is 32 rows, both operands
// are MN-major tiles [32 K rows][128 M / N columns] per piece (off(r, g) = r * 16 + g * 512; LBO 128, SBO 512), CTA c
// of the pair produces the dZ2 columns j and the H1 columns i of ITS half, so nothing is computed twice.
//   worker warp w (0..15): the dZ2^T chunk and the H1 chunk of column group w (lane = row of the stage), gb2 on the way
// Pairs [0, n_pi) work on the policy network, the rest on the value network; a pair owns the stages
// pr, pr + npairs, ... of its network and keeps its 256 x 256 accumulator in tensor memory until the end.
// Inputs of the 32 rows of one stage (lane = row), fetched ONCE per CTA by a loader warp (warps 17, 18): the 16 worker
// warps all need the same rows, and the buffer coordinates of a row (a division) and its D strided observation loads
// were a quarter of the kernel's instructions when every warp fetched its own copy.
struct WInSlot {
  float4 o0[32];  // obs[0..3]
  float4 o1[32];  // obs[4..6], dOut[0] (value network)
  float4 dd[32];  // policy network: s_d * dOut[0..3]
  uint4 mm[32];   // mask2 words of this CTA's 128-unit half
};
constexpr int kWIn = 6;
constexpr int kWThreads = kXThreads + 64;  // + two loader warps
template <int NPB>
struct SmemXW {
  static constexpr int kStages = NPB == 2 ? 6 : 4;
  StageX<NPB> ring[kStages];  // 196608
  float w1t[8][H];           //   8192
  float w3[kMaxPT][H];       //   4096
  float red[32];
  uint2 lut[16];             //    128  mask nibble -> four 16-bit {0, 1}
  WInSlot in[kWIn];          //  12288  inputs of the next stages' rows, fetched by the loader warp
  uint64_t full[kStages], empty[kStages], flush_full, flush_empty, in_full[kWIn], in_empty[kWIn];
  uint32_t tmem_base;
};
// The tensor pipe truncates its fp32 accumulator after every instruction (measured: the error of a long
// accumulation grows linearly with the number of instructions, tools/check_split_accuracy.py), and this kernel
// accumulates over ALL rows of a pair.  So (1) the leading products a0 b0 go to one accumulator and the five small
// correction products to another (256 columns each; the corrections are 2^-8 of the result, their truncation does
// not matter), and (2) every kFlushStages stages both are added in fp32 (round to nearest) into gW2 and restarted:
// the leading accumulator sees at most 2 * kFlushStages truncations.
constexpr int kFlushStages = 128;
static_assert(sizeof(SmemXW<2>) <= 227 * 1024 && sizeof(SmemXW<3>) <= 227 * 1024, "SmemXW exceeds the 227 KB CTA limit");

// fp16 pieces: s_d scales dOut (as in the input-gradient kernel), [W1 | b1] in shared memory carries the H1 scale;
// inv_scale = 1 / (s_d s_h) multiplies the accumulators in the flush, 1 / s_d the column sums (gb2).
template <int PN, int NPB>
__device__ __forceinline__ void update_w_workers(SmemXW<NPB>& s, const NetParams& np, const UpdXArgs& a, int net,
                                                 int64_t pr, int64_t npairs, uint32_t rank, float s_d,
                                                 float inv_scale) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = np.D;
  const int64_t nstages = (a.Mc + kXKc - 1) / kXKc;
  const int64_t n_my = pr < nstages ? (nstages - pr + npairs - 1) / npairs : 0;
  constexpr int kWStages = SmemXW<NPB>::kStages;
  // Every worker warp w produces column group w of BOTH operands for the 32 rows of a stage (lane = row): one dZ2^T
  // chunk and one H1 chunk -- the same work in every warp (with 8 dZ2 warps and 8 H1 warps the lighter role spent a
  // fifth of its time waiting for ring slots).
  const int g = warp;
  float gb2_acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) gb2_acc[e] = 0.0f;
  // inputs of this lane's row of a stage (from the loader warp's slot): s_d dOut[4], the mask word of unit group
  // g / 4, obs[7]
  auto read_inputs = [&](int64_t k, float* d4, uint32_t& m, float* ob) {
    const int slot = (int)(k % kWIn);
    mbar_wait(&s.in_full[slot], (uint32_t)((k / kWIn) & 1));
    const WInSlot& in = s.in[slot];
    const float4 d = in.dd[lane];
    m = reinterpret_cast<const uint32_t*>(&in.mm[lane])[g >> 2];
    const float4 o0 = in.o0[lane], o1 = in.o1[lane];
    d4[0] = d.x, d4[1] = d.y, d4[2] = d.z, d4[3] = d.w;
    ob[0] = o0.x, ob[1] = o0.y, ob[2] = o0.z, ob[3] = o0.w, ob[4] = o1.x, ob[5] = o1.y, ob[6] = o1.z;
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.in_empty[slot]);
  };
  // accumulator flush f covers the stages [f F, min((f + 1) F, n_my)); the workers run it once they have produced the
  // kWStages - 1 stages after its last one (what the ring holds without the tensor pipe moving on)
  const int64_t nflush = (n_my + kFlushStages - 1) / kFlushStages;
  int64_t next_flush = 0;
  auto flush = [&](int64_t f) {
    mbar_wait_cluster(&s.flush_full, (uint32_t)(f & 1));
    fence_after_sync();
    const int q = warp & 3, cq = warp >> 2;
    const int j = 128 * (int)rank + q * 32 + lane;  // lane = unit j of this CTA's half, 256 columns i
    float* dst = a.gw2[net] + (int64_t)j * H + cq * 64;
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      float v[16], c[16];
      const uint32_t at = s.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64 + h * 16);
      tmem_ld16_nowait(at, v);
      tmem_ld16_nowait(at + H, c);
      tmem_wait_ld();
      reg_fence16f(v);
      reg_fence16f(c);
      if (h == 3) {
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&s.flush_empty, 0);
      }
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        red_add_v4(dst + h * 16 + e, inv_scale * (v[e] + c[e]), inv_scale * (v[e + 1] + c[e + 1]),
                   inv_scale * (v[e + 2] + c[e + 2]), inv_scale * (v[e + 3] + c[e + 3]));
    }
  };
  // Two stages per generic -> async proxy fence: the fence (and the warp barrier behind it) is a serial latency of
  // every warp and stage -- the stage period of this kernel did not depend on the amount of producer work.  Stage
  // pairs start at even k (kWStages and kFlushStages are even): a pair never wraps the ring or straddles a flush.
  static_assert(kWStages % 2 == 0 && kFlushStages % 2 == 0, "stage pairs");
  for (int64_t k0 = 0; k0 < n_my; k0 += 2) {
    const int cnt = k0 + 1 < n_my ? 2 : 1;
    const int st0 = (int)(k0 % kWStages);
    const uint32_t use = (uint32_t)(k0 / kWStages);
    if (cnt == 2) {
      // both stages of the pair at once: [W1 | b1] of input group g is read from shared memory once for the two rows
      float d0[4], d1[4], o0[7], o1[7];
      uint32_t m0, m1;
      read_inputs(k0, d0, m0, o0);
      read_inputs(k0 + 1, d1, m1, o1);
      if (use > 0) {
        mbar_wait_cluster_sleep(&s.empty[st0], (use - 1) & 1);
        mbar_wait_cluster_sleep(&s.empty[st0 + 1], (use - 1) & 1);
      }
      if (!(X3_ABL(a) & 32)) {
        float v[8], w[8];
        uint8_t* tiles[NPB];
        // A: dZ2^T of unit group g
        dz2_chunk<PN>(s.w3, d0, (m0 >> (8 * (g & 3))) & 0xffu, 128 * (int)rank + 8 * g, v);
        dz2_chunk<PN>(s.w3, d1, (m1 >> (8 * (g & 3))) & 0xffu, 128 * (int)rank + 8 * g, w);
#pragma unroll
        for (int e = 0; e < 8; ++e) gb2_acc[e] += v[e] + w[e];
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0].a[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), v);
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0 + 1].a[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), w);
        // B: H1 of input group g  (-s_h H1: inv_scale is negative)
        uint32_t bits0, bits1;
        h1_chunk2<true>(s.w1t, obs_pairs(o0), obs_pairs(o1), D, 128 * (int)rank + 8 * g, v, w, bits0, bits1);
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0].b[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), v);
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0 + 1].b[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), w);
      }
    } else {
      float cur_d[4], cur_o[7];
      uint32_t cur_m;
      read_inputs(k0, cur_d, cur_m, cur_o);
      if (use > 0) mbar_wait_cluster_sleep(&s.empty[st0], (use - 1) & 1);
      if (!(X3_ABL(a) & 32)) {
        float v[8];
        uint8_t* tiles[NPB];
        dz2_chunk<PN>(s.w3, cur_d, (cur_m >> (8 * (g & 3))) & 0xffu, 128 * (int)rank + 8 * g, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) gb2_acc[e] += v[e];
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0].a[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), v);
        h1_chunk<true>(s.w1t, obs_pairs(cur_o), D, 128 * (int)rank + 8 * g, v);
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st0].b[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), v);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_cluster(&s.full[st0], 0);
      if (cnt == 2) mbar_arrive_cluster(&s.full[st0 + 1], 0);
    }
    const int64_t k = k0 + cnt - 1;  // the last stage produced
    while (next_flush < nflush) {
      const int64_t last = (next_flush + 1) * kFlushStages < n_my ? (next_flush + 1) * kFlushStages - 1 : n_my - 1;
      const int64_t due = last + kWStages - 1 < n_my - 1 ? last + kWStages - 1 : n_my - 1;
      if (k < due) break;
      flush(next_flush++);
    }
  }
  if (n_my > 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float w = warp_sum(gb2_acc[e]);
      if (lane == 0) atomicAdd(a.gb2[net] + 128 * rank + 8 * g + e, w * (1.0f / s_d));
    }
  }
}

// The value network (one head output): dZ2[row][j] = dOut[row] mask2[row][j] W3[j], so
//     gW2[j][i] = W3[j] * sum_rows mask2[row][j] * (dOut[row] H1[row][i]),   gb2[j] = W3[j] * sum_rows dOut[row] mask2[row][j]
// -- the A operand is the mask itself (exact in one bf16 piece: NPB piece products instead of NPB (NPB + 1) / 2, one
// 16-byte store per chunk instead of a three-way split), dOut scales the H1 row before it is split, and W3[j] multiplies
// the accumulator row in the flush.  Every worker warp w produces column group w of BOTH operands for the 32 rows of a
// stage (lane = row).
// fp16 pieces: s_d here scales dOut to at most 1, so that dOut s_d (H1 s_h) stays within the H1 bound; inv_scale =
// 1 / (s_d s_h); gb2 sums the unscaled dOut.
template <int NPB>
__device__ __forceinline__ void update_w_workers_vnet(SmemXW<NPB>& s, const NetParams& np, const UpdXArgs& a,
                                                      int64_t pr, int64_t npairs, uint32_t rank, float s_d,
                                                      float inv_scale) {
  constexpr int net = 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = np.D;
  const int64_t nstages = (a.Mc + kXKc - 1) / kXKc;
  const int64_t n_my = pr < nstages ? (nstages - pr + npairs - 1) / npairs : 0;
  constexpr int kWStages = SmemXW<NPB>::kStages;
  const int g = warp;  // column group of this CTA's 128-column half, both operands
  float gb2_acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) gb2_acc[e] = 0.0f;
  // inputs of this lane's row of a stage (from the loader warp's slot): obs[0..6], dOut in slot 7, the mask word of
  // unit group g / 4
  auto read_inputs = [&](int64_t k, float* f, uint32_t& m) {
    const int slot = (int)(k % kWIn);
    mbar_wait(&s.in_full[slot], (uint32_t)((k / kWIn) & 1));
    const WInSlot& in = s.in[slot];
    const float4 o0 = in.o0[lane], o1 = in.o1[lane];
    f[0] = o0.x, f[1] = o0.y, f[2] = o0.z, f[3] = o0.w, f[4] = o1.x, f[5] = o1.y, f[6] = o1.z, f[7] = o1.w;
    m = reinterpret_cast<const uint32_t*>(&in.mm[lane])[g >> 2];
    __syncwarp();
    if (lane == 0) mbar_arrive(&s.in_empty[slot]);
  };
  const int64_t nflush = (n_my + kFlushStages - 1) / kFlushStages;
  int64_t next_flush = 0;
  auto flush = [&](int64_t f) {
    mbar_wait_cluster(&s.flush_full, (uint32_t)(f & 1));
    fence_after_sync();
    const int q = warp & 3, cq = warp >> 2;
    const int j = 128 * (int)rank + q * 32 + lane;  // lane = unit j of this CTA's half, 256 columns i
    const float w3j = s.w3[0][j] * inv_scale;  // (a power of two: exact)
    float* dst = a.gw2[net] + (int64_t)j * H + cq * 64;
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      float v[16], c[16];
      const uint32_t at = s.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64 + h * 16);
      tmem_ld16_nowait(at, v);
      tmem_ld16_nowait(at + H, c);
      tmem_wait_ld();
      reg_fence16f(v);
      reg_fence16f(c);
      if (h == 3) {
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&s.flush_empty, 0);
      }
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        red_add_v4(dst + h * 16 + e, w3j * (v[e] + c[e]), w3j * (v[e + 1] + c[e + 1]), w3j * (v[e + 2] + c[e + 2]),
                   w3j * (v[e + 3] + c[e + 3]));
    }
  };
  for (int64_t k0 = 0; k0 < n_my; k0 += 2) {  // (stage pairs, see update_w_workers)
    const int cnt = k0 + 1 < n_my ? 2 : 1;
    const int st0 = (int)(k0 % kWStages);
    const uint32_t use = (uint32_t)(k0 / kWStages);
    float f0[8], f1[8];
    uint32_t m0, m1 = 0u;
    read_inputs(k0, f0, m0);
    if (cnt == 2) read_inputs(k0 + 1, f1, m1);
    if (use > 0) {
      mbar_wait_cluster_sleep(&s.empty[st0], (use - 1) & 1);
      if (cnt == 2) mbar_wait_cluster_sleep(&s.empty[st0 + 1], (use - 1) & 1);
    }
    if (!(X3_ABL(a) & 32)) {
      // A: the mask bits of unit group g  (tile a[0] only); gb2 on the way
      auto mask_chunk = [&](int st, uint32_t m, float d) {
        const uint32_t byte = (m >> (8 * (g & 3))) & 0xffu;
        *reinterpret_cast<uint4*>(s.ring[st].a[0] + lane * 16 + g * 512) = mask_byte_lut(s.lut, byte);
#pragma unroll
        for (int e = 0; e < 8; ++e) gb2_acc[e] += (byte >> e) & 1u ? d : 0.0f;
      };
      // B: dOut[row] * H1[row][input group g], split  (-s_h H1: inv_scale is negative)
      auto store_b = [&](int st, float* v, float ds) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= ds;
        uint8_t* tiles[NPB];
#pragma unroll
        for (int p = 0; p < NPB; ++p) tiles[p] = s.ring[st].b[p];
        store_split_chunk<NPB, kUF16>(tiles, (uint32_t)(lane * 16 + g * 512), v);
      };
      float v[8], w[8];
      mask_chunk(st0, m0, f0[7]);
      if (cnt == 2) {
        // both stages of the pair at once: [W1 | b1] of input group g is read once for the two rows
        mask_chunk(st0 + 1, m1, f1[7]);
        uint32_t bits0, bits1;
        h1_chunk2<true>(s.w1t, obs_pairs(f0), obs_pairs(f1), D, 128 * (int)rank + 8 * g, v, w, bits0, bits1);
        store_b(st0, v, f0[7] * s_d);
        store_b(st0 + 1, w, f1[7] * s_d);
      } else {
        h1_chunk<true>(s.w1t, obs_pairs(f0), D, 128 * (int)rank + 8 * g, v);
        store_b(st0, v, f0[7] * s_d);
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_cluster(&s.full[st0], 0);
      if (cnt == 2) mbar_arrive_cluster(&s.full[st0 + 1], 0);
    }
    const int64_t k = k0 + cnt - 1;  // the last stage produced
    while (next_flush < nflush) {
      const int64_t last = (next_flush + 1) * kFlushStages < n_my ? (next_flush + 1) * kFlushStages - 1 : n_my - 1;
      const int64_t due = last + kWStages - 1 < n_my - 1 ? last + kWStages - 1 : n_my - 1;
      if (k < due) break;
      flush(next_flush++);
    }
  }
  if (n_my > 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float w = warp_sum(gb2_acc[e]);
      const int j = 128 * (int)rank + 8 * g + e;
      if (lane == 0) atomicAdd(a.gb2[net] + j, s.w3[0][j] * w);
    }
  }
}

template <int P, int NPB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWThreads, 1)
x3_update_w_kernel(NetParams np_pi, NetParams np_vf, UpdXArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemXW<NPB>& s = *reinterpret_cast<SmemXW<NPB>*>(smem_raw);
  constexpr int kWStages = SmemXW<NPB>::kStages;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, pairs_all = gridDim.x >> 1;
  const int n_pi = a.n_pi_w;
  const int net = pair < n_pi ? 0 : 1;
  const int64_t pr = net ? pair - n_pi : pair, npairs = net ? pairs_all - n_pi : n_pi;
  const NetParams np = net ? np_vf : np_pi;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&s.full[i], 32);
      mbar_init(&s.empty[i], 1);
    }
    mbar_init(&s.flush_full, 1);
    mbar_init(&s.flush_empty, 32);
#pragma unroll
    for (int i = 0; i < kWIn; ++i) {
      mbar_init(&s.in_full[i], 1);
      mbar_init(&s.in_empty[i], 16);
    }
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc_pair(&s.tmem_base, 512);
  float s_h = 1.0f, s_d = 1.0f, inv_scale = 1.0f;
  if constexpr (kUF16) {
    s_h = pow2_scale_for(h1_bound(np, __uint_as_float(a.sc->omax), s.red));
    const float dmax = __uint_as_float(a.sc->dmax[net]);
    if (net == 0) s_d = pow2_scale_for(dmax * w3_colsum_bound(np, s.red));
    else s_d = pow2_scale_for(dmax) * (1.0f / 16384.0f);  // max |dOut| s_d <= 1
    inv_scale = (1.0f / s_d) * (1.0f / s_h);
  }
  inv_scale = -inv_scale;  // the H1 operand is staged negated (h1_chunk<true>)
  stage_w1t(s.w1t, np, -s_h);
  fill_mask_lut<kUF16>(s.lut);
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  if (warp < 16) {
    if (net == 0) update_w_workers<P, NPB>(s, np, a, 0, pr, npairs, rank, s_d, inv_scale);
    else update_w_workers_vnet<NPB>(s, np, a, pr, npairs, rank, s_d, inv_scale);
  } else if (warp >= 17) {
    // loader warps (two in both CTAs): the inputs of the rows of stage k, lane = row, up to kWIn stages ahead of the
    // workers
    const int lane = tid & 31;
    const int64_t nstages = (a.Mc + kXKc - 1) / kXKc;
    const int64_t n_my = pr < nstages ? (nstages - pr + npairs - 1) / npairs : 0;
    // Each loader warp keeps four stages' loads in flight (the loads of its stage k + 8 are issued right after its stage
    // k has been handed over): with one warp and one or two stages in flight the whole kernel ran at the pace of the
    // loader's global-load and index-arithmetic latency, whatever the workers did.
    struct Row {
      float ob[7];
      float4 d;
      uint4 m;
    };
    auto issue = [&](int64_t k, Row& r) {
      const int64_t rowl = (pr + k * npairs) * kXKc + lane;
      const bool valid = load_row_obs(a, rowl, np.D, r.ob, nullptr);
      r.d = make_float4(0.0f, 0.0f, 0.0f, 0.0f), r.m = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        if (net == 0) r.d = *reinterpret_cast<const float4*>(a.dout[0] + rowl * 4);
        else r.d.x = a.dout[1][rowl * 4];
        r.m = *reinterpret_cast<const uint4*>(a.mask2[net] + rowl * 8 + 4 * rank);
      }
    };
    auto hand_over = [&](int64_t k, const Row& r) {
      const int slot = (int)(k % kWIn);
      if (k >= kWIn) mbar_wait(&s.in_empty[slot], (uint32_t)((k / kWIn - 1) & 1));
      WInSlot& in = s.in[slot];
      in.o0[lane] = make_float4(r.ob[0], r.ob[1], r.ob[2], r.ob[3]);
      in.o1[lane] = make_float4(r.ob[4], r.ob[5], r.ob[6], r.d.x);
      if (net == 0) in.dd[lane] = make_float4(r.d.x * s_d, r.d.y * s_d, r.d.z * s_d, r.d.w * s_d);
      in.mm[lane] = r.m;
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.in_full[slot]);
    };
    // loader warp l (0 / 1) owns the stages k = l (mod 2): k, k + 2, k + 4, k + 6 in flight
    const int64_t l = warp - 17;
    Row ra, rb, rc, rd;
    if (l < n_my) issue(l, ra);
    if (l + 2 < n_my) issue(l + 2, rb);
    if (l + 4 < n_my) issue(l + 4, rc);
    if (l + 6 < n_my) issue(l + 6, rd);
    for (int64_t k = l; k < n_my; k += 8) {
      hand_over(k, ra);
      if (k + 8 < n_my) issue(k + 8, ra);
      if (k + 2 < n_my) {
        hand_over(k + 2, rb);
        if (k + 10 < n_my) issue(k + 10, rb);
      }
      if (k + 4 < n_my) {
        hand_over(k + 4, rc);
        if (k + 12 < n_my) issue(k + 12, rc);
      }
      if (k + 6 < n_my) {
        hand_over(k + 6, rd);
        if (k + 14 < n_my) issue(k + 14, rd);
      }
    }
  } else if (rank == 0) {
    const int64_t nstages = (a.Mc + kXKc - 1) / kXKc;
    const int64_t n_my = pr < nstages ? (nstages - pr + npairs - 1) / npairs : 0;
    const uint32_t idesc = upd_idesc(1, 1);
    using T = Terms<NPB>;
    int64_t nf = 0;  // flushes committed so far
    for (int64_t k = 0; k < n_my; ++k) {
      const int st = (int)(k % kWStages);
      const bool fresh = k % kFlushStages == 0;  // first stage of a flush interval: both accumulators restart
      if (fresh && k > 0) mbar_wait_cluster(&s.flush_empty, (uint32_t)((nf - 1) & 1));
      mbar_wait_cluster_sleep(&s.full[st], (uint32_t)((k / kWStages) & 1));
      fence_after_sync();
      if (net == 1) {  // (warp-uniform: elect.sync needs the whole warp)
        if (elect_one()) {
        // value network: A = the mask (tile a[0]), NPB products a0 b_i per K step, a0 b0 -> leading accumulator
#pragma unroll
        for (int ks = 0; ks < kXKc / 16; ++ks) {
          const uint64_t ad = smem_desc(smem_u32(s.ring[st].a[0]) + ks * 256, 128, 512);
#pragma unroll
          for (int i = NPB - 1; i >= 0; --i) {
            const uint64_t bd = smem_desc(smem_u32(s.ring[st].b[i]) + ks * 256, 128, 512);
            const bool first = fresh && ks == 0 && (i == 0 || i == NPB - 1);
            mma_bf16_pair(s.tmem_base + (i == 0 ? 0u : (uint32_t)H), ad, bd, idesc, first ? 0u : 1u);
          }
        }
        mma_commit_pair(&s.empty[st]);
        if ((k + 1) % kFlushStages == 0 || k == n_my - 1) mma_commit_pair(&s.flush_full);
        }
      } else if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kXKc / 16; ++ks) {
#pragma unroll
          for (int i = 0; i < T::n; ++i) {
            // MN-major tiles: LBO 128 (next 8 K rows), SBO 512 (next 8 M / N columns), 16 K rows = 256 bytes
            const uint64_t ad = smem_desc(smem_u32(s.ring[st].a[T::a(i)]) + ks * 256, 128, 512);
            const uint64_t bd = smem_desc(smem_u32(s.ring[st].b[T::b(i)]) + ks * 256, 128, 512);
            const bool lead = T::a(i) == 0 && T::b(i) == 0;  // a0 b0 -> leading accumulator, the rest -> corrections
            const bool first = fresh && ks == 0 && (lead || i == 0);
            mma_bf16_pair(s.tmem_base + (lead ? 0u : (uint32_t)H), ad, bd, idesc, first ? 0u : 1u);
          }
        }
        mma_commit_pair(&s.empty[st]);
        if ((k + 1) % kFlushStages == 0 || k == n_my - 1) mma_commit_pair(&s.flush_full);
      }
      __syncwarp();
      if ((k + 1) % kFlushStages == 0 || k == n_my - 1) ++nf;
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 16) tmem_dealloc_pair(s.tmem_base, 512);
}

// ---- host ------------------------------------------------------------------------------------------------------------
static unsigned long long* g_x3_dbg = nullptr;
extern "C" int rl8_x3_debug_buffer(unsigned long long* device_counters) {
  g_x3_dbg = device_counters;
  return RL8_OK;
}
static int x3_stages() {  // development switch: bit 0 forward / loss, bit 1 input-gradient, bit 2 weight-gradient kernel
  const char* e = getenv("RL8_X3_STAGES");
  return e ? atoi(e) : 7;
}
static int x3_backward_pieces() {  // bf16 pieces -- 3: six piece products in the gradient contractions too; 2: three
  if (kUF16) return 2;             // (fp16 pieces come in twos)
  const char* e = getenv("RL8_X3_BACKWARD_PIECES");
  return e && atoi(e) == 2 ? 2 : 3;
}
static int x3_policy_pairs(int pairs) {
  const char* e = getenv("RL8_X3_POLICY_PAIRS");  // tuning knob: the policy network's epilogue is the heavier one
  int n = e ? atoi(e) : (pairs * 20 + 18) / 37;   // 40 of 74 (measured: 40 -> 2.23, 42 -> 2.30, 44 -> 2.44 ms)
  if (n < 1) n = 1;
  if (n > pairs - 1) n = pairs - 1;
  return n;
}

static int x3_gradient_policy_pairs(const char* env, int pairs, int of74) {
  const char* e = getenv(env);  // tuning knob
  int n = e ? atoi(e) : (of74 * pairs + 36) / 74;
  if (n < 1) n = 1;
  if (n > pairs - 1) n = pairs - 1;
  return n;
}

int64_t ppo_x3_workspace(const rl8_model*, int64_t max_rows) {
  const int64_t chunk = max_rows < kXChunkRows ? max_rows : kXChunkRows;
  // forward + transposed piece images of both networks, the operand magnitudes, scratch of both networks
  return 4 * (int64_t)kXImgBytes + 256 + 2 * chunk * 80 + 256;
}

int ppo_minibatch_x3(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch, const int64_t* rows,
                     int64_t row_begin, int64_t M, double mean_denominator, const rl8_ppo_hparams* hp,
                     double* loss_sums, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (model->H != H || model->P > kMaxPT || model->D > 7) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < ppo_x3_workspace(model, M)) return RL8_ERR_WORKSPACE;
  const int64_t chunk = M < kXChunkRows ? M : kXChunkRows;
  uint8_t* p = (uint8_t*)workspace;
  uint8_t* img_f[2] = {p, p + kXImgBytes};
  p += 2 * (int64_t)kXImgBytes;
  uint8_t* img_b[2] = {p, p + kXImgBytes};
  p += 2 * (int64_t)kXImgBytes;
  X3Scales* sc = (X3Scales*)p;
  p += 256;
  const int npb = x3_backward_pieces();
  int rc;
  if (kUF16) {
    cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(X3Scales), st);
    if (e != cudaSuccess) {
      set_last_error("cudaMemsetAsync", e);
      return RL8_ERR_CUDA;
    }
    // max |obs| over the T slabs the minibatch rows come from: bounds H1 (h1_bound)
    if ((rc = launch_absmax_bits(batch->obs, (int64_t)batch->T * model->D * batch->N, &sc->omax, st))) return rc;
  }
  {
    // the four piece images of the call in one launch; with bf16 pieces the forward image has three pieces and the
    // transposed ones may have two (RL8_X3_BACKWARD_PIECES): then two launches
    PackJobs fwd{}, bwd{};
    for (int net = 0; net < 2; ++net) {
      const float* w2 = net ? model->vf_w2 : model->pi_w2;
      fwd.j[net] = PackJob{w2, img_f[net], nullptr, &sc->w2f[net], 0};
      bwd.j[net] = PackJob{w2, img_b[net], net ? model->vf_w3 : nullptr, &sc->w2b[net], 1};
    }
    if (kUF16 || npb == 3) {
      PackJobs all = fwd;
      all.j[2] = bwd.j[0], all.j[3] = bwd.j[1];
      if ((rc = launch_pack_w2_jobs(all, 4, kUF16 ? -2 : 3, st))) return rc;
    } else {
      if ((rc = launch_pack_w2_jobs(fwd, 2, 3, st))) return rc;
      if ((rc = launch_pack_w2_jobs(bwd, 2, npb, st))) return rc;
    }
  }
  UpdXArgs a;
  a.sc = sc;
  for (int net = 0; net < 2; ++net) {
    a.mask1[net] = (uint32_t*)p, p += chunk * 32;
    a.mask2[net] = (uint32_t*)p, p += chunk * 32;
    a.dout[net] = (float*)p, p += chunk * 16;
  }
  a.obs = batch->obs, a.actions = batch->actions, a.logp = batch->logp;
  a.adv = batch->advantages, a.ret = batch->returns;
  a.rows = rows, a.row_begin = row_begin, a.M = M, a.N = batch->N, a.T = batch->T;
  a.slab_env0 = 0, a.slab_nenv = 0;
  if (!rows && batch->T > 0 && row_begin % batch->T == 0 && M % batch->T == 0) {
    a.slab_env0 = row_begin / batch->T;
    a.slab_nenv = M / batch->T;
  }
  a.small = (M < (1ll << 31) && (int64_t)(batch->T + 1) * batch->N < (1ll << 31)) ? 1 : 0;
  a.dist_kind = batch->dist_kind, a.hp = *hp;
  a.inv_denom = (float)((double)hp->loss_scale / mean_denominator);
  a.gw1[0] = (float*)grads->pi_w1, a.gb1[0] = (float*)grads->pi_b1, a.gw2[0] = (float*)grads->pi_w2;
  a.gb2[0] = (float*)grads->pi_b2, a.gw3[0] = (float*)grads->pi_w3, a.gb3[0] = (float*)grads->pi_b3;
  a.gw1[1] = (float*)grads->vf_w1, a.gb1[1] = (float*)grads->vf_b1, a.gw2[1] = (float*)grads->vf_w2;
  a.gb2[1] = (float*)grads->vf_b2, a.gw3[1] = (float*)grads->vf_w3, a.gb3[1] = (float*)grads->vf_b3;
  a.sums = loss_sums;
  {
    const char* e = getenv("RL8_X3_ABL");
    a.abl = e ? atoi(e) : 0;  // read by the kernels only in RL8_X3_ABLATE builds
  }
  a.dbg = g_x3_dbg;
  const NetParams np_pi = net_params(model, 0, img_f[0]), np_vf = net_params(model, 1, img_f[1]);
  const int stages = x3_stages();
  for (int64_t off = 0; off < M; off += chunk) {
    a.row_off = off;
    a.Mc = M - off < chunk ? M - off : chunk;
    const int64_t ntiles = ceil_div(a.Mc, 256);
    int pairs = (int)(2 * ntiles < kNumSMs / 2 ? 2 * ntiles : kNumSMs / 2);
    if (pairs < 2) pairs = 2;
    a.n_pi = pairs == kNumSMs / 2 ? x3_policy_pairs(pairs) : pairs / 2;
    // value pairs of the gradient kernels issue half the piece products: the policy network gets 42 of 74 pairs
    // (measured, b: 38 -> 2.16, 40 -> 2.05, 42 -> 1.98, 44 -> 2.08 ms)
    a.n_pi_b = x3_gradient_policy_pairs("RL8_X3_POLICY_PAIRS_B", pairs, 42);
    if (stages & 1) {
#define RL8_UPDF(PV)                                                                                  \
  case PV:                                                                                            \
    if ((rc = set_smem((const void*)x3_update_f_kernel<PV>, sizeof(SmemXF)))) return rc;               \
    x3_update_f_kernel<PV><<<2 * pairs, kXThreads, sizeof(SmemXF), st>>>(np_pi, np_vf, a);             \
    break;
      switch (model->P) {
        RL8_UPDF(2) RL8_UPDF(3) RL8_UPDF(4)
        default: return RL8_ERR_UNSUPPORTED;
      }
#undef RL8_UPDF
      if ((rc = check_launch("x3_update_f"))) return rc;
    }
    if (stages & 2) {
      const NetParams nb_pi = net_params(model, 0, img_b[0]), nb_vf = net_params(model, 1, img_b[1]);
#define RL8_UPDB(PV)                                                                                            \
  case PV:                                                                                                      \
    if (npb == 2) {                                                                                             \
      if ((rc = set_smem((const void*)x3_update_b_kernel<PV, 2>, sizeof(SmemXB<2>)))) return rc;                 \
      x3_update_b_kernel<PV, 2><<<2 * pairs, kXThreads, sizeof(SmemXB<2>), st>>>(nb_pi, nb_vf, a);               \
    } else {                                                                                                    \
      if ((rc = set_smem((const void*)x3_update_b_kernel<PV, kNPBMax>, sizeof(SmemXB<kNPBMax>)))) return rc;     \
      x3_update_b_kernel<PV, kNPBMax><<<2 * pairs, kXThreads, sizeof(SmemXB<kNPBMax>), st>>>(nb_pi, nb_vf, a);   \
    }                                                                                                           \
    break;
      switch (model->P) {
        RL8_UPDB(2) RL8_UPDB(3) RL8_UPDB(4)
        default: return RL8_ERR_UNSUPPORTED;
      }
#undef RL8_UPDB
      if ((rc = check_launch("x3_update_b"))) return rc;
    }
    if (stages & 4) {
      const int64_t nst = ceil_div(a.Mc, kXKc);
      int wpairs = (int)(nst < kNumSMs / 2 ? nst : kNumSMs / 2);
      if (wpairs < 2) wpairs = 2;
      a.n_pi_w = x3_gradient_policy_pairs("RL8_X3_POLICY_PAIRS_W", wpairs, 39);  // 36 -> 1.55, 38 -> 1.47, 40 -> 1.52 ms
#define RL8_UPDW(PV)                                                                                            \
  case PV:                                                                                                      \
    if (npb == 2) {                                                                                             \
      if ((rc = set_smem((const void*)x3_update_w_kernel<PV, 2>, sizeof(SmemXW<2>)))) return rc;                 \
      x3_update_w_kernel<PV, 2><<<2 * wpairs, kWThreads, sizeof(SmemXW<2>), st>>>(np_pi, np_vf, a);              \
    } else {                                                                                                    \
      if ((rc = set_smem((const void*)x3_update_w_kernel<PV, kNPBMax>, sizeof(SmemXW<kNPBMax>)))) return rc;     \
      x3_update_w_kernel<PV, kNPBMax><<<2 * wpairs, kWThreads, sizeof(SmemXW<kNPBMax>), st>>>(np_pi, np_vf, a);  \
    }                                                                                                           \
    break;
      switch (model->P) {
        RL8_UPDW(2) RL8_UPDW(3) RL8_UPDW(4)
        default: return RL8_ERR_UNSUPPORTED;
      }
#undef RL8_UPDW
      if ((rc = check_launch("x3_update_w"))) return rc;
    }
  }
  return RL8_OK;
}

}  // namespace rl8
