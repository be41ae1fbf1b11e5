// Recurrent (LSTM) policy.  RL8_PREC_FP32 (reference `enable_amp=False`): every GEMM in fp32 on CUDA
// cores, bit-comparable with the reference.  RL8_PREC_BF16 (`enable_amp=True`): the three 256 x 1024
// contractions (gates = h W_hh^T, dh = dG W_hh, gW_hh += dG^T h) run on tcgen05 with bf16 operands and
// fp32 accumulation (gemm_tc.cu) -- the forward one with the cell non-linearity in its epilogue
// (tc_lstm_cell_kernel) -- everything else is unchanged fp32.
//
//   rollout  (src/rl8/algorithms/_recurrent.py:356-445): per step  SGEMM h.W_hh^T -> cell kernel
//            (adds x.W_ih^T + biases, gate non-linearities, writes the state slabs) -> heads ->
//            the feedforward path's fused [sample + logp + env.step + buffer writes] tail.
//   update   (:517-600): minibatches of seq_len-step sequences replayed from the stored
//            chunk-start states; hand-derived back-propagation through time.
//
// Gate packing follows torch.nn.LSTM: [i | f | g | o] blocks of H rows; c' = sig(f) c + sig(i)
// tanh(g), h' = sig(o) tanh(c') (restated in oracle/recurrent_oracle.py:lstm_cell).
#include "dist.cuh"
#include "lstm_tc.cuh"
#include "mlp_fp32.cuh"
#include "ppo_loss.cuh"

namespace rl8 {

// collect.cu
int validate_rollout_dims(int mD, int mH, int mP, const rl8_rollout* ro);
int collect_tail(const rl8_rollout* ro, int t, const float* feat, cudaStream_t st, uint32_t* omax_next = nullptr,
                 uint32_t* omax_all = nullptr);

// The recurrent path's GEMM: CUDA-core fp32 or tcgen05 bf16 by precision.
static int lstm_gemm(int prec, bool a_kmajor, bool b_kmajor, int epi, const float* A, const float* B, float* C,
                     int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int splits,
                     cudaStream_t st) {
  if (prec == RL8_PREC_BF16)
    return launch_tc_gemm(a_kmajor, b_kmajor, epi, A, B, C, M, N, K, lda, ldb, ldc, splits, st);
  return launch_sgemm(a_kmajor, b_kmajor, epi, A, B, C, M, N, K, lda, ldb, ldc, nullptr, splits, st);
}

constexpr int kLH = 256;   // hidden width of the default recurrent models
constexpr int kLD = 8;     // widest observation on the fused path

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- forward cell --------------------------------------------------------------------------------
// HBM-bound stream: per row 4 KB of gate pre-activations in, 4 KB of activations + c + h out.  A thread
// owns FOUR consecutive hidden units of one row (128-bit loads and stores throughout); a block of 256
// threads walks 4 rows per iteration; W_ih (transposed, [gate][d][unit]) and the biases sit in shared
// memory.  G[r][4H] holds h_prev.W_hh^T on entry; on exit `act` (may alias G) holds the gate ACTIVATIONS
// i, f, g, o (kept for the backward pass; NULL in the rollout).  Arithmetic order per element:
// pre = (b_ih + sum_d x_d w_d) + (G + b_hh), as oracle/recurrent_oracle.py:lstm_cell.
__global__ void __launch_bounds__(256, 3)
lstm_cell_fwd_kernel(const float* G, RowMap xmap, int D, int64_t rows,
                     const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                     const float* __restrict__ b_hh, const float* c_prev, float* act, float* c_out,
                     float* __restrict__ h_out) {
  __shared__ __align__(16) float wt[4][kLD][kLH];  // 32 KB
  __shared__ __align__(16) float bi[4][kLH], bh[4][kLH];  // 8 KB
  for (int i = threadIdx.x; i < 4 * kLD * kLH; i += blockDim.x) {
    const int g = i / (kLD * kLH), d = (i / kLH) % kLD, j = i % kLH;
    wt[g][d][j] = d < D ? w_ih[(g * kLH + j) * D + d] : 0.0f;
  }
  for (int i = threadIdx.x; i < 4 * kLH; i += blockDim.x) {
    (&bi[0][0])[i] = b_ih[i];
    (&bh[0][0])[i] = b_hh[i];
  }
  __syncthreads();
  const int rb = threadIdx.x >> 6, j0 = (threadIdx.x & 63) * 4;
  const int64_t ds = xmap.dstride();
  for (int64_t r = (int64_t)blockIdx.x * 4 + rb; r < rows; r += (int64_t)gridDim.x * 4) {
    const float* Gr = G + r * 4 * kLH + j0;
    float4 gq[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) gq[g] = *reinterpret_cast<const float4*>(Gr + g * kLH);
    const float4 cp = *reinterpret_cast<const float4*>(c_prev + r * kLH + j0);
    const int64_t xo = xmap.offset(r);
    float x[kLD];
#pragma unroll
    for (int d = 0; d < kLD; ++d) x[d] = d < D ? xmap.obs[xo + d * ds] : 0.0f;
    float pre[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 b = *reinterpret_cast<const float4*>(&bi[g][j0]);
      float a[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int d = 0; d < kLD; ++d) {
        if (d < D) {
          const float4 w = *reinterpret_cast<const float4*>(&wt[g][d][j0]);
          a[0] = fmaf(x[d], w.x, a[0]), a[1] = fmaf(x[d], w.y, a[1]);
          a[2] = fmaf(x[d], w.z, a[2]), a[3] = fmaf(x[d], w.w, a[3]);
        }
      }
      const float4 h4 = *reinterpret_cast<const float4*>(&bh[g][j0]);
      pre[g][0] = a[0] + (gq[g].x + h4.x), pre[g][1] = a[1] + (gq[g].y + h4.y);
      pre[g][2] = a[2] + (gq[g].z + h4.z), pre[g][3] = a[3] + (gq[g].w + h4.w);
    }
    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
    float ig[4], fg[4], gg[4], og[4], c[4], h[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ig[u] = sigmoidf_(pre[0][u]), fg[u] = sigmoidf_(pre[1][u]);
      gg[u] = tanhf(pre[2][u]), og[u] = sigmoidf_(pre[3][u]);
      c[u] = fg[u] * cpv[u] + ig[u] * gg[u];
      h[u] = og[u] * tanhf(c[u]);
    }
    if (act) {
      float* ar = act + r * 4 * kLH + j0;
      *reinterpret_cast<float4*>(ar) = make_float4(ig[0], ig[1], ig[2], ig[3]);
      *reinterpret_cast<float4*>(ar + kLH) = make_float4(fg[0], fg[1], fg[2], fg[3]);
      *reinterpret_cast<float4*>(ar + 2 * kLH) = make_float4(gg[0], gg[1], gg[2], gg[3]);
      *reinterpret_cast<float4*>(ar + 3 * kLH) = make_float4(og[0], og[1], og[2], og[3]);
    }
    *reinterpret_cast<float4*>(c_out + r * kLH + j0) = make_float4(c[0], c[1], c[2], c[3]);
    *reinterpret_cast<float4*>(h_out + r * kLH + j0) = make_float4(h[0], h[1], h[2], h[3]);
  }
}

static int launch_cell_fwd(const rl8_lstm_model* m, const float* G, const RowMap& xmap, int64_t rows,
                           const float* c_prev, float* act, float* c_out, float* h_out,
                           cudaStream_t st) {
  const int64_t groups = ceil_div(rows, 4);  // 4 rows per block iteration
  int grid = (int)(groups < (int64_t)kNumSMs * 8 ? groups : (int64_t)kNumSMs * 8);
  lstm_cell_fwd_kernel<<<grid, 256, 0, st>>>(G, xmap, m->D, rows, m->w_ih, m->b_ih, m->b_hh, c_prev,
                                             act, c_out, h_out);
  return check_launch("lstm_cell_fwd");
}

// One LSTM step + heads.  G: [rows][4H] scratch.  hb_in / hb_out (tensor-core path): bf16 T128 images of h_in / h_out
// (hb_in null: h_in is packed into the scratch first; hb_out null: not written).
static bool lstm_step_on_tc(int prec, int64_t rows, const float* h_in, const float* h_out, const float* c_in,
                            const float* c_out) {
  // the fused kernels use 256-bit accesses on the state rows: every row base must be 32-byte aligned
  auto al32 = [](const void* p) { return ((uintptr_t)p & 31u) == 0; };
  return prec == RL8_PREC_BF16 && rows >= 512 && h_in != h_out && c_in != c_out && al32(h_out) && al32(c_in) &&
         al32(c_out);
}
static int lstm_step_fp32(const rl8_lstm_model* m, const RowMap& xmap, int64_t rows,
                          const float* h_in, const float* c_in, float* h_out, float* c_out,
                          float* act, float* G, float* features, float* values, int tanh_col1,
                          int prec, cudaStream_t st, const uint8_t* hb_in = nullptr, uint8_t* hb_out = nullptr,
                          uint8_t* zb = nullptr) {
  int rc;
  if (lstm_step_on_tc(prec, rows, h_in, h_out, c_in, c_out)) {
    // tensor-core path: the gate GEMM with the cell in its epilogue (no pre-activation round trip through HBM)
    if (!hb_in) {  // [rows][4H] floats of scratch hold the [rows_pad][H] bf16 image (rows >= 512)
      if ((rc = launch_pack_t128(h_in, rows, (uint8_t*)G, st))) return rc;
      hb_in = (const uint8_t*)G;
    }
    // (the fused kernel keeps bf16 pre-activations `zb` for the backward pass instead of fp32 activations `act`)
    if ((rc = launch_lstm_cell_tc(hb_in, m->w_hh, m->w_ih, m->b_ih, m->b_hh, c_in, xmap, m->D, rows, zb, c_out,
                                  h_out, hb_out, st)))
      return rc;
  } else {
    if ((rc = lstm_gemm(prec, true, true, EPI_STORE, h_in, m->w_hh, G, rows, 4 * kLH, kLH, kLH, kLH, 4 * kLH, 1,
                        st)))
      return rc;
    if ((rc = launch_cell_fwd(m, G, xmap, rows, c_in, act, c_out, h_out, st))) return rc;
  }
  if (features && values)  // both heads in one pass over h' (same arithmetic per output)
    return launch_lstm_heads_fwd(h_out, rows, m->P, m->pi_w, m->pi_b, m->vf_w, m->vf_b, features, values, tanh_col1, st);
  if (features &&
      (rc = launch_head_fwd(h_out, rows, kLH, m->P, m->pi_w, m->pi_b, features, tanh_col1, st)))
    return rc;
  if (values && (rc = launch_head_fwd(h_out, rows, kLH, 1, m->vf_w, m->vf_b, values, 0, st)))
    return rc;
  return RL8_OK;
}

static int check_model(const rl8_lstm_model* m) {
  if (!m || !m->w_ih || !m->w_hh || !m->b_ih || !m->b_hh || !m->pi_w || !m->pi_b || !m->vf_w ||
      !m->vf_b)
    return RL8_ERR_ARG;
  if (m->H != kLH || m->D < 1 || m->D > kLD || m->P < 2 || m->P > kMaxP) return RL8_ERR_UNSUPPORTED;
  return RL8_OK;
}

// Does step t of this collect start from re-initialised (zero) states?  (:384-392)
static bool state_reset_at(const rl8_recurrent_rollout* rro, int t) {
  if (t % rro->seq_len) return false;
  const int64_t seqs = rro->seqs + t / rro->seq_len;
  if (seqs && rro->seqs_per_state_reset < 0) return false;
  return (seqs % rro->seqs_per_state_reset) == 0;
}

int lstm_collect_fp32(const rl8_lstm_model* m, const rl8_recurrent_rollout* rro, int prec, void* workspace,
                      int64_t workspace_bytes, cudaStream_t st) {
  const rl8_rollout* ro = &rro->ro;
  const int64_t N = ro->N;
  const int T = ro->T, D = m->D;
  const int64_t need = (N * 4 * kLH + 2 * N * kLH + N * kMaxP) * 4;
  if (!workspace || workspace_bytes < need) return RL8_ERR_WORKSPACE;
  float* G = (float*)workspace;
  float* hs = G + N * 4 * kLH;  // scratch state for the bootstrap value
  float* cs = hs + N * kLH;
  float* feat = cs + N * kLH;
  const bool continuous = ro->dist_kind != RL8_DIST_CATEGORICAL;
  RowMap map{};
  map.mode = 0, map.stride_r = 1, map.stride_d = N, map.D = D;
  const int64_t slab = N * kLH;
  // tensor-core path: h travels between steps as a bf16 T128 image too (two of them, ping-pong, inside the gate scratch
  // the fused kernel does not need); an image is (re)built from the fp32 slab at t = 0 and after a state reset
  const bool tc = lstm_step_on_tc(prec, N, nullptr, rro->hidden, nullptr, rro->cell) && (slab * 4) % 32 == 0;
  uint8_t* hb[2] = {(uint8_t*)G, (uint8_t*)G + t128_bytes(N, kLH)};
  bool hb_valid = false;
  for (int t = 0; t < T; ++t) {
    float* h_t = rro->hidden + (int64_t)t * slab;
    float* c_t = rro->cell + (int64_t)t * slab;
    if (state_reset_at(rro, t)) {
      cudaMemsetAsync(h_t, 0, slab * 4, st);
      cudaMemsetAsync(c_t, 0, slab * 4, st);
      hb_valid = false;
    }
    int rc;
    if (tc && !hb_valid && (rc = launch_pack_t128(h_t, N, hb[t & 1], st))) return rc;
    map.obs = ro->obs + (int64_t)t * D * N;
    rc = lstm_step_fp32(m, map, N, h_t, c_t, h_t + slab, c_t + slab, nullptr, G, feat,
                        ro->values + (int64_t)t * N, continuous, prec, st, tc ? hb[t & 1] : nullptr,
                        tc ? hb[(t + 1) & 1] : nullptr);
    hb_valid = tc;
    if (rc) return rc;
    if ((rc = collect_tail(ro, t, feat, st))) return rc;
  }
  // bootstrap value from the last observation and the final states (:433-445)
  map.obs = ro->obs + (int64_t)T * D * N;
  return lstm_step_fp32(m, map, N, rro->hidden + (int64_t)T * slab, rro->cell + (int64_t)T * slab,
                        hs, cs, nullptr, G, nullptr, ro->values + (int64_t)T * N, 0, prec, st,
                        tc && hb_valid ? hb[T & 1] : nullptr, nullptr);
}

// ---- update ----------------------------------------------------------------------------------------

// rows_k[k*C + r] = seq(r)*L + k (flattened transition, row = n*T + t);  h0/c0[r] = stored
// chunk-start state of sequence r.
__global__ void __launch_bounds__(256)
seq_setup_kernel(const int64_t* __restrict__ seqs, int64_t seq_begin, int64_t C, int L, int T,
                 int64_t N, const float* __restrict__ hidden, const float* __restrict__ cell,
                 int64_t* __restrict__ rows_k, float* __restrict__ h0, float* __restrict__ c0) {
  const int Q = kLH / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < C * Q;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Q;
    const int q = (int)(i - r * Q);
    const int64_t s = seqs ? seqs[r] : seq_begin + r;
    if (q < L) rows_k[(int64_t)q * C + r] = s * L + q;
    if (q == 0)
      for (int k = Q; k < L; ++k) rows_k[(int64_t)k * C + r] = s * L + k;  // L > 64 (rare)
    const int64_t g = s * L, n = g / T, t = g - n * T;
    const int64_t src = (t * N + n) * kLH + q * 4;
    *reinterpret_cast<float4*>(h0 + r * kLH + q * 4) = *reinterpret_cast<const float4*>(hidden + src);
    *reinterpret_cast<float4*>(c0 + r * kLH + q * 4) = *reinterpret_cast<const float4*>(cell + src);
  }
}

// dh[r][j] = sum_p dpi[r][p] pi_w[p][j] + dvf[r] vf_w[j] (+ dh[r][j] when accumulate)
template <int P>
__global__ void __launch_bounds__(256)
lstm_dh_kernel(const float* __restrict__ dpi, const float* __restrict__ dvf, int64_t rows,
               const float* __restrict__ pi_w, const float* __restrict__ vf_w, int accumulate,
               float* __restrict__ dh) {
  constexpr int Q = kLH / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * Q;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Q;
    const int j = (int)(i - r * Q) * 4;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float d = dpi[r * P + p];
      const float4 w = *reinterpret_cast<const float4*>(pi_w + p * kLH + j);
      g[0] = fmaf(d, w.x, g[0]), g[1] = fmaf(d, w.y, g[1]);
      g[2] = fmaf(d, w.z, g[2]), g[3] = fmaf(d, w.w, g[3]);
    }
    const float dv = dvf[r];
    const float4 wv = *reinterpret_cast<const float4*>(vf_w + j);
    g[0] = fmaf(dv, wv.x, g[0]), g[1] = fmaf(dv, wv.y, g[1]);
    g[2] = fmaf(dv, wv.z, g[2]), g[3] = fmaf(dv, wv.w, g[3]);
    float4* out = reinterpret_cast<float4*>(dh + r * kLH + j);
    if (accumulate) {
      const float4 o = *out;
      g[0] += o.x, g[1] += o.y, g[2] += o.z, g[3] += o.w;
    }
    *out = make_float4(g[0], g[1], g[2], g[3]);
  }
}

static int launch_dh(int P, const float* dpi, const float* dvf, int64_t rows, const float* pi_w,
                     const float* vf_w, int accumulate, float* dh, cudaStream_t st) {
  int grid = grid_for(rows * 64, 256, 8, 4);
#define RL8_DH(PV)                                                                          \
  case PV:                                                                                  \
    lstm_dh_kernel<PV><<<grid, 256, 0, st>>>(dpi, dvf, rows, pi_w, vf_w, accumulate, dh);    \
    break;
  switch (P) {
    RL8_DH(2) RL8_DH(3) RL8_DH(4) RL8_DH(5) RL8_DH(6) RL8_DH(7) RL8_DH(8)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_DH
  return check_launch("lstm_dh");
}

// Cell backward, one (row, unit) per thread.  act holds i,f,g,o and is overwritten with the
// PRE-activation gate gradients; dc (in/out) carries dL/dc across steps (has_dc_in = 0 for the
// last step of a sequence).
__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(float* __restrict__ act, const float* __restrict__ c, const float* __restrict__ c_prev,
                     const float* __restrict__ dh, float* __restrict__ dc, int has_dc_in,
                     int64_t rows) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * kLH;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kLH;
    const int j = (int)(i - r * kLH);
    float* a = act + r * 4 * kLH + j;
    const float ig = a[0], fg = a[kLH], gg = a[2 * kLH], og = a[3 * kLH];
    const float tc = tanhf(c[i]);
    const float dhv = dh[i];
    float dcv = dhv * og * (1.0f - tc * tc);
    if (has_dc_in) dcv += dc[i];
    a[0] = dcv * gg * ig * (1.0f - ig);
    a[kLH] = dcv * c_prev[i] * fg * (1.0f - fg);
    a[2 * kLH] = dcv * ig * (1.0f - gg * gg);
    a[3 * kLH] = dhv * tc * og * (1.0f - og);
    dc[i] = dcv * fg;
  }
}

// gw_ih[g][d] += sum_r dG[r][g] x[r][d];  gb_ih[g], gb_hh[g] += sum_r dG[r][g]   (g < 4H)
constexpr int kGrRows = 64;
__global__ void __launch_bounds__(256)
lstm_gate_reduce_kernel(const float* __restrict__ dG, int64_t rows, RowMap xmap, int D,
                        float* __restrict__ gw_ih, float* __restrict__ gb_ih,
                        float* __restrict__ gb_hh, int64_t rows_per_block) {
  __shared__ float sx[kGrRows][kLD];
  const int g = blockIdx.y * 256 + threadIdx.x;
  float acc[kLD], accb = 0.0f;
#pragma unroll
  for (int d = 0; d < kLD; ++d) acc[d] = 0.0f;
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = min(rows, rbeg + rows_per_block);
  const int64_t ds = xmap.dstride();
  for (int64_t r0 = rbeg; r0 < rend; r0 += kGrRows) {
    const int nr = (int)min((int64_t)kGrRows, rend - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < kGrRows * D; i += blockDim.x) {
      const int d = i / kGrRows, r = i - d * kGrRows;
      sx[r][d] = r < nr ? xmap.obs[xmap.offset(r0 + r) + d * ds] : 0.0f;
    }
    __syncthreads();
    // four rows per trip: their loads are in flight together (the loop is latency-, not bandwidth-bound);
    // the accumulation order over rows is unchanged
    for (int rq = 0; rq < nr; rq += 4) {
      float v4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v4[i] = rq + i < nr ? dG[(r0 + rq + i) * 4 * kLH + g] : 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (rq + i < nr) {
          accb += v4[i];
#pragma unroll
          for (int d = 0; d < kLD; ++d)
            if (d < D) acc[d] = fmaf(v4[i], sx[rq + i][d], acc[d]);
        }
      }
    }
  }
#pragma unroll
  for (int d = 0; d < kLD; ++d)
    if (d < D) atomicAdd(gw_ih + (int64_t)g * D + d, acc[d]);
  atomicAdd(gb_ih + g, accb);
  atomicAdd(gb_hh + g, accb);
}

static int launch_gate_reduce(const float* dG, int64_t rows, const RowMap& xmap, int D, float* gw_ih,
                              float* gb_ih, float* gb_hh, cudaStream_t st) {
  int64_t blocks = min(ceil_div(rows, kGrRows), (int64_t)kNumSMs * 2);
  int64_t rpb = round_up(ceil_div(rows, blocks), kGrRows);
  blocks = ceil_div(rows, rpb);
  lstm_gate_reduce_kernel<<<dim3((unsigned)blocks, 4), 256, 0, st>>>(dG, rows, xmap, D, gw_ih, gb_ih,
                                                                    gb_hh, rpb);
  return check_launch("lstm_gate_reduce");
}

// Sequences resident per update chunk: bounds the activation workspace (6H floats per row and step:
// 1.6 GB at the cap, sized for 180 GB of HBM) while keeping every kernel of a step at >= 512 row tiles.
static int64_t lstm_chunk_seqs(int64_t max_seqs, int L) {
  int64_t cap = 262144 / L;
  if (cap < 1) cap = 1;
  return max_seqs < cap ? max_seqs : cap;
}

int64_t lstm_ppo_fp32_workspace(int64_t max_seqs, int L) {
  const int64_t C = lstm_chunk_seqs(max_seqs, L);
  // per step: act [C][4H], c [C][H], h [C][H], out_pi/dout_pi [C][kMaxP] x2, out_vf/dout_vf [C] x2
  const int64_t per_step = C * (6 * kLH + 2 * kMaxP + 2);
  // h0, c0, dh, dc [C][H] each; rows_k [L][C] int64; bf16 T128 images of h_0 .. h_{L-1} (tensor-core path)
  // ... and of the gate gradients dG_k [C][4H] and [x | 1] [C][16] of every step
  return (L * per_step + 4 * C * kLH) * 4 + L * C * 8 + 64 + 12 * 256 /* array alignment */ +
         L * (t128_bytes(C, kLH) + t128_bytes(C, 4 * kLH) + t128_bytes(C, 16)) + 256;
}

int lstm_ppo_minibatch_fp32(const rl8_lstm_model* m, const rl8_lstm_model* g,
                            const rl8_recurrent_batch* rb, const int64_t* seqs, int64_t seq_begin,
                            int64_t M, double denom, const rl8_ppo_hparams* hp, double* loss_sums,
                            int prec, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const rl8_batch* b = &rb->b;
  const int L = rb->seq_len, D = m->D, P = m->P;
  const int64_t C = lstm_chunk_seqs(M, L);
  if (!workspace || workspace_bytes < lstm_ppo_fp32_workspace(M, L)) return RL8_ERR_WORKSPACE;
  float* p = (float*)workspace;
  auto take = [&](int64_t n) { float* q = p; p += (n + 63) / 64 * 64; return q; };  // 256-byte aligned arrays
  float* act = take((int64_t)L * C * 4 * kLH);
  float* cbuf = take((int64_t)L * C * kLH);
  float* hbuf = take((int64_t)L * C * kLH);
  float* out_pi = take((int64_t)L * C * kMaxP);
  float* dout_pi = take((int64_t)L * C * kMaxP);
  float* out_vf = take((int64_t)L * C);
  float* dout_vf = take((int64_t)L * C);
  float* h0 = take(C * kLH);
  float* c0 = take(C * kLH);
  float* dh = take(C * kLH);
  float* dc = take(C * kLH);
  int64_t* rows_k = (int64_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  uint8_t* hb = (uint8_t*)(((uintptr_t)(rows_k + (int64_t)L * C) + 255) & ~(uintptr_t)255);  // [L] T128 images
  const int64_t hb_bytes = t128_bytes(C, kLH), dgb_bytes = t128_bytes(C, 4 * kLH), xb_bytes = t128_bytes(C, 16);
  uint8_t* dgb = hb + (int64_t)L * hb_bytes;
  uint8_t* xb = dgb + (int64_t)L * dgb_bytes;
  uint8_t* zb = (uint8_t*)act;  // tensor-core path: bf16 pre-activation images live where the fp32 activations would
  const bool continuous = b->dist_kind != RL8_DIST_CATEGORICAL;
  const int splits = 64;

  for (int64_t s0 = 0; s0 < M; s0 += C) {
    const int64_t R = (M - s0) < C ? (M - s0) : C;
    seq_setup_kernel<<<grid_for(R * 64, 256, 8, 4), 256, 0, st>>>(
        seqs ? seqs + s0 : nullptr, seq_begin + s0, R, L, b->T, b->N, rb->hidden, rb->cell, rows_k,
        h0, c0);
    int rc = check_launch("seq_setup");
    if (rc) return rc;
    RowMap map{};
    map.obs = b->obs, map.mode = 1, map.T = b->T, map.D = D, map.N = b->N;
    const bool tc = lstm_step_on_tc(prec, R, h0, hbuf, c0, cbuf);
    if (tc && (rc = launch_pack_t128(h0, R, hb, st))) return rc;

    // ---- forward through the sequence + per-step losses ------------------------------------
    for (int k = 0; k < L; ++k) {
      map.rows = rows_k + (int64_t)k * R;
      float* act_k = act + (int64_t)k * C * 4 * kLH;
      float* c_k = cbuf + (int64_t)k * C * kLH;
      float* h_k = hbuf + (int64_t)k * C * kLH;
      const float* h_prev = k ? h_k - C * kLH : h0;
      const float* c_prev = k ? c_k - C * kLH : c0;
      // (no heads here: nothing in the replay depends on them, they follow for all L steps in one pass over hbuf)
      if ((rc = lstm_step_fp32(m, map, R, h_prev, c_prev, h_k, c_k, act_k, act_k, nullptr, nullptr, continuous, prec,
                               st, tc ? hb + (int64_t)k * hb_bytes : nullptr,
                               tc && k + 1 < L ? hb + (int64_t)(k + 1) * hb_bytes : nullptr,
                               tc ? zb + (int64_t)k * dgb_bytes : nullptr)))
        return rc;
    }
    {
      HeadBlocks blocks;
      blocks.steps = L, blocks.h_stride = C, blocks.pi_stride = C * kMaxP, blocks.vf_stride = C;
      if ((rc = launch_lstm_heads_fwd(hbuf, R, m->P, m->pi_w, m->pi_b, m->vf_w, m->vf_b, out_pi, out_vf, continuous, st,
                                      blocks)))
        return rc;
    }
    {  // the losses of all L steps of the chunk in one launch (nothing in the forward replay depends on them)
      LossArgs la{};
      la.dist_kind = b->dist_kind, la.P = P, la.M = R;
      la.out_pi = out_pi, la.out_vf = out_vf;
      la.actions = b->actions, la.logp_old = b->logp, la.advantages = b->advantages;
      la.returns = b->returns, la.rows = rows_k, la.row_begin = 0;
      la.T = b->T, la.N = b->N, la.hp = *hp;
      la.inv_denom = (float)((double)hp->loss_scale / denom);
      la.dout_pi = dout_pi, la.dout_vf = dout_vf;
      la.gb3_pi = (float*)g->pi_b, la.gb3_vf = (float*)g->vf_b, la.sums = loss_sums;
      la.steps = L, la.pi_stride = C * kMaxP, la.vf_stride = C, la.rows_stride = R;
      if ((rc = launch_ppo_loss(la, st))) return rc;
    }

    // ---- back-propagation through time ------------------------------------------------------
    for (int k = L - 1; k >= 0; --k) {
      map.rows = rows_k + (int64_t)k * R;
      float* act_k = act + (int64_t)k * C * 4 * kLH;
      float* c_k = cbuf + (int64_t)k * C * kLH;
      float* h_k = hbuf + (int64_t)k * C * kLH;
      const float* h_prev = k ? h_k - C * kLH : h0;
      const float* c_prev = k ? c_k - C * kLH : c0;
      const float* dpi = dout_pi + (int64_t)k * C * kMaxP;
      const float* dvf = dout_vf + (int64_t)k * C;
      const int last = k == L - 1;
      if (tc) {
        // tensor-core path (lstm_tc.cu): one fused elementwise kernel (heads' term of dL/dh, cell backward, dG and
        // [x | 1] as bf16 T128 images, head weight gradients), then dh_{k-1} = dG_k W_hh on tcgen05
        LstmBwdArgs ba{};
        ba.zb = zb + (int64_t)k * dgb_bytes, ba.c_prev = c_prev, ba.h = h_k, ba.dh_rec = last ? nullptr : dh, ba.dc = dc;
        ba.dpi = dpi, ba.dvf = dvf, ba.pi_w = m->pi_w, ba.vf_w = m->vf_w;
        ba.dGb = dgb + (int64_t)k * dgb_bytes, ba.xb = xb + (int64_t)k * xb_bytes;
        ba.gpi_w = (float*)g->pi_w, ba.gvf_w = (float*)g->vf_w;
        ba.xmap = map, ba.D = D, ba.has_dc_in = !last, ba.rows = R;
        if ((rc = launch_lstm_cell_bwd_tc(ba, P, st))) return rc;
        if (k && (rc = launch_lstm_dh_tc(ba.dGb, m->w_hh, dh, R, st))) return rc;
        continue;
      }
      // dL/dh_k = heads' contribution (+ the recurrent term left in dh by step k+1)
      if ((rc = launch_dh(P, dpi, dvf, R, m->pi_w, m->vf_w, !last, dh, st))) return rc;
      // head weight gradients
      if ((rc = launch_thin_reduce(h_k, R, kLH, dpi, nullptr, P, (float*)g->pi_w, kLH, 1, nullptr, st)))
        return rc;
      if ((rc = launch_thin_reduce(h_k, R, kLH, dvf, nullptr, 1, (float*)g->vf_w, kLH, 1, nullptr, st)))
        return rc;
      lstm_cell_bwd_kernel<<<grid_for(R * kLH, 256, 8, 4), 256, 0, st>>>(act_k, c_k, c_prev, dh, dc,
                                                                         !last, R);
      if ((rc = check_launch("lstm_cell_bwd"))) return rc;
      // gw_hh[g][j] += sum_r dG[r][g] h_prev[r][j]   (split-K over rows)
      if ((rc = lstm_gemm(prec, false, false, EPI_ATOMIC, act_k, h_prev, (float*)g->w_hh, 4 * kLH, kLH, R,
                          4 * kLH, kLH, kLH, splits, st)))
        return rc;
      if ((rc = launch_gate_reduce(act_k, R, map, D, (float*)g->w_ih, (float*)g->b_ih,
                                   (float*)g->b_hh, st)))
        return rc;
      // dL/dh_{k-1} (recurrent term) = dG . W_hh; nothing flows into the stored chunk-start state
      if (k && (rc = lstm_gemm(prec, true, false, EPI_STORE, act_k, m->w_hh, dh, R, kLH, 4 * kLH, 4 * kLH,
                               kLH, kLH, 1, st)))
        return rc;
    }
    if (tc) {  // [gW_hh | gW_ih | gb] += dG^T [h_prev | x | 1] over the L steps of the chunk
      LstmWgArgs wa{};
      wa.dGb = dgb, wa.hb = hb, wa.xb = xb;
      wa.dGb_stride = dgb_bytes, wa.hb_stride = hb_bytes, wa.xb_stride = xb_bytes;
      wa.L = L, wa.row_tiles = ceil_div(R, (int64_t)128);
      wa.gw_hh = (float*)g->w_hh, wa.gw_ih = (float*)g->w_ih, wa.gb_ih = (float*)g->b_ih, wa.gb_hh = (float*)g->b_hh;
      wa.D = D;
      if ((rc = launch_lstm_wgrad_tc(wa, st))) return rc;
    }
  }
  return RL8_OK;
}

}  // namespace rl8

using namespace rl8;

extern "C" int64_t rl8_lstm_collect_workspace(const rl8_lstm_model* model, int64_t N, int32_t T,
                                              int precision) {
  if (check_model(model) || N <= 0 || T <= 0) return RL8_ERR_ARG;
  if (precision != RL8_PREC_FP32 && precision != RL8_PREC_BF16) return RL8_ERR_UNSUPPORTED;
  return (N * 4 * kLH + 2 * N * kLH + N * kMaxP) * 4;
}

extern "C" int rl8_lstm_collect(const rl8_lstm_model* model, const rl8_recurrent_rollout* rro,
                                int precision, void* workspace, int64_t workspace_bytes,
                                rl8_stream_t stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if (!rro || !rro->hidden || !rro->cell || rro->seq_len <= 0 || rro->seqs_per_state_reset == 0 ||
      rro->seqs < 0)
    return RL8_ERR_ARG;
  if ((rc = validate_rollout_dims(model->D, model->H, model->P, &rro->ro))) return rc;
  if (rro->ro.T % rro->seq_len) return RL8_ERR_ARG;
  if (precision != RL8_PREC_FP32 && precision != RL8_PREC_BF16) return RL8_ERR_UNSUPPORTED;
  return lstm_collect_fp32(model, rro, precision, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int rl8_lstm_forward(const rl8_lstm_model* model, const float* obs, int64_t obs_stride_r,
                                int64_t obs_stride_d, const float* h_in, const float* c_in,
                                float* h_out, float* c_out, float* features, float* values,
                                int64_t B, int apply_tanh_log_std, int precision, void* workspace,
                                int64_t workspace_bytes, rl8_stream_t stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if (!obs || !h_in || !c_in || !h_out || !c_out || B <= 0) return RL8_ERR_ARG;
  if (precision != RL8_PREC_FP32 && precision != RL8_PREC_BF16) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < B * 4 * kLH * 4) return RL8_ERR_WORKSPACE;
  RowMap map{};
  map.obs = obs, map.mode = 0, map.stride_r = obs_stride_r, map.stride_d = obs_stride_d;
  map.D = model->D;
  return lstm_step_fp32(model, map, B, h_in, c_in, h_out, c_out, nullptr, (float*)workspace,
                        features, values, apply_tanh_log_std, precision, (cudaStream_t)stream);
}

extern "C" int64_t rl8_lstm_ppo_workspace(const rl8_lstm_model* model, int64_t max_seqs,
                                          int32_t seq_len, int precision) {
  if (check_model(model) || max_seqs <= 0 || seq_len <= 0) return RL8_ERR_ARG;
  if (precision != RL8_PREC_FP32 && precision != RL8_PREC_BF16) return RL8_ERR_UNSUPPORTED;
  return lstm_ppo_fp32_workspace(max_seqs, seq_len);
}

extern "C" int rl8_lstm_ppo_minibatch(const rl8_lstm_model* model, const rl8_lstm_model* grads,
                                      const rl8_recurrent_batch* batch, const int64_t* seqs,
                                      int64_t seq_begin, int64_t M, double mean_denominator,
                                      const rl8_ppo_hparams* hp, double* loss_sums, int precision,
                                      void* workspace, int64_t workspace_bytes,
                                      rl8_stream_t stream) {
  int rc = check_model(model);
  if (rc) return rc;
  if ((rc = check_model(grads))) return rc;
  if (!batch || !hp || !loss_sums || M <= 0 || mean_denominator <= 0) return RL8_ERR_ARG;
  const rl8_batch* b = &batch->b;
  if (!b->obs || !b->actions || !b->logp || !b->advantages || !b->returns || !batch->hidden ||
      !batch->cell || batch->seq_len <= 0 || b->T % batch->seq_len)
    return RL8_ERR_ARG;
  if (b->dist_kind != RL8_DIST_CATEGORICAL && model->P != 2) return RL8_ERR_UNSUPPORTED;
  if (b->dist_kind == RL8_DIST_SQUASHED_NORMAL && hp->entropy_coeff != 0.0f)
    return RL8_ERR_UNSUPPORTED;
  if (precision != RL8_PREC_FP32 && precision != RL8_PREC_BF16) return RL8_ERR_UNSUPPORTED;
  return lstm_ppo_minibatch_fp32(model, grads, batch, seqs, seq_begin, M, mean_denominator, hp,
                                 loss_sums, precision, workspace, workspace_bytes, (cudaStream_t)stream);
}
