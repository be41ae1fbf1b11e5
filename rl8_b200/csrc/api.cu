// C-ABI entry points that orchestrate several launches (include/rl8_b200.h).
#include "mlp_fp32.cuh"
#include "ppo_loss.cuh"
#include "dist.cuh"

namespace rl8 {

static thread_local char g_last_error[256] = "";

void set_last_error(const char* where, cudaError_t err) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(err));
}

// collect.cu
int validate_rollout(const rl8_model* model, const rl8_rollout* ro);
int mlp_forward_fp32(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out,
                     int tanh_col1, float* h1, float* h2, cudaStream_t st);
int collect_fp32(const rl8_model* model, const rl8_rollout* ro, void* workspace,
                 int64_t workspace_bytes, cudaStream_t st);
// tensor-core path (mlp_tc.cu)
int64_t collect_tc_workspace(const rl8_model* model, int64_t N, int32_t T);
int collect_tc(const rl8_model* model, const rl8_rollout* ro, void* workspace,
               int64_t workspace_bytes, cudaStream_t st);
int64_t ppo_tc_workspace(const rl8_model* model, int64_t max_rows);
int ppo_minibatch_tc(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch,
                     const int64_t* rows, int64_t row_begin, int64_t M, double mean_denominator,
                     const rl8_ppo_hparams* hp, double* loss_sums, void* workspace,
                     int64_t workspace_bytes, cudaStream_t st);
int mlp_forward_tc(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out,
                   int tanh_col1, void* workspace, int64_t workspace_bytes, cudaStream_t st);
// fp32-accurate tensor-core path (split_tc.cu)
int64_t forward_x3_workspace();
int mlp_forward_x3(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out, int tanh_col1,
                   void* workspace, int64_t workspace_bytes, cudaStream_t st);
int64_t ppo_x3_workspace(const rl8_model* model, int64_t max_rows);
int ppo_minibatch_x3(const rl8_model* model, const rl8_model* grads, const rl8_batch* batch, const int64_t* rows,
                     int64_t row_begin, int64_t M, double mean_denominator, const rl8_ppo_hparams* hp,
                     double* loss_sums, void* workspace, int64_t workspace_bytes, cudaStream_t st);
int64_t collect_x3_workspace(const rl8_model* model, int64_t N, int32_t T);
int collect_x3(const rl8_model* model, const rl8_rollout* ro, void* workspace, int64_t workspace_bytes,
               cudaStream_t st);

constexpr int64_t kChunkRows = 65536;  // rows of activations resident per update chunk (fp32)

static int64_t ppo_fp32_chunk(int64_t max_rows) { return max_rows < kChunkRows ? max_rows : kChunkRows; }

static int ppo_minibatch_fp32(const rl8_model* m, const rl8_model* g, const rl8_batch* b,
                              const int64_t* rows, int64_t row_begin, int64_t M, double denom,
                              const rl8_ppo_hparams* hp, double* loss_sums, void* workspace,
                              int64_t workspace_bytes, cudaStream_t st) {
  const int H = m->H, D = m->D, P = m->P;
  const int64_t C = ppo_fp32_chunk(M);
  const int64_t need = (5 * C * H + C * (2 * kMaxP + 2)) * 4;
  if (!workspace || workspace_bytes < need) return RL8_ERR_WORKSPACE;
  float* h1p = (float*)workspace;
  float* h2p = h1p + C * H;
  float* h1v = h2p + C * H;
  float* h2v = h1v + C * H;
  float* dz2 = h2v + C * H;
  float* out_pi = dz2 + C * H;
  float* dout_pi = out_pi + C * kMaxP;
  float* out_vf = dout_pi + C * kMaxP;
  float* dout_vf = out_vf + C;
  const bool continuous = b->dist_kind != RL8_DIST_CATEGORICAL;
  const int splits = 64;

  for (int64_t c0 = 0; c0 < M; c0 += C) {
    const int64_t R = (M - c0) < C ? (M - c0) : C;
    RowMap map{};
    map.obs = b->obs, map.mode = 1, map.rows = rows ? rows + c0 : nullptr;
    map.row_begin = row_begin + c0, map.T = b->T, map.D = D, map.N = b->N;
    int rc;
    if ((rc = mlp_forward_fp32(m, 0, map, R, out_pi, continuous, h1p, h2p, st))) return rc;
    if ((rc = mlp_forward_fp32(m, 1, map, R, out_vf, 0, h1v, h2v, st))) return rc;

    LossArgs la{};
    la.dist_kind = b->dist_kind, la.P = P, la.M = R;
    la.out_pi = out_pi, la.out_vf = out_vf;
    la.actions = b->actions, la.logp_old = b->logp, la.advantages = b->advantages;
    la.returns = b->returns, la.rows = map.rows, la.row_begin = map.row_begin;
    la.T = b->T, la.N = b->N, la.hp = *hp;
    la.inv_denom = (float)((double)hp->loss_scale / denom);
    la.dout_pi = dout_pi, la.dout_vf = dout_vf;
    la.gb3_pi = (float*)g->pi_b3, la.gb3_vf = (float*)g->vf_b3, la.sums = loss_sums;
    if ((rc = launch_ppo_loss(la, st))) return rc;

    for (int which = 0; which < 2; ++which) {
      float* h1 = which ? h1v : h1p;
      float* h2 = which ? h2v : h2p;
      const float* dout = which ? dout_vf : dout_pi;
      const int Pn = which ? 1 : P;
      const float* w2 = which ? m->vf_w2 : m->pi_w2;
      const float* w3 = which ? m->vf_w3 : m->pi_w3;
      float* gw1 = (float*)(which ? g->vf_w1 : g->pi_w1);
      float* gb1 = (float*)(which ? g->vf_b1 : g->pi_b1);
      float* gw2 = (float*)(which ? g->vf_w2 : g->pi_w2);
      float* gb2 = (float*)(which ? g->vf_b2 : g->pi_b2);
      float* gw3 = (float*)(which ? g->vf_w3 : g->pi_w3);
      // dz2 = (dout @ w3) * (h2 > 0)
      if ((rc = launch_head_bwd(h2, dout, R, H, Pn, w3, dz2, st))) return rc;
      // gw3[p][j] += sum_r dout[r][p] * h2[r][j]
      if ((rc = launch_thin_reduce(h2, R, H, dout, nullptr, Pn, gw3, H, 1, nullptr, st))) return rc;
      // gb2[j] += sum_r dz2[r][j]
      if ((rc = launch_thin_reduce(dz2, R, H, nullptr, nullptr, 0, nullptr, 0, 0, gb2, st))) return rc;
      // gw2[j][i] += sum_r dz2[r][j] * h1[r][i]   (split-K over rows)
      if ((rc = launch_sgemm(false, false, EPI_ATOMIC, dz2, h1, gw2, H, H, R, H, H, H, nullptr,
                             splits, st)))
        return rc;
      // dz1 = (dz2 @ w2) * (h1 > 0), in place over h1
      if ((rc = launch_sgemm(true, false, EPI_MASK_INPLACE, dz2, w2, h1, R, H, H, H, H, H, nullptr,
                             1, st)))
        return rc;
      // gw1[i][d] += sum_r dz1[r][i] * obs[r][d];  gb1[i] += sum_r dz1[r][i]
      if ((rc = launch_thin_reduce(h1, R, H, nullptr, &map, D, gw1, 1, D, gb1, st))) return rc;
    }
  }
  return RL8_OK;
}

}  // namespace rl8

using namespace rl8;

extern "C" int rl8_abi_version(void) { return 1; }
extern "C" const char* rl8_last_error(void) { return g_last_error; }

extern "C" int64_t rl8_mlp_forward_workspace(int32_t H, int64_t rows) {
  const int64_t simt = 2 * rows * (int64_t)H * 4, split = forward_x3_workspace();
  return simt > split ? simt : split;
}

extern "C" int rl8_mlp_forward(const rl8_model* model, int which, const float* obs,
                               int64_t obs_stride_r, int64_t obs_stride_d, int64_t rows,
                               float* out, int apply_tanh_log_std, int precision, void* workspace,
                               int64_t workspace_bytes, rl8_stream_t stream) {
  if (!model || !obs || !out || rows <= 0 || (which != 0 && which != 1)) return RL8_ERR_ARG;
  if (model->H != 256) return RL8_ERR_UNSUPPORTED;
  RowMap map{};
  map.obs = obs, map.mode = 0, map.stride_r = obs_stride_r, map.stride_d = obs_stride_d;
  map.D = model->D;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RL8_PREC_BF16)
    return mlp_forward_tc(model, which, map, rows, out, apply_tanh_log_std, workspace,
                          workspace_bytes, st);
  if (precision == RL8_PREC_FP32_TC)
    return mlp_forward_x3(model, which, map, rows, out, apply_tanh_log_std, workspace, workspace_bytes, st);
  if (precision != RL8_PREC_FP32) return RL8_ERR_ARG;
  if (!workspace || workspace_bytes < rl8_mlp_forward_workspace(model->H, rows))
    return RL8_ERR_WORKSPACE;
  float* h1 = (float*)workspace;
  float* h2 = h1 + rows * model->H;
  return mlp_forward_fp32(model, which, map, rows, out, apply_tanh_log_std, h1, h2, st);
}

extern "C" int64_t rl8_collect_workspace(const rl8_model* model, int64_t N, int32_t T,
                                         int precision) {
  if (!model) return RL8_ERR_ARG;
  if (precision == RL8_PREC_BF16) return collect_tc_workspace(model, N, T);
  if (precision == RL8_PREC_FP32_TC) return collect_x3_workspace(model, N, T);
  return 2 * N * (int64_t)model->H * 4 + N * 8 * 4;
}

extern "C" int rl8_collect(const rl8_model* model, const rl8_rollout* ro, int precision,
                           void* workspace, int64_t workspace_bytes, rl8_stream_t stream) {
  int rc = validate_rollout(model, ro);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RL8_PREC_FP32) return collect_fp32(model, ro, workspace, workspace_bytes, st);
  if (precision == RL8_PREC_BF16) return collect_tc(model, ro, workspace, workspace_bytes, st);
  if (precision == RL8_PREC_FP32_TC) return collect_x3(model, ro, workspace, workspace_bytes, st);
  return RL8_ERR_ARG;
}

extern "C" int64_t rl8_ppo_workspace(const rl8_model* model, int64_t max_rows, int precision) {
  if (!model || max_rows <= 0) return RL8_ERR_ARG;
  if (precision == RL8_PREC_BF16) return ppo_tc_workspace(model, max_rows);
  if (precision == RL8_PREC_FP32_TC) return ppo_x3_workspace(model, max_rows);
  const int64_t C = ppo_fp32_chunk(max_rows);
  return (5 * C * model->H + C * (2 * 8 + 2)) * 4;
}

extern "C" int rl8_ppo_minibatch(const rl8_model* model, const rl8_model* grads,
                                 const rl8_batch* batch, const int64_t* rows, int64_t row_begin,
                                 int64_t M, double mean_denominator, const rl8_ppo_hparams* hp,
                                 double* loss_sums, int precision, void* workspace,
                                 int64_t workspace_bytes, rl8_stream_t stream) {
  if (!model || !grads || !batch || !hp || !loss_sums || M <= 0 || mean_denominator <= 0)
    return RL8_ERR_ARG;
  if (!batch->obs || !batch->actions || !batch->logp || !batch->advantages || !batch->returns)
    return RL8_ERR_ARG;
  if (model->H != 256 || model->P < 2 || model->P > 8) return RL8_ERR_UNSUPPORTED;
  if (batch->dist_kind != RL8_DIST_CATEGORICAL && model->P != 2) return RL8_ERR_UNSUPPORTED;
  if (batch->dist_kind == RL8_DIST_SQUASHED_NORMAL && hp->entropy_coeff != 0.0f)
    return RL8_ERR_UNSUPPORTED;  // SquashedNormal.entropy raises (distributions.py:153-157)
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RL8_PREC_FP32)
    return ppo_minibatch_fp32(model, grads, batch, rows, row_begin, M, mean_denominator, hp,
                              loss_sums, workspace, workspace_bytes, st);
  if (precision == RL8_PREC_BF16)
    return ppo_minibatch_tc(model, grads, batch, rows, row_begin, M, mean_denominator, hp,
                            loss_sums, workspace, workspace_bytes, st);
  if (precision == RL8_PREC_FP32_TC)
    return ppo_minibatch_x3(model, grads, batch, rows, row_begin, M, mean_denominator, hp,
                            loss_sums, workspace, workspace_bytes, st);
  return RL8_ERR_ARG;
}

static int ppo_losses_impl(int log_std_direct, int dist_kind, const float* features, int32_t P, const float* values,
                           const void* actions, const float* logp_old, const float* advantages,
                              const float* returns, int64_t B, double mean_denominator,
                              const rl8_ppo_hparams* hp, double* loss_sums, float* d_features,
                              float* d_values, rl8_stream_t stream) {
  if (!features || !values || !actions || !logp_old || !advantages || !returns || !hp ||
      !loss_sums || B <= 0 || mean_denominator <= 0)
    return RL8_ERR_ARG;
  if (dist_kind != RL8_DIST_CATEGORICAL && P != 2) return RL8_ERR_UNSUPPORTED;
  if (dist_kind == RL8_DIST_SQUASHED_NORMAL && hp->entropy_coeff != 0.0f) return RL8_ERR_UNSUPPORTED;
  LossArgs la{};
  la.dist_kind = dist_kind, la.P = P, la.M = B;
  la.out_pi = features, la.out_vf = values;
  la.actions = actions, la.logp_old = logp_old, la.advantages = advantages, la.returns = returns;
  la.rows = nullptr, la.row_begin = 0, la.T = 1, la.N = B;  // row g -> element g
  la.hp = *hp;
  la.inv_denom = (float)((double)hp->loss_scale / mean_denominator);
  la.dout_pi = d_features, la.dout_vf = d_values;
  la.gb3_pi = nullptr, la.gb3_vf = nullptr, la.sums = loss_sums;
  la.log_std_direct = log_std_direct;
  return launch_ppo_loss(la, (cudaStream_t)stream);
}

extern "C" int rl8_ppo_losses(int dist_kind, const float* features, int32_t P, const float* values,
                              const void* actions, const float* logp_old, const float* advantages,
                              const float* returns, int64_t B, double mean_denominator,
                              const rl8_ppo_hparams* hp, double* loss_sums, float* d_features,
                              float* d_values, rl8_stream_t stream) {
  return ppo_losses_impl(0, dist_kind, features, P, values, actions, logp_old, advantages, returns, B,
                         mean_denominator, hp, loss_sums, d_features, d_values, stream);
}

extern "C" int rl8_ppo_losses_direct(int dist_kind, const float* features, int32_t P, const float* values,
                                     const void* actions, const float* logp_old, const float* advantages,
                                     const float* returns, int64_t B, double mean_denominator,
                                     const rl8_ppo_hparams* hp, double* loss_sums, float* d_features,
                                     float* d_values, rl8_stream_t stream) {
  return ppo_losses_impl(1, dist_kind, features, P, values, actions, logp_old, advantages, returns, B,
                         mean_denominator, hp, loss_sums, d_features, d_values, stream);
}
