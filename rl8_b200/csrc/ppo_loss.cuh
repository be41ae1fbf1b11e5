// PPO clipped policy / value / entropy losses with KL and the hand-derived backward
// (src/rl8/nn/functional.py:259-363, src/rl8/algorithms/_feedforward.py:552-559).
#pragma once
#include "mlp_fp32.cuh"

namespace rl8 {

struct LossArgs {
  int dist_kind, P;
  int64_t M;           // rows in this launch
  const float* out_pi; // [M][P] logits | {mean, log_std}
  const float* out_vf; // [M]
  // horizon-major batch fields, gathered through (rows | row_begin) like RowMap mode 1
  const void* actions;
  const float* logp_old;
  const float* advantages;
  const float* returns;
  const int64_t* rows;
  int64_t row_begin;
  int32_t T;
  int64_t N;
  rl8_ppo_hparams hp;
  float inv_denom;     // loss_scale / mean_denominator
  float* dout_pi;      // [M][P]  d(loss)/d(head output, pre-tanh for log_std)
  float* dout_vf;      // [M]
  float* gb3_pi;       // [P]  += sum_r dout_pi
  float* gb3_vf;       // [1]
  double* sums;        // [5] += entropy, policy, vf, kl, count
  int log_std_direct;  // continuous head: dout_pi[:, 1] is d/d(log_std) itself, not d/d(pre-tanh output)
  // steps > 1: the launch covers `steps` blocks of M rows (the time steps of a TBPTT chunk); block k reads / writes
  // out_pi + k * pi_stride, out_vf + k * vf_stride (dout_* likewise) and rows + k * rows_stride.  0 means 1.
  int steps;
  int64_t pi_stride, vf_stride, rows_stride;
};

int launch_ppo_loss(const LossArgs& a, cudaStream_t st);

}  // namespace rl8
