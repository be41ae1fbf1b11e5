// Standalone PPO loss + output-gradient kernel (fp32 path): one transition per thread.
#include "ppo_loss.cuh"
#include "ppo_loss_math.cuh"

namespace rl8 {

template <int P>
__global__ void __launch_bounds__(256) ppo_loss_kernel(LossArgs a) {
  __shared__ double red[32];
  double s_ent = 0, s_pol = 0, s_vf = 0, s_kl = 0;
  float gb[P], gbv = 0.0f;
#pragma unroll
  for (int k = 0; k < P; ++k) gb[k] = 0.0f;
  const bool discrete = a.dist_kind == RL8_DIST_CATEGORICAL;
  const int steps = a.steps > 1 ? a.steps : 1;
  for (int64_t ra = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ra < a.M * steps;
       ra += (int64_t)gridDim.x * blockDim.x) {
    const int64_t blk = steps > 1 ? ra / a.M : 0, r = ra - blk * a.M;
    const float* out_pi = a.out_pi + blk * a.pi_stride;
    const int64_t g = a.rows ? a.rows[blk * a.rows_stride + r] : a.row_begin + r;
    const int64_t n = g / a.T, t = g - n * a.T;
    const int64_t idx = t * a.N + n;
    float o[P], d_o[P], d_v;
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = out_pi[r * P + k];
    const float act = discrete ? (float)((const long long*)a.actions)[idx]
                               : ((const float*)a.actions)[idx];
    RowLoss L = ppo_row<P>(a.dist_kind, o, a.out_vf[blk * a.vf_stride + r], act, a.logp_old[idx], a.advantages[idx],
                           a.returns[idx], a.hp, a.inv_denom, d_o, &d_v, !a.log_std_direct);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if (a.dout_pi) a.dout_pi[blk * a.pi_stride + r * P + k] = d_o[k];
      gb[k] += d_o[k];
    }
    if (a.dout_vf) a.dout_vf[blk * a.vf_stride + r] = d_v;
    gbv += d_v;
    s_ent += L.entropy, s_pol += L.policy, s_vf += L.vf, s_kl += L.kl;
  }
  double v[4] = {s_ent, s_pol, s_vf, s_kl};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = block_sum(v[i], red);
    if (threadIdx.x == 0) atomicAdd(a.sums + i, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.sums + 4, (double)(a.M * steps));
  if (a.gb3_pi) {
#pragma unroll
    for (int k = 0; k < P; ++k) {
      double s = block_sum((double)gb[k], red);
      if (threadIdx.x == 0) atomicAdd(a.gb3_pi + k, (float)s);
    }
  }
  if (a.gb3_vf) {
    double s = block_sum((double)gbv, red);
    if (threadIdx.x == 0) atomicAdd(a.gb3_vf, (float)s);
  }
}

int launch_ppo_loss(const LossArgs& a, cudaStream_t st) {
  int grid = grid_for(a.M * (a.steps > 1 ? a.steps : 1), 256, 4, 2);
  switch (a.P) {
    case 2: ppo_loss_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 3: ppo_loss_kernel<3><<<grid, 256, 0, st>>>(a); break;
    case 4: ppo_loss_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 5: ppo_loss_kernel<5><<<grid, 256, 0, st>>>(a); break;
    case 6: ppo_loss_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 7: ppo_loss_kernel<7><<<grid, 256, 0, st>>>(a); break;
    case 8: ppo_loss_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: return RL8_ERR_UNSUPPORTED;
  }
  return check_launch("ppo_loss");
}

}  // namespace rl8
