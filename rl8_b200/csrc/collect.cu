// collect(): the T-step rollout (src/rl8/algorithms/_feedforward.py:359-408).
//
// fp32 path: per step  layer1 -> SGEMM(+bias,ReLU) -> head -> [sample + logp + env.step +
// buffer writes] (one fused tail kernel, one env per lane, env state in registers), then one
// batched value pass over all T+1 observation slabs.  The value network never influences
// the trajectory, so it is hoisted out of the step loop.
// bf16 path: see collect_tc.cu (persistent tcgen05 kernel).
#include "dist.cuh"
#include "mlp_fp32.cuh"

namespace rl8 {

// One env per thread: features -> action/logp -> env transition -> buffer slabs.
template <int KIND, int P>
__global__ void __launch_bounds__(256)
sample_step_store_kernel(rl8_env_cfg cfg, int dist_kind, int deterministic,
                         const float* __restrict__ feat, const float* __restrict__ noise,
                         float* __restrict__ state, float* __restrict__ obs_next,
                         void* __restrict__ action_out, float* __restrict__ logp_out,
                         float* __restrict__ reward_out, const float* __restrict__ rdr_prev,
                         float* __restrict__ rdr_next, float gamma, int64_t N,
                         uint32_t* __restrict__ omax_next, uint32_t* __restrict__ omax_all) {
  using Tr = EnvTraits<KIND>;
  float omax = 0.0f;  // max |obs| this thread writes (the fp16-piece forward derives its H1 scale from it)
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    float o[P];
#pragma unroll
    for (int k = 0; k < P; ++k) o[k] = feat[n * P + k];
    float act, lp;
    if constexpr (Tr::discrete) {
      float norm[P], probs[P];
      categorical_norm<P>(o, norm, probs);
      int a;
      if (deterministic) {
        a = categorical_mode<P>(probs);
      } else {
        float q[P];
#pragma unroll
        for (int k = 0; k < P; ++k) q[k] = noise[n * P + k];
        a = categorical_sample<P>(probs, q);
      }
      lp = norm[0];
#pragma unroll
      for (int k = 1; k < P; ++k) lp = (a == k) ? norm[k] : lp;
      act = (float)a;
      ((long long*)action_out)[n] = a;
    } else {
      const float mean = o[0], scale = expf(o[1]);
      float x = deterministic ? mean : add(mul(noise[n], scale), mean);
      if (dist_kind == RL8_DIST_SQUASHED_NORMAL) {
        x = tanhf(x);
        lp = squashed_logp(mean, scale, x, nullptr);
      } else {
        lp = normal_logp(mean, scale, x);
      }
      act = x;
      ((float*)action_out)[n] = x;
    }
    logp_out[n] = lp;
    float s[Tr::S], ob[Tr::D], r;
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) s[i] = state[(int64_t)i * N + n];
    env_step<KIND>(cfg, s, act, ob, r);
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) state[(int64_t)i * N + n] = s[i];
#pragma unroll
    for (int i = 0; i < Tr::D; ++i) obs_next[(int64_t)i * N + n] = ob[i], omax = fmaxf(omax, fabsf(ob[i]));
    reward_out[n] = r;
    // rdr[t+1] = gamma * rdr[t] + reward   (:378-383)
    if (rdr_next) rdr_next[n] = add(mul(gamma, rdr_prev[n]), r);
  }
  if (omax_next) {  // (uniform) max |obs| of the slab just written, as float bits; NaNs are skipped by fmaxf
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) omax = fmaxf(omax, __shfl_xor_sync(0xffffffffu, omax, o));
    if ((threadIdx.x & 31) == 0 && omax > 0.0f) {
      atomicMax(omax_next, __float_as_uint(omax));
      atomicMax(omax_all, __float_as_uint(omax));
    }
  }
}

template <int KIND, int P>
static int launch_tail(const rl8_rollout* ro, int t, const float* feat, cudaStream_t st, uint32_t* omax_next,
                       uint32_t* omax_all) {
  using Tr = EnvTraits<KIND>;
  const int64_t N = ro->N;
  const size_t asz = Tr::discrete ? 8 : 4;
  const float* noise = nullptr;
  if (!ro->deterministic) noise = ro->noise + (int64_t)t * N * (Tr::discrete ? P : 1);
  sample_step_store_kernel<KIND, P><<<grid_for(N, 256), 256, 0, st>>>(
      ro->env_cfg, ro->dist_kind, ro->deterministic, feat, noise, ro->env_state,
      ro->obs + (int64_t)(t + 1) * Tr::D * N, (char*)ro->actions + (size_t)t * N * asz,
      ro->logp + (int64_t)t * N, ro->rewards + (int64_t)t * N,
      ro->rdr ? ro->rdr + (int64_t)t * N : nullptr,
      ro->rdr ? ro->rdr + (int64_t)(t + 1) * N : nullptr, ro->gamma, N, omax_next, omax_all);
  return check_launch("sample_step_store");
}

int env_dims(int env_kind, int* S, int* D, int* P_discrete) {
  switch (env_kind) {
    case RL8_ENV_DISCRETE_DUMMY: *S = 1, *D = 1, *P_discrete = 2; return RL8_OK;
    case RL8_ENV_CONTINUOUS_DUMMY: *S = 1, *D = 1, *P_discrete = 0; return RL8_OK;
    case RL8_ENV_CARTPOLE: *S = 4, *D = 5, *P_discrete = 3; return RL8_OK;
    case RL8_ENV_MOUNTAIN_CAR: *S = 2, *D = 2, *P_discrete = 3; return RL8_OK;
    case RL8_ENV_PENDULUM: *S = 2, *D = 3, *P_discrete = 0; return RL8_OK;
  }
  return RL8_ERR_ARG;
}

int validate_rollout_dims(int mD, int mH, int mP, const rl8_rollout* ro) {
  if (!ro || !ro->env_state || !ro->obs || !ro->actions || !ro->logp || !ro->values ||
      !ro->rewards || ro->N <= 0 || ro->T <= 0)
    return RL8_ERR_ARG;
  if (!ro->deterministic && !ro->noise) return RL8_ERR_ARG;
  int S, D, Pd;
  if (env_dims(ro->env_kind, &S, &D, &Pd)) return RL8_ERR_ARG;
  if (mD != D || mH != 256) return RL8_ERR_UNSUPPORTED;
  if (Pd) {
    if (ro->dist_kind != RL8_DIST_CATEGORICAL || mP != Pd) return RL8_ERR_UNSUPPORTED;
  } else {
    if (ro->dist_kind == RL8_DIST_CATEGORICAL || mP != 2) return RL8_ERR_UNSUPPORTED;
  }
  return RL8_OK;
}

int validate_rollout(const rl8_model* model, const rl8_rollout* ro) {
  if (!model) return RL8_ERR_ARG;
  return validate_rollout_dims(model->D, model->H, model->P, ro);
}

// omax_next / omax_all (both or neither): receive the bits of max |obs| of slab t + 1 (atomicMax)
int collect_tail(const rl8_rollout* ro, int t, const float* feat, cudaStream_t st, uint32_t* omax_next = nullptr,
                 uint32_t* omax_all = nullptr) {
  if ((omax_next == nullptr) != (omax_all == nullptr)) return RL8_ERR_ARG;
  switch (ro->env_kind) {
    case RL8_ENV_DISCRETE_DUMMY: return launch_tail<RL8_ENV_DISCRETE_DUMMY, 2>(ro, t, feat, st, omax_next, omax_all);
    case RL8_ENV_CONTINUOUS_DUMMY: return launch_tail<RL8_ENV_CONTINUOUS_DUMMY, 2>(ro, t, feat, st, omax_next, omax_all);
    case RL8_ENV_CARTPOLE: return launch_tail<RL8_ENV_CARTPOLE, 3>(ro, t, feat, st, omax_next, omax_all);
    case RL8_ENV_MOUNTAIN_CAR: return launch_tail<RL8_ENV_MOUNTAIN_CAR, 3>(ro, t, feat, st, omax_next, omax_all);
    case RL8_ENV_PENDULUM: return launch_tail<RL8_ENV_PENDULUM, 2>(ro, t, feat, st, omax_next, omax_all);
  }
  return RL8_ERR_ARG;
}

// One network forward on `rows` rows, fp32.  ws: h1[rows][H], h2[rows][H].
int mlp_forward_fp32(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out,
                     int tanh_col1, float* h1, float* h2, cudaStream_t st) {
  const float* w1 = which ? m->vf_w1 : m->pi_w1;
  const float* b1 = which ? m->vf_b1 : m->pi_b1;
  const float* w2 = which ? m->vf_w2 : m->pi_w2;
  const float* b2 = which ? m->vf_b2 : m->pi_b2;
  const float* w3 = which ? m->vf_w3 : m->pi_w3;
  const float* b3 = which ? m->vf_b3 : m->pi_b3;
  const int P = which ? 1 : m->P;
  int rc = launch_layer1_fwd(map, rows, m->D, m->H, w1, b1, h1, st);
  if (rc) return rc;
  rc = launch_sgemm(true, true, EPI_BIAS_RELU, h1, w2, h2, rows, m->H, m->H, m->H, m->H, m->H, b2,
                    1, st);
  if (rc) return rc;
  return launch_head_fwd(h2, rows, m->H, P, w3, b3, out, tanh_col1, st);
}

int collect_fp32(const rl8_model* model, const rl8_rollout* ro, void* workspace,
                 int64_t workspace_bytes, cudaStream_t st) {
  const int64_t N = ro->N;
  const int H = model->H, D = model->D;
  const int64_t need = 2 * N * H * 4 + N * kMaxP * 4;
  if (!workspace || workspace_bytes < need) return RL8_ERR_WORKSPACE;
  float* h1 = (float*)workspace;
  float* h2 = h1 + N * H;
  float* feat = h2 + N * H;
  const bool continuous = ro->dist_kind != RL8_DIST_CATEGORICAL;
  RowMap map{};
  map.mode = 0, map.stride_r = 1, map.stride_d = N;
  for (int t = 0; t < ro->T; ++t) {
    map.obs = ro->obs + (int64_t)t * D * N;
    int rc = mlp_forward_fp32(model, 0, map, N, feat, continuous, h1, h2, st);
    if (rc) return rc;
    rc = collect_tail(ro, t, feat, st);
    if (rc) return rc;
  }
  // values for all T+1 observation slabs (bootstrap value included, :396-408)
  for (int t = 0; t <= ro->T; ++t) {
    map.obs = ro->obs + (int64_t)t * D * N;
    int rc = mlp_forward_fp32(model, 1, map, N, ro->values + (int64_t)t * N, 0, h1, h2, st);
    if (rc) return rc;
  }
  return RL8_OK;
}

}  // namespace rl8
