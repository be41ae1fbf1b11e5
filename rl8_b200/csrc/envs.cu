// Standalone Env.reset / Env.step kernels (the Env protocol boundary, src/rl8/env.py:100-128).
//
// HBM-bound streaming kernels: one env per lane, VEC=4 envs per thread through 128-bit
// loads/stores on the SoA state / obs / reward rows, state held in registers between the
// load and the store.  Algorithmic bytes per env-step: cartpole 64 B,
// mountain car 36 B, pendulum 36 B, discrete dummy 24 B, continuous dummy 20 B.
#include "envs.cuh"

namespace rl8 {

template <int VEC>
struct VecF;
template <>
struct VecF<1> {
  float v[1];
  __device__ __forceinline__ void load(const float* p) { v[0] = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};
template <>
struct VecF<4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = ld_stream4(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    st_stream4(p, make_float4(v[0], v[1], v[2], v[3]));
  }
};

template <int KIND, int VEC>
__global__ void __launch_bounds__(256)
env_step_kernel(rl8_env_cfg cfg, float* __restrict__ state, const void* __restrict__ action,
                float* __restrict__ obs, int64_t obs_stride_n, int64_t obs_stride_d,
                float* __restrict__ reward, int64_t N) {
  using Tr = EnvTraits<KIND>;
  const int64_t groups = N / VEC;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n0 = g * VEC;
    VecF<VEC> s[Tr::S], o[Tr::D], r;
    float a[VEC];
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) s[i].load(state + (int64_t)i * N + n0);
    if constexpr (Tr::discrete) {
      const long long* ap = (const long long*)action + n0;
      if constexpr (VEC == 4) {
        longlong2 a01 = *(const longlong2*)ap, a23 = *(const longlong2*)(ap + 2);
        a[0] = (float)a01.x, a[1] = (float)a01.y, a[2] = (float)a23.x, a[3] = (float)a23.y;
      } else {
        a[0] = (float)ap[0];
      }
    } else {
      VecF<VEC> av;
      av.load((const float*)action + n0);
#pragma unroll
      for (int j = 0; j < VEC; ++j) a[j] = av.v[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float sj[Tr::S], oj[Tr::D], rj;
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) sj[i] = s[i].v[j];
      env_step<KIND>(cfg, sj, a[j], oj, rj);
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) s[i].v[j] = sj[i];
#pragma unroll
      for (int i = 0; i < Tr::D; ++i) o[i].v[j] = oj[i];
      r.v[j] = rj;
    }
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) s[i].store(state + (int64_t)i * N + n0);
    if (VEC == 1 || obs_stride_n == 1) {
#pragma unroll
      for (int i = 0; i < Tr::D; ++i) o[i].store(obs + n0 * obs_stride_n + (int64_t)i * obs_stride_d);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
#pragma unroll
        for (int i = 0; i < Tr::D; ++i)
          obs[(n0 + j) * obs_stride_n + (int64_t)i * obs_stride_d] = o[i].v[j];
    }
    r.store(reward + n0);
  }
}

template <int KIND, int VEC>
__global__ void __launch_bounds__(256)
env_reset_kernel(rl8_env_cfg cfg, const float* __restrict__ noise, float* __restrict__ state,
                 float* __restrict__ obs, int64_t obs_stride_n, int64_t obs_stride_d, int64_t N) {
  using Tr = EnvTraits<KIND>;
  const int64_t groups = N / VEC;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n0 = g * VEC;
    VecF<VEC> z[Tr::S], s[Tr::S], o[Tr::D];
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) z[i].load(noise + (int64_t)i * N + n0);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float zj[Tr::S], sj[Tr::S], oj[Tr::D];
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) zj[i] = z[i].v[j];
      env_reset<KIND>(cfg, zj, sj);
      env_observe<KIND>(sj, oj);
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) s[i].v[j] = sj[i];
#pragma unroll
      for (int i = 0; i < Tr::D; ++i) o[i].v[j] = oj[i];
    }
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) s[i].store(state + (int64_t)i * N + n0);
    if (VEC == 1 || obs_stride_n == 1) {
#pragma unroll
      for (int i = 0; i < Tr::D; ++i) o[i].store(obs + n0 * obs_stride_n + (int64_t)i * obs_stride_d);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j)
#pragma unroll
        for (int i = 0; i < Tr::D; ++i)
          obs[(n0 + j) * obs_stride_n + (int64_t)i * obs_stride_d] = o[i].v[j];
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(256)
env_observe_kernel(const float* __restrict__ state, float* __restrict__ obs, int64_t obs_stride_n,
                   int64_t obs_stride_d, int64_t N) {
  using Tr = EnvTraits<KIND>;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    float s[Tr::S], o[Tr::D];
#pragma unroll
    for (int i = 0; i < Tr::S; ++i) s[i] = state[(int64_t)i * N + n];
    env_observe<KIND>(s, o);
#pragma unroll
    for (int i = 0; i < Tr::D; ++i) obs[n * obs_stride_n + (int64_t)i * obs_stride_d] = o[i];
  }
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

template <int KIND>
static int launch_step(const rl8_env_cfg* cfg, float* state, const void* action, float* obs,
                       int64_t sn, int64_t sd, float* reward, int64_t N, cudaStream_t st) {
  bool vec = (N % 4 == 0) && aligned16(state) && aligned16(action) && aligned16(reward) &&
             aligned16(obs) && (sn != 1 || sd % 4 == 0);
  if (vec) {
    int grid = grid_for(N / 4, 256);
    env_step_kernel<KIND, 4><<<grid, 256, 0, st>>>(*cfg, state, action, obs, sn, sd, reward, N);
  } else {
    int grid = grid_for(N, 256);
    env_step_kernel<KIND, 1><<<grid, 256, 0, st>>>(*cfg, state, action, obs, sn, sd, reward, N);
  }
  return check_launch("rl8_env_step");
}

template <int KIND>
static int launch_reset(const rl8_env_cfg* cfg, const float* noise, float* state, float* obs,
                        int64_t sn, int64_t sd, int64_t N, cudaStream_t st) {
  bool vec = (N % 4 == 0) && aligned16(state) && aligned16(noise) && aligned16(obs) &&
             (sn != 1 || sd % 4 == 0);
  if (vec) {
    int grid = grid_for(N / 4, 256);
    env_reset_kernel<KIND, 4><<<grid, 256, 0, st>>>(*cfg, noise, state, obs, sn, sd, N);
  } else {
    int grid = grid_for(N, 256);
    env_reset_kernel<KIND, 1><<<grid, 256, 0, st>>>(*cfg, noise, state, obs, sn, sd, N);
  }
  return check_launch("rl8_env_reset");
}

}  // namespace rl8

using namespace rl8;

extern "C" int rl8_env_reset(int env_kind, const rl8_env_cfg* cfg, const float* noise,
                             float* state, float* obs, int64_t obs_stride_n, int64_t obs_stride_d,
                             int64_t N, rl8_stream_t stream) {
  if (!cfg || !noise || !state || !obs || N <= 0) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (env_kind) {
    case RL8_ENV_DISCRETE_DUMMY:
      return launch_reset<RL8_ENV_DISCRETE_DUMMY>(cfg, noise, state, obs, obs_stride_n, obs_stride_d, N, st);
    case RL8_ENV_CONTINUOUS_DUMMY:
      return launch_reset<RL8_ENV_CONTINUOUS_DUMMY>(cfg, noise, state, obs, obs_stride_n, obs_stride_d, N, st);
    case RL8_ENV_CARTPOLE:
      return launch_reset<RL8_ENV_CARTPOLE>(cfg, noise, state, obs, obs_stride_n, obs_stride_d, N, st);
    case RL8_ENV_MOUNTAIN_CAR:
      return launch_reset<RL8_ENV_MOUNTAIN_CAR>(cfg, noise, state, obs, obs_stride_n, obs_stride_d, N, st);
    case RL8_ENV_PENDULUM:
      return launch_reset<RL8_ENV_PENDULUM>(cfg, noise, state, obs, obs_stride_n, obs_stride_d, N, st);
  }
  return RL8_ERR_ARG;
}

extern "C" int rl8_env_step(int env_kind, const rl8_env_cfg* cfg, float* state, const void* action,
                            float* obs, int64_t obs_stride_n, int64_t obs_stride_d, float* reward,
                            int64_t N, rl8_stream_t stream) {
  if (!cfg || !state || !action || !obs || !reward || N <= 0) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (env_kind) {
    case RL8_ENV_DISCRETE_DUMMY:
      return launch_step<RL8_ENV_DISCRETE_DUMMY>(cfg, state, action, obs, obs_stride_n, obs_stride_d, reward, N, st);
    case RL8_ENV_CONTINUOUS_DUMMY:
      return launch_step<RL8_ENV_CONTINUOUS_DUMMY>(cfg, state, action, obs, obs_stride_n, obs_stride_d, reward, N, st);
    case RL8_ENV_CARTPOLE:
      return launch_step<RL8_ENV_CARTPOLE>(cfg, state, action, obs, obs_stride_n, obs_stride_d, reward, N, st);
    case RL8_ENV_MOUNTAIN_CAR:
      return launch_step<RL8_ENV_MOUNTAIN_CAR>(cfg, state, action, obs, obs_stride_n, obs_stride_d, reward, N, st);
    case RL8_ENV_PENDULUM:
      return launch_step<RL8_ENV_PENDULUM>(cfg, state, action, obs, obs_stride_n, obs_stride_d, reward, N, st);
  }
  return RL8_ERR_ARG;
}

extern "C" int rl8_env_observe(int env_kind, const float* state, float* obs, int64_t obs_stride_n,
                               int64_t obs_stride_d, int64_t N, rl8_stream_t stream) {
  if (!state || !obs || N <= 0) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(N, 256);
  switch (env_kind) {
    case RL8_ENV_DISCRETE_DUMMY:
      env_observe_kernel<RL8_ENV_DISCRETE_DUMMY><<<grid, 256, 0, st>>>(state, obs, obs_stride_n, obs_stride_d, N);
      break;
    case RL8_ENV_CONTINUOUS_DUMMY:
      env_observe_kernel<RL8_ENV_CONTINUOUS_DUMMY><<<grid, 256, 0, st>>>(state, obs, obs_stride_n, obs_stride_d, N);
      break;
    case RL8_ENV_CARTPOLE:
      env_observe_kernel<RL8_ENV_CARTPOLE><<<grid, 256, 0, st>>>(state, obs, obs_stride_n, obs_stride_d, N);
      break;
    case RL8_ENV_MOUNTAIN_CAR:
      env_observe_kernel<RL8_ENV_MOUNTAIN_CAR><<<grid, 256, 0, st>>>(state, obs, obs_stride_n, obs_stride_d, N);
      break;
    case RL8_ENV_PENDULUM:
      env_observe_kernel<RL8_ENV_PENDULUM><<<grid, 256, 0, st>>>(state, obs, obs_stride_n, obs_stride_d, N);
      break;
    default:
      return RL8_ERR_ARG;
  }
  return check_launch("rl8_env_observe");
}
