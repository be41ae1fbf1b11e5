// Standalone action-distribution kernels: Distribution.sample / deterministic_sample /
// logp / entropy (src/rl8/distributions.py:98-170).  One row per thread.
#include "dist.cuh"

namespace rl8 {

template <int P>
__global__ void __launch_bounds__(256)
categorical_sample_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                          int deterministic, long long* __restrict__ action,
                          float* __restrict__ logp, int64_t B) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B;
       r += (int64_t)gridDim.x * blockDim.x) {
    float l[P], norm[P], probs[P];
#pragma unroll
    for (int k = 0; k < P; ++k) l[k] = logits[r * P + k];
    categorical_norm<P>(l, norm, probs);
    int a;
    if (deterministic) {
      a = categorical_mode<P>(probs);
    } else {
      float q[P];
#pragma unroll
      for (int k = 0; k < P; ++k) q[k] = noise[r * P + k];
      a = categorical_sample<P>(probs, q);
    }
    action[r] = a;
    if (logp) {
      float lp = norm[0];
#pragma unroll
      for (int k = 1; k < P; ++k) lp = (a == k) ? norm[k] : lp;
      logp[r] = lp;
    }
  }
}

template <int P>
__global__ void __launch_bounds__(256)
categorical_logp_entropy_kernel(const float* __restrict__ logits,
                                const long long* __restrict__ action, float* __restrict__ logp,
                                float* __restrict__ entropy, int64_t B) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B;
       r += (int64_t)gridDim.x * blockDim.x) {
    float l[P], norm[P], probs[P];
#pragma unroll
    for (int k = 0; k < P; ++k) l[k] = logits[r * P + k];
    categorical_norm<P>(l, norm, probs);
    if (logp) {
      int a = (int)action[r];
      float lp = norm[0];
#pragma unroll
      for (int k = 1; k < P; ++k) lp = (a == k) ? norm[k] : lp;
      logp[r] = lp;
    }
    if (entropy) entropy[r] = categorical_entropy<P>(norm, probs);
  }
}

__global__ void __launch_bounds__(256)
normal_sample_kernel(const float* __restrict__ feats, const float* __restrict__ noise,
                     int deterministic, int squashed, float* __restrict__ action,
                     float* __restrict__ logp, int64_t B) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B;
       r += (int64_t)gridDim.x * blockDim.x) {
    float mean = feats[2 * r], scale = expf(feats[2 * r + 1]);
    // torch.normal(loc, scale) = z * scale, then + loc (two roundings, Appendix A.10)
    float x = deterministic ? mean : add(mul(noise[r], scale), mean);
    if (squashed) x = tanhf(x);
    action[r] = x;
    if (logp) logp[r] = squashed ? squashed_logp(mean, scale, x, nullptr) : normal_logp(mean, scale, x);
  }
}

__global__ void __launch_bounds__(256)
normal_logp_entropy_kernel(const float* __restrict__ feats, const float* __restrict__ action,
                           int squashed, float* __restrict__ logp, float* __restrict__ entropy,
                           int64_t B) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < B;
       r += (int64_t)gridDim.x * blockDim.x) {
    float mean = feats[2 * r], scale = expf(feats[2 * r + 1]);
    if (logp) {
      float x = action[r];
      logp[r] = squashed ? squashed_logp(mean, scale, x, nullptr) : normal_logp(mean, scale, x);
    }
    if (entropy) entropy[r] = normal_entropy(scale);
  }
}

}  // namespace rl8

using namespace rl8;

#define RL8_DISPATCH_P(P, ...)            \
  switch (P) {                            \
    case 2: { constexpr int kP = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int kP = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int kP = 4; __VA_ARGS__; } break; \
    case 5: { constexpr int kP = 5; __VA_ARGS__; } break; \
    case 6: { constexpr int kP = 6; __VA_ARGS__; } break; \
    case 7: { constexpr int kP = 7; __VA_ARGS__; } break; \
    case 8: { constexpr int kP = 8; __VA_ARGS__; } break; \
    default: return RL8_ERR_UNSUPPORTED;  \
  }

extern "C" int rl8_dist_sample(int dist_kind, const float* features, int32_t P, const float* noise,
                               int deterministic, void* action, float* logp, int64_t B,
                               rl8_stream_t stream) {
  if (!features || !action || B <= 0 || (!deterministic && !noise)) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(B, 256);
  if (dist_kind == RL8_DIST_CATEGORICAL) {
    RL8_DISPATCH_P(P, categorical_sample_kernel<kP><<<grid, 256, 0, st>>>(
                          features, noise, deterministic, (long long*)action, logp, B));
  } else if (dist_kind == RL8_DIST_NORMAL || dist_kind == RL8_DIST_SQUASHED_NORMAL) {
    if (P != 2) return RL8_ERR_UNSUPPORTED;
    normal_sample_kernel<<<grid, 256, 0, st>>>(features, noise, deterministic,
                                               dist_kind == RL8_DIST_SQUASHED_NORMAL,
                                               (float*)action, logp, B);
  } else {
    return RL8_ERR_ARG;
  }
  return check_launch("rl8_dist_sample");
}

extern "C" int rl8_dist_logp_entropy(int dist_kind, const float* features, int32_t P,
                                     const void* action, float* logp, float* entropy, int64_t B,
                                     rl8_stream_t stream) {
  if (!features || B <= 0 || (logp && !action)) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(B, 256);
  if (dist_kind == RL8_DIST_CATEGORICAL) {
    RL8_DISPATCH_P(P, categorical_logp_entropy_kernel<kP><<<grid, 256, 0, st>>>(
                          features, (const long long*)action, logp, entropy, B));
  } else if (dist_kind == RL8_DIST_NORMAL || dist_kind == RL8_DIST_SQUASHED_NORMAL) {
    if (P != 2) return RL8_ERR_UNSUPPORTED;
    if (entropy && dist_kind == RL8_DIST_SQUASHED_NORMAL) return RL8_ERR_UNSUPPORTED;
    normal_logp_entropy_kernel<<<grid, 256, 0, st>>>(features, (const float*)action,
                                                     dist_kind == RL8_DIST_SQUASHED_NORMAL, logp,
                                                     entropy, B);
  } else {
    return RL8_ERR_ARG;
  }
  return check_launch("rl8_dist_logp_entropy");
}
