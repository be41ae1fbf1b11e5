// Action-distribution math, one row per thread.  Op order follows
// torch.distributions as wrapped by src/rl8/distributions.py (see oracle/ppo_oracle.py
// `_Bound`).  No FMA contraction: explicit round-to-nearest intrinsics.
#pragma once
#include "envs.cuh"

namespace rl8 {

constexpr int kMaxP = 8;                          // widest policy head on the fused path
constexpr float kLogSqrt2Pi = 0.91893853320467267f;  // log(sqrt(2*pi))
constexpr float kF32Eps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps

// Normalised logits (logits - logsumexp) and probabilities as torch.distributions.Categorical
// builds them: logsumexp = log(sum(exp(x - max))) + max; probs = softmax(norm).
template <int P>
__device__ __forceinline__ void categorical_norm(const float* logits, float* norm, float* probs) {
  float m = logits[0];
#pragma unroll
  for (int k = 1; k < P; ++k) m = fmaxf(m, logits[k]);
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < P; ++k) s = add(s, expf(sub(logits[k], m)));
  float lse = add(logf(s), m);
  float m2 = -INFINITY;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    norm[k] = sub(logits[k], lse);
    m2 = fmaxf(m2, norm[k]);
  }
  float s2 = 0.0f;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    probs[k] = expf(sub(norm[k], m2));
    s2 = add(s2, probs[k]);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) probs[k] = dvd(probs[k], s2);
}

// multinomial(probs, 1) == argmax(probs / q), q ~ Exp(1), first index wins ties
// (SURVEY.md Appendix A.9).
template <int P>
__device__ __forceinline__ int categorical_sample(const float* probs, const float* q) {
  int best = 0;
  float bv = dvd(probs[0], q[0]);
#pragma unroll
  for (int k = 1; k < P; ++k) {
    float v = dvd(probs[k], q[k]);
    if (v > bv) bv = v, best = k;
  }
  return best;
}
template <int P>
__device__ __forceinline__ int categorical_mode(const float* probs) {
  int best = 0;
#pragma unroll
  for (int k = 1; k < P; ++k)
    if (probs[k] > probs[best]) best = k;
  return best;
}
template <int P>
__device__ __forceinline__ float categorical_entropy(const float* norm, const float* probs) {
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < P; ++k) s = add(s, mul(fmaxf(norm[k], -3.4028234663852886e38f), probs[k]));
  return -s;
}

// Normal(loc=mean, scale=exp(log_std)).log_prob(x)
__device__ __forceinline__ float normal_logp(float mean, float scale, float x) {
  float var = mul(scale, scale);
  float d = sub(x, mean);
  return sub(sub(dvd(-mul(d, d), mul(2.0f, var)), logf(scale)), kLogSqrt2Pi);
}
__device__ __forceinline__ float normal_entropy(float scale) {
  // 0.5 + 0.5*log(2*pi) + log(scale)
  return add(1.4189385332046727f, logf(scale));
}
// SquashedNormal.logp (src/rl8/distributions.py:159-167); `clamped` reports whether the
// inner clamp to [-100, 100] was active (gradient is zero there).
__device__ __forceinline__ float squashed_logp(float mean, float scale, float x, bool* inside) {
  float xc = clampf(x, add(-1.0f, kF32Eps), sub(1.0f, kF32Eps));
  float inv = mul(0.5f, sub(log1pf(xc), log1pf(-xc)));
  float lp = normal_logp(mean, scale, inv);
  if (inside) *inside = (lp >= -100.0f) && (lp <= 100.0f);
  lp = clampf(lp, -100.0f, 100.0f);
  return sub(lp, logf(add(sub(1.0f, mul(x, x)), kF32Eps)));
}
__device__ __forceinline__ float squashed_inverse(float x) {
  float xc = clampf(x, add(-1.0f, kF32Eps), sub(1.0f, kF32Eps));
  return mul(0.5f, sub(log1pf(xc), log1pf(-xc)));
}

}  // namespace rl8
