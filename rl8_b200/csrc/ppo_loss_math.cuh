// Per-row PPO loss + gradient w.r.t. the network outputs.  Shared by the standalone loss
// kernel (fp32 path) and the fused tcgen05 update kernel.
//
// Forward follows src/rl8/nn/functional.py:320-352 op for op; the backward is the
// autograd derivative of that graph:
//   torch.min(a, b)       -> grad to the smaller operand, split 1/2 : 1/2 on ties
//   torch.clamp(x, lo, hi) -> grad passes iff lo <= x <= hi
//   smooth_l1(beta=1)      -> d if |d| < 1 else sign(d)
//   torch.where / torch.max as usual.
#pragma once
#include "dist.cuh"

namespace rl8 {

struct RowLoss {
  float entropy, policy, vf, kl;
};

// d(policy surrogate)/d(ratio) for one row, and the surrogate value.
__device__ __forceinline__ float ppo_surrogate(float adv, float ratio, float clip,
                                               float dual_clip, float* dratio) {
  const float lo = sub(1.0f, clip), hi = add(1.0f, clip);
  const float s1 = mul(adv, ratio);
  const float rc = clampf(ratio, lo, hi);
  const float s2 = mul(adv, rc);
  const bool clamp_pass = (ratio >= lo) && (ratio <= hi);
  // min(s1, s2)
  float w1, w2;
  if (s1 < s2) w1 = 1.0f, w2 = 0.0f;
  else if (s1 > s2) w1 = 0.0f, w2 = 1.0f;
  else w1 = 0.5f, w2 = 0.5f;
  float c1 = fminf(s1, s2);
  float dc1 = adv * (w1 + (clamp_pass ? w2 : 0.0f));  // d c1 / d ratio
  float out = c1, dout = dc1;
  if (dual_clip > 0.0f) {
    const float bound = mul(dual_clip, adv);
    const float c2 = fmaxf(c1, bound);
    float g;  // d c2 / d c1
    if (c1 > bound) g = 1.0f;
    else if (c1 < bound) g = 0.0f;
    else g = 0.5f;
    if (adv < 0.0f) out = c2, dout = dc1 * g;
  }
  *dratio = dout;
  return out;
}

// Value loss clamp(smooth_l1(v, ret), 0, vf_clip) and its derivative w.r.t. v.
__device__ __forceinline__ float ppo_value_loss(float v, float ret, float vf_clip, float* dv) {
  const float d = sub(v, ret);
  const float ad = fabsf(d);
  float l, g;
  if (ad < 1.0f) {
    l = mul(mul(0.5f, d), d);  // 0.5 * d * d / beta, beta = 1
    g = d;
  } else {
    l = sub(ad, 0.5f);
    g = d > 0.0f ? 1.0f : -1.0f;
  }
  const bool pass = (l >= 0.0f) && (l <= vf_clip);
  *dv = pass ? g : 0.0f;
  return clampf(l, 0.0f, vf_clip);
}

// Policy half of one row.  `o` = policy head outputs (logits, or {mean, log_std}); `act` =
// stored action as float (discrete index or continuous value).  Writes d(scaled total)/d(o)
// (for the continuous head: w.r.t. {mean, raw log_std before tanh}); returns entropy, the
// surrogate and the KL term in L.
template <int P>
__device__ __forceinline__ void ppo_policy_row(int dist_kind, const float* o, float act,
                                               float logp_old, float adv,
                                               const rl8_ppo_hparams& hp, float inv_denom,
                                               float* d_o, RowLoss& L, bool tanh_chain = true) {
  float logp_new, dratio;
  const bool want_ent = hp.entropy_coeff != 0.0f;
  L.entropy = 0.0f;
  if (dist_kind == RL8_DIST_CATEGORICAL) {
    float norm[P], probs[P];
    categorical_norm<P>(o, norm, probs);
    const int a = (int)act;
    logp_new = norm[0];
#pragma unroll
    for (int k = 1; k < P; ++k) logp_new = (a == k) ? norm[k] : logp_new;
    const float lr = sub(logp_new, logp_old);
    const float ratio = expf(lr);
    L.policy = ppo_surrogate(adv, ratio, hp.clip_param, hp.dual_clip_param, &dratio);
    L.kl = sub(sub(ratio, 1.0f), lr);
    // total = vf_coeff*vf - policy - ent_coeff*entropy  ->  d/dlogp = -dratio*ratio
    const float dlogp = -inv_denom * dratio * ratio;
    float ent = 0.0f;
    if (want_ent) ent = categorical_entropy<P>(norm, probs);
    L.entropy = ent;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      float g = dlogp * ((a == k ? 1.0f : 0.0f) - probs[k]);
      if (want_ent) g += inv_denom * hp.entropy_coeff * probs[k] * (norm[k] + ent);
      d_o[k] = g;
    }
  } else {
    const float mean = o[0], ls = o[1];
    const float scale = expf(ls);
    float x = act;
    bool inside = true;
    if (dist_kind == RL8_DIST_SQUASHED_NORMAL) {
      logp_new = squashed_logp(mean, scale, act, &inside);
      x = squashed_inverse(act);
    } else {
      logp_new = normal_logp(mean, scale, act);
    }
    const float lr = sub(logp_new, logp_old);
    const float ratio = expf(lr);
    L.policy = ppo_surrogate(adv, ratio, hp.clip_param, hp.dual_clip_param, &dratio);
    L.kl = sub(sub(ratio, 1.0f), lr);
    float dlogp = -inv_denom * dratio * ratio;
    if (!inside) dlogp = 0.0f;
    const float z = (x - mean) / scale;  // standardised residual
    float dmean = dlogp * (z / scale);
    float dls = dlogp * (z * z - 1.0f);
    if (want_ent) {  // Normal only; the squashed normal has no entropy (distributions.py:153-157)
      L.entropy = normal_entropy(scale);
      dls += -inv_denom * hp.entropy_coeff;
    }
    d_o[0] = dmean;
    // default models: log_std = tanh(raw) and the gradient is w.r.t. raw; custom models get it w.r.t. log_std
    d_o[1] = tanh_chain ? dls * (1.0f - ls * ls) : dls;
#pragma unroll
    for (int k = 2; k < P; ++k) d_o[k] = 0.0f;
  }
}

// Value half of one row.
__device__ __forceinline__ void ppo_value_row(float v, float ret, const rl8_ppo_hparams& hp,
                                              float inv_denom, float* d_v, RowLoss& L) {
  float dv;
  L.vf = ppo_value_loss(v, ret, hp.vf_clip_param, &dv);
  *d_v = inv_denom * hp.vf_coeff * dv;
}

template <int P>
__device__ __forceinline__ RowLoss ppo_row(int dist_kind, const float* o, float v, float act,
                                           float logp_old, float adv, float ret,
                                           const rl8_ppo_hparams& hp, float inv_denom, float* d_o,
                                           float* d_v, bool tanh_chain = true) {
  RowLoss L;
  ppo_policy_row<P>(dist_kind, o, act, logp_old, adv, hp, inv_denom, d_o, L, tanh_chain);
  ppo_value_row(v, ret, hp, inv_denom, d_v, L);
  return L;
}

}  // namespace rl8
