// Per-environment transition functions, one env per thread, state in registers.
//
// Every arithmetic step uses the round-to-nearest intrinsics (__fmul_rn/__fadd_rn/...)
// so that nvcc never contracts a multiply-add into an FMA: the reference evaluates each
// torch op with its own rounding, and these functions follow its op order one for one
// (see oracle/ppo_oracle.py, which is pinned bit-for-bit to the upstream code).
#pragma once
#include "common.cuh"

namespace rl8 {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  // torch.clamp: min(max(x, lo), hi); NaN propagates in torch, irrelevant here.
  return fminf(fmaxf(x, lo), hi);
}

template <int KIND>
struct EnvTraits;
template <>
struct EnvTraits<RL8_ENV_DISCRETE_DUMMY> {
  static constexpr int S = 1, D = 1;
  static constexpr bool discrete = true;
};
template <>
struct EnvTraits<RL8_ENV_CONTINUOUS_DUMMY> {
  static constexpr int S = 1, D = 1;
  static constexpr bool discrete = false;
};
template <>
struct EnvTraits<RL8_ENV_CARTPOLE> {
  static constexpr int S = 4, D = 5;
  static constexpr bool discrete = true;
};
template <>
struct EnvTraits<RL8_ENV_MOUNTAIN_CAR> {
  static constexpr int S = 2, D = 2;
  static constexpr bool discrete = true;
};
template <>
struct EnvTraits<RL8_ENV_PENDULUM> {
  static constexpr int S = 2, D = 3;
  static constexpr bool discrete = false;
};

// Action passed as a float for both kinds: discrete actions are small integers and
// (a - 1) * force etc. are exact in f32, exactly like the reference's int64 -> f32 cast.

// ---- observation of a state (used by reset) -------------------------------------------
template <int KIND>
__device__ __forceinline__ void env_observe(const float* s, float* obs) {
  if constexpr (KIND == RL8_ENV_DISCRETE_DUMMY || KIND == RL8_ENV_CONTINUOUS_DUMMY) {
    obs[0] = s[0];
  } else if constexpr (KIND == RL8_ENV_CARTPOLE) {
    // examples/cartpole/env.py:133-136
    obs[0] = s[0];
    obs[1] = s[1];
    obs[2] = cosf(s[2]);
    obs[3] = sinf(s[2]);
    obs[4] = s[3];
  } else if constexpr (KIND == RL8_ENV_MOUNTAIN_CAR) {
    obs[0] = s[0];
    obs[1] = s[1];
  } else {
    // examples/pendulum/env.py:104-106
    obs[0] = cosf(s[0]);
    obs[1] = sinf(s[0]);
    obs[2] = s[1];
  }
}

// ---- reset: noise -> state ---------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ void env_reset(const rl8_env_cfg& c, const float* z, float* s) {
  if constexpr (KIND == RL8_ENV_DISCRETE_DUMMY || KIND == RL8_ENV_CONTINUOUS_DUMMY) {
    // src/rl8/env.py:197-203  uniform_(-b, b) = u * (b - (-b)) + (-b)
    float b = c.p[0];
    s[0] = add(mul(z[0], sub(b, -b)), -b);
  } else if constexpr (KIND == RL8_ENV_CARTPOLE) {
    // examples/cartpole/env.py:128-132  normal(0, 0.01)
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = add(mul(z[i], 0.01f), 0.0f);
  } else if constexpr (KIND == RL8_ENV_MOUNTAIN_CAR) {
    // examples/mountain_car/env.py:92-102  p ~ N(-0.5, 0.05), v ~ N(0, 0.05)
    s[0] = add(mul(z[0], 0.05f), -0.5f);
    s[1] = add(mul(z[1], 0.05f), 0.0f);
  } else {
    // examples/pendulum/env.py:94-103  th ~ U(-pi, pi), thdot ~ U(-1, 1)
    float pi = c.p[5];
    s[0] = add(mul(z[0], sub(pi, -pi)), -pi);
    s[1] = add(mul(z[1], 2.0f), -1.0f);
  }
}

// ---- step: (state, action) -> (state', obs, reward) -----------------------------------------
template <int KIND>
__device__ __forceinline__ void env_step(const rl8_env_cfg& c, float* s, float a, float* obs,
                                         float& reward) {
  if constexpr (KIND == RL8_ENV_DISCRETE_DUMMY) {
    // src/rl8/env.py:253-259  state += 2a - 1
    s[0] = add(s[0], sub(mul(2.0f, a), 1.0f));
    obs[0] = s[0];
    reward = -fabsf(s[0]);
  } else if constexpr (KIND == RL8_ENV_CONTINUOUS_DUMMY) {
    // src/rl8/env.py:224-230  state += a
    s[0] = add(s[0], a);
    obs[0] = s[0];
    reward = -fabsf(s[0]);
  } else if constexpr (KIND == RL8_ENV_CARTPOLE) {
    // examples/cartpole/env.py:12-64
    const float force_mag = c.p[0], gravity = c.p[1], length = c.p[2], pole_mass = c.p[3],
                pml = c.p[4], total_mass = c.p[5], tau = c.p[6], four_thirds = c.p[7];
    float x = s[0], xd = s[1], th = s[2], thd = s[3];
    float push = mul(sub(a, 1.0f), force_mag);
    float cth = cosf(th), sth = sinf(th);
    float tmp = dvd(add(push, mul(mul(pml, mul(thd, thd)), sth)), total_mass);
    float th_acc = dvd(sub(mul(gravity, sth), mul(cth, tmp)),
                       mul(length, sub(four_thirds, dvd(mul(pole_mass, mul(cth, cth)), total_mass))));
    float x_acc = sub(tmp, dvd(mul(mul(pml, th_acc), cth), total_mass));
    if (c.p[8] == 0.0f) {  // "euler"
      x = add(x, mul(tau, xd));
      xd = add(xd, mul(tau, x_acc));
      th = add(th, mul(tau, thd));
      thd = add(thd, mul(tau, th_acc));
    } else {
      xd = add(xd, mul(tau, x_acc));
      x = add(x, mul(tau, xd));
      thd = add(thd, mul(tau, th_acc));
      th = add(th, mul(tau, thd));
    }
    s[0] = x, s[1] = xd, s[2] = th, s[3] = thd;
    float cn = cosf(th), sn = sinf(th);
    obs[0] = x, obs[1] = xd, obs[2] = cn, obs[3] = sn, obs[4] = thd;
    float ang = add(fabsf(sub(cn, 1.0f)), fabsf(sub(sn, 0.0f)));
    float oth = add(add(fabsf(x), fabsf(xd)), fabsf(thd));
    reward = -add(ang, oth);
  } else if constexpr (KIND == RL8_ENV_MOUNTAIN_CAR) {
    // examples/mountain_car/env.py:12-38
    const float force_mag = c.p[0], goal_p = c.p[1], goal_v = c.p[2], gravity = c.p[3],
                max_p = c.p[4], max_speed = c.p[5], min_p = c.p[6];
    float p = s[0], v = s[1];
    v = add(v, sub(mul(sub(a, 1.0f), force_mag), mul(gravity, cosf(mul(3.0f, p)))));
    v = clampf(v, -max_speed, max_speed);
    p = add(p, v);
    p = clampf(p, min_p, max_p);
    if (p == min_p && v < 0.0f) v = 0.0f;
    float r = mul(fabsf(sub(p, goal_p)), -1.0f);
    if (p >= goal_p && v >= goal_v) r = 1.0f;
    s[0] = p, s[1] = v;
    obs[0] = p, obs[1] = v;
    reward = r;
  } else {
    // examples/pendulum/env.py:12-39
    const float dt = c.p[0], c_sin = c.p[1], c_u = c.p[2], max_speed = c.p[3],
                max_torque = c.p[4], pi = c.p[5], two_pi = c.p[6];
    float th = s[0], thd = s[1];
    float u = clampf(a, -max_torque, max_torque);
    // torch.remainder: fmod, then shift into the divisor's sign.
    float w = fmodf(add(th, pi), two_pi);
    if (w != 0.0f && ((w < 0.0f) != (two_pi < 0.0f))) w = add(w, two_pi);
    w = sub(w, pi);
    float cost = add(add(mul(w, w), mul(0.1f, mul(thd, thd))), mul(0.001f, mul(u, u)));
    float nthd = add(thd, mul(add(mul(c_sin, sinf(th)), mul(c_u, u)), dt));
    nthd = clampf(nthd, -max_speed, max_speed);
    float nth = add(th, mul(nthd, dt));
    s[0] = nth, s[1] = nthd;
    obs[0] = cosf(nth), obs[1] = sinf(nth), obs[2] = nthd;
    reward = -cost;
  }
}

}  // namespace rl8
