// Generalized Advantage Estimation (src/rl8/nn/functional.py:50-123) and the collect()
// statistics (src/rl8/algorithms/_feedforward.py:410-436).
//
// HBM-bound: per transition 8 B read (r, V) + 12 B written (scaled r, A, ret) in the scan,
// +8 B in the normalisation pass (read A, write A).  Three layouts:
//   horizon-major (stride_n == 1): one env per lane, 4 envs per thread via 128-bit
//     accesses, sequential in t (exactly the reference's op order -> bit-comparable);
//   env-major (stride_t == 1, the reference layout [N, T+1]): one env row per warp; the
//     row is loaded coalesced in 32-step chunks and the recurrence
//     A_t = delta_t + c*A_{t+1} is solved with a 5-step warp shuffle scan (fmaf and a reassociated sum with
//     powers c^1..c^16: NOT the reference's rounding sequence -- within 2e-6 / 5e-6 absolute of the oracle on
//     returns / advantages for N(0,1) inputs up to T = 70, tests/test_gpu_kernels.py; Algorithm.step() never uses it);
//   anything else: strided sequential.
// There are no `done` flags in the reference: every env bootstraps from
// V_T; an optional done mask would only zero `c` and `gamma` per element.
#include "envs.cuh"

namespace rl8 {

struct GaeConsts {
  float gamma;      // f32(gamma)
  float gl;         // f32(gamma * gae_lambda)  (double product, one rounding)
  float denom;      // f32(reward_scale + 1e-8)
  const float* denom_dev;  // when set: the same value in device memory (rl8_reward_scale), read by the kernel
  int keep_rewards;        // horizon-major kernel: do not write the scaled rewards back (Algorithm.step drops them)
};
// the reward divisor of this launch: the host value, or the one a previous kernel left on the device
__device__ __forceinline__ float gae_denom(const GaeConsts& c) { return c.denom_dev ? __ldg(c.denom_dev) : c.denom; }

// ---- horizon-major, VEC envs per thread ----------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
gae_scan_hm_kernel(float* __restrict__ rewards, const float* __restrict__ values,
                   float* __restrict__ adv, float* __restrict__ ret, int64_t N, int T,
                   int64_t st, GaeConsts c, double* __restrict__ moments) {
  const float denom = gae_denom(c);
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0;
  const int64_t groups = N / VEC;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n0 = g * VEC;
    float vnext[VEC], prev[VEC], tmp[VEC];
    auto load = [&](const float* p, float* out) {
      if constexpr (VEC == 4) {
        float4 t = ld_stream4(p);
        out[0] = t.x, out[1] = t.y, out[2] = t.z, out[3] = t.w;
      } else {
        out[0] = *p;
      }
    };
    auto store = [&](float* p, const float* in) {
      if constexpr (VEC == 4) st_stream4(p, make_float4(in[0], in[1], in[2], in[3]));
      else *p = in[0];
    };
    load(values + (int64_t)T * st + n0, vnext);
#pragma unroll
    for (int j = 0; j < VEC; ++j) prev[j] = 0.0f, tmp[j] = 0.0f;
    store(adv + (int64_t)T * st + n0, tmp);               // A_T = 0
    if (ret) store(ret + (int64_t)T * st + n0, vnext);    // ret_T = 0 + V_T
    // Scaled reward of slot T (the reference divides the whole tensor, :106).
    load(rewards + (int64_t)T * st + n0, tmp);
#pragma unroll
    for (int j = 0; j < VEC; ++j) tmp[j] = dvd(tmp[j], denom);
    if (!c.keep_rewards) store(rewards + (int64_t)T * st + n0, tmp);

    constexpr int U = 4;  // software pipeline: U time steps of loads in flight
    int t = T - 1;
    for (; t >= U - 1; t -= U) {
      float r[U][VEC], v[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        load(rewards + (int64_t)(t - u) * st + n0, r[u]);
        load(values + (int64_t)(t - u) * st + n0, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a[VEC], rt[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          r[u][j] = dvd(r[u][j], denom);
          float delta = add(r[u][j], sub(mul(c.gamma, vnext[j]), v[u][j]));
          a[j] = add(delta, mul(c.gl, prev[j]));
          prev[j] = a[j];
          rt[j] = add(a[j], v[u][j]);
          vnext[j] = v[u][j];
          s1 += (double)a[j];
          s2 += (double)a[j] * (double)a[j];
        }
        if (!c.keep_rewards) store(rewards + (int64_t)(t - u) * st + n0, r[u]);
        store(adv + (int64_t)(t - u) * st + n0, a);
        if (ret) store(ret + (int64_t)(t - u) * st + n0, rt);
      }
    }
    for (; t >= 0; --t) {
      float r[VEC], v[VEC], a[VEC], rt[VEC];
      load(rewards + (int64_t)t * st + n0, r);
      load(values + (int64_t)t * st + n0, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        r[j] = dvd(r[j], denom);
        float delta = add(r[j], sub(mul(c.gamma, vnext[j]), v[j]));
        a[j] = add(delta, mul(c.gl, prev[j]));
        prev[j] = a[j];
        rt[j] = add(a[j], v[j]);
        vnext[j] = v[j];
        s1 += (double)a[j];
        s2 += (double)a[j] * (double)a[j];
      }
      if (!c.keep_rewards) store(rewards + (int64_t)t * st + n0, r);
      store(adv + (int64_t)t * st + n0, a);
      if (ret) store(ret + (int64_t)t * st + n0, rt);
    }
  }
  if (moments) {
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
      atomicAdd(moments + 0, s1);
      atomicAdd(moments + 1, s2);
      if (blockIdx.x == 0) atomicAdd(moments + 2, (double)N * (double)T);
    }
  }
}

// ---- generic strides, sequential -----------------------------------------------------------
__global__ void __launch_bounds__(256)
gae_scan_strided_kernel(float* __restrict__ rewards, const float* __restrict__ values,
                        float* __restrict__ adv, float* __restrict__ ret, int64_t N, int T,
                        int64_t sn, int64_t st, GaeConsts c, double* __restrict__ moments) {
  const float denom = gae_denom(c);
  __shared__ double red[32];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t base = n * sn;
    float vnext = values[base + (int64_t)T * st], prev = 0.0f;
    adv[base + (int64_t)T * st] = 0.0f;
    if (ret) ret[base + (int64_t)T * st] = vnext;
    rewards[base + (int64_t)T * st] = dvd(rewards[base + (int64_t)T * st], denom);
    for (int t = T - 1; t >= 0; --t) {
      const int64_t i = base + (int64_t)t * st;
      float r = dvd(rewards[i], denom), v = values[i];
      float delta = add(r, sub(mul(c.gamma, vnext), v));
      float a = add(delta, mul(c.gl, prev));
      rewards[i] = r;
      adv[i] = a;
      if (ret) ret[i] = add(a, v);
      prev = a, vnext = v;
      s1 += (double)a;
      s2 += (double)a * (double)a;
    }
  }
  if (moments) {
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
      atomicAdd(moments + 0, s1);
      atomicAdd(moments + 1, s2);
      if (blockIdx.x == 0) atomicAdd(moments + 2, (double)N * (double)T);
    }
  }
}

// ---- env-major: one env row per warp, warp-level reverse scan ------------------------------
// Row n occupies [n*sn, n*sn + T] (T+1 contiguous floats).  Chunks of 32 time steps are
// processed from the end of the horizon; within a chunk lane l holds t = t0 + l and
//   x_l <- x_l + c^d * x_{l+d}  for d = 1, 2, 4, 8, 16   (Kogge-Stone, constant ratio c)
// then the carry of the later chunk enters as c^(32-l) * A_{t0+32}.
__global__ void __launch_bounds__(256)
gae_scan_em_kernel(float* __restrict__ rewards, const float* __restrict__ values,
                   float* __restrict__ adv, float* __restrict__ ret, int64_t N, int T,
                   int64_t sn, GaeConsts c, double* __restrict__ moments) {
  const float denom = gae_denom(c);
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // c^1, c^2, c^4, c^8, c^16 and the per-lane carry weight c^(32-lane).
  float cp[5];
  cp[0] = c.gl;
#pragma unroll
  for (int i = 1; i < 5; ++i) cp[i] = cp[i - 1] * cp[i - 1];
  float cw = 1.0f;
  for (int i = 0; i < 32 - lane; ++i) cw *= c.gl;

  double s1 = 0.0, s2 = 0.0;
  for (int64_t n = warp; n < N; n += nwarps) {
    const int64_t base = n * sn;
    float carry = 0.0f;  // A at the first step of the later chunk
    if (lane == 0) {
      float vT = values[base + T];
      adv[base + T] = 0.0f;
      if (ret) ret[base + T] = vT;
      rewards[base + T] = dvd(rewards[base + T], denom);
    }
    for (int t0 = ((T - 1) / 32) * 32; t0 >= 0; t0 -= 32) {
      const int t = t0 + lane;
      const bool live = t < T;
      float r = 0.0f, v = 0.0f, vn = 0.0f;
      if (live) {
        r = dvd(rewards[base + t], denom);
        v = values[base + t];
        vn = values[base + t + 1];
      }
      float x = live ? add(r, sub(mul(c.gamma, vn), v)) : 0.0f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float y = __shfl_down_sync(0xffffffffu, x, 1 << i);
        if (lane + (1 << i) < 32) x = fmaf(cp[i], y, x);
      }
      x = fmaf(cw, carry, x);
      carry = __shfl_sync(0xffffffffu, x, 0);
      if (live) {
        rewards[base + t] = r;
        adv[base + t] = x;
        if (ret) ret[base + t] = add(x, v);
        s1 += (double)x;
        s2 += (double)x * (double)x;
      }
    }
  }
  if (moments) {
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) {
      atomicAdd(moments + 0, s1);
      atomicAdd(moments + 1, s2);
      if (blockIdx.x == 0) atomicAdd(moments + 2, (double)N * (double)T);
    }
  }
}

// ---- normalisation: A[:, :T] <- (A - mean) / (std + 1e-8), unbiased std --------------------
__global__ void __launch_bounds__(256)
gae_normalize_kernel(float* __restrict__ adv, int64_t N, int T, int64_t sn, int64_t st,
                     const double* __restrict__ moments) {
  const double cnt = moments[2];
  const double mean_d = moments[0] / cnt;
  double var = (moments[1] - moments[0] * mean_d) / (cnt - 1.0);
  if (var < 0.0) var = 0.0;
  const float mean = (float)mean_d;
  const float denom = add((float)sqrt(var), 1e-8f);
  const int64_t total = N * (int64_t)T;
  if (sn == 1 && st == N && (N % 4 == 0) && (((uintptr_t)adv & 15u) == 0)) {
    // contiguous [T][N] prefix -> 128-bit streaming
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total;
         i += (int64_t)gridDim.x * blockDim.x * 4) {
      float4 a = ld_stream4(adv + i);
      a.x = dvd(sub(a.x, mean), denom);
      a.y = dvd(sub(a.y, mean), denom);
      a.z = dvd(sub(a.z, mean), denom);
      a.w = dvd(sub(a.w, mean), denom);
      st_stream4(adv + i, a);
    }
  } else if (st == 1) {
    // env-major rows: consecutive threads walk t within a row
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      int64_t n = i / T, t = i - n * T;
      float* p = adv + n * sn + t;
      *p = dvd(sub(*p, mean), denom);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
      int64_t t = i / N, n = i - t * N;
      float* p = adv + n * sn + t * st;
      *p = dvd(sub(*p, mean), denom);
    }
  }
}

// ---- collect statistics --------------------------------------------------------------------
// Horizon-major rewards[T+1][N], rdr[T+1][N]; one env per thread, coalesced along n.
__global__ void __launch_bounds__(256)
collect_stats_kernel(const float* __restrict__ rewards, const float* __restrict__ rdr, int64_t N,
                     int T, int t0, double* __restrict__ acc) {
  __shared__ double red[32];
  double sr = 0, sr2 = 0, sR = 0, sR2 = 0, sd = 0, sd2 = 0;
  double mnr = INFINITY, mxr = -INFINITY, mnR = INFINITY, mxR = -INFINITY;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (int64_t)gridDim.x * blockDim.x) {
    float R = 0.0f;  // torch.sum(rewards, dim=1): sequential f32 accumulation over t
    for (int t = t0; t < T; ++t) {
      float r = rewards[(int64_t)t * N + n];
      R = add(R, r);
      sr += r;
      sr2 += (double)r * r;
      mnr = fmin(mnr, (double)r);
      mxr = fmax(mxr, (double)r);
    }
    sR += R;
    sR2 += (double)R * R;
    mnR = fmin(mnR, (double)R);
    mxR = fmax(mxR, (double)R);
    if (rdr) {
      for (int t = 1; t <= T; ++t) {
        float d = rdr[(int64_t)t * N + n];
        sd += d;
        sd2 += (double)d * d;
      }
    }
  }
  double v[6] = {sr, sr2, sR, sR2, sd, sd2};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = block_sum(v[i], red);
    if (threadIdx.x == 0) atomicAdd(acc + i, s);
  }
  // min / max
  double m[4] = {mnr, -mxr, mnR, -mxR};  // reduce all four as minima
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double w = warp_min(m[i]);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) {
      int nw = (blockDim.x + 31) >> 5;
      w = threadIdx.x < nw ? red[threadIdx.x] : INFINITY;
      w = warp_min(w);
      if (threadIdx.x == 0) {
        if (i & 1) atomic_max_double(acc + 6 + i, -w);
        else atomic_min_double(acc + 6 + i, w);
      }
    }
  }
}

}  // namespace rl8

using namespace rl8;

static int gae_scan_impl(float* rewards, const float* values, float* advantages, float* returns, int64_t N,
                         int32_t T, int64_t stride_n, int64_t stride_t, double gamma, double gae_lambda,
                         double reward_scale, const float* denom_dev, int keep_rewards, double* moments,
                         rl8_stream_t stream) {
  if (!rewards || !values || !advantages || N <= 0 || T <= 0) return RL8_ERR_ARG;
  if (keep_rewards && stride_n != 1) return RL8_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  GaeConsts c{(float)gamma, (float)(gamma * gae_lambda), (float)(reward_scale + 1e-8), denom_dev, keep_rewards};
  auto al = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  if (stride_n == 1) {
    bool vec = (N % 4 == 0) && (stride_t % 4 == 0) && al(rewards) && al(values) &&
               al(advantages) && (!returns || al(returns));
    if (vec) {
      gae_scan_hm_kernel<4><<<grid_for(N / 4, 256, 4, 2), 256, 0, st>>>(
          rewards, values, advantages, returns, N, T, stride_t, c, moments);
    } else {
      gae_scan_hm_kernel<1><<<grid_for(N, 256, 4, 2), 256, 0, st>>>(
          rewards, values, advantages, returns, N, T, stride_t, c, moments);
    }
  } else if (stride_t == 1) {
    gae_scan_em_kernel<<<grid_for(N * 32, 256, 8, 2), 256, 0, st>>>(
        rewards, values, advantages, returns, N, T, stride_n, c, moments);
  } else {
    gae_scan_strided_kernel<<<grid_for(N, 256, 4, 2), 256, 0, st>>>(
        rewards, values, advantages, returns, N, T, stride_n, stride_t, c, moments);
  }
  return check_launch("rl8_gae_scan");
}

extern "C" int rl8_gae_scan(float* rewards, const float* values, float* advantages, float* returns,
                            int64_t N, int32_t T, int64_t stride_n, int64_t stride_t, double gamma,
                            double gae_lambda, double reward_scale, double* moments,
                            rl8_stream_t stream) {
  return gae_scan_impl(rewards, values, advantages, returns, N, T, stride_n, stride_t, gamma, gae_lambda,
                       reward_scale, nullptr, 0, moments, stream);
}

extern "C" int rl8_gae_scan_dev(float* rewards, const float* values, float* advantages, float* returns,
                                int64_t N, int32_t T, int64_t stride_n, int64_t stride_t, double gamma,
                                double gae_lambda, const float* reward_scale_dev, int write_scaled_rewards,
                                double* moments, rl8_stream_t stream) {
  if (!reward_scale_dev) return RL8_ERR_ARG;
  return gae_scan_impl(rewards, values, advantages, returns, N, T, stride_n, stride_t, gamma, gae_lambda, 1.0,
                       reward_scale_dev + 1, write_scaled_rewards ? 0 : 1, moments, stream);
}

// out[0] = f32(unbiased std of the reversed discounted returns) from the (all-reduced) accumulator of
// rl8_collect_stats (acc[4] = sum, acc[5] = sum of squares over `count` elements), 1 when !normalize_rewards;
// out[1] = f32(out[0] + 1e-8), the divisor rl8_gae_scan_dev uses.
__global__ void reward_scale_kernel(const double* __restrict__ acc, double count, int normalize_rewards,
                                    float* __restrict__ out) {
  float scale = 1.0f;
  if (normalize_rewards) {
    const double mean = acc[4] / count;
    const double var = count > 1.0 ? (acc[5] - acc[4] * mean) / (count - 1.0) : nan("");
    scale = (float)sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
  }
  out[0] = scale;
  out[1] = (float)((double)scale + 1e-8);
}

extern "C" int rl8_reward_scale(const double* acc, double count, int normalize_rewards, float* out,
                                rl8_stream_t stream) {
  if (!acc || !out || count <= 0) return RL8_ERR_ARG;
  reward_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(acc, count, normalize_rewards, out);
  return check_launch("rl8_reward_scale");
}

extern "C" int rl8_gae_normalize(float* advantages, int64_t N, int32_t T, int64_t stride_n,
                                 int64_t stride_t, const double* moments, rl8_stream_t stream) {
  if (!advantages || !moments || N <= 0 || T <= 0) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  gae_normalize_kernel<<<grid_for(N * T / 4 + 1, 256, 8, 2), 256, 0, st>>>(advantages, N, T, stride_n,
                                                                         stride_t, moments);
  return check_launch("rl8_gae_normalize");
}

extern "C" int rl8_collect_stats_from(const float* rewards, const float* rdr, int64_t N,
                                      int32_t T, int32_t reward_t0, double* acc,
                                      rl8_stream_t stream) {
  if (!rewards || !acc || N <= 0 || T <= 0 || reward_t0 < 0 || reward_t0 >= T) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  collect_stats_kernel<<<grid_for(N, 256, 4, 2), 256, 0, st>>>(rewards, rdr, N, T, reward_t0, acc);
  return check_launch("rl8_collect_stats");
}

extern "C" int rl8_collect_stats(const float* rewards, const float* rdr, int64_t N, int32_t T,
                                 double* acc, rl8_stream_t stream) {
  return rl8_collect_stats_from(rewards, rdr, N, T, 0, acc, stream);
}
