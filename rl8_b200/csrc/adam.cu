// clip_grad_norm_ + Adam on one flat fp32 parameter buffer
// (src/rl8/algorithms/_feedforward.py:586-593; torch.optim.Adam defaults,
// torch.nn.utils.clip_grad_norm_ with error_if_nonfinite=False).
//
// HBM-bound and tiny (135 684 parameters in the CartPole model): 16 B read + 12 B written
// per parameter.  Two launches: a one-CTA norm reduction, then the fused clip+Adam update
// that reads the norm from device memory -- no host synchronisation.
#include <math.h>

#include "envs.cuh"

namespace rl8 {

__global__ void __launch_bounds__(1024)
grad_norm_kernel(const float* __restrict__ g, int64_t count, float* __restrict__ norm_out) {
  __shared__ double red[32];
  double s = 0.0;
  // 128-bit loads, four in flight per thread: one CTA (a fixed summation order) without a latency chain
  const int64_t n4 = ((uintptr_t)g % 16 == 0) ? count / 4 : 0;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 4 * (int64_t)blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = i0 + k * (int64_t)blockDim.x;
      v[k] = i < n4 ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      s += (double)v[k].x * v[k].x + (double)v[k].y * v[k].y + (double)v[k].z * v[k].z + (double)v[k].w * v[k].w;
  }
  for (int64_t i = 4 * n4 + threadIdx.x; i < count; i += blockDim.x) s += (double)g[i] * (double)g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) norm_out[0] = (float)sqrt(s);
}

struct AdamConsts {
  float max_norm, w1 /* 1-beta1 */, beta2, w2 /* 1-beta2 */, bc2_sqrt, eps, neg_step_size;
};

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, int64_t count, AdamConsts c,
                 const float* __restrict__ norm) {
  // clip_coef = max_norm / (total_norm + 1e-6), clamped to 1.0; grads are always scaled.
  const float coef = fminf(dvd(c.max_norm, add(norm[0], 1e-6f)), 1.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = mul(g[i], coef);
    g[i] = gi;
    // exp_avg.lerp_(grad, 1 - beta1)
    const float mi = add(m[i], mul(c.w1, sub(gi, m[i])));
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float vi = add(mul(v[i], c.beta2), mul(mul(c.w2, gi), gi));
    m[i] = mi;
    v[i] = vi;
    // denom = (sqrt(v) / bias_correction2_sqrt) + eps; param += -step_size * (m / denom)
    const float denom = add(dvd(sqrtf(vi), c.bc2_sqrt), c.eps);
    p[i] = add(p[i], mul(c.neg_step_size, dvd(mi, denom)));
  }
}

// Device-resident scalars (CUDA-graph replays: nothing of the update may be a launch argument that changes from one
// optimizer step to the next): one thread advances the step counter and forms the step's constants in double, exactly
// as rl8_clip_adam forms them on the host.
__global__ void adam_consts_kernel(long long* __restrict__ step, const double* __restrict__ lr, double max_norm,
                                   double beta1, double beta2, double eps, AdamConsts* __restrict__ out) {
  const long long s = step[0] + 1;
  step[0] = s;
  const double bc1 = 1.0 - pow(beta1, (double)s);
  const double bc2 = 1.0 - pow(beta2, (double)s);
  AdamConsts c;
  c.max_norm = (float)max_norm;
  c.w1 = (float)(1.0 - beta1);
  c.beta2 = (float)beta2;
  c.w2 = (float)(1.0 - beta2);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.eps = (float)eps;
  c.neg_step_size = (float)(-(lr[0] / bc1));
  *out = c;
}
__global__ void __launch_bounds__(256)
clip_adam_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     int64_t count, const AdamConsts* __restrict__ cp, const float* __restrict__ norm) {
  const AdamConsts c = *cp;
  const float coef = fminf(dvd(c.max_norm, add(norm[0], 1e-6f)), 1.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = mul(g[i], coef);
    g[i] = gi;
    const float mi = add(m[i], mul(c.w1, sub(gi, m[i])));
    const float vi = add(mul(v[i], c.beta2), mul(mul(c.w2, gi), gi));
    m[i] = mi;
    v[i] = vi;
    const float denom = add(dvd(sqrtf(vi), c.bc2_sqrt), c.eps);
    p[i] = add(p[i], mul(c.neg_step_size, dvd(mi, denom)));
  }
}

__global__ void __launch_bounds__(256)
clip_scale_kernel(float* __restrict__ g, int64_t count, float max_norm, const float* __restrict__ norm) {
  const float coef = fminf(dvd(max_norm, add(norm[0], 1e-6f)), 1.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    g[i] = mul(g[i], coef);
}

}  // namespace rl8

using namespace rl8;

extern "C" int rl8_clip_grads(float* grads, int64_t count, double max_norm, float* norm_out, rl8_stream_t stream) {
  if (!grads || !norm_out || count <= 0) return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  grad_norm_kernel<<<1, 1024, 0, st>>>(grads, count, norm_out);
  int rc = check_launch("grad_norm");
  if (rc) return rc;
  clip_scale_kernel<<<grid_for(count, 256), 256, 0, st>>>(grads, count, (float)max_norm, norm_out);
  return check_launch("clip_scale");
}

extern "C" int rl8_clip_adam(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                             int64_t count, double max_norm, double lr, double beta1,
                             double beta2, double eps, int64_t step, float* norm_out,
                             rl8_stream_t stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !norm_out || count <= 0 || step <= 0)
    return RL8_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  grad_norm_kernel<<<1, 1024, 0, st>>>(grads, count, norm_out);
  int rc = check_launch("grad_norm");
  if (rc) return rc;
  // Scalars exactly as torch.optim.adam._single_tensor_adam forms them (Python doubles).
  const double b1 = beta1, b2 = beta2;
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  AdamConsts c;
  c.max_norm = (float)max_norm;
  c.w1 = (float)(1.0 - b1);
  c.beta2 = (float)beta2;
  c.w2 = (float)(1.0 - b2);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.eps = (float)eps;
  c.neg_step_size = (float)(-(lr / bc1));
  clip_adam_kernel<<<grid_for(count, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, count,
                                                        c, norm_out);
  return check_launch("clip_adam");
}

extern "C" int rl8_clip_adam_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t count,
                                 double max_norm, const double* lr_dev, double beta1, double beta2, double eps,
                                 long long* step_dev, float* scratch, rl8_stream_t stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !scratch || count <= 0)
    return RL8_ERR_ARG;
  static_assert(sizeof(AdamConsts) <= 8 * sizeof(float), "AdamConsts must fit scratch[8..15]");
  cudaStream_t st = (cudaStream_t)stream;
  AdamConsts* consts = reinterpret_cast<AdamConsts*>(scratch + 8);
  grad_norm_kernel<<<1, 1024, 0, st>>>(grads, count, scratch);
  int rc = check_launch("grad_norm");
  if (rc) return rc;
  adam_consts_kernel<<<1, 1, 0, st>>>(step_dev, lr_dev, max_norm, beta1, beta2, eps, consts);
  if ((rc = check_launch("adam_consts"))) return rc;
  clip_adam_dev_kernel<<<grid_for(count, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, count, consts,
                                                            scratch);
  return check_launch("clip_adam_dev");
}
