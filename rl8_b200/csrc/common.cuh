// Shared helpers for the rl8_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rl8_b200.h"

namespace rl8 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Thread-local text of the last CUDA failure (rl8_last_error()).
void set_last_error(const char* where, cudaError_t err);

inline int check_launch(const char* where) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_last_error(where, err);
    return RL8_ERR_CUDA;
  }
  return RL8_OK;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Grid for a grid-stride kernel: enough CTAs for `waves` full waves at `ctas_per_sm`,
// never more than the work needs.
inline int grid_for(int64_t work_items, int block, int ctas_per_sm = 8, int waves = 4) {
  int64_t need = ceil_div(work_items, block);
  int64_t cap = (int64_t)kNumSMs * ctas_per_sm * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- streaming 128-bit accesses (bypass L1 allocation for touch-once data) ----------
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of `v` (blockDim.x <= 1024); result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 */) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? smem[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// atomic min / max on doubles via CAS (used once per block).
__device__ __forceinline__ void atomic_min_double(double* addr, double val) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) <= val) break;
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_double(double* addr, double val) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= val) break;
    old = atomicCAS(a, assumed, __double_as_longlong(val));
  } while (assumed != old);
}

}  // namespace rl8
