// Shared pieces of the tensor-core kernels (mlp_tc.cu, update_tc.cu).
#pragma once
#include "dist.cuh"
#include "mlp_fp32.cuh"
#include "tc.cuh"

namespace rl8 {

using namespace tc;

constexpr int H = 256;         // hidden width
constexpr int TILE = 128;      // rows (envs / transitions) per CTA tile == UMMA M
constexpr int kW2Bytes = H * H * 2;
constexpr int kTileBytes = TILE * H * 2;
constexpr int kMaxPT = 4;      // widest head on the tensor-core path

struct NetParams {
  const uint8_t* w2_img;  // bf16, chunked [j = 256 rows][i = 256 cols]
  const float *w1, *b1, *b2, *w3, *b3;
  int D, P;
};

// Shared-memory plan of the rollout kernel (dynamic smem, 128-byte aligned).
struct Smem {
  uint8_t w2[kW2Bytes];           // 131072
  uint8_t a_tile[kTileBytes];     //  65536  H1 tile: K-major A operand
  uint8_t w1aug[H * 32];          //   8192  tf32 [W1 | b1]: off(i, d) = i*16 + (d/4)*4096 + (d%4)*4
  uint8_t aug32[TILE * 32];       //   4096  tf32 [obs, 1]:  off(r, d) = r*16 + (d/4)*2048 + (d%4)*4
  float b2[H];                    //   1024
  float w3[kMaxPT][H];            //   4096
  float part[2][4][TILE][kMaxPT]; //  16384  head partial sums per column quarter, one set per accumulator
  uint64_t bar_w, bar_z, bar_mma[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Smem) <= 227 * 1024, "smem plan exceeds the 227 KB CTA limit");

// ---- weight packing: fp32 [256][256] row-major -> bf16 chunked image ------------------------
static __global__ void __launch_bounds__(256) pack_w2_kernel(const float* __restrict__ w2, uint8_t* __restrict__ img) {
  // one thread per 16-byte chunk: row j, column group c (8 consecutive i)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 256*32
  if (idx >= H * (H / 8)) return;
  const int j = idx % H, c = idx / H;
  float v[8];
  const float4 lo = *reinterpret_cast<const float4*>(w2 + j * H + c * 8);
  const float4 hi = *reinterpret_cast<const float4*>(w2 + j * H + c * 8 + 4);
  v[0] = lo.x, v[1] = lo.y, v[2] = lo.z, v[3] = lo.w, v[4] = hi.x, v[5] = hi.y, v[6] = hi.z, v[7] = hi.w;
  store_chunk(img, chunk_offset<H>(j, c), v);
}

static inline int launch_pack_w2(const float* w2, uint8_t* img, cudaStream_t st) {
  pack_w2_kernel<<<H * (H / 8) / 256, 256, 0, st>>>(w2, img);
  return check_launch("pack_w2");
}

// ---- shared device pieces ------------------------------------------------------------------------
// Barriers, TMEM, the resident W2 image (one bulk-async copy) and the per-network constants; works on any
// plan with the members w2, w1aug, b2, w3, bar_w, bar_z, bar_mma[2], tmem_base.
template <class S>
__device__ __forceinline__ void cta_setup(S& s, const NetParams& np, uint32_t tmem_cols) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s.bar_w, 1);
    mbar_init(&s.bar_z, 1);
    mbar_init(&s.bar_mma[0], 1);
    mbar_init(&s.bar_mma[1], 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, tmem_cols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_expect_tx(&s.bar_w, kW2Bytes);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      bulk_g2s(s.w2 + i * (kW2Bytes / 8), np.w2_img + i * (kW2Bytes / 8), kW2Bytes / 8, &s.bar_w);
  }
  // [W1 | b1] rounded to tf32: the B operand of the layer-1 MMA  Z1 = [obs, 1] * [W1, b1]^T  (K = 8)
  for (int e = tid; e < H * 8; e += blockDim.x) {
    const int i = e >> 3, d = e & 7;
    float v = 0.0f;
    if (d < np.D) v = np.w1[i * np.D + d];
    else if (d == np.D) v = np.b1[i];
    *reinterpret_cast<float*>(s.w1aug + i * 16 + (d >> 2) * (H * 16) + (d & 3) * 4) = tf32_round(v);
  }
  for (int i = tid; i < H; i += blockDim.x) s.b2[i] = np.b2[i];
  for (int i = tid; i < kMaxPT * H; i += blockDim.x) {
    const int p = i / H, c = i - p * H;
    s.w3[p][c] = p < np.P ? np.w3[p * H + c] : 0.0f;
  }
  __syncthreads();
  mbar_wait(&s.bar_w, 0);
}

constexpr int kFwdThreads = 512;  // 16 warps: thread -> row 32 * (warp % 4) + lane, column quarter warp / 4

// Z1 = [obs, 1] * [W1, b1]^T into the accumulator at acc_tmem (elected lane; commits to s.bar_z)
template <class S>
__device__ __forceinline__ void issue_layer1(S& s, uint32_t acc_tmem) {
  mma_tf32(acc_tmem, smem_desc(smem_u32(s.aug32), TILE * 16, 128), smem_desc(smem_u32(s.w1aug), H * 16, 128),
           instr_desc_tf32(TILE, H), 0u);
  mma_commit(&s.bar_z);
}
// H1 = relu(Z1): accumulator -> bf16 chunks of the A tile (this thread's row and column quarter)
template <class S>
__device__ __forceinline__ void h1_epilogue(S& s, uint32_t acc_tmem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, part = warp >> 2;
  const int r = q * 32 + lane;
  const uint32_t base = acc_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64);
  float v0[32], v1[32];
  tmem_ld32_nowait(base, v0);
  tmem_ld32_nowait(base + 32, v1);
  tmem_wait_ld();
  reg_fence32(v0);
  reg_fence32(v1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    store_chunk_relu(s.a_tile, chunk_offset<TILE>(r, part * 8 + k), v0 + 8 * k);
    store_chunk_relu(s.a_tile, chunk_offset<TILE>(r, part * 8 + 4 + k), v1 + 8 * k);
  }
}

// Head partial sums of one accumulator: thread -> row 32*(warp%4)+lane, column quarter warp/4.
// dot[p] = sum_{j in quarter} relu(z[r][j] + b2[j]) * w3[p][j]  -> s.part[quarter][r][p]
// Per-column constants are read as 128-bit warp broadcasts.
template <int P, class S>
__device__ __forceinline__ void head_partials(S& s, uint32_t acc_tmem, int set = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, part = warp >> 2;
  const int r = q * 32 + lane;
  float dot[P];
#pragma unroll
  for (int p = 0; p < P; ++p) dot[p] = 0.0f;
  float v0[32], v1[32];
  tmem_ld32_nowait(acc_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64), v0);
  tmem_ld32_nowait(acc_tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64 + 32), v1);
  tmem_wait_ld();
  reg_fence32(v0);
  reg_fence32(v1);
#pragma unroll
  for (int c2 = 0; c2 < 2; ++c2) {
    const int col0 = part * 64 + c2 * 32;
    const float* v = c2 ? v1 : v0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(&s.b2[col0 + j]);
      const float h0 = fmaxf(v[j] + b.x, 0.0f), h1 = fmaxf(v[j + 1] + b.y, 0.0f);
      const float h2 = fmaxf(v[j + 2] + b.z, 0.0f), h3 = fmaxf(v[j + 3] + b.w, 0.0f);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float4 w = *reinterpret_cast<const float4*>(&s.w3[p][col0 + j]);
        dot[p] = fmaf(h0, w.x, dot[p]);
        dot[p] = fmaf(h1, w.y, dot[p]);
        dot[p] = fmaf(h2, w.z, dot[p]);
        dot[p] = fmaf(h3, w.w, dot[p]);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) s.part[set][part][r][p] = dot[p];
}
// sum of the four column quarters' partial sums of row r (in a fixed order)
template <class S>
__device__ __forceinline__ float head_sum(const S& s, int r, int p, int set = 0) {
  return (s.part[set][0][r][p] + s.part[set][1][r][p]) + (s.part[set][2][r][p] + s.part[set][3][r][p]);
}

}  // namespace rl8

namespace rl8 {

static inline int set_smem(const void* fn, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute", e);
    return RL8_ERR_CUDA;
  }
  return RL8_OK;
}

static inline NetParams net_params(const rl8_model* m, int which, const uint8_t* img) {
  NetParams np;
  np.w2_img = img;
  np.w1 = which ? m->vf_w1 : m->pi_w1;
  np.b1 = which ? m->vf_b1 : m->pi_b1;
  np.b2 = which ? m->vf_b2 : m->pi_b2;
  np.w3 = which ? m->vf_w3 : m->pi_w3;
  np.b3 = which ? m->vf_b3 : m->pi_b3;
  np.D = m->D;
  np.P = which ? 1 : m->P;
  return np;
}

}  // namespace rl8
