// fp32-accurate tensor-core path (RL8_PREC_FP32): split-bf16 operands, tcgen05.mma.cta_group::2 (see split_tc.cuh).
//
//   tc3_selftest_kernel   D[256][256] = A[256][K] * B[256][K]^T with a selectable set of piece products: pins the
//                         pair plumbing (cluster launch, pair TMEM allocation, N-split B operand, multicast commit)
//                         and measures what each term set costs in accuracy.
#include "split_tc.cuh"

namespace rl8 {

using namespace tc;

// ---- pair selftest ---------------------------------------------------------------------------------------------
// One cluster of two CTAs, 256 threads each, synchronous K loop in chunks of 32: stage -> cluster barrier ->
// leader issues -> multicast commit -> both CTAs wait.
struct SmemSelf {
  uint8_t a[3][128 * 4 * 16];  // piece tiles [128 rows][4 column groups of 8 K]: off(r, c) = r*16 + c*2048
  uint8_t b[3][128 * 4 * 16];
  uint64_t bar;
  uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
tc3_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K,
                    int terms) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemSelf& s = *reinterpret_cast<SmemSelf*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&s.bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_pair(&s.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  uint8_t* const a_tiles[3] = {s.a[0], s.a[1], s.a[2]};
  uint8_t* const b_tiles[3] = {s.b[0], s.b[1], s.b[2]};
  const uint32_t idesc = instr_desc(256, 256, 0, 0);
  uint32_t parity = 0;
  bool first = true;
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = tid & 127, cg = (tid >> 7) + 2 * i;
      const int64_t g = (int64_t)(128 * rank + row) * K + k0 + cg * 8;
      float v[8];
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(A + g);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(A + g + 4);
      store_split_chunk<3>(a_tiles, (uint32_t)(row * 16 + cg * 2048), v);
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(B + g);
      *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(B + g + 4);
      store_split_chunk<3>(b_tiles, (uint32_t)(row * 16 + cg * 2048), v);
    }
    fence_async_smem();
    fence_before_sync();
    cluster_sync_all();
    if (rank == 0 && warp == 0 && elect_one()) {
      fence_after_sync();
      // (a piece, b piece) of term bit i
      const int ta[6] = {0, 0, 1, 1, 0, 2}, tb[6] = {0, 1, 0, 1, 2, 0};
      for (int ks = 0; ks < 2; ++ks) {
        for (int i = 5; i >= 0; --i) {
          if (!((terms >> i) & 1)) continue;
          const uint64_t ad = smem_desc(smem_u32(s.a[ta[i]]) + ks * 4096, 2048, 128);
          const uint64_t bd = smem_desc(smem_u32(s.b[tb[i]]) + ks * 4096, 2048, 128);
          mma_bf16_pair(tmem, ad, bd, idesc, first ? 0u : 1u);
          first = false;
        }
      }
      mma_commit_pair(&s.bar);
    }
    mbar_wait_cluster(&s.bar, parity);
    parity ^= 1u;
    fence_after_sync();
  }
  {
    const int q = warp & 3, half = warp >> 2;
    const int64_t row = 128 * rank + q * 32 + lane;
    for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) D[row * 256 + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tmem, 512);
}

}  // namespace rl8

using namespace rl8;

// Test hook: the pair MMA with split operands (tests/test_gpu_split.py).  terms: bit 0 a0b0, 1 a0b1, 2 a1b0,
// 3 a1b1, 4 a0b2, 5 a2b0.
extern "C" int rl8_tc3_selftest(const float* A, const float* B, float* D, int32_t K, int32_t terms,
                                rl8_stream_t stream) {
  if (!A || !B || !D || K < 32 || (K % 32) || terms < 1 || terms > 63) return RL8_ERR_ARG;
  cudaError_t e = cudaFuncSetAttribute(tc3_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(SmemSelf));
  if (e != cudaSuccess) {
    set_last_error("cudaFuncSetAttribute", e);
    return RL8_ERR_CUDA;
  }
  tc3_selftest_kernel<<<2, 256, sizeof(SmemSelf), (cudaStream_t)stream>>>(A, B, D, K, terms);
  return check_launch("tc3_selftest");
}
