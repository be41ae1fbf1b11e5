// fp32 CUDA-core kernels of the default two-MLP model (parity path).
#include "mlp_fp32.cuh"

namespace rl8 {

// ---- layer 1: h1 = relu(b1 + obs @ w1^T), K = D <= 8 ------------------------------------
// Block = H threads (one hidden unit each, weights in registers); obs rows staged in smem.
constexpr int kL1Rows = 64;
constexpr int kMaxD = 8;

__global__ void __launch_bounds__(256)
layer1_fwd_kernel(RowMap map, int64_t rows, int D, int H, const float* __restrict__ w1,
                  const float* __restrict__ b1, float* __restrict__ h1) {
  __shared__ float sobs[kL1Rows][kMaxD];
  const int j = threadIdx.x;
  float w[kMaxD], b = 0.0f;
#pragma unroll
  for (int d = 0; d < kMaxD; ++d) w[d] = (j < H && d < D) ? w1[j * D + d] : 0.0f;
  if (j < H) b = b1[j];
  const int64_t ds = map.dstride();
  for (int64_t r0 = (int64_t)blockIdx.x * kL1Rows; r0 < rows; r0 += (int64_t)gridDim.x * kL1Rows) {
    __syncthreads();
    for (int i = threadIdx.x; i < kL1Rows * D; i += blockDim.x) {
      int d = i / kL1Rows, r = i - d * kL1Rows;  // r fastest: coalesced for SoA obs
      sobs[r][d] = (r0 + r < rows) ? map.obs[map.offset(r0 + r) + d * ds] : 0.0f;
    }
    __syncthreads();
    if (j < H) {
      const int nr = (int)min((int64_t)kL1Rows, rows - r0);
      for (int r = 0; r < nr; ++r) {
        float acc = b;
#pragma unroll
        for (int d = 0; d < kMaxD; ++d)
          if (d < D) acc = fmaf(sobs[r][d], w[d], acc);
        h1[(r0 + r) * H + j] = fmaxf(acc, 0.0f);
      }
    }
  }
}

int launch_layer1_fwd(const RowMap& map, int64_t rows, int D, int H, const float* w1,
                      const float* b1, float* h1, cudaStream_t st) {
  if (D > kMaxD || H > 256) return RL8_ERR_UNSUPPORTED;
  int grid = (int)min(ceil_div(rows, kL1Rows), (int64_t)kNumSMs * 8);
  layer1_fwd_kernel<<<grid, 256, 0, st>>>(map, rows, D, H, w1, b1, h1);
  return check_launch("layer1_fwd");
}

// ---- SGEMM 128x128x8, 256 threads, 8x8 register tile, double-buffered smem -----------------
constexpr int BM = 128, BN = 128, BK = 8;

template <bool KMAJOR>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t ld, int64_t mn0,
                                          int64_t mn_lim, int64_t k0, int64_t k_lim, float4& v,
                                          int& s_k, int& s_mn) {
  // Fetch this thread's 4 elements of a [128 (m or n)] x [8 (k)] operand tile.
  const int tid = threadIdx.x;
  v = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (KMAJOR) {
    // element (mn, k) at P[mn*ld + k]; thread -> (mn = tid/2, k4 = (tid&1)*4)
    s_mn = tid >> 1;
    s_k = (tid & 1) * 4;
    int64_t mn = mn0 + s_mn, k = k0 + s_k;
    if (mn < mn_lim) {
      const float* p = P + mn * ld + k;
      if (k + 3 < k_lim) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        if (k + 0 < k_lim) v.x = p[0];
        if (k + 1 < k_lim) v.y = p[1];
        if (k + 2 < k_lim) v.z = p[2];
      }
    }
  } else {
    // element (mn, k) at P[k*ld + mn]; thread -> (k = tid/32, mn4 = (tid&31)*4)
    s_k = tid >> 5;
    s_mn = (tid & 31) * 4;
    int64_t mn = mn0 + s_mn, k = k0 + s_k;
    if (k < k_lim) {
      const float* p = P + k * ld + mn;
      if (mn + 3 < mn_lim) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        if (mn + 0 < mn_lim) v.x = p[0];
        if (mn + 1 < mn_lim) v.y = p[1];
        if (mn + 2 < mn_lim) v.z = p[2];
      }
    }
  }
}

template <bool KMAJOR>
__device__ __forceinline__ void store_tile(float (*S)[BM], const float4& v, int s_k, int s_mn) {
  if constexpr (KMAJOR) {
    S[s_k + 0][s_mn] = v.x;
    S[s_k + 1][s_mn] = v.y;
    S[s_k + 2][s_mn] = v.z;
    S[s_k + 3][s_mn] = v.w;
  } else {
    *reinterpret_cast<float4*>(&S[s_k][s_mn]) = v;
  }
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
__global__ void __launch_bounds__(256, 2)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
             const float* __restrict__ bias, int64_t kchunk) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  const int64_t kend = min(K, kbeg + kchunk);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float4 va, vb;
  int ak, am, bk, bn;
  load_tile<A_KMAJOR>(A, lda, m0, M, kbeg, kend, va, ak, am);
  load_tile<B_KMAJOR>(B, ldb, n0, N, kbeg, kend, vb, bk, bn);
  store_tile<A_KMAJOR>(As[0], va, ak, am);
  store_tile<B_KMAJOR>(Bs[0], vb, bk, bn);
  __syncthreads();

  int buf = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = (k0 + BK) < kend;
    if (has_next) {
      load_tile<A_KMAJOR>(A, lda, m0, M, k0 + BK, kend, va, ak, am);
      load_tile<B_KMAJOR>(B, ldb, n0, N, k0 + BK, kend, vb, bk, bn);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tile<A_KMAJOR>(As[buf ^ 1], va, ak, am);
      store_tile<B_KMAJOR>(Bs[buf ^ 1], vb, bk, bn);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= N) continue;
      float* cp = C + m * ldc + n;
      float r[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if constexpr (EPI == EPI_ATOMIC) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) atomicAdd(cp + j, r[j]);
      } else {
        if constexpr (EPI == EPI_BIAS_RELU) {
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = (n + j < N) ? fmaxf(r[j] + bias[n + j], 0.0f) : 0.0f;
        }
        if (n + 3 < N) {
          if constexpr (EPI == EPI_MASK_INPLACE) {
            float4 old = *reinterpret_cast<const float4*>(cp);
            r[0] = old.x > 0.0f ? r[0] : 0.0f;
            r[1] = old.y > 0.0f ? r[1] : 0.0f;
            r[2] = old.z > 0.0f ? r[2] : 0.0f;
            r[3] = old.w > 0.0f ? r[3] : 0.0f;
          }
          *reinterpret_cast<float4*>(cp) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < N) {
              if constexpr (EPI == EPI_MASK_INPLACE) r[j] = cp[j] > 0.0f ? r[j] : 0.0f;
              cp[j] = r[j];
            }
        }
      }
    }
  }
}

template <bool AK, bool BKM>
static int launch_sgemm_epi(int epi, const float* A, const float* B, float* C, int64_t M, int N,
                            int64_t K, int64_t lda, int64_t ldb, int64_t ldc, const float* bias,
                            int splits, cudaStream_t st) {
  if (splits < 1) splits = 1;
  int64_t kchunk = round_up(ceil_div(K, splits), BK);
  splits = (int)ceil_div(K, kchunk);
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), (unsigned)splits);
  switch (epi) {
    case EPI_STORE:
      sgemm_kernel<AK, BKM, EPI_STORE><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias, kchunk);
      break;
    case EPI_BIAS_RELU:
      sgemm_kernel<AK, BKM, EPI_BIAS_RELU><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias, kchunk);
      break;
    case EPI_MASK_INPLACE:
      sgemm_kernel<AK, BKM, EPI_MASK_INPLACE><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias, kchunk);
      break;
    case EPI_ATOMIC:
      sgemm_kernel<AK, BKM, EPI_ATOMIC><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, bias, kchunk);
      break;
    default:
      return RL8_ERR_ARG;
  }
  return check_launch("sgemm");
}

int launch_sgemm(bool a_kmajor, bool b_kmajor, int epi, const float* A, const float* B, float* C,
                 int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                 const float* bias, int splits, cudaStream_t st) {
  if ((lda % 4) || (ldb % 4) || (ldc % 4) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15) ||
      ((uintptr_t)C & 15))
    return RL8_ERR_ARG;
  if (epi != EPI_ATOMIC && splits > 1) return RL8_ERR_ARG;
  if (a_kmajor && b_kmajor)
    return launch_sgemm_epi<true, true>(epi, A, B, C, M, N, K, lda, ldb, ldc, bias, splits, st);
  if (a_kmajor && !b_kmajor)
    return launch_sgemm_epi<true, false>(epi, A, B, C, M, N, K, lda, ldb, ldc, bias, splits, st);
  if (!a_kmajor && !b_kmajor)
    return launch_sgemm_epi<false, false>(epi, A, B, C, M, N, K, lda, ldb, ldc, bias, splits, st);
  return RL8_ERR_UNSUPPORTED;
}

// ---- heads -------------------------------------------------------------------------------------
// out[r][p] = b3[p] + sum_j h2[r][j] * w3[p][j]; one warp per row, H = 256 -> 8 floats / lane.
template <int P>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ h2, int64_t rows, const float* __restrict__ w3,
                const float* __restrict__ b3, float* __restrict__ out, int tanh_col1) {
  constexpr int H = 256;
  const int lane = threadIdx.x & 31;
  float w[P][8];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    float4 w0 = *reinterpret_cast<const float4*>(w3 + p * H + lane * 8);
    float4 w1 = *reinterpret_cast<const float4*>(w3 + p * H + lane * 8 + 4);
    w[p][0] = w0.x, w[p][1] = w0.y, w[p][2] = w0.z, w[p][3] = w0.w;
    w[p][4] = w1.x, w[p][5] = w1.y, w[p][6] = w1.z, w[p][7] = w1.w;
  }
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    float4 x0 = ld_stream4(h2 + r * H + lane * 8);
    float4 x1 = ld_stream4(h2 + r * H + lane * 8 + 4);
    float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    float acc[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(x[i], w[p][i], s);
      acc[p] = warp_sum(s);
    }
    if (lane == 0) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float v = acc[p] + b3[p];
        if (tanh_col1 && p == 1) v = tanhf(v);
        out[r * P + p] = v;
      }
    }
  }
}

int launch_head_fwd(const float* h2, int64_t rows, int H, int P, const float* w3, const float* b3,
                    float* out, int tanh_col1, cudaStream_t st) {
  if (H != 256) return RL8_ERR_UNSUPPORTED;
  int grid = grid_for(rows * 32, 256, 8, 2);
  switch (P) {
    case 1: head_fwd_kernel<1><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 2: head_fwd_kernel<2><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 3: head_fwd_kernel<3><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 4: head_fwd_kernel<4><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 5: head_fwd_kernel<5><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 6: head_fwd_kernel<6><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 7: head_fwd_kernel<7><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    case 8: head_fwd_kernel<8><<<grid, 256, 0, st>>>(h2, rows, w3, b3, out, tanh_col1); break;
    default: return RL8_ERR_UNSUPPORTED;
  }
  return check_launch("head_fwd");
}

// dz2[r][j] = h2[r][j] > 0 ? sum_p dout[r][p] * w3[p][j] : 0; thread -> 4 consecutive j.
template <int P>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ h2, const float* __restrict__ dout, int64_t rows,
                const float* __restrict__ w3, float* __restrict__ dz2) {
  constexpr int H = 256, Q = H / 4;
  const int64_t total = rows * Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Q;
    const int j = (int)(i - r * Q) * 4;
    float4 h = ld_stream4(h2 + r * H + j);
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float d = dout[r * P + p];
      float4 w = *reinterpret_cast<const float4*>(w3 + p * H + j);
      g[0] = fmaf(d, w.x, g[0]);
      g[1] = fmaf(d, w.y, g[1]);
      g[2] = fmaf(d, w.z, g[2]);
      g[3] = fmaf(d, w.w, g[3]);
    }
    float4 o = make_float4(h.x > 0.f ? g[0] : 0.f, h.y > 0.f ? g[1] : 0.f, h.z > 0.f ? g[2] : 0.f,
                           h.w > 0.f ? g[3] : 0.f);
    *reinterpret_cast<float4*>(dz2 + r * H + j) = o;
  }
}

int launch_head_bwd(const float* h2, const float* dout, int64_t rows, int H, int P, const float* w3,
                    float* dz2, cudaStream_t st) {
  if (H != 256) return RL8_ERR_UNSUPPORTED;
  int grid = grid_for(rows * 64, 256, 8, 4);
  switch (P) {
    case 1: head_bwd_kernel<1><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 2: head_bwd_kernel<2><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 3: head_bwd_kernel<3><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 4: head_bwd_kernel<4><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 5: head_bwd_kernel<5><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 6: head_bwd_kernel<6><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 7: head_bwd_kernel<7><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    case 8: head_bwd_kernel<8><<<grid, 256, 0, st>>>(h2, dout, rows, w3, dz2); break;
    default: return RL8_ERR_UNSUPPORTED;
  }
  return check_launch("head_bwd");
}

// ---- thin weight-gradient reductions --------------------------------------------------------------
constexpr int kThinRows = 64;
constexpr int kMaxS = 8;

template <int S>
__global__ void __launch_bounds__(256)
thin_reduce_kernel(const float* __restrict__ X, int64_t rows, int H, const float* __restrict__ Y,
                   RowMap ymap, int use_map, float* __restrict__ gw, int64_t gw_stride_s,
                   int64_t gw_stride_c, float* __restrict__ gb, int64_t rows_per_block) {
  __shared__ float sy[kThinRows][S > 0 ? S : 1];
  const int c = threadIdx.x;
  float acc[S > 0 ? S : 1], accb = 0.0f;
#pragma unroll
  for (int s = 0; s < (S > 0 ? S : 1); ++s) acc[s] = 0.0f;
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = min(rows, rbeg + rows_per_block);
  const int64_t ds = use_map ? ymap.dstride() : 1;
  for (int64_t r0 = rbeg; r0 < rend; r0 += kThinRows) {
    const int nr = (int)min((int64_t)kThinRows, rend - r0);
    if constexpr (S > 0) {
      __syncthreads();
      for (int i = threadIdx.x; i < kThinRows * S; i += blockDim.x) {
        int r, s;
        if (use_map) {
          s = i / kThinRows, r = i - s * kThinRows;
        } else {
          r = i / S, s = i - r * S;
        }
        float v = 0.0f;
        if (r < nr) v = use_map ? ymap.obs[ymap.offset(r0 + r) + s * ds] : Y[(r0 + r) * S + s];
        sy[r][s] = v;
      }
      __syncthreads();
    }
    if (c < H) {
      for (int r = 0; r < nr; ++r) {
        float x = X[(r0 + r) * H + c];
        accb += x;
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(x, sy[r][s], acc[s]);
      }
    }
  }
  if (c < H) {
#pragma unroll
    for (int s = 0; s < S; ++s) atomicAdd(gw + s * gw_stride_s + c * gw_stride_c, acc[s]);
    if (gb) atomicAdd(gb + c, accb);
  }
}

int launch_thin_reduce(const float* X, int64_t rows, int H, const float* Y, const RowMap* ymap,
                       int S, float* gw, int64_t gw_stride_s, int64_t gw_stride_c, float* gb,
                       cudaStream_t st) {
  if (H > 256 || S > kMaxS) return RL8_ERR_UNSUPPORTED;
  int64_t blocks = min(ceil_div(rows, kThinRows), (int64_t)kNumSMs * 4);
  int64_t rpb = round_up(ceil_div(rows, blocks), kThinRows);
  blocks = ceil_div(rows, rpb);
  RowMap m{};
  if (ymap) m = *ymap;
  int use_map = ymap != nullptr;
#define RL8_THIN(SV)                                                                          \
  case SV:                                                                                    \
    thin_reduce_kernel<SV><<<(int)blocks, 256, 0, st>>>(X, rows, H, Y, m, use_map, gw,         \
                                                        gw_stride_s, gw_stride_c, gb, rpb);   \
    break;
  switch (S) {
    RL8_THIN(0) RL8_THIN(1) RL8_THIN(2) RL8_THIN(3) RL8_THIN(4)
    RL8_THIN(5) RL8_THIN(6) RL8_THIN(7) RL8_THIN(8)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_THIN
  return check_launch("thin_reduce");
}

}  // namespace rl8
