// fp32 (CUDA-core) building blocks of the default two-MLP model: the parity path
// (RL8_PREC_FP32, reference `enable_amp=False`).  Launch wrappers only; kernels live in
// mlp_fp32.cu.  All matrices are row-major fp32; H is the hidden width (256).
#pragma once
#include "common.cuh"

namespace rl8 {

// How a kernel finds observation element (r, d) of its r-th row.
//   mode 0: obs[r*stride_r + d*stride_d]
//   mode 1: flattened transition g = rows ? rows[r] : row_begin + r  (row = n*T + t) inside
//           the horizon-major buffer obs[T+1][D][N]:  obs[(t*D + d)*N + n]
//   mode 2: slab-major rows r = t*N + n of the same buffer (all T+1 slabs, value pass)
struct RowMap {
  const float* obs;
  int mode;
  int64_t stride_r, stride_d;
  const int64_t* rows;
  int64_t row_begin;
  int32_t T, D;
  int64_t N;
  __host__ __device__ __forceinline__ int64_t offset(int64_t r) const {
    if (mode == 0) return r * stride_r;
    if (mode == 2) {
      int64_t t2 = r / N;
      return t2 * (int64_t)D * N + (r - t2 * N);
    }
    int64_t g = rows ? rows[r] : row_begin + r;
    int64_t n = g / T, t = g - n * T;
    return t * (int64_t)D * N + n;
  }
  __host__ __device__ __forceinline__ int64_t dstride() const { return mode == 0 ? stride_d : N; }
};

// h1[rows][H] = relu(b1 + obs @ w1^T)
int launch_layer1_fwd(const RowMap& map, int64_t rows, int D, int H, const float* w1,
                      const float* b1, float* h1, cudaStream_t st);

enum { EPI_STORE = 0, EPI_BIAS_RELU = 1, EPI_MASK_INPLACE = 2, EPI_ATOMIC = 3 };
// C[M][N] (ldc) = epilogue(sum_k A(m,k) * B(k,n)).
//   a_kmajor: A(m,k) = A[m*lda + k] else A[k*lda + m];  b_kmajor: B(k,n) = B[n*ldb + k] else B[k*ldb + n]
//   EPI_BIAS_RELU: relu(acc + bias[n]);  EPI_MASK_INPLACE: C[m,n] = C[m,n] > 0 ? acc : 0;
//   EPI_ATOMIC: split-K over gridDim.z, atomicAdd into C.
int launch_sgemm(bool a_kmajor, bool b_kmajor, int epi, const float* A, const float* B, float* C,
                 int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
                 const float* bias, int splits, cudaStream_t st);

// The same contract on tcgen05 (gemm_tc.cu): operands converted to bf16 while staged, fp32 accumulate;
// EPI_STORE / EPI_ATOMIC only; 16-byte aligned pointers, lda / ldb / ldc / N multiples of 4.
int launch_tc_gemm(bool a_kmajor, bool b_kmajor, int epi, const float* A, const float* B, float* C,
                   int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int splits,
                   cudaStream_t st);

// One LSTM step on tensor cores (gemm_tc.cu): gates = h W_hh^T (bf16 operands, fp32 accumulate) with the cell
// non-linearity in the epilogue; writes c_out, h_out [rows][H] and, for the backward pass, the gate PRE-activations as
// bf16 T128 (zb [rows_pad][4H], may be NULL).  H = 256.  h arrives as bf16 in the T128 layout (hb_in:
// launch_pack_t128 of the fp32 state, or a previous step's hb_out).
int64_t t128_bytes(int64_t rows, int cols);
int launch_pack_t128(const float* x, int64_t rows, uint8_t* out, cudaStream_t st);
int launch_lstm_cell_tc(const uint8_t* hb_in, const float* w_hh, const float* w_ih, const float* b_ih, const float* b_hh,
                        const float* c_prev, const RowMap& xmap, int D, int64_t rows, uint8_t* zb, float* c_out,
                        float* h_out, uint8_t* hb_out, cudaStream_t st);

// out[rows][P] = b3 + h2 @ w3^T; column 1 is tanh'ed when tanh_col1 (continuous log_std).
int launch_head_fwd(const float* h2, int64_t rows, int H, int P, const float* w3, const float* b3,
                    float* out, int tanh_col1, cudaStream_t st);

// dz2[r][j] = h2[r][j] > 0 ? sum_p dout[r][p] * w3[p][j] : 0
int launch_head_bwd(const float* h2, const float* dout, int64_t rows, int H, int P, const float* w3,
                    float* dz2, cudaStream_t st);

// Thin weight-gradient reductions over rows (atomicAdd into the gradient buffers):
//   gw[s*gw_stride_s + c*gw_stride_c] += sum_r X[r][c] * Y(r, s)   s < S
//   gb[c]                             += sum_r X[r][c]             (gb may be NULL)
// Y is either a dense [rows][S] matrix (ymap == NULL) or observations through a RowMap.
int launch_thin_reduce(const float* X, int64_t rows, int H, const float* Y, const RowMap* ymap,
                       int S, float* gw, int64_t gw_stride_s, int64_t gw_stride_c, float* gb,
                       cudaStream_t st);

}  // namespace rl8
