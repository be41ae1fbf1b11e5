// Backward kernels of the recurrent path on tensor cores (lstm_tc.cu); operands in the bf16 T128 layout (tc.cuh).
#pragma once
#include "mlp_fp32.cuh"

namespace rl8 {

struct LstmBwdArgs {
  const uint8_t* zb;    // bf16 T128 [R_pad][1024] gate PRE-activations of step k (tc_lstm_cell_kernel)
  const float* c_prev;  // [R][256]
  const float* h;       // [R][256]    h_k (head weight gradients)
  const float* dh_rec;  // [R][256]    dG_{k+1} W_hh, or null at the last step of the sequence
  float* dc;            // [R][256]    dL/dc: read when has_dc_in, written for step k-1
  const float* dpi;     // [R][P]      dL/d(policy head outputs)
  const float* dvf;     // [R]
  const float* pi_w;    // [P][256]
  const float* vf_w;    // [256]
  uint8_t* dGb;         // bf16 T128 [R_pad][1024]
  uint8_t* xb;          // bf16 T128 [R_pad][16]: x_0 .. x_{D-1}, 0.., then 1 at column 8
  float* gpi_w;         // [P][256] +=
  float* gvf_w;         // [256] +=
  RowMap xmap;
  int D, has_dc_in;
  int64_t rows, rows_pad, rows_per_block;  // rows_per_block: multiple of 32
};
// out_pi[rows][P] = pi_b + h pi_w^T (column 1 tanh'ed when tanh_col1), out_vf[rows] = vf_b + h vf_w^T: one pass over h.
// HeadBlocks: steps > 1 -> the launch covers `steps` blocks of `rows` rows (the time steps of a TBPTT chunk); block k
// reads h + k * h_stride rows and writes out_pi + k * pi_stride, out_vf + k * vf_stride (strides in elements of each).
struct HeadBlocks {
  int steps = 1;
  int64_t h_stride = 0, pi_stride = 0, vf_stride = 0;
};
int launch_lstm_heads_fwd(const float* h, int64_t rows, int P, const float* pi_w, const float* pi_b, const float* vf_w,
                          const float* vf_b, float* out_pi, float* out_vf, int tanh_col1, cudaStream_t st,
                          HeadBlocks hb = HeadBlocks());

// dG_k (bf16 T128), dL/dc for step k-1, [x | 1] (bf16 T128) and the head weight gradients of one BPTT step
int launch_lstm_cell_bwd_tc(const LstmBwdArgs& args, int P, cudaStream_t st);

// dh[R][256] = dG[R][1024] W_hh[1024][256]  (dG as bf16 T128, W_hh fp32 -> bf16 in the kernel)
int launch_lstm_dh_tc(const uint8_t* dGb, const float* w_hh, float* dh, int64_t rows, cudaStream_t st);

struct LstmWgArgs {
  const uint8_t *dGb, *hb, *xb;              // first step's images
  int64_t dGb_stride, hb_stride, xb_stride;  // bytes between consecutive steps
  int L;
  int64_t row_tiles;  // 128-row tiles per step
  float *gw_hh, *gw_ih, *gb_ih, *gb_hh;
  int D;
};
// gw_hh[1024][256] += sum_k dG_k^T h_{k-1};  gw_ih[1024][D] += dG^T x;  gb_ih, gb_hh [1024] += column sums of dG
int launch_lstm_wgrad_tc(const LstmWgArgs& g, cudaStream_t st);

}  // namespace rl8
