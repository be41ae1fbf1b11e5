// Recurrent (LSTM) policy, RL8_PREC_BF16: the backward pass of one truncated-BPTT step on tensor cores.
//
// Hand-derived counterpart of autograd through nn.LSTM (src/rl8/models/_recurrent.py:259-341) inside the reference's
// update loop (src/rl8/algorithms/_recurrent.py:517-600); restated on the CPU in oracle/recurrent_oracle.py.
// Three kernels per step (lstm_fp32.cu holds the host loop), all operands of the two contractions travel through HBM
// as bf16 in the T128 layout (tc.cuh: operand-format tiles), so that staging a tile is one bulk copy:
//
//   lstm_cell_bwd_tc_kernel   elementwise, HBM-bound.  dL/dh = heads' term (+ recurrent term), cell backward, gate
//       pre-activation gradients dG -> bf16 T128 (never stored in fp32), dL/dc carried, [x | 1] -> bf16 T128 (the
//       operand that turns the input-weight / bias gradients into GEMM columns), head weight gradients accumulated per
//       thread.  Per row and step: 28 B x 256 read, 12 B x 256 written.
//   lstm_dh_tc_kernel         dh_{k-1} = dG_k W_hh: a CTA keeps the 64 hidden columns it owns of W_hh in shared memory
//       (bf16, MN-major, 128 KB) and streams 128-row x 128-k tiles of dG through a 3-stage bulk-copy ring.
//   lstm_wgrad_tc_kernel      [gW_hh | gW_ih | gb] += dG^T [h_{k-1} | x | 1] over ALL steps of the chunk in one launch:
//       a CTA owns 128 gate rows x (256 + 16) columns of accumulators in tensor memory for its share of the row tiles
//       (contraction over rows: both operands MN-major views of the T128 tiles), one vector-atomic flush at the end.
#include "dist.cuh"
#include "lstm_tc.cuh"
#include "tc.cuh"

namespace rl8 {

using namespace tc;

constexpr int kTH = 256;  // hidden width of the default recurrent models

__device__ __forceinline__ void ldg8f(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg8f(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// ---- both heads in one pass over h (head_fwd_kernel's arithmetic per output; one warp per row) -----------------------
template <int P>
__global__ void __launch_bounds__(256)
lstm_heads_fwd_kernel(const float* __restrict__ h, int64_t rows, const float* __restrict__ pi_w,
                      const float* __restrict__ pi_b, const float* __restrict__ vf_w, const float* __restrict__ vf_b,
                      float* __restrict__ out_pi, float* __restrict__ out_vf, int tanh_col1, HeadBlocks hb) {
  const int lane = threadIdx.x & 31;
  float w[P + 1][8];
#pragma unroll
  for (int p = 0; p <= P; ++p) {
    const float* src = (p < P ? pi_w + p * kTH : vf_w) + lane * 8;
    const float4 w0 = *reinterpret_cast<const float4*>(src), w1 = *reinterpret_cast<const float4*>(src + 4);
    w[p][0] = w0.x, w[p][1] = w0.y, w[p][2] = w0.z, w[p][3] = w0.w;
    w[p][4] = w1.x, w[p][5] = w1.y, w[p][6] = w1.z, w[p][7] = w1.w;
  }
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // two rows per trip: both rows' loads are in flight before the first dot product (the streaming loads are
  // volatile asm, the compiler does not hoist them across iterations by itself)
  // row ra of the launch = row r of block ra / rows (the time steps of a TBPTT chunk, or one block)
  auto locate = [&](int64_t ra, int64_t& hrow, int64_t& pi_at, int64_t& vf_at) {
    const int64_t blk = hb.steps > 1 ? ra / rows : 0, r = ra - blk * rows;
    hrow = blk * hb.h_stride + r, pi_at = blk * hb.pi_stride + r * P, vf_at = blk * hb.vf_stride + r;
  };
  const int64_t total = rows * (hb.steps > 1 ? hb.steps : 1);
  for (int64_t r0 = warp; r0 < total; r0 += 2 * nwarps) {
    const int64_t r1 = r0 + nwarps;
    const bool two = r1 < total;
    int64_t hr[2] = {0, 0}, pa[2] = {0, 0}, va[2] = {0, 0};
    locate(r0, hr[0], pa[0], va[0]);
    if (two) locate(r1, hr[1], pa[1], va[1]);
    float4 xa[2], xb[2];
    xa[0] = ld_stream4(h + hr[0] * kTH + lane * 8), xb[0] = ld_stream4(h + hr[0] * kTH + lane * 8 + 4);
    xa[1] = xb[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (two) xa[1] = ld_stream4(h + hr[1] * kTH + lane * 8), xb[1] = ld_stream4(h + hr[1] * kTH + lane * 8 + 4);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const float x[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
      float acc[P + 1];
#pragma unroll
      for (int p = 0; p <= P; ++p) {
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum = fmaf(x[i], w[p][i], sum);
        acc[p] = warp_sum(sum);
      }
      if (lane == 0) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
          float v = acc[p] + pi_b[p];
          if (tanh_col1 && p == 1) v = tanhf(v);
          out_pi[pa[u] + p] = v;
        }
        out_vf[va[u]] = acc[P] + vf_b[0];
      }
    }
  }
}

int launch_lstm_heads_fwd(const float* h, int64_t rows, int P, const float* pi_w, const float* pi_b, const float* vf_w,
                          const float* vf_b, float* out_pi, float* out_vf, int tanh_col1, cudaStream_t st,
                          HeadBlocks hb) {
  // one resident wave: the head weights are loaded once per warp
  const int grid = grid_for(rows * (hb.steps > 1 ? hb.steps : 1) * 32, 256, 8, 1);
#define RL8_HEADS(PV)                                                                                             \
  case PV:                                                                                                        \
    lstm_heads_fwd_kernel<PV><<<grid, 256, 0, st>>>(h, rows, pi_w, pi_b, vf_w, vf_b, out_pi, out_vf, tanh_col1, hb); \
    break;
  switch (P) {
    RL8_HEADS(2) RL8_HEADS(3) RL8_HEADS(4) RL8_HEADS(5) RL8_HEADS(6) RL8_HEADS(7) RL8_HEADS(8)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_HEADS
  return check_launch("lstm_heads_fwd");
}

// ---- cell backward ----------------------------------------------------------------------------------------------

constexpr int kBwdGroup = 32;  // rows per transposition pass

// lane <-> hidden unit 32 w + lane of warp w; a warp walks its block's rows in groups of 32.  The group's gate
// pre-activations of the warp's 32 units (4 gates x 4 column groups of the bf16 T128 image) arrive as 512-byte runs
// (lane = row) in the WARP'S OWN 8 KB of shared memory; every thread then reads the values of ITS unit, recomputes the
// gate activations and c_k from them (cheaper than 6 KB per row of stored fp32 activations), and writes the gate
// gradient back IN PLACE; the transposed block leaves as 512-byte runs again.  No block-wide barrier anywhere: the 8
// warps of a block (and the blocks of an SM) drift apart, so the load bursts of some overlap the arithmetic of others.
template <int P>
__global__ void __launch_bounds__(256, 3) lstm_cell_bwd_tc_kernel(LstmBwdArgs a) {
  extern __shared__ __align__(16) uint8_t dgs_all[];  // per warp: [4 gates x 4 column groups][32 rows][16 B] = 8 KB
  const int j = threadIdx.x, warp = j >> 5, lane = j & 31;
  uint8_t* dgs = dgs_all + warp * (16 * kBwdGroup * 16);
  float wpi[P], gpi[P], wvf = a.vf_w[j], gvf = 0.0f;
#pragma unroll
  for (int p = 0; p < P; ++p) wpi[p] = a.pi_w[p * kTH + j], gpi[p] = 0.0f;
  const int64_t rbeg = (int64_t)blockIdx.x * a.rows_per_block;
  const int64_t rend = rbeg + a.rows_per_block < a.rows_pad ? rbeg + a.rows_per_block : a.rows_pad;
  const int cg = lane >> 3;  // column group of this lane inside the warp's 32 units
  for (int64_t r0 = rbeg; r0 < rend; r0 += kBwdGroup) {
    // pre-activations in: 16 (gate, column group) runs, lane = row; row slot xor-swizzled by the column group so that
    // the per-unit 2-byte accesses below (4 column groups per warp instruction) hit different banks
#pragma unroll
    for (int s0 = 0; s0 < 16; s0 += 8) {
      uint4 zq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int seg = s0 + e;
        zq[e] = __ldg(reinterpret_cast<const uint4*>(a.zb + t128_offset(r0 + lane, (seg >> 2) * 32 + warp * 4 + (seg & 3), 128)));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int seg = s0 + e;
        *reinterpret_cast<uint4*>(dgs + (size_t)(seg * kBwdGroup + (lane ^ ((seg & 3) << 1))) * 16) = zq[e];
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int rr0 = 0; rr0 < kBwdGroup; rr0 += 4) {
      float cp[4], dhr[4], dci[4], hk[4], dp[4][P], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = r0 + rr0 + u;
        const bool live = r < a.rows;
        const int64_t i = r * kTH + j;
        cp[u] = live ? __ldg(a.c_prev + i) : 0.0f;
        hk[u] = live ? __ldg(a.h + i) : 0.0f;
        dhr[u] = (live && a.dh_rec) ? __ldg(a.dh_rec + i) : 0.0f;
        dci[u] = (live && a.has_dc_in) ? a.dc[i] : 0.0f;
#pragma unroll
        for (int p = 0; p < P; ++p) dp[u][p] = live ? __ldg(a.dpi + r * P + p) : 0.0f;
        dv[u] = live ? __ldg(a.dvf + r) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = r0 + rr0 + u;
        const int rr = rr0 + u;
        const bool live = r < a.rows;
        const uint32_t off = (uint32_t)((cg * kBwdGroup + (rr ^ (cg << 1))) * 16 + (lane & 7) * 2);
        constexpr uint32_t gate_stride = 4 * kBwdGroup * 16;  // 4 column groups per gate
        __nv_bfloat16* zi = reinterpret_cast<__nv_bfloat16*>(dgs + off);
        __nv_bfloat16* zf = reinterpret_cast<__nv_bfloat16*>(dgs + gate_stride + off);
        __nv_bfloat16* zg = reinterpret_cast<__nv_bfloat16*>(dgs + 2 * gate_stride + off);
        __nv_bfloat16* zo = reinterpret_cast<__nv_bfloat16*>(dgs + 3 * gate_stride + off);
        // gates and cell state of the forward step, recomputed (tc_lstm_cell_kernel's functions)
        const float ig = sigmoid_fast(__bfloat162float(*zi)), fg = sigmoid_fast(__bfloat162float(*zf));
        const float gg = tanh_fast(__bfloat162float(*zg)), og = sigmoid_fast(__bfloat162float(*zo));
        const float tc_ = tanh_fast(fg * cp[u] + ig * gg);
        // dL/dh_k = heads' term (+ the recurrent term), in the op order of lstm_dh_kernel
        float dhv = 0.0f;
#pragma unroll
        for (int p = 0; p < P; ++p) dhv = fmaf(dp[u][p], wpi[p], dhv);
        dhv = fmaf(dv[u], wvf, dhv);
        if (a.dh_rec) dhv += dhr[u];
        // head weight gradients
#pragma unroll
        for (int p = 0; p < P; ++p) gpi[p] = fmaf(dp[u][p], hk[u], gpi[p]);
        gvf = fmaf(dv[u], hk[u], gvf);
        // cell backward (lstm_cell_bwd_kernel)
        float dcv = dhv * og * (1.0f - tc_ * tc_);
        if (a.has_dc_in) dcv += dci[u];
        const float di = dcv * gg * ig * (1.0f - ig);
        const float df = dcv * cp[u] * fg * (1.0f - fg);
        const float dg = dcv * ig * (1.0f - gg * gg);
        const float dO = dhv * tc_ * og * (1.0f - og);
        if (live) a.dc[r * kTH + j] = dcv * fg;
        // rows past the end carry a zero gradient whatever their (unwritten) pre-activations hold
        *zi = __float2bfloat16_rn(live ? di : 0.0f);
        *zf = __float2bfloat16_rn(live ? df : 0.0f);
        *zg = __float2bfloat16_rn(live ? dg : 0.0f);
        *zo = __float2bfloat16_rn(live ? dO : 0.0f);
      }
    }
    __syncwarp();
    // 16 (gate, column group) runs of 32 rows x 16 B: lane = row, 512 contiguous bytes of the T128 image per store
#pragma unroll
    for (int seg = 0; seg < 16; ++seg) {
      const int gate = seg >> 2, c = seg & 3;
      *reinterpret_cast<uint4*>(a.dGb + t128_offset(r0 + lane, gate * 32 + warp * 4 + c, 128)) =
          *reinterpret_cast<const uint4*>(dgs + (size_t)(seg * kBwdGroup + (lane ^ (c << 1))) * 16);
    }
    if (warp == 0) {  // [x | 1] of the group's rows (zeros past the last row)
      const int64_t r = r0 + lane;
      float x[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) x[d] = 0.0f;
      float one = 0.0f;
      if (r < a.rows) {
        const int64_t xo = a.xmap.offset(r), ds = a.xmap.dstride();
#pragma unroll
        for (int d = 0; d < 8; ++d)
          if (d < a.D) x[d] = a.xmap.obs[xo + d * ds];
        one = 1.0f;
      }
      uint4 q0, q1;
      q0.x = pack_bf16x2(x[0], x[1]), q0.y = pack_bf16x2(x[2], x[3]);
      q0.z = pack_bf16x2(x[4], x[5]), q0.w = pack_bf16x2(x[6], x[7]);
      q1.x = pack_bf16x2(one, 0.0f), q1.y = q1.z = q1.w = 0u;
      *reinterpret_cast<uint4*>(a.xb + t128_offset(r, 0, 2)) = q0;
      *reinterpret_cast<uint4*>(a.xb + t128_offset(r, 1, 2)) = q1;
    }
    __syncwarp();
  }
#pragma unroll
  for (int p = 0; p < P; ++p) atomicAdd(a.gpi_w + p * kTH + j, gpi[p]);
  atomicAdd(a.gvf_w + j, gvf);
}

int launch_lstm_cell_bwd_tc(const LstmBwdArgs& args, int P, cudaStream_t st) {
  LstmBwdArgs a = args;
  a.rows_pad = (a.rows + 127) / 128 * 128;
  // 3 resident blocks per SM (register-bound), every block a whole number of 32-row groups
  int64_t blocks = (int64_t)kNumSMs * 3;
  int64_t rpb = ceil_div(ceil_div(a.rows_pad, blocks), (int64_t)kBwdGroup) * kBwdGroup;
  blocks = ceil_div(a.rows_pad, rpb);
  a.rows_per_block = rpb;
  constexpr int smem = 8 * 16 * kBwdGroup * 16;
#define RL8_BWD(PV)                                                                                          \
  case PV:                                                                                                   \
    cudaFuncSetAttribute(lstm_cell_bwd_tc_kernel<PV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);     \
    lstm_cell_bwd_tc_kernel<PV><<<(unsigned)blocks, 256, smem, st>>>(a);                                      \
    break;
  switch (P) {
    RL8_BWD(2) RL8_BWD(3) RL8_BWD(4) RL8_BWD(5) RL8_BWD(6) RL8_BWD(7) RL8_BWD(8)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_BWD
  return check_launch("lstm_cell_bwd_tc");
}

// ---- dh_{k-1} = dG_k W_hh ---------------------------------------------------------------------------------------------
constexpr int kDhStages = 3;
constexpr int kDhStageBytes = 128 * 128 * 2;  // 128 rows x 128 k
constexpr int kDhThreads = 256;
struct SmemDh {
  uint8_t b[1024 * 64 * 2];               // 131072  W_hh[:, 64 nb .. 64 nb + 64): MN-major, [8 column groups][1024 k][16 B]
  uint8_t a[kDhStages][kDhStageBytes];    //  98304
  uint64_t full[kDhStages], free_[kDhStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
struct LstmDhArgs {
  const uint8_t* dGb;  // bf16 T128 [R_pad][1024]
  const float* w_hh;   // [1024][256]
  float* dh;           // [R][256]
  int64_t rows;
};

__global__ void __launch_bounds__(kDhThreads, 1) lstm_dh_tc_kernel(LstmDhArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemDh& s = *reinterpret_cast<SmemDh*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = (int)(blockIdx.x & 3);
  const int64_t mt0 = blockIdx.x >> 2, mstride = gridDim.x >> 2;
  const int64_t mtiles = (g.rows + 127) / 128;
  if (tid == 0) {
    for (int i = 0; i < kDhStages; ++i) mbar_init(&s.full[i], 1), mbar_init(&s.free_[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&s.acc_full[i], 1), mbar_init(&s.acc_empty[i], 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&s.tmem_base, 128);
  // W_hh block: chunk (column group c8, k) = W_hh[k][64 nb + 8 c8 .. + 8)
  for (int e = tid; e < 8 * 1024; e += kDhThreads) {
    const int c8 = e & 7, k = e >> 3;
    const float4 lo = __ldg(reinterpret_cast<const float4*>(g.w_hh + (int64_t)k * kTH + nb * 64 + c8 * 8));
    const float4 hi = __ldg(reinterpret_cast<const float4*>(g.w_hh + (int64_t)k * kTH + nb * 64 + c8 * 8) + 1);
    const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    store_chunk(s.b, (uint32_t)((c8 * 1024 + k) * 16), v);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;

  if (warp == 0) {
    // ---- loader: eight 32 KB K chunks per row tile --------------------------------------------------------------
    uint32_t cnt = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride) {
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc, ++cnt) {
        const int st = (int)(cnt % kDhStages);
        const uint32_t use = cnt / kDhStages;
        if (use > 0) mbar_wait(&s.free_[st], (use - 1) & 1);
        if (elect_one()) {
          mbar_expect_tx(&s.full[st], kDhStageBytes);
          bulk_g2s(s.a[st], g.dGb + t128_offset(mt * 128, kc * 16, 128), kDhStageBytes, &s.full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- issuer -----------------------------------------------------------------------------------------------------
    uint32_t cnt = 0;
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
      const int buf = it & 1;
      const uint32_t acc = tmem + (uint32_t)(buf * 64);
      if (it >= 2) mbar_wait(&s.acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc, ++cnt) {
        const int st = (int)(cnt % kDhStages);
        mbar_wait(&s.full[st], (cnt / kDhStages) & 1);
        fence_after_sync();
        if (elect_one()) {
          // A: K-major tile of 128 rows; B: MN-major, rows = k (1024 of them), this chunk starts at k = 128 kc
          issue_gemm(acc, smem_u32(s.a[st]), 128, false, smem_u32(s.b) + (uint32_t)(kc * 128 * 16), 1024, true, 128, 64,
                     128, kc > 0);
          mma_commit(&s.free_[st]);
          if (kc == 7) mma_commit(&s.acc_full[buf]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ---- epilogue: warp (q = warp & 3) -> row 32 q + lane, the CTA's 64 columns ----------------------------------------
    const int q = warp & 3;
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
      const int buf = it & 1;
      const int64_t r = mt * 128 + q * 32 + lane;
      mbar_wait(&s.acc_full[buf], (uint32_t)((it >> 1) & 1));
      fence_after_sync();
      const uint32_t base = tmem + (uint32_t)(buf * 64) + ((uint32_t)(q * 32) << 16);
      float v[2][32];
      tmem_ld32(base, v[0]);
      tmem_ld32(base + 32, v[1]);
      fence_before_sync();
      __syncwarp();
      if (elect_one()) mbar_arrive(&s.acc_empty[buf]);
      __syncwarp();
      if (r < g.rows) {
        float* dst = g.dh + r * kTH + nb * 64;
#pragma unroll
        for (int e = 0; e < 64; e += 8) stg8f(dst + e, &v[e >> 5][e & 31]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

int launch_lstm_dh_tc(const uint8_t* dGb, const float* w_hh, float* dh, int64_t rows, cudaStream_t st) {
  if (!dGb || !w_hh || !dh || rows <= 0 || ((uintptr_t)dh & 31u) || ((uintptr_t)w_hh & 15u)) return RL8_ERR_ARG;
  LstmDhArgs g{dGb, w_hh, dh, rows};
  const int64_t mtiles = ceil_div(rows, (int64_t)128);
  int64_t per_block = kNumSMs / 4;
  if (per_block > mtiles) per_block = mtiles;
  cudaFuncSetAttribute(lstm_dh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDh));
  lstm_dh_tc_kernel<<<(unsigned)(per_block * 4), kDhThreads, sizeof(SmemDh), st>>>(g);
  return check_launch("lstm_dh_tc");
}

// ---- [gW_hh | gW_ih | gb] += dG^T [h_prev | x | 1] ----------------------------------------------------------------------
constexpr int kWgThreads = 256;
constexpr int kWgABytes = 128 * 128 * 2, kWgBBytes = 128 * 256 * 2, kWgXBytes = 128 * 16 * 2;
struct WgStage {
  uint8_t a[kWgABytes];  // dG tile: 128 rows (K) x this CTA's 128 gate columns (M), MN-major
  uint8_t b[kWgBBytes];  // h_prev tile: 128 rows (K) x 256 (N), MN-major
  uint8_t x[kWgXBytes];  // [x | 1] tile: 128 rows (K) x 16 (N), MN-major
};
struct SmemWg {
  WgStage st[2];
  uint64_t full[2], free_[2], done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kWgThreads, 1) lstm_wgrad_tc_kernel(LstmWgArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemWg& s = *reinterpret_cast<SmemWg*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = (int)(blockIdx.x & 7);  // gate rows [128 mt, 128 mt + 128)
  const int64_t ks = blockIdx.x >> 3, nks = gridDim.x >> 3;
  const int64_t tiles = (int64_t)g.L * g.row_tiles;
  const int64_t n_my = ks < tiles ? (tiles - ks + nks - 1) / nks : 0;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&s.full[i], 1), mbar_init(&s.free_[i], 1);
    mbar_init(&s.done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&s.tmem_base, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  if (n_my > 0) {
    if (warp == 0) {
      for (int64_t i = 0; i < n_my; ++i) {
        const int st = (int)(i & 1);
        if (i >= 2) mbar_wait(&s.free_[st], (uint32_t)(((i >> 1) - 1) & 1));
        if (elect_one()) {
          const int64_t t = ks + i * nks, k = t / g.row_tiles, rt = t - k * g.row_tiles;
          mbar_expect_tx(&s.full[st], kWgABytes + kWgBBytes + kWgXBytes);
          bulk_g2s(s.st[st].a, g.dGb + k * g.dGb_stride + t128_offset(rt * 128, mt * 16, 128), kWgABytes, &s.full[st]);
          bulk_g2s(s.st[st].b, g.hb + k * g.hb_stride + t128_offset(rt * 128, 0, 32), kWgBBytes, &s.full[st]);
          bulk_g2s(s.st[st].x, g.xb + k * g.xb_stride + t128_offset(rt * 128, 0, 2), kWgXBytes, &s.full[st]);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      for (int64_t i = 0; i < n_my; ++i) {
        const int st = (int)(i & 1);
        mbar_wait(&s.full[st], (uint32_t)((i >> 1) & 1));
        fence_after_sync();
        if (elect_one()) {
          // contraction over the tile's 128 rows: both operands MN-major (tile rows = K)
          issue_gemm(tmem, smem_u32(s.st[st].a), 128, true, smem_u32(s.st[st].b), 128, true, 128, 256, 128, i > 0);
          issue_gemm(tmem + 256, smem_u32(s.st[st].a), 128, true, smem_u32(s.st[st].x), 128, true, 128, 16, 128, i > 0);
          mma_commit(&s.free_[st]);
          if (i == n_my - 1) mma_commit(&s.done);
        }
        __syncwarp();
      }
    }
    // ---- flush: warp (q = warp & 3, half = warp >> 2) -> gate row 128 mt + 32 q + lane, 128 hidden columns -------------
    mbar_wait(&s.done, 0);
    fence_after_sync();
    const int q = warp & 3, half = warp >> 2;
    const int row = mt * 128 + q * 32 + lane;
    const uint32_t base = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(base + (uint32_t)(half * 128 + c0), v);
      float* dst = g.gw_hh + (int64_t)row * kTH + half * 128 + c0;
#pragma unroll
      for (int e = 0; e < 32; e += 4) red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
    }
    if (half == 0) {
      float v[16];
      tmem_ld16(base + 256, v);
      for (int d = 0; d < g.D; ++d) atomicAdd(g.gw_ih + (int64_t)row * g.D + d, v[d]);
      atomicAdd(g.gb_ih + row, v[8]);
      atomicAdd(g.gb_hh + row, v[8]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int launch_lstm_wgrad_tc(const LstmWgArgs& g, cudaStream_t st) {
  if (!g.dGb || !g.hb || !g.xb || g.L <= 0 || g.row_tiles <= 0) return RL8_ERR_ARG;
  const int64_t tiles = (int64_t)g.L * g.row_tiles;
  int64_t nks = kNumSMs / 8;
  if (nks > tiles) nks = tiles;
  cudaFuncSetAttribute(lstm_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemWg));
  lstm_wgrad_tc_kernel<<<(unsigned)(nks * 8), kWgThreads, sizeof(SmemWg), st>>>(g);
  return check_launch("lstm_wgrad_tc");
}

}  // namespace rl8
