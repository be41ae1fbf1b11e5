// Generic tensor-core GEMM for the recurrent path's three contractions (RL8_PREC_BF16):
//
//     C[M][N] (ldc)  =  or  +=   sum_k A(m, k) * B(k, n)        fp32 in global memory, either major
//
// The operands are fp32 in HBM (LSTM states, gate activations / gradients, W_hh); a CTA converts
// each K chunk to bf16 WHILE staging it into the library's chunked shared-memory format -- eight
// consecutive elements along the operand's unit-stride axis are one 16-byte chunk in either major,
// so both layouts are read with 128-bit loads and written with 128-bit stores -- and one elected
// lane feeds tcgen05 (fp32 accumulation in TMEM).  Two stages: the loads of chunk k+1 run under
// the MMAs of chunk k.  Grid = (M tiles of 128, N tiles of 256, K splits); split-K accumulates
// with vector reductions (EPI_ATOMIC), otherwise the tile is stored (EPI_STORE).
//
// Replaces, for `enable_amp=True`, the cuBLAS / ATen GEMMs inside nn.LSTM's forward and autograd
// backward that the reference runs (src/rl8/models/_recurrent.py:259-341 through
// src/rl8/algorithms/_recurrent.py:517-600): gates = h W_hh^T, dh = dG W_hh, gW_hh += dG^T h.
#include "mlp_fp32.cuh"
#include "tc.cuh"

namespace rl8 {

using namespace tc;

constexpr int kGM = 128;        // rows of C per CTA (MMA M)
constexpr int kGN = 256;        // columns of C per CTA (MMA N)
constexpr int kGK = 64;         // K elements per stage
constexpr int kGThreads = 256;  // 8 warps: all stage operands; warp 0 elects the issuer

struct SmemGemm {
  uint8_t a[2][kGM * kGK * 2];  // 2 x 16 KB
  uint8_t b[2][kGN * kGK * 2];  // 2 x 32 KB
  uint64_t bar[2];
  uint32_t tmem_base;
};

struct GemmArgs {
  const float *A, *B;
  float* C;
  int64_t M, K, lda, ldb, ldc, kchunk;
  int N;
  unsigned ntn;  // column tiles
};

// Eight consecutive fp32 elements at p (the first `valid` of them inside the matrix) -> one bf16 chunk.
struct Chunk8 {
  float4 lo, hi;
};
__device__ __forceinline__ Chunk8 load8(const float* p, bool in_range, int64_t valid) {
  Chunk8 c;
  c.lo = make_float4(0.f, 0.f, 0.f, 0.f), c.hi = c.lo;
  if (in_range) {
    if (valid >= 8) {
      c.lo = __ldg(reinterpret_cast<const float4*>(p));
      c.hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
    } else {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = j < valid ? p[j] : 0.0f;
      c.lo = make_float4(v[0], v[1], v[2], v[3]), c.hi = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  return c;
}
__device__ __forceinline__ uint4 pack8(const Chunk8& c) {
  uint4 q;
  q.x = pack_bf16x2(c.lo.x, c.lo.y), q.y = pack_bf16x2(c.lo.z, c.lo.w);
  q.z = pack_bf16x2(c.hi.x, c.hi.y), q.w = pack_bf16x2(c.hi.z, c.hi.w);
  return q;
}
constexpr int kGBatch = 4;  // chunks a thread has in flight: the staging loop is latency-, not issue-bound

// Stage ROWS x kGK of an operand whose unit-stride axis is K (X(i, k) = X[i * ld + k]) as a K-major tile.
template <int ROWS>
__device__ __forceinline__ void stage_kmajor(uint8_t* tile, const float* X, int64_t ld, int64_t i0, int64_t imax,
                                             int64_t k0, int64_t kmax) {
  // 8 chunks of 8 k per row: consecutive threads read 256 contiguous bytes of one row
  constexpr int kTasks = ROWS * (kGK / 8);
  static_assert(kTasks % (kGBatch * kGThreads) == 0, "tile tasks must divide evenly");
  for (int e0 = threadIdx.x; e0 < kTasks; e0 += kGBatch * kGThreads) {
    Chunk8 c[kGBatch];
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kGThreads, row = e >> 3, kc = e & 7;
      const int64_t i = i0 + row, k = k0 + kc * 8;
      c[b] = load8(X + i * ld + k, i < imax && k < kmax, kmax - k);
    }
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kGThreads, row = e >> 3, kc = e & 7;
      *reinterpret_cast<uint4*>(tile + chunk_offset<ROWS>(row, kc)) = pack8(c[b]);
    }
  }
}
// Stage kGK x COLS of an operand whose unit-stride axis is M/N (X(i, k) = X[k * ld + i]) as an MN-major tile
// (tile rows = K, 16-byte chunks = 8 consecutive i).
template <int COLS>
__device__ __forceinline__ void stage_mnmajor(uint8_t* tile, const float* X, int64_t ld, int64_t i0, int64_t imax,
                                              int64_t k0, int64_t kmax) {
  constexpr int CG = COLS / 8;  // column groups per K row: consecutive threads read one row of X contiguously
  constexpr int kTasks = kGK * CG;
  static_assert(kTasks % (kGBatch * kGThreads) == 0, "tile tasks must divide evenly");
  for (int e0 = threadIdx.x; e0 < kTasks; e0 += kGBatch * kGThreads) {
    Chunk8 c[kGBatch];
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kGThreads, kr = e / CG, cg = e - kr * CG;
      const int64_t k = k0 + kr, i = i0 + cg * 8;
      c[b] = load8(X + k * ld + i, k < kmax && i < imax, imax - i);
    }
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kGThreads, kr = e / CG, cg = e - kr * CG;
      *reinterpret_cast<uint4*>(tile + chunk_offset<kGK>(kr, cg)) = pack8(c[b]);
    }
  }
}

template <bool AK, bool BK, bool ATOMIC>
__global__ void __launch_bounds__(kGThreads, 2) tc_gemm_kernel(GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemGemm& s = *reinterpret_cast<SmemGemm*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // column tiles are the fast block index: the CTAs that share a 128-row strip of A are co-resident, so the
  // strip comes from HBM once and from L2 for the other column tiles
  const int64_t m0 = (int64_t)(blockIdx.x / g.ntn) * kGM;
  const int64_t n0 = (int64_t)(blockIdx.x % g.ntn) * kGN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
  const int64_t kend = kbeg + g.kchunk < g.K ? kbeg + g.kchunk : g.K;
  if (tid == 0) {
    mbar_init(&s.bar[0], 1);
    mbar_init(&s.bar[1], 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, kGN);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;
  const int nchunks = (int)((kend - kbeg + kGK - 1) / kGK);
  for (int c = 0; c < nchunks; ++c) {
    const int st = c & 1;
    const int64_t k0 = kbeg + (int64_t)c * kGK;
    if (c >= 2) {  // the MMAs of chunk c-2 have read this stage
      mbar_wait(&s.bar[st], (uint32_t)(((c - 2) >> 1) & 1));
      fence_after_sync();
    }
    if constexpr (AK) stage_kmajor<kGM>(s.a[st], g.A, g.lda, m0, g.M, k0, kend);
    else stage_mnmajor<kGM>(s.a[st], g.A, g.lda, m0, g.M, k0, kend);
    if constexpr (BK) stage_kmajor<kGN>(s.b[st], g.B, g.ldb, n0, g.N, k0, kend);
    else stage_mnmajor<kGN>(s.b[st], g.B, g.ldb, n0, g.N, k0, kend);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (cta_issuer()) {
      fence_after_sync();
      issue_gemm(tmem, smem_u32(s.a[st]), AK ? kGM : kGK, !AK, smem_u32(s.b[st]), BK ? kGN : kGK, !BK, kGM, kGN,
                 kGK, c > 0);
      mma_commit(&s.bar[st]);
    }
  }
  if (nchunks > 0) {  // the last commit covers every earlier MMA
    mbar_wait(&s.bar[(nchunks - 1) & 1], (uint32_t)(((nchunks - 1) >> 1) & 1));
    fence_after_sync();
    // epilogue: warp (q = warp & 3, half = warp >> 2) -> rows 32q + lane, columns [128 half, 128 half + 128)
    const int q = warp & 3, half = warp >> 2;
    const int64_t m = m0 + q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      float v[32];
      const int col = half * 128 + j * 32;
      tmem_ld32(tmem + lane_base + (uint32_t)col, v);
      if (m < g.M) {
        float* dst = g.C + m * g.ldc + n0 + col;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          if (n0 + col + e < g.N) {
            if constexpr (ATOMIC) red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
            else *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, kGN);
}

// ---- B-resident variant: C[M][N] = A[M][256] * B[N][256]^T, both K-major, store --------------------------------
// (the LSTM gate pre-activations h W_hh^T: M = rows, N = 4H = 1024, K = H = 256).  A CTA keeps ONE 256-column
// block of B in shared memory (128 KB bf16, converted once) and walks 128-row tiles of A: per tile it converts
// 128 x 256 of A in two K halves (each half is multiplied as soon as it is staged), accumulates in one of two
// 256-column TMEM buffers and streams the PREVIOUS tile's accumulator to global memory under the MMAs of the
// current one.  Per tile: 128 KB of A read, 128 KB of C written, 2 056 tensor cycles.
constexpr int kRK = 256;         // K of the resident variant
constexpr int kRThreads = 512;
struct SmemGemmR {
  uint8_t b[kGN * kRK * 2];      // 131072  B block, K-major tile of 256 rows
  uint8_t a[2][kGM * 128 * 2];   //  65536  A tile in two K halves of 128 (K-major tiles of 128 rows, 16 chunks)
  uint64_t bar_h[2], bar_acc[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kRThreads, 1) tc_gemm_bres_kernel(GemmArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemGemmR& s = *reinterpret_cast<SmemGemmR*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = (int)(blockIdx.x % g.ntn);            // column block of this CTA
  const int64_t n0 = (int64_t)nb * kGN;
  const int64_t mt0 = blockIdx.x / g.ntn, mstride = gridDim.x / g.ntn;
  const int64_t mtiles = (g.M + kGM - 1) / kGM;
  if (tid == 0) {
    mbar_init(&s.bar_h[0], 1), mbar_init(&s.bar_h[1], 1);
    mbar_init(&s.bar_acc[0], 1), mbar_init(&s.bar_acc[1], 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, 512);
  // B block: 256 rows x 32 chunks of 8 k
  for (int e0 = tid; e0 < kGN * (kRK / 8); e0 += kGBatch * kRThreads) {
    Chunk8 c[kGBatch];
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kRThreads, row = e >> 5, kc = e & 31;
      c[b] = load8(g.B + (n0 + row) * g.ldb + kc * 8, n0 + row < g.N, 8);
    }
#pragma unroll
    for (int b = 0; b < kGBatch; ++b) {
      const int e = e0 + b * kRThreads, row = e >> 5, kc = e & 31;
      *reinterpret_cast<uint4*>(s.b + chunk_offset<kGN>(row, kc)) = pack8(c[b]);
    }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;

  // K half h of the A tile at rows m0: 128 rows x 16 chunks = 2048 tasks, 4 per thread, all loads in flight
  auto stage_a = [&](int h, int64_t m0) {
    Chunk8 c[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int e = tid + b * kRThreads, row = e >> 4, kc = e & 15;
      c[b] = load8(g.A + (m0 + row) * g.lda + h * 128 + kc * 8, m0 + row < g.M, 8);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int e = tid + b * kRThreads, row = e >> 4, kc = e & 15;
      *reinterpret_cast<uint4*>(s.a[h] + chunk_offset<kGM>(row, kc)) = pack8(c[b]);
    }
  };
  // accumulator buf -> C rows of tile mt: warp (q = warp & 3, part = warp >> 2) -> rows 32q + lane, 64 columns
  auto epilogue = [&](int64_t mt, int buf) {
    const int q = warp & 3, part = warp >> 2;
    const int64_t m = mt * kGM + q * 32 + lane;
    const uint32_t base = tmem + (uint32_t)(buf * kGN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64);
#pragma unroll 1
    for (int j = 0; j < 2; ++j) {
      float v[32];
      tmem_ld32(base + (uint32_t)(j * 32), v);
      if (m < g.M) {
        float* dst = g.C + m * g.ldc + n0 + part * 64 + j * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          if (n0 + part * 64 + j * 32 + e < g.N)
            *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  };

  int it = 0;
  int64_t prev = -1;
  for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
    const int buf = it & 1;
    const uint32_t acc = tmem + (uint32_t)(buf * kGN);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (it > 0) {  // the MMAs of the previous tile have read this K half
        mbar_wait(&s.bar_h[h], (uint32_t)((it - 1) & 1));
        fence_after_sync();
      }
      stage_a(h, mt * kGM);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (cta_issuer()) {
        fence_after_sync();
        // 8 K steps of this half against the resident block: B advances 8 chunk groups of 256 rows per half
        issue_gemm(acc, smem_u32(s.a[h]), kGM, false, smem_u32(s.b) + h * 16 * (kGN * 16), kGN, false, kGM, kGN, 128,
                   h > 0);
        mma_commit(&s.bar_h[h]);
        if (h == 1) mma_commit(&s.bar_acc[buf]);
      }
      if (h == 0 && prev >= 0) {  // the previous tile's accumulator: its MMAs were committed one tile ago
        mbar_wait(&s.bar_acc[buf ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
        fence_after_sync();
        epilogue(prev, buf ^ 1);
        fence_before_sync();  // its TMEM reads precede the next tile's MMAs into that buffer (barrier below)
      }
    }
    prev = mt;
  }
  if (prev >= 0) {
    mbar_wait(&s.bar_acc[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
    fence_after_sync();
    epilogue(prev, (it - 1) & 1);
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// ---- fused LSTM step: gates = h W_hh^T on tcgen05 with the cell in the epilogue -----------------------------
// (nn.LSTM's forward cell, src/rl8/models/_recurrent.py:259-341; restated in oracle/recurrent_oracle.py:lstm_cell)
// Same structure as the B-resident GEMM, but a CTA's 256 accumulator columns are the FOUR gates of 64 hidden
// units (column g * 64 + u <-> W_hh row g * 256 + 64 nb + u), so the epilogue owns everything a unit needs:
//   pre_g = (b_ih + x W_ih^T) + (acc + b_hh);  c' = sig(f) c + sig(i) tanh(g);  h' = sig(o) tanh(c')
// and writes the gate activations (kept for the backward pass), c' and h' -- the 4 KB / row of pre-activations
// never go to HBM and the separate cell kernel disappears.
constexpr int kCellD = 8;  // widest observation
struct SmemLstm {
  uint8_t b[kGN * kRK * 2];      // 131072  W_hh block (4 gates x 64 units), K-major tile of 256 rows
  uint8_t a[2][kGM * 128 * 2];   //  65536  h tile in two K halves
  uint8_t wx[kGN * 32 * 2];      //  16384  [W_ih | biases] of the block as 32 more k (see the kernel), K-major, 256 rows
  uint8_t xa[2][kGM * 32 * 2];   //  16384  the matching A columns [x pieces | 1] of two row tiles
  uint64_t full[2], free_[2], acc_full[2], acc_empty[2], xa_full[2], xa_free[2];
  uint32_t tmem_base;
};
struct LstmCellArgs {
  const uint8_t* hb_in;   // h as bf16 in the T128 layout (below): one bulk copy per K half stages the A operand
  uint8_t* hb_out;        // h' in the same layout for the next step / the weight-gradient GEMM (may be null)
  const float *w_hh, *w_ih, *b_ih, *b_hh, *c_prev;
  float *c_out, *h_out;
  uint8_t* zb;            // gate PRE-activations as bf16 T128 [rows_pad][4H] for the backward pass (null in the rollout)
  RowMap xmap;
  int64_t rows;
  int D;
};

// Warp roles: warps 0..15 run the cell epilogue, warp 16 stages the h tile (one bulk copy per K half from the T128
// image of h: no registers, no LSU), warp 17 issues the MMAs.  The three run decoupled behind mbarriers: a K half is
// refilled as soon as the MMAs that read it have completed, the issuer starts a tile as soon as its operands and an
// accumulator buffer are there, and the epilogue of tile k-1 (whose c_prev / observation loads are issued BEFORE it
// waits for the accumulator) runs under the loads and MMAs of tiles k, k+1 -- the kernel is bound by HBM traffic, so
// what matters is that loads are in flight all the time, not only between two block-wide barriers.
constexpr int kLcEpiWarps = 16, kLcLoadWarps = 1;  // + the issuing warp and two idle ones: 20 warps (register allocation is per 4 warps)
constexpr int kLcThreads = 640;

int64_t t128_bytes(int64_t rows, int cols) { return ((rows + 127) / 128) * 128 * (int64_t)cols * 2; }

// 256-bit global accesses (sm_100: LDG / STG .256): one full 32-byte sector per lane and HALF the LSU wavefronts of two
// 128-bit accesses when every lane of the warp addresses a different row
__device__ __forceinline__ void ldg8(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld8s(const float* p, float* v) {  // eight consecutive floats of shared memory
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// fp32 X[rows][256] (row-major) -> bf16 T128; rows of the last tile past `rows` are zero
__global__ void __launch_bounds__(256) pack_t128_kernel(const float* __restrict__ x, int64_t rows, int64_t rows_pad,
                                                        uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows_pad * 32; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = i >> 12;
    const int c8 = (int)(i >> 7) & 31;
    const int64_t r = tile * 128 + (i & 127);
    float v[8];
    if (r < rows) {
      ldg8(x + r * 256 + c8 * 8, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    }
    uint4 q;
    q.x = pack_bf16x2(v[0], v[1]), q.y = pack_bf16x2(v[2], v[3]), q.z = pack_bf16x2(v[4], v[5]), q.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + i * 16) = q;
  }
}
int launch_pack_t128(const float* x, int64_t rows, uint8_t* out, cudaStream_t st) {
  const int64_t rows_pad = (rows + 127) / 128 * 128;
  pack_t128_kernel<<<grid_for(rows_pad * 32, 256, 8, 1), 256, 0, st>>>(x, rows, rows_pad, out);
  return check_launch("pack_t128");
}

template <int D>  // observation width (compile time: the per-unit input FMAs unroll without predicates)
__global__ void __launch_bounds__(kLcThreads, 1) tc_lstm_cell_kernel(LstmCellArgs g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemLstm& s = *reinterpret_cast<SmemLstm*>(smem_raw);
  constexpr int LH = 256;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = (int)(blockIdx.x & 3);  // unit block: units [64 nb, 64 nb + 64)
  const int64_t mt0 = blockIdx.x >> 2, mstride = gridDim.x >> 2;
  const int64_t mtiles = (g.rows + kGM - 1) / kGM;
  if (tid == 0) {
    for (int h = 0; h < 2; ++h) {
      mbar_init(&s.full[h], 1), mbar_init(&s.free_[h], 1);
      mbar_init(&s.acc_full[h], 1), mbar_init(&s.acc_empty[h], kLcEpiWarps);
      mbar_init(&s.xa_full[h], 1), mbar_init(&s.xa_free[h], 1);
    }
    fence_mbar_init();
  }
  if (warp == kLcEpiWarps + kLcLoadWarps) tmem_alloc(&s.tmem_base, 512);
  // W_hh block: tile row c = gate * 64 + u  <-  W_hh row gate * 256 + 64 nb + u
  for (int e0 = tid; e0 < kGN * (kRK / 8); e0 += 4 * kLcThreads) {
    Chunk8 c[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int e = e0 + b * kLcThreads, row = e >> 5, kc = e & 31;
      const int src = (row >> 6) * LH + nb * 64 + (row & 63);
      c[b] = load8(g.w_hh + (int64_t)src * LH + kc * 8, e < kGN * (kRK / 8), 8);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int e = e0 + b * kLcThreads, row = e >> 5, kc = e & 31;
      if (e < kGN * (kRK / 8)) *reinterpret_cast<uint4*>(s.b + chunk_offset<kGN>(row, kc)) = pack8(c[b]);
    }
  }
  // The input term and both biases ride on the tensor core as 32 more k, fp32-accurate through hi / lo bf16 pieces
  // (x = xh + xl, W = Wh + Wl; the dropped xl Wl is 2^-18 of the product):
  //     A columns  [ xh(8) | xh(8) | xl(8) | 1 1 1 1 0 0 0 0 ]
  //     B columns  [ Wh(8) | Wl(8) | Wh(8) | b_ih hi, lo, b_hh hi, lo, 0 0 0 0 ]
  // so the epilogue reads finished pre-activations and carries no per-unit weights at all.
  for (int row = tid; row < kGN; row += kLcThreads) {
    const int src = (row >> 6) * LH + nb * 64 + (row & 63);
    float wh[8], wl[8], bp[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const float w = d < D ? g.w_ih[(int64_t)src * D + d] : 0.0f;
      wh[d] = __bfloat162float(__float2bfloat16_rn(w));
      wl[d] = w - wh[d];
      bp[d] = 0.0f;
    }
    const float bi = g.b_ih[src], bh = g.b_hh[src];
    bp[0] = __bfloat162float(__float2bfloat16_rn(bi)), bp[1] = bi - bp[0];
    bp[2] = __bfloat162float(__float2bfloat16_rn(bh)), bp[3] = bh - bp[2];
    store_chunk(s.wx, chunk_offset<kGN>(row, 0), wh);
    store_chunk(s.wx, chunk_offset<kGN>(row, 1), wl);
    store_chunk(s.wx, chunk_offset<kGN>(row, 2), wh);
    store_chunk(s.wx, chunk_offset<kGN>(row, 3), bp);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = s.tmem_base;

  if (warp == kLcEpiWarps) {
    // ---- loader: one bulk copy per K half -- the tile's 128 rows x 128 k are 32 KB contiguous in the T128 layout and
    //      already in the operand format, so nothing passes through registers or the LSU ----------------------------
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (it > 0) mbar_wait(&s.free_[h], (uint32_t)((it - 1) & 1));  // the previous tile's MMAs have read it
        if (elect_one()) {
          mbar_expect_tx(&s.full[h], kGM * 128 * 2);
          bulk_g2s(s.a[h], g.hb_in + t128_offset(mt * kGM, h * 16, LH / 8), kGM * 128 * 2, &s.full[h]);
        }
        __syncwarp();
      }
    }
  } else if (warp == kLcEpiWarps + kLcLoadWarps) {
    // ---- issuer ---------------------------------------------------------------------------------------------------
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
      const int buf = it & 1;
      const uint32_t acc = tmem + (uint32_t)(buf * kGN);
      if (it >= 2) mbar_wait(&s.acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));  // epilogue of tile it-2 done
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&s.full[h], (uint32_t)(it & 1));
        fence_after_sync();
        if (elect_one()) {
          issue_gemm(acc, smem_u32(s.a[h]), kGM, false, smem_u32(s.b) + h * 16 * (kGN * 16), kGN, false, kGM, kGN, 128,
                     h > 0);
          mma_commit(&s.free_[h]);
        }
        __syncwarp();
      }
      mbar_wait(&s.xa_full[buf], (uint32_t)((it >> 1) & 1));
      fence_after_sync();
      if (elect_one()) {
        issue_gemm(acc, smem_u32(s.xa[buf]), kGM, false, smem_u32(s.wx), kGN, false, kGM, kGN, 32, true);
        mma_commit(&s.xa_free[buf]);
        mma_commit(&s.acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == kLcEpiWarps + kLcLoadWarps + 1) {
    // ---- observation columns: lane l builds rows l, l + 32, l + 64, l + 96 of the tile's [x pieces | 1] operand -----
    const int64_t ds = g.xmap.dstride();
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
      const int buf = it & 1;
      float x[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = mt * kGM + i * 32 + lane;
        const bool live = r < g.rows;
        int64_t xo = 0;
        if (live) xo = g.xmap.offset(r);
#pragma unroll
        for (int d = 0; d < 8; ++d) x[i][d] = (live && d < D) ? g.xmap.obs[xo + d * ds] : 0.0f;
      }
      if (it >= 2) mbar_wait(&s.xa_free[buf], (uint32_t)(((it >> 1) - 1) & 1));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float xh[8], xl[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          xh[d] = __bfloat162float(__float2bfloat16_rn(x[i][d]));
          xl[d] = x[i][d] - xh[d];
        }
        const float ones[8] = {1.0f, 1.0f, 1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const int row = i * 32 + lane;
        store_chunk(s.xa[buf], chunk_offset<kGM>(row, 0), xh);
        store_chunk(s.xa[buf], chunk_offset<kGM>(row, 1), xh);
        store_chunk(s.xa[buf], chunk_offset<kGM>(row, 2), xl);
        store_chunk(s.xa[buf], chunk_offset<kGM>(row, 3), ones);
      }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) mbar_arrive(&s.xa_full[buf]);
      __syncwarp();
    }
  } else if (warp < kLcEpiWarps) {
    // ---- cell epilogue: warp (q = warp & 3, part = warp >> 2) -> row 32q + lane, units [16 part, 16 part + 16) of this
    //      block, in two groups of 8 units ----------------------------------------------------------------------------
    const int q = warp & 3, part = warp >> 2;
    int it = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstride, ++it) {
      const int buf = it & 1;
      const int64_t r = mt * kGM + q * 32 + lane;
      const bool live = r < g.rows;
      // everything the cell needs from global memory, requested before the wait on the accumulator
      float cp[2][8];
      {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int j0 = nb * 64 + part * 16 + half * 8;
          if (live) {
            ldg8(g.c_prev + r * LH + j0, cp[half]);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) cp[half][u] = 0.0f;
          }
        }
      }
      mbar_wait(&s.acc_full[buf], (uint32_t)((it >> 1) & 1));
      fence_after_sync();
      const uint32_t base = tmem + (uint32_t)(buf * kGN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 16);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int u0 = part * 16 + half * 8;  // first unit of the group inside the block
        const int j0 = nb * 64 + u0;          // ... and in the layer
        float pre[4][8];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) tmem_ld8(base + (uint32_t)(gate * 64 + half * 8), pre[gate]);
        if (half == 1) {  // the accumulator buffer is free as soon as it has been read
          fence_before_sync();
          __syncwarp();
          if (elect_one()) mbar_arrive(&s.acc_empty[buf]);
          __syncwarp();
        }
        if (!live) {  // rows of the last tile past the end: a finite (zero) image for the GEMMs that contract over rows
          if (g.hb_out)
            *reinterpret_cast<uint4*>(g.hb_out + t128_offset(r, j0 >> 3, LH / 8)) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        const float* cph = cp[half];
        // pre[gate][u]: finished pre-activations (h W_hh^T + x W_ih^T + b_ih + b_hh, all accumulated by the tensor core)
        if (g.zb) {  // kept for the backward pass as bf16: lanes = consecutive rows, 512 contiguous bytes per store
#pragma unroll
          for (int gate = 0; gate < 4; ++gate) {
            uint4 qz;
            qz.x = pack_bf16x2(pre[gate][0], pre[gate][1]), qz.y = pack_bf16x2(pre[gate][2], pre[gate][3]);
            qz.z = pack_bf16x2(pre[gate][4], pre[gate][5]), qz.w = pack_bf16x2(pre[gate][6], pre[gate][7]);
            *reinterpret_cast<uint4*>(g.zb + t128_offset(r, gate * 32 + (j0 >> 3), 4 * LH / 8)) = qz;
          }
        }
        float cn[8], hn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float ig = sigmoid_fast(pre[0][u]), fg = sigmoid_fast(pre[1][u]);
          const float gg = tanh_fast(pre[2][u]), og = sigmoid_fast(pre[3][u]);
          cn[u] = fg * cph[u] + ig * gg;
          hn[u] = og * tanh_fast(cn[u]);
        }
        stg8(g.c_out + r * LH + j0, cn);
        stg8(g.h_out + r * LH + j0, hn);
        if (g.hb_out) {  // lanes = consecutive rows: 512 contiguous bytes per warp
          uint4 qh;
          qh.x = pack_bf16x2(hn[0], hn[1]), qh.y = pack_bf16x2(hn[2], hn[3]);
          qh.z = pack_bf16x2(hn[4], hn[5]), qh.w = pack_bf16x2(hn[6], hn[7]);
          *reinterpret_cast<uint4*>(g.hb_out + t128_offset(r, j0 >> 3, LH / 8)) = qh;
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == kLcEpiWarps + kLcLoadWarps) tmem_dealloc(tmem, 512);
}

// h' / c' / gate activations of one LSTM step for `rows` rows (H = 256).  hb_in: h as bf16 T128 (launch_pack_t128 or
// a previous call's hb_out); hb_out (may be null): h' in the same form.
int launch_lstm_cell_tc(const uint8_t* hb_in, const float* w_hh, const float* w_ih, const float* b_ih, const float* b_hh,
                        const float* c_prev, const RowMap& xmap, int D, int64_t rows, uint8_t* zb, float* c_out,
                        float* h_out, uint8_t* hb_out, cudaStream_t st) {
  if (D < 1 || D > kCellD || rows <= 0 || !hb_in) return RL8_ERR_ARG;
  LstmCellArgs g;
  g.hb_in = hb_in, g.hb_out = hb_out, g.w_hh = w_hh, g.w_ih = w_ih, g.b_ih = b_ih, g.b_hh = b_hh, g.c_prev = c_prev;
  g.zb = zb, g.c_out = c_out, g.h_out = h_out, g.xmap = xmap, g.rows = rows, g.D = D;
  const int64_t mtiles = ceil_div(rows, kGM);
  int64_t per_block = kNumSMs / 4;
  if (per_block > mtiles) per_block = mtiles;
  const unsigned grid = (unsigned)(per_block * 4);
#define RL8_LSTM_CELL(DV)                                                                                         \
  case DV:                                                                                                         \
    cudaFuncSetAttribute(tc_lstm_cell_kernel<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLstm)); \
    tc_lstm_cell_kernel<DV><<<grid, kLcThreads, sizeof(SmemLstm), st>>>(g);                                         \
    break;
  switch (D) {
    RL8_LSTM_CELL(1) RL8_LSTM_CELL(2) RL8_LSTM_CELL(3) RL8_LSTM_CELL(4)
    RL8_LSTM_CELL(5) RL8_LSTM_CELL(6) RL8_LSTM_CELL(7) RL8_LSTM_CELL(8)
    default: return RL8_ERR_ARG;
  }
#undef RL8_LSTM_CELL
  return check_launch("tc_lstm_cell");
}

// Same contract as launch_sgemm (mlp_fp32.cuh) for EPI_STORE / EPI_ATOMIC, bf16 operands on tcgen05.
// Requires 16-byte aligned operands, lda / ldb / ldc % 4 == 0 and N % 4 == 0 (the vector epilogue).
int launch_tc_gemm(bool a_kmajor, bool b_kmajor, int epi, const float* A, const float* B, float* C, int64_t M,
                   int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int splits, cudaStream_t st) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) return RL8_ERR_ARG;
  if ((lda | ldb | ldc | N) & 3) return RL8_ERR_UNSUPPORTED;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return RL8_ERR_UNSUPPORTED;
  if (epi != EPI_STORE && epi != EPI_ATOMIC) return RL8_ERR_UNSUPPORTED;
  if (epi == EPI_STORE) splits = 1;
  if (splits < 1) splits = 1;
  GemmArgs g;
  g.A = A, g.B = B, g.C = C, g.M = M, g.N = N, g.K = K, g.lda = lda, g.ldb = ldb, g.ldc = ldc;
  if (a_kmajor && b_kmajor && epi == EPI_STORE && K == kRK && M >= 4 * kGM) {
    // the LSTM gate GEMM: one 256-column block of B resident per CTA, rows streamed
    g.ntn = (unsigned)ceil_div(N, kGN);
    g.kchunk = K;
    const int64_t mtiles = ceil_div(M, kGM);
    int64_t per_block = kNumSMs / g.ntn;  // CTAs per column block
    if (per_block < 1) per_block = 1;
    if (per_block > mtiles) per_block = mtiles;
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(tc_gemm_bres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemGemmR));
      attr = true;
    }
    tc_gemm_bres_kernel<<<(unsigned)(per_block * g.ntn), kRThreads, sizeof(SmemGemmR), st>>>(g);
    return check_launch("tc_gemm_bres");
  }
  g.kchunk = round_up(ceil_div(K, splits), kGK);
  splits = (int)ceil_div(K, g.kchunk);
  g.ntn = (unsigned)ceil_div(N, kGN);
  dim3 grid((unsigned)(ceil_div(M, kGM) * g.ntn), 1, (unsigned)splits);
  const size_t smem = sizeof(SmemGemm);
#define RL8_GEMM(AKV, BKV, ATV)                                                                          \
  {                                                                                                      \
    static bool attr = false;                                                                            \
    if (!attr) {                                                                                         \
      cudaFuncSetAttribute(tc_gemm_kernel<AKV, BKV, ATV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                           (int)smem);                                                                   \
      attr = true;                                                                                       \
    }                                                                                                    \
    tc_gemm_kernel<AKV, BKV, ATV><<<grid, kGThreads, smem, st>>>(g);                                     \
  }
  const bool at = epi == EPI_ATOMIC;
  if (a_kmajor && b_kmajor && !at) RL8_GEMM(true, true, false)
  else if (a_kmajor && b_kmajor && at) RL8_GEMM(true, true, true)
  else if (a_kmajor && !b_kmajor && !at) RL8_GEMM(true, false, false)
  else if (a_kmajor && !b_kmajor && at) RL8_GEMM(true, false, true)
  else if (!a_kmajor && !b_kmajor && !at) RL8_GEMM(false, false, false)
  else if (!a_kmajor && !b_kmajor && at) RL8_GEMM(false, false, true)
  else return RL8_ERR_UNSUPPORTED;
#undef RL8_GEMM
  return check_launch("tc_gemm");
}

}  // namespace rl8

// Test hook (C ABI): the generic tensor-core GEMM with the launch_sgemm contract.
extern "C" int rl8_tc_gemm(int a_kmajor, int b_kmajor, int accumulate, const float* A, const float* B, float* C,
                           int64_t M, int32_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int32_t splits,
                           rl8_stream_t stream) {
  return rl8::launch_tc_gemm(a_kmajor != 0, b_kmajor != 0, accumulate ? rl8::EPI_ATOMIC : rl8::EPI_STORE, A, B, C,
                             M, N, K, lda, ldb, ldc, splits, (cudaStream_t)stream);
}
