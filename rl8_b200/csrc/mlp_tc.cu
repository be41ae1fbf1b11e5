// Tensor-core path (RL8_PREC_BF16): tcgen05.mma with bf16 operands staged in shared memory
// in the chunked format of tc.cuh, fp32 accumulators in TMEM, epilogues on CUDA cores.
//
//   tc_forward_kernel   one network over any number of rows (value pass of collect(),
//                       rl8_mlp_forward): persistent CTAs, 128-row tiles, W2 resident in smem
//                       (one 128 KB bulk-async copy per CTA), TMEM double-buffered so the
//                       epilogue of tile i overlaps the MMAs of tile i+1.
//   tc_rollout_kernel   the whole T-step rollout of TWO 128-env tiles inside one persistent CTA,
//                       half a step apart: layer 1 (one tf32 MMA) -> 256x256 layer on tcgen05 ->
//                       head dot products -> sampling, log-prob, env transition (state in
//                       registers) and the horizon-major buffer writes, the last stage of one
//                       tile running under the MMAs of the other.  No grid-wide sync: envs are
//                       independent.
//   tc_selftest_kernel  one 128xNxK GEMM with either operand major, used by the parity tests to
//                       pin the descriptor encodings.
//
// The first layer is one kind::tf32 instruction ([obs, 1] * [W1, b1]^T, K = 8); the heads (N = P <= 4)
// are not GEMM-shaped and stay on CUDA cores in fp32.  Only the 256x256 contraction runs in bf16.
#include "mlp_tc.cuh"

namespace rl8 {

using namespace tc;

// ---- forward kernel ---------------------------------------------------------------------------------
// Per 128-row tile k (accumulator buf = k & 1):
//   Z1 = [obs, 1] * [W1, b1]^T   one kind::tf32 instruction (K = 8) into acc[buf] -- the same layer-1
//                                arithmetic as the update kernels, so stored values and their later
//                                recomputation agree bit for bit
//   H1 = relu(Z1) -> bf16 tile;  Z2 = H1 * W2^T into acc[buf];  head epilogue of tile k-1 (acc[buf ^ 1])
//   runs under Z2(k);  then Z1(k+1) is issued into acc[buf ^ 1] behind Z2(k) on the in-order tensor pipe.
struct SmemF {
  uint8_t w2[kW2Bytes];           // 131072
  uint8_t a_tile[kTileBytes];     //  65536  H1 tile: K-major A operand
  uint8_t w1aug[H * 32];          //   8192  tf32 [W1 | b1]: off(i, d) = i*16 + (d/4)*4096 + (d%4)*4
  uint8_t aug32[TILE * 32];       //   4096  tf32 [obs, 1]:  off(r, d) = r*16 + (d/4)*2048 + (d%4)*4
  float b2[H];                    //   1024
  float w3[kMaxPT][H];            //   4096
  float part[1][4][TILE][kMaxPT]; //   8192  head partial sums per column quarter
  uint64_t bar_w, bar_z, bar_mma[2];
  uint32_t tmem_base;
};
static_assert(sizeof(SmemF) <= 227 * 1024, "smem plan exceeds the 227 KB CTA limit");

template <int P>
__global__ void __launch_bounds__(kFwdThreads, 1)
tc_forward_kernel(NetParams np, RowMap map, int64_t rows, float* __restrict__ out, int tanh_col1) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemF& s = *reinterpret_cast<SmemF*>(smem_raw);
  const int tid = threadIdx.x;
  cta_setup(s, np, 512);

  const uint32_t tmem = s.tmem_base;
  const int64_t ntiles = (rows + TILE - 1) / TILE;
  const int64_t ds = map.dstride();
  const int D = np.D;
  const float b3[kMaxPT] = {np.b3[0], P > 1 ? np.b3[1] : 0.f, P > 2 ? np.b3[2] : 0.f, P > 3 ? np.b3[3] : 0.f};
  // this thread stages slots d0 and d0 + 4 of [obs, 1] of row rr
  const int rr = tid & (TILE - 1), d0 = tid >> 7;
  auto load_obs = [&](int64_t tile, float& v0, float& v1) {
    v0 = d0 == D ? 1.0f : 0.0f;
    v1 = d0 + 4 == D ? 1.0f : 0.0f;
    const int64_t row = tile * TILE + rr;
    if (row < rows) {
      const float* base = map.obs + map.offset(row);
      if (d0 < D) v0 = __ldg(base + (int64_t)d0 * ds);
      if (d0 + 4 < D) v1 = __ldg(base + (int64_t)(d0 + 4) * ds);
    }
  };
  auto store_aug = [&](float v0, float v1) {
    *reinterpret_cast<float*>(s.aug32 + rr * 16 + d0 * 4) = tf32_round(v0);
    *reinterpret_cast<float*>(s.aug32 + rr * 16 + TILE * 16 + d0 * 4) = tf32_round(v1);
  };
  auto issue_z1 = [&](int buf) { issue_layer1(s, tmem + (uint32_t)(buf * H)); };  // elected lane
  auto epilogue = [&](int64_t tile, int buf) {
    head_partials<P>(s, tmem + (uint32_t)(buf * H));
    fence_before_sync();
    __syncthreads();
    if (tid < TILE) {
      const int64_t row = tile * TILE + tid;
      if (row < rows) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
          float v = head_sum(s, tid, p) + b3[p];
          if (tanh_col1 && p == 1) v = tanhf(v);
          out[row * P + p] = v;
        }
      }
    }
  };

  int64_t tile = blockIdx.x;
  if (tile < ntiles) {
    float v0, v1;
    load_obs(tile, v0, v1);
    store_aug(v0, v1);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (cta_issuer()) {
      fence_after_sync();
      issue_z1(0);
    }
  }
  int it = 0;
  int64_t prev_tile = -1;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const bool has_next = tile + gridDim.x < ntiles;
    float n0 = 0.0f, n1 = 0.0f;
    if (has_next) load_obs(tile + gridDim.x, n0, n1);  // in flight until the staging below
    mbar_wait(&s.bar_z, (uint32_t)(it & 1));  // Z1 of this tile is in acc[buf]
    fence_after_sync();
    h1_epilogue(s, tmem + (uint32_t)(buf * H));
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (cta_issuer()) {
      fence_after_sync();
      issue_gemm(tmem + (uint32_t)(buf * H), smem_u32(s.a_tile), TILE, false, smem_u32(s.w2), H, false, TILE, H, H,
                 false);
      mma_commit(&s.bar_mma[buf]);
    }
    if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1);  // under the MMAs just issued
    if (has_next) store_aug(n0, n1);                   // Z1 of this tile has read aug32
    fence_async_smem();
    fence_before_sync();
    __syncthreads();  // acc[buf ^ 1] has been read by the head epilogue: the next Z1 may overwrite it
    if (has_next && cta_issuer()) {
      fence_after_sync();
      issue_z1(buf ^ 1);
    }
    mbar_wait(&s.bar_mma[buf], (uint32_t)((it >> 1) & 1));  // Z2 done: the H1 tile may be rewritten
    fence_after_sync();
    prev_tile = tile;
  }
  if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1);
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// ---- rollout kernel -----------------------------------------------------------------------------------
struct RolloutArgs {
  rl8_env_cfg cfg;
  int dist_kind, deterministic, T;
  int64_t N;
  float gamma;
  float* state;        // [S][N]
  float* obs;          // [T+1][D][N]
  void* actions;       // [T+1][N]
  float* logp;         // [T+1][N]
  float* rewards;      // [T+1][N]
  float* rdr;          // [T+1][N] or null
  const float* noise;  // [T][N][P] | [T][N]
};

// Two env tiles per CTA, half a step apart: while the tensor core works on one tile's 256x256 layer, the
// 128 owner threads of the OTHER tile sample its actions and step its environments, so the serial chain
// layer 1 -> MMA -> head -> sample / env-step of a tile is hidden behind its partner's.
//   threads   0..127 own the rows of tile A (accumulator columns   0..255, s.part[0])
//   threads 128..255 own the rows of tile B (accumulator columns 256..511, s.part[1])
// One aug32 / a_tile serves both: every producer runs strictly between the consumers (see the cycle below).
template <int KIND, int P>
__global__ void __launch_bounds__(kFwdThreads, 1) tc_rollout_kernel(NetParams np, RolloutArgs a) {
  using Tr = EnvTraits<KIND>;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  cta_setup(s, np, 512);
  const uint32_t tmem = s.tmem_base;
  const int tid = threadIdx.x;
  const int slot = tid >> 7, row = tid & (TILE - 1);  // slot 0 / 1: owner of tile A / B; 2, 3: helpers
  const int64_t N = a.N;
  const int64_t ntiles = (N + TILE - 1) / TILE;
  const int64_t npairs = (ntiles + 1) / 2;
  float b3[P];
#pragma unroll
  for (int p = 0; p < P; ++p) b3[p] = np.b3[p];
  uint32_t phase[2] = {0u, 0u}, zphase = 0u;
  constexpr int kNz = Tr::discrete ? P : 1;  // noise values per env and step

  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const bool has_b = 2 * pair + 1 < ntiles;
    const int64_t n = (2 * pair + slot) * TILE + row;  // env of this thread when it is an owner
    const bool owner = slot < 2 && n < N;
    float st[Tr::S], rdr_prev = 0.0f, nz[kNz];
    if (owner) {
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) st[i] = a.state[(int64_t)i * N + n];
      if (a.rdr) rdr_prev = a.rdr[n];
    }
    auto load_noise = [&](int t) {
      if (owner && !a.deterministic && t < a.T) {
        const float* src = a.noise + ((int64_t)t * N + n) * kNz;
#pragma unroll
        for (int k = 0; k < kNz; ++k) nz[k] = src[k];
      }
    };
    // [obs, 1] of this slot's row in tf32 -> aug32 (the A operand of the layer-1 MMA); visible to the
    // tensor core after the fence + the CTA barrier that always follows
    auto put_obs = [&](const float* ob) {
      float v[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) v[d] = d < Tr::D ? tf32_round(ob[d]) : (d == Tr::D ? 1.0f : 0.0f);
      *reinterpret_cast<float4*>(s.aug32 + row * 16) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(s.aug32 + row * 16 + TILE * 16) = make_float4(v[4], v[5], v[6], v[7]);
      fence_async_smem();
    };
    auto stage_obs = [&](int t) {  // this slot's observations of step t: buffer -> aug32
      float ob[Tr::D];
#pragma unroll
      for (int d = 0; d < Tr::D; ++d) ob[d] = n < N ? a.obs[((int64_t)t * Tr::D + d) * N + n] : 0.0f;
      put_obs(ob);
    };
    // sample the action of step t from this slot's head outputs, step the env, write slab t (+ obs, rdr of t+1)
    auto env_phase = [&](int t) {
      if (owner) {
        float o[P];
#pragma unroll
        for (int p = 0; p < P; ++p) o[p] = head_sum(s, row, p, slot) + b3[p];
        float act, lp;
        if constexpr (Tr::discrete) {
          float norm[P], probs[P];
          categorical_norm<P>(o, norm, probs);
          const int ai = a.deterministic ? categorical_mode<P>(probs) : categorical_sample<P>(probs, nz);
          lp = norm[0];
#pragma unroll
          for (int k = 1; k < P; ++k) lp = (ai == k) ? norm[k] : lp;
          act = (float)ai;
          ((long long*)a.actions)[(int64_t)t * N + n] = ai;
        } else {
          const float mean = o[0], scale = expf(tanhf(o[1]));
          float x = a.deterministic ? mean : add(mul(nz[0], scale), mean);
          if (a.dist_kind == RL8_DIST_SQUASHED_NORMAL) {
            x = tanhf(x);
            lp = squashed_logp(mean, scale, x, nullptr);
          } else {
            lp = normal_logp(mean, scale, x);
          }
          act = x;
          ((float*)a.actions)[(int64_t)t * N + n] = x;
        }
        a.logp[(int64_t)t * N + n] = lp;
        float ob[Tr::D], r;
        env_step<KIND>(a.cfg, st, act, ob, r);
        a.rewards[(int64_t)t * N + n] = r;
        if (a.rdr) {
          rdr_prev = add(mul(a.gamma, rdr_prev), r);
          a.rdr[(int64_t)(t + 1) * N + n] = rdr_prev;
        }
#pragma unroll
        for (int d = 0; d < Tr::D; ++d) a.obs[((int64_t)(t + 1) * Tr::D + d) * N + n] = ob[d];
        put_obs(ob);
      } else if (slot < 2) {
        float ob[Tr::D];
#pragma unroll
        for (int d = 0; d < Tr::D; ++d) ob[d] = 0.0f;
        put_obs(ob);
      }
    };
    // layers 1 and 2 of slot sl from aug32: Z1 (tf32 MMA) -> H1 tile -> Z2 issued into the slot's accumulator
    auto layers = [&](int sl) {
      const uint32_t acc = tmem + (uint32_t)(sl * H);
      if (cta_issuer()) {
        fence_after_sync();
        issue_layer1(s, acc);
      }
      mbar_wait(&s.bar_z, zphase);
      zphase ^= 1u;
      fence_after_sync();
      h1_epilogue(s, acc);
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (cta_issuer()) {
        fence_after_sync();
        issue_gemm(acc, smem_u32(s.a_tile), TILE, false, smem_u32(s.w2), H, false, TILE, H, H, false);
        mma_commit(&s.bar_mma[sl]);
      }
    };
    auto wait_head = [&](int sl) {  // accumulator of slot sl -> head partial sums
      mbar_wait(&s.bar_mma[sl], phase[sl]);
      phase[sl] ^= 1u;
      fence_after_sync();
      head_partials<P>(s, tmem + (uint32_t)(sl * H), sl);
    };

    if (slot == 0) {
      stage_obs(0);
      load_noise(0);
    }
    __syncthreads();
    // cycle of step t:  a. layers(A)  b. env(B, t-1) under MMA(A)  c. head(A)  d. layers(B)  e. env(A, t) under MMA(B)  f. head(B)
    for (int t = 0; t < a.T; ++t) {
      layers(0);  // a. aug32 holds A's observations of step t
      if (slot == 1) {  // b. B's observations of step t: from its previous env step, or the buffer at t = 0
        if (t == 0) stage_obs(0);
        else env_phase(t - 1);
        load_noise(t);
      }
      wait_head(0);  // c.
      fence_before_sync();
      __syncthreads();
      if (has_b) layers(1);  // d. aug32 holds B's observations of step t
      if (slot == 0) {  // e.
        env_phase(t);
        load_noise(t + 1);
      }
      if (has_b) wait_head(1);  // f.
      fence_before_sync();
      __syncthreads();
    }
    if (slot == 1 && has_b) env_phase(a.T - 1);  // B's last step
    if (owner) {
#pragma unroll
      for (int i = 0; i < Tr::S; ++i) a.state[(int64_t)i * N + n] = st[i];
    }
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// ---- descriptor self-test --------------------------------------------------------------------------------
// D[128][N] = A[128][K] * B[N][K]^T with A, B given row-major in fp32; each operand is staged
// either K-major (tile rows = M/N index) or MN-major (tile rows = K index).
__global__ void __launch_bounds__(256, 1)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                   int N, int K, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* a_tile = smem_raw;                  // up to 64 KB
  uint8_t* b_tile = smem_raw + 65536;          // up to 128 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 65536 + 131072);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(tmem_ptr, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_ptr;
  const int a_rows = a_mn ? K : 128, b_rows = b_mn ? K : N;
  for (int i = tid; i < 128 * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    const int row = a_mn ? k : m, col = a_mn ? m : k;
    const uint32_t off = row * 16 + (col / 8) * (a_rows * 16) + (col % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(a_tile + off) = __float2bfloat16(A[i]);
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    const int row = b_mn ? k : n, col = b_mn ? n : k;
    const uint32_t off = row * 16 + (col / 8) * (b_rows * 16) + (col % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(b_tile + off) = __float2bfloat16(B[i]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    issue_gemm(tmem, smem_u32(a_tile), a_rows, a_mn != 0, smem_u32(b_tile), b_rows, b_mn != 0, 128, N,
               K, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  if (tid < 128) {
    const int warp = tid >> 5, lane = tid & 31;
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int j = 0; j < 8; ++j) D[r * N + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 256);
}

// D[128][N] = A[128][8] * B[N][8]^T through kind::tf32 (one instruction, both operands K-major,
// fp32 containers rounded onto the tf32 grid) -- the layer-1 contraction of the update kernels.
__global__ void __launch_bounds__(128, 1)
tc_selftest_tf32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                        int N) {
  __shared__ __align__(128) uint8_t a_tile[128 * 32];
  __shared__ __align__(128) uint8_t b_tile[256 * 32];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&tmem_ptr, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  for (int i = tid; i < 128 * 8; i += blockDim.x) {
    const int m = i >> 3, k = i & 7;
    *reinterpret_cast<float*>(a_tile + m * 16 + (k >> 2) * (128 * 16) + (k & 3) * 4) = tf32_round(A[i]);
  }
  for (int i = tid; i < N * 8; i += blockDim.x) {
    const int n = i >> 3, k = i & 7;
    *reinterpret_cast<float*>(b_tile + n * 16 + (k >> 2) * (N * 16) + (k & 3) * 4) = tf32_round(B[i]);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    fence_after_sync();
    mma_tf32(tmem, smem_desc(smem_u32(a_tile), 128 * 16, 128), smem_desc(smem_u32(b_tile), N * 16, 128),
             instr_desc_tf32(128, N), 0u);
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      for (int j = 0; j < 16; ++j) D[r * N + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 256);
}

// TMEM round trip of raw 32-bit words next to a live accumulator: out[r][c] = in[r][c] for the
// 32x32b.x32 store / load pair the update kernel parks its packed H1 copy with.
__global__ void __launch_bounds__(128, 1)
tc_selftest_tmem_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc(&tmem_ptr, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  const int warp = tid >> 5;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = in[tid * 128 + c0 + j];
    tmem_st32_raw(tmem + 256 + lane_base + (uint32_t)c0, v);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32_raw(tmem + 256 + lane_base + (uint32_t)c0, v);
    for (int j = 0; j < 32; ++j) out[tid * 128 + c0 + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// Register layout of the 16x256b TMEM load (the shape whose per-thread rows / columns match the warp-level MMA
// fragments): in[128][32] is stored with 32x32b (thread = row), then every warp reads its 32 lanes x 32 columns as
// two 16x256b.x4 loads (lanes +0..15 and +16..31); out[tid][0..15] / out[tid][16..31] are the registers of the two
// loads in order.  tests/test_gpu_tc.py pins the mapping x3_update_f_kernel's gW3 pass relies on.
__global__ void __launch_bounds__(128, 1)
tc_selftest_tmem_16x256b_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc(&tmem_ptr, 32);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  const int warp = tid >> 5;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = in[tid * 32 + j];
    tmem_st32_raw(tmem + lane_base, v);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[16];
    tmem_ld_16x256b_x4(tmem + lane_base + ((uint32_t)(16 * h) << 16), v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) out[tid * 32 + 16 * h + j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 32);
}

// Microbenchmark: cycles for `iters` rounds of tensor-memory reads by `blockDim.x / 32` warps,
// each warp reading 32 lanes x (32 * width) columns per round.  mode 0: one 32x32b.x32 load +
// wait per round; mode 1: two loads in flight per wait; mode 2: x16 loads.
__global__ void __launch_bounds__(512, 1) tc_bench_tmem_kernel(long long* out, int iters, int mode) {
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 32) tmem_alloc(&tmem_ptr, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_ptr;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  float acc = 0.0f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      float v[32];
      tmem_ld32(base + (uint32_t)((i & 1) * 32), v);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v[j];
    } else if (mode == 1) {
      float v0[32], v1[32];
      tmem_ld32_nowait(base, v0);
      tmem_ld32_nowait(base + 32, v1);
      tmem_wait_ld();
      reg_fence32(v0);
      reg_fence32(v1);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v0[j] + v1[j];
    } else {
      float v[16];
      tmem_ld16(base + (uint32_t)((i & 3) * 16), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += v[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  if (acc == 123.456f) out[1] = 1;
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// Microbenchmark: cycles for `reps` back-to-back GEMMs D[128][N] (+)= A[128][k_total] * B[N][k_total]^T issued by
// one thread with ONE commit at the end, operands in this library's chunked shared-memory format (zeros).
// out[0] = issue -> mbarrier-wait round trip, out[1] = cycles the issuing loop itself took.
struct SmemMmaBench {
  uint8_t b[kW2Bytes];
  uint8_t a[kTileBytes];
  uint64_t bar;
  uint32_t tmem_base;
};
__global__ void __launch_bounds__(128, 1)
tc_bench_mma_kernel(long long* out, int N, int k_total, int reps, int a_mn, int b_mn) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemMmaBench& s = *reinterpret_cast<SmemMmaBench*>(smem_raw);
  const int tid = threadIdx.x;
  for (int i = tid; i < (kW2Bytes + kTileBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(s.b)[i] = 0u;
  if (tid == 0) {
    mbar_init(&s.bar, 1);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(&s.tmem_base, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid < 32 && elect_one()) {
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (k_total == 8)  // kind::tf32, K = 8: the layer-1 instruction of the tensor-core kernels
        mma_tf32(s.tmem_base, smem_desc(smem_u32(s.a), TILE * 16, 128), smem_desc(smem_u32(s.b), N * 16, 128),
                 instr_desc_tf32(TILE, N), r > 0 ? 1u : 0u);
      else
        issue_gemm(s.tmem_base, smem_u32(s.a), TILE, a_mn != 0, smem_u32(s.b), b_mn ? k_total : N, b_mn != 0, TILE,
                   N, k_total, r > 0);
    }
    mma_commit(&s.bar);
    const long long t1 = clock64();
    mbar_wait(&s.bar, 0);
    const long long t2 = clock64();
    out[0] = t2 - t0;
    out[1] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(s.tmem_base, 512);
}

// ---- host side ------------------------------------------------------------------------------------------
static int launch_forward(const NetParams& np, const RowMap& map, int64_t rows, float* out,
                          int tanh_col1, cudaStream_t st) {
  const int64_t ntiles = ceil_div(rows, TILE);
  const int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  int rc;
#define RL8_FWD(PV)                                                                       \
  case PV:                                                                                \
    if ((rc = set_smem((const void*)tc_forward_kernel<PV>, sizeof(SmemF)))) return rc;     \
    tc_forward_kernel<PV><<<grid, kFwdThreads, sizeof(SmemF), st>>>(np, map, rows, out, tanh_col1); \
    break;
  switch (np.P) {
    RL8_FWD(1) RL8_FWD(2) RL8_FWD(3) RL8_FWD(4)
    default: return RL8_ERR_UNSUPPORTED;
  }
#undef RL8_FWD
  return check_launch("tc_forward");
}

int mlp_forward_tc(const rl8_model* m, int which, const RowMap& map, int64_t rows, float* out,
                   int tanh_col1, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  if (m->H != H || m->D > 7 || m->P > kMaxPT) return RL8_ERR_UNSUPPORTED;  // D + 1 <= 8 = K of the tf32 MMA
  if (!workspace || workspace_bytes < kW2Bytes) return RL8_ERR_WORKSPACE;
  uint8_t* img = (uint8_t*)workspace;
  int rc = launch_pack_w2(which ? m->vf_w2 : m->pi_w2, img, st);
  if (rc) return rc;
  return launch_forward(net_params(m, which, img), map, rows, out, tanh_col1, st);
}

int64_t collect_tc_workspace(const rl8_model*, int64_t, int32_t) { return 2 * kW2Bytes; }

template <int KIND, int P>
static int launch_rollout(const NetParams& np, const rl8_rollout* ro, cudaStream_t st) {
  RolloutArgs a;
  a.cfg = ro->env_cfg;
  a.dist_kind = ro->dist_kind, a.deterministic = ro->deterministic, a.T = ro->T, a.N = ro->N;
  a.gamma = ro->gamma;
  a.state = ro->env_state, a.obs = ro->obs, a.actions = ro->actions, a.logp = ro->logp;
  a.rewards = ro->rewards, a.rdr = ro->rdr, a.noise = ro->noise;
  int rc = set_smem((const void*)tc_rollout_kernel<KIND, P>, sizeof(Smem));
  if (rc) return rc;
  const int64_t npairs = ceil_div(ceil_div(ro->N, TILE), 2);  // a CTA rolls two env tiles out at a time
  const int grid = (int)(npairs < kNumSMs ? npairs : kNumSMs);
  tc_rollout_kernel<KIND, P><<<grid, kFwdThreads, sizeof(Smem), st>>>(np, a);
  return check_launch("tc_rollout");
}

int collect_tc(const rl8_model* model, const rl8_rollout* ro, void* workspace,
               int64_t workspace_bytes, cudaStream_t st) {
  if (model->H != H || model->P > kMaxPT) return RL8_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < 2 * kW2Bytes) return RL8_ERR_WORKSPACE;
  uint8_t* img_pi = (uint8_t*)workspace;
  uint8_t* img_vf = img_pi + kW2Bytes;
  int rc;
  if ((rc = launch_pack_w2(model->pi_w2, img_pi, st))) return rc;
  if ((rc = launch_pack_w2(model->vf_w2, img_vf, st))) return rc;
  const NetParams np = net_params(model, 0, img_pi);
  switch (ro->env_kind) {
    case RL8_ENV_DISCRETE_DUMMY: rc = launch_rollout<RL8_ENV_DISCRETE_DUMMY, 2>(np, ro, st); break;
    case RL8_ENV_CONTINUOUS_DUMMY: rc = launch_rollout<RL8_ENV_CONTINUOUS_DUMMY, 2>(np, ro, st); break;
    case RL8_ENV_CARTPOLE: rc = launch_rollout<RL8_ENV_CARTPOLE, 3>(np, ro, st); break;
    case RL8_ENV_MOUNTAIN_CAR: rc = launch_rollout<RL8_ENV_MOUNTAIN_CAR, 3>(np, ro, st); break;
    case RL8_ENV_PENDULUM: rc = launch_rollout<RL8_ENV_PENDULUM, 2>(np, ro, st); break;
    default: return RL8_ERR_ARG;
  }
  if (rc) return rc;
  // values of all T+1 observation slabs in one launch (rows r = t*N + n)
  RowMap map{};
  map.obs = ro->obs, map.mode = 2, map.D = model->D, map.N = ro->N, map.T = ro->T;
  return launch_forward(net_params(model, 1, img_vf), map, (int64_t)(ro->T + 1) * ro->N, ro->values,
                        0, st);
}

}  // namespace rl8

using namespace rl8;

// Test hook: one tcgen05 GEMM through the operand format of tc.cuh (see tests/test_gpu_tc.py).
extern "C" int rl8_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K,
                               int a_mn_major, int b_mn_major, rl8_stream_t stream) {
  if (!A || !B || !D || N < 8 || N > 256 || (N % 8) || K < 16 || K > 256 || (K % 16)) return RL8_ERR_ARG;
  const size_t bytes = 65536 + 131072 + 64;
  int rc = set_smem((const void*)tc_selftest_kernel, bytes);
  if (rc) return rc;
  tc_selftest_kernel<<<1, 256, bytes, (cudaStream_t)stream>>>(A, B, D, N, K, a_mn_major, b_mn_major);
  return check_launch("tc_selftest");
}

// Test hooks for the tf32 layer-1 instruction and the raw TMEM store / load pair.
extern "C" int rl8_tc_selftest_tf32(const float* A, const float* B, float* D, int32_t N,
                                    rl8_stream_t stream) {
  if (!A || !B || !D || N < 16 || N > 256 || (N % 16)) return RL8_ERR_ARG;
  tc_selftest_tf32_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(A, B, D, N);
  return check_launch("tc_selftest_tf32");
}
extern "C" int rl8_tc_selftest_tmem_16x256b(const uint32_t* in, uint32_t* out, rl8_stream_t stream) {
  if (!in || !out) return RL8_ERR_ARG;
  tc_selftest_tmem_16x256b_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(in, out);
  return check_launch("tc_selftest_tmem_16x256b");
}
extern "C" int rl8_tc_selftest_tmem(const uint32_t* in, uint32_t* out, rl8_stream_t stream) {
  if (!in || !out) return RL8_ERR_ARG;
  tc_selftest_tmem_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(in, out);
  return check_launch("tc_selftest_tmem");
}

extern "C" int rl8_tc_bench_mma(long long* out_cycles, int32_t N, int32_t k_total, int32_t reps, int a_mn_major,
                                int b_mn_major, rl8_stream_t stream) {
  if (!out_cycles || N < 16 || N > 256 || N % 16 || reps < 1) return RL8_ERR_ARG;
  if (k_total != 8 && (k_total < 16 || k_total > 256 || k_total % 16)) return RL8_ERR_ARG;  // 8: one tf32 instruction
  if (a_mn_major && k_total > TILE) return RL8_ERR_ARG;
  int rc;
  if ((rc = set_smem((const void*)tc_bench_mma_kernel, sizeof(SmemMmaBench)))) return rc;
  tc_bench_mma_kernel<<<1, 128, sizeof(SmemMmaBench), (cudaStream_t)stream>>>(out_cycles, N, k_total, reps,
                                                                           a_mn_major, b_mn_major);
  return check_launch("tc_bench_mma");
}

// Microbenchmark hook: SM cycles for `iters` rounds of TMEM reads with `nwarps` warps (see kernel).
extern "C" int rl8_tc_bench_tmem(long long* out_cycles, int32_t nwarps, int32_t iters, int32_t mode,
                                 rl8_stream_t stream) {
  if (!out_cycles || nwarps < 1 || nwarps > 16 || iters < 1) return RL8_ERR_ARG;
  tc_bench_tmem_kernel<<<1, 32 * nwarps, 0, (cudaStream_t)stream>>>(out_cycles, iters, mode);
  return check_launch("tc_bench_tmem");
}
