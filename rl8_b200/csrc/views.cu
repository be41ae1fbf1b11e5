// View requirements: sliding windows over the time axis of a rollout field (src/rl8/views.py).
//
// Every view of the reference -- RollingWindow.apply_all (:158-193), PaddedRollingWindow.apply_all
// (:245-281), pad_last_sequence (:57-88), pad_whole_sequence (:91-118) -- is one primitive:
//
//     out[b][w][s][f] = x[b][t_first + w + s][f]   (0 and mask = 1 where that time index is < 0)
//
// for w < count, s < size.  The reference materialises it with unfold + permute + reshape
// (a strided gather, size x amplification of the item).  Here a CTA stages the time steps a block
// of windows needs for a few sequences in shared memory ONCE -- read coalesced along whichever of
// the sequence / feature axes is unit-stride, so the horizon-major buffer ([T+1][F][N], b fastest)
// and the reference's env-major tensors ([N][T+1][F], f fastest) are both read at full width -- and
// then streams every sequence's windows out as one contiguous run.  HBM traffic is the algorithmic
// minimum: the item read once (+ a halo of size-1 steps per window block), the windows written once.
#include "common.cuh"

namespace rl8 {

constexpr int kViewThreads = 256;
constexpr int kViewSmemBytes = 64 * 1024;  // staging tile budget: 3 CTAs / SM

struct ViewArgs {
  const void* x;
  void* out;
  uint8_t* mask;
  int64_t B, T, F;
  int64_t sb, st, sf;  // element strides of x
  int64_t t_first, count;
  int size;
  int tb;     // sequences per CTA
  int tw;     // windows per CTA
  int pitch;  // staged elements per sequence (odd, >= (tw + size - 1) * F)
  int tb_shift;       // log2(tb), or -1 when tb is not a power of two
  uint32_t f_magic;   // floor(2^32 / F) + 1: n / F == umulhi(n, f_magic) for n * F < 2^32 (n < 2^14 here); 0 when F == 1
  int ld_vec;    // 128-bit staging loads: unit sequence stride, 4-byte elements, 16-byte aligned rows
  int mask_vec;  // every (sequence, window block) run of the mask is a whole number of aligned 32-bit words
  int vec;    // 128-bit store path: out 16-byte aligned and size * F a multiple of the vector width
};

template <typename U>
__global__ void __launch_bounds__(kViewThreads) view_windows_kernel(ViewArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  U* tile = reinterpret_cast<U*>(smem_raw);
  const U* __restrict__ x = static_cast<const U*>(a.x);
  U* __restrict__ out = static_cast<U*>(a.out);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * a.tb;
  const int64_t w0 = (int64_t)blockIdx.y * a.tw;
  const int nb = (int)min((int64_t)a.tb, a.B - b0);
  const int nw = (int)min((int64_t)a.tw, a.count - w0);
  const int F = (int)a.F;
  const int rows = nw + a.size - 1;         // time steps staged
  const int64_t t0 = a.t_first + w0;        // time index of staged row 0 (may be negative)
  const int per_seq = rows * F;

  // ---- stage x[b0 .. b0+nb)[t0 .. t0+rows)[:] -> tile[bl * pitch + tl * F + f] -----------------------
  // (kLd independent loads per thread in flight: the staging loop is latency-, not issue-bound)
  constexpr int kLd = 8;
  if (a.ld_vec) {
    // horizon-major buffer, 4-byte elements, everything 16-byte aligned: each lane reads four consecutive
    // sequences with one 128-bit load (a quarter of the index arithmetic per element)
    if constexpr (sizeof(U) == 4) {
      const int tb4 = a.tb >> 2;
      const int total = per_seq * tb4;
      for (int e0 = tid; e0 < total; e0 += kLd * kViewThreads) {
        uint4 v[kLd];
        int dst[kLd];
#pragma unroll
        for (int k = 0; k < kLd; ++k) {
          const int e = e0 + k * kViewThreads;
          v[k] = make_uint4(0u, 0u, 0u, 0u), dst[k] = -1;
          if (e < total) {
            const int tf = a.tb_shift >= 2 ? e >> (a.tb_shift - 2) : e / tb4, bl = (e - tf * tb4) << 2;
            const int tl = a.f_magic ? (int)__umulhi((uint32_t)tf, a.f_magic) : tf, f = tf - tl * F;
            const int64_t t = t0 + tl;
            dst[k] = bl * a.pitch + tf;
            if (bl < nb && t >= 0)
              v[k] = __ldg(reinterpret_cast<const uint4*>(x + (b0 + bl) + t * a.st + (int64_t)f * a.sf));
          }
        }
#pragma unroll
        for (int k = 0; k < kLd; ++k)
          if (dst[k] >= 0) {
            tile[dst[k]] = v[k].x, tile[dst[k] + a.pitch] = v[k].y;
            tile[dst[k] + 2 * a.pitch] = v[k].z, tile[dst[k] + 3 * a.pitch] = v[k].w;
          }
      }
    }
  } else if (a.sb <= a.sf) {
    // sequence axis is the fast one (horizon-major buffer): lanes walk b
    const int total = per_seq * a.tb;
    for (int e0 = tid; e0 < total; e0 += kLd * kViewThreads) {
      U v[kLd];
      int dst[kLd];
#pragma unroll
      for (int k = 0; k < kLd; ++k) {
        const int e = e0 + k * kViewThreads;
        v[k] = 0, dst[k] = -1;
        if (e < total) {
          int bl, tf;
          if (a.tb_shift >= 0) {
            bl = e & (a.tb - 1), tf = e >> a.tb_shift;
          } else {
            bl = e % a.tb, tf = e / a.tb;
          }
          const int tl = a.f_magic ? (int)__umulhi((uint32_t)tf, a.f_magic) : tf, f = tf - tl * F;
          const int64_t t = t0 + tl;
          dst[k] = bl * a.pitch + tf;
          if (bl < nb && t >= 0) v[k] = x[(b0 + bl) * a.sb + t * a.st + (int64_t)f * a.sf];
        }
      }
#pragma unroll
      for (int k = 0; k < kLd; ++k)
        if (dst[k] >= 0) tile[dst[k]] = v[k];
    }
  } else {
    // feature axis is the fast one (env-major tensors): lanes walk (t, f) of one sequence
    for (int bl = warp; bl < nb; bl += kViewThreads / 32) {
      const U* xb = x + (b0 + bl) * a.sb;
      U* tb_ = tile + bl * a.pitch;
      for (int tf0 = lane; tf0 < per_seq; tf0 += kLd * 32) {
        U v[kLd];
#pragma unroll
        for (int k = 0; k < kLd; ++k) {
          const int tf = tf0 + k * 32;
          v[k] = 0;
          if (tf < per_seq) {
            const int tl = a.f_magic ? (int)__umulhi((uint32_t)tf, a.f_magic) : tf, f = tf - tl * F;
            const int64_t t = t0 + tl;
            if (t >= 0) v[k] = xb[t * a.st + (int64_t)f * a.sf];
          }
        }
#pragma unroll
        for (int k = 0; k < kLd; ++k)
          if (tf0 + k * 32 < per_seq) tb_[tf0 + k * 32] = v[k];
      }
    }
  }
  __syncthreads();

  // ---- stream the windows of each sequence: out run of nw * size * F elements ----------------------------------
  // element e = w * (size * F) + rem  reads  tile[w * F + rem]   ((w + s) * F + f with rem = s * F + f)
  const int SF = a.size * F;
  const int run = nw * SF;
  constexpr int kVec = 16 / (int)sizeof(U);  // elements per 128-bit store
  if (a.vec) {
    // every sequence's run starts 16-byte aligned and a window is a whole number of vectors
    const int vpw = SF / kVec;  // vectors per window
    const int nvec = nw * vpw;
    const int q32 = 32 / vpw, r32 = 32 - q32 * vpw;
    for (int bl = warp; bl < nb; bl += kViewThreads / 32) {
      uint4* dst = reinterpret_cast<uint4*>(out + ((b0 + bl) * a.count + w0) * (int64_t)SF);
      const U* src = tile + bl * a.pitch;
      int w = lane / vpw, j = lane - w * vpw;
      for (int v = lane; v < nvec; v += 32) {
        const U* p = src + w * F + j * kVec;
        uint4 o;
        if constexpr (sizeof(U) == 4) {
          o = make_uint4(p[0], p[1], p[2], p[3]);
        } else {
          o = make_uint4((uint32_t)p[0], (uint32_t)(p[0] >> 32), (uint32_t)p[1], (uint32_t)(p[1] >> 32));
        }
        asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst + v), "r"(o.x),
                     "r"(o.y), "r"(o.z), "r"(o.w));
        w += q32, j += r32;
        if (j >= vpw) j -= vpw, ++w;
      }
    }
  } else {
    const int q32 = 32 / SF, r32 = 32 - q32 * SF;
    for (int bl = warp; bl < nb; bl += kViewThreads / 32) {
      U* dst = out + ((b0 + bl) * a.count + w0) * (int64_t)SF;
      const U* src = tile + bl * a.pitch;
      int w = lane / SF, rem = lane - w * SF;
      for (int e = lane; e < run; e += 32) {
        dst[e] = src[w * F + rem];
        w += q32, rem += r32;
        if (rem >= SF) rem -= SF, ++w;
      }
    }
  }
  // ---- padding mask: true where the window element lies before the start of the sequence -----------------------
  // (a function of (w, s) only: one 32-bit word of four flags per store when every sequence's run is word-aligned)
  if (a.mask) {
    const int per = nw * a.size;  // mask bytes of one sequence in this window block
    if (a.mask_vec) {
      const int words = per >> 2;
      for (int e = tid; e < nb * words; e += kViewThreads) {
        const int bl = e / words, i = e - bl * words;
        int w = (4 * i) / a.size, sft = 4 * i - w * a.size;
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          word |= (uint32_t)((t0 + w + sft) < 0) << (8 * j);
          if (++sft == a.size) sft = 0, ++w;
        }
        reinterpret_cast<uint32_t*>(a.mask + ((b0 + bl) * a.count + w0) * a.size)[i] = word;
      }
    } else {
      for (int e = tid; e < nb * per; e += kViewThreads) {
        const int bl = e / per, ws = e - bl * per;
        const int w = ws / a.size, sft = ws - w * a.size;
        a.mask[((b0 + bl) * a.count + w0) * a.size + ws] = (t0 + w + sft) < 0 ? 1 : 0;
      }
    }
  }
}

}  // namespace rl8

using namespace rl8;

extern "C" int rl8_view_windows(const void* x, int32_t elem_bytes, int64_t B, int64_t T, int64_t F,
                                int64_t stride_b, int64_t stride_t, int64_t stride_f, int32_t size,
                                int64_t t_first, int64_t count, void* out, uint8_t* mask,
                                rl8_stream_t stream) {
  if (B < 0 || T < 0 || F < 0 || size < 1 || count < 0) return RL8_ERR_ARG;
  if (B == 0 || F == 0 || count == 0) return RL8_OK;
  if (!x || !out) return RL8_ERR_ARG;
  if (t_first + count - 1 + size - 1 >= T) return RL8_ERR_ARG;  // windows never run past the sequence end
  if (elem_bytes != 4 && elem_bytes != 8) return RL8_ERR_UNSUPPORTED;
  if (F * size > (1 << 20)) return RL8_ERR_UNSUPPORTED;
  ViewArgs a;
  a.x = x, a.out = out, a.mask = mask;
  a.B = B, a.T = T, a.F = F, a.sb = stride_b, a.st = stride_t, a.sf = stride_f;
  a.t_first = t_first, a.count = count, a.size = size;
  // tile shape: up to 32 sequences x as many windows as the staging budget holds
  const int64_t budget = kViewSmemBytes / elem_bytes;  // elements
  int tb = 32;
  int64_t tw = 0;
  for (;; tb /= 2) {
    tw = (budget / tb - 1) / F - (size - 1);
    if (tw >= 1 || tb == 1) break;
  }
  if (tw < 1) return RL8_ERR_UNSUPPORTED;  // one window of one sequence exceeds the staging tile
  if (tw > count) tw = count;
  if (tb > B) tb = (int)B;
  a.tb = tb, a.tw = (int)tw;
  int64_t pitch = (tw + size - 1) * F;
  pitch |= 1;  // odd: conflict-free when lanes walk the sequence axis
  a.pitch = (int)pitch;
  a.tb_shift = -1;
  for (int sft = 0; sft < 6; ++sft)
    if ((1 << sft) == tb) a.tb_shift = sft;
  a.f_magic = F == 1 ? 0u : (uint32_t)((1ull << 32) / (uint64_t)F + 1);
  a.vec = ((uintptr_t)out % 16 == 0 && (size * F) % (16 / elem_bytes) == 0) ? 1 : 0;
  a.ld_vec = (elem_bytes == 4 && stride_b == 1 && tb % 4 == 0 && B % 4 == 0 && stride_t % 4 == 0 &&
              stride_f % 4 == 0 && (uintptr_t)x % 16 == 0) ? 1 : 0;
  a.mask_vec = (mask && (uintptr_t)mask % 4 == 0 && (count * size) % 4 == 0 && (tw * size) % 4 == 0) ? 1 : 0;
  const int64_t gx = ceil_div(B, tb), gy = ceil_div(count, tw);
  if (gy > 65535) return RL8_ERR_UNSUPPORTED;
  const size_t smem = (size_t)tb * pitch * elem_bytes;
  dim3 grid((unsigned)gx, (unsigned)gy);
  cudaStream_t st = (cudaStream_t)stream;
  if (elem_bytes == 4) {
    static bool attr4 = false;
    if (!attr4) {
      cudaFuncSetAttribute(view_windows_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kViewSmemBytes + 1024);
      attr4 = true;
    }
    view_windows_kernel<uint32_t><<<grid, kViewThreads, smem, st>>>(a);
  } else {
    static bool attr8 = false;
    if (!attr8) {
      cudaFuncSetAttribute(view_windows_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           kViewSmemBytes + 1024);
      attr8 = true;
    }
    view_windows_kernel<uint64_t><<<grid, kViewThreads, smem, st>>>(a);
  }
  return check_launch("view_windows");
}
