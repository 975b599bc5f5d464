// SalUn gradient-saliency mask and the masked AdamW update (HBM-bound byte/integer work).
//
// Reference: methods/salun/generate_mask.py:17-89 (`save_gradient_ratio`):
//   :41-44  gradients[name] += param.grad            -> saliency_accumulate  (bf16 += bf16)
//   :47-49  abs_().cpu().to(float16)                  -> sal_key(): the fp16 bit pattern of |g|
//   :61-78  per tensor: sort descending, top int(th*numel) -> 1 (int8)
//                                                     -> histogram over the 32768 non-negative fp16
//                                                        patterns, select the k-th largest pattern and
//                                                        the tie budget, write the mask
// Order statistics over fp16 keys need no sort: non-negative fp16 bit patterns order like the values,
// so the k-th largest element is found from a 32768-bin histogram, and the reference's stable tie
// rule (lowest flat index first among equal keys) is an exclusive scan over "== tau".
// The masked update is the upstream SalUn semantic `grad *= mask` followed by AdamW
// in torch's `_single_tensor_adam` operation order; fp32 arithmetic is spelled with explicit
// round-to-nearest intrinsics so that no FMA contraction can change a bit relative to the oracle.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.h"
#include "adam_common.cuh"

namespace clear {
namespace {

constexpr int kBins = 32768;
constexpr int kChunk = 8192;        // elements per mask chunk (256 threads x 4 vectors x 8)
constexpr int kChunkThreads = 256;
constexpr int kChunkIters = kChunk / (kChunkThreads * 8);

__device__ __forceinline__ uint32_t sal_key(__nv_bfloat16 x) {
  const float f = fabsf(__bfloat162float(x));
  return (uint32_t)__half_as_ushort(__float2half_rn(f)) & 0x7FFFu;
}

__device__ __forceinline__ void keys8(const uint4& raw, uint32_t (&k)[8]) {
  const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
  for (int j = 0; j < 8; ++j) k[j] = sal_key(h[j]);
}

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------------
// torch's bf16 add on a packed pair: operands widened to fp32, added, rounded to nearest-even bf16
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t g) {
  const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a));
  const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g));
  const __nv_bfloat162 r = __floats2bfloat162_rn(__fadd_rn(x.x, y.x), __fadd_rn(x.y, y.y));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// kWide: 256-bit accesses (sm_100 LDG/STG.E.ENL2.256; both pointers 32-byte aligned), else 128-bit
template <bool kWide>
__global__ void __launch_bounds__(256) accumulate_kernel(__nv_bfloat16* acc, const __nv_bfloat16* grad, int64_t n) {
  constexpr int kElems = kWide ? 16 : 8;
  const int64_t nvec = n / kElems;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    if constexpr (kWide) {
      uint32_t a[8], g[8];
      asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                   : "l"(acc + v * 16));
      asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]), "=r"(g[4]), "=r"(g[5]), "=r"(g[6]), "=r"(g[7])
                   : "l"(grad + v * 16));
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = add_bf16x2(a[j], g[j]);
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(acc + v * 16), "r"(a[0]), "r"(a[1]), "r"(a[2]),
                   "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                   : "memory");
    } else {
      uint4 a = reinterpret_cast<const uint4*>(acc)[v];
      const uint4 g = ld_stream(reinterpret_cast<const uint4*>(grad) + v);
      a.x = add_bf16x2(a.x, g.x);
      a.y = add_bf16x2(a.y, g.y);
      a.z = add_bf16x2(a.z, g.z);
      a.w = add_bf16x2(a.w, g.w);
      reinterpret_cast<uint4*>(acc)[v] = a;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec * kElems; i < n; ++i)
      acc[i] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(acc[i]), __bfloat162float(grad[i])));
  }
}

// ------------------------------------------------------------------------------------------------
// Histogram: each CTA owns a contiguous slice of the flat buffer and a private 32768-bin histogram in
// shared memory (128 KB); it flushes (sparse global atomics) whenever the slice crosses into the next
// tensor and at the end.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_segment(const int64_t* off, int n_segs, int64_t pos) {
  int lo = 0, hi = n_segs - 1;  // largest s with off[s] <= pos
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= pos) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// The histogram is taken over the 32768 bf16 MAGNITUDE patterns (raw bits & 0x7FFF: no conversion per element), and the
// bins are mapped to their fp16 keys when a CTA flushes: the key is a monotone function of the pattern (many patterns share
// key 0 and the key of infinity), so  hist_key[key(m)] += hist_pattern[m]  gives exactly the histogram of the keys.
constexpr int kHistThreads = 1024;
constexpr int kHistUnroll = 4;  // 128-bit loads in flight per thread
// Lane-private window.  A warp's 32 shared-memory atomics into random bins cost an expected 3.4 bank-conflict passes (plus
// same-address serialisation: gradients populate only a few hundred bins).  The bulk of a tensor's |g| values lies within
// a few octaves, so the kWinSlots patterns around the tensor's typical magnitude (estimated per CTA and tensor from the
// first 1024 vectors of the slice; any estimate is correct, a poor one is only slow) get one counter PER LANE: pattern m
// of lane l lives in bank l, and the atomic of an in-window pattern never conflicts within the warp.  Patterns outside the
// window take the shared bins.
constexpr int kWinSlots = 768;                       // 6 octaves of bf16 magnitudes (128 patterns per octave)
constexpr int kWinWords = kWinSlots * 32;            // 96 KB (+ 128 KB of shared bins: 224 of the 227 KB an SM offers)
constexpr int kHistSmemBytes = (kBins + kWinWords) * 4;

__global__ void __launch_bounds__(kHistThreads) histogram_kernel(const __nv_bfloat16* acc, const int64_t* seg_off, int n_segs,
                                                                 int64_t n_total, int64_t per_block, uint32_t* hist, int win_slots) {
  extern __shared__ uint32_t sh[];
  uint32_t* win = sh + kBins;
  __shared__ unsigned long long est[2];  // (sum of sampled patterns, number of samples)
  __shared__ int win_base_sh;
  for (int b = threadIdx.x; b < kBins + kWinWords; b += kHistThreads) sh[b] = 0;
  if (threadIdx.x < 2) est[threadIdx.x] = 0;
  __syncthreads();
  int64_t pos = (int64_t)blockIdx.x * per_block;
  const int64_t end = min(pos + per_block, n_total);
  if (pos >= end) return;
  const uint32_t lane = threadIdx.x & 31;
  int seg = find_segment(seg_off, n_segs, pos);
  while (pos < end) {
    while (seg_off[seg + 1] <= pos) ++seg;  // skips empty tensors
    const int64_t stop = min(end, seg_off[seg + 1]);
    const int64_t v1 = stop >> 3;
    {  // window placement: mean pattern of one sampled element per thread (zeros excluded), minus half a window
      const int64_t v = (pos >> 3) + threadIdx.x;
      unsigned ssum = 0, scnt = 0;
      if (v < v1) {
        const uint32_t m = (uint32_t)__bfloat16_as_ushort(acc[v << 3]) & 0x7FFFu;
        if (m) { ssum = m; scnt = 1; }
      }
      ssum = __reduce_add_sync(0xffffffffu, ssum);
      scnt = __reduce_add_sync(0xffffffffu, scnt);
      if (lane == 0 && scnt) { atomicAdd(&est[0], (unsigned long long)ssum); atomicAdd(&est[1], (unsigned long long)scnt); }
      __syncthreads();
      if (threadIdx.x == 0) {
        int base = est[1] ? (int)(est[0] / est[1]) - kWinSlots / 2 : 0;
        base = max(0, min(base, kBins - kWinSlots));
        win_base_sh = base;
        est[0] = est[1] = 0;
      }
      __syncthreads();
    }
    const uint32_t win_base = (uint32_t)win_base_sh;
    // index of pattern m: the lane's private counter (bank = lane) inside the window, the shared bin outside; no branch
    auto count = [&](uint32_t m) {
      const uint32_t rel = m - win_base;
      const uint32_t idx = (rel < (uint32_t)win_slots) ? (uint32_t)kBins + rel * 32u + lane : m;
      atomicAdd(&sh[idx], 1u);
    };
    for (int64_t v0 = (pos >> 3) + threadIdx.x; v0 < v1; v0 += (int64_t)kHistThreads * kHistUnroll) {
      uint4 raw[kHistUnroll];
#pragma unroll
      for (int u = 0; u < kHistUnroll; ++u) {
        const int64_t v = v0 + (int64_t)u * kHistThreads;
        if (v < v1) raw[u] = ld_stream(reinterpret_cast<const uint4*>(acc) + v);
      }
#pragma unroll
      for (int u = 0; u < kHistUnroll; ++u) {
        const int64_t v = v0 + (int64_t)u * kHistThreads;
        if (v < v1) {
          const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            count(w[q] & 0x7FFFu);
            count((w[q] >> 16) & 0x7FFFu);
          }
        }
      }
    }
    __syncthreads();
    // fold the window into the shared bins (thread t sums the 32 lane counters of slot t, skewed so that a warp's reads
    // of one step fall into 32 different banks), then flush the populated bins to the bins of their fp16 keys
    for (int t = threadIdx.x; t < kWinSlots; t += kHistThreads) {
      uint32_t c = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        const int ll = (l + t) & 31;
        c += win[t * 32 + ll];
        win[t * 32 + ll] = 0;
      }
      if (c) sh[win_base + t] += c;
    }
    __syncthreads();
    uint32_t* out = hist + (size_t)seg * kBins;
    for (int b = threadIdx.x; b < kBins; b += kHistThreads) {
      const uint32_t c = sh[b];
      if (c) { atomicAdd(&out[sal_key(__ushort_as_bfloat16((unsigned short)b))], c); sh[b] = 0; }
    }
    __syncthreads();
    pos = stop;
  }
}

// ------------------------------------------------------------------------------------------------
// Select: one CTA per tensor.  Descending cumulative counts over the bins; for every threshold the bin
// tau with  #{key > tau} < k <= #{key >= tau}  and ties = k - #{key > tau}.
// Output per (tensor, threshold): (tau, ties, lo, hi, total) where total = #{key == tau} and [lo, hi] is the range of bf16 magnitude bit
// patterns whose key equals tau (the key is monotone in the magnitude pattern), so that the mask pass
// compares raw 16-bit patterns and never converts:  key > tau  <=>  |x|bits > hi.
// ------------------------------------------------------------------------------------------------
constexpr int kSelFields = 5;
__device__ __forceinline__ uint32_t key_of_pattern(uint32_t m) {
  return sal_key(__ushort_as_bfloat16((unsigned short)m));
}
__device__ void write_select(int64_t* out, int64_t tau, int64_t ties, int64_t total) {
  out[0] = tau;
  out[1] = ties;
  out[4] = total;  // elements of the tensor whose key is tau
  if (tau < 0) {  // nothing selected: no pattern is above hi, none inside [lo, hi]
    out[2] = 0x8000;
    out[3] = 0x7FFF;
    return;
  }
  uint32_t lo = 0, hi = 0x7FFF;  // smallest m with key(m) >= tau
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (key_of_pattern(mid) >= (uint32_t)tau) hi = mid; else lo = mid + 1;
  }
  out[2] = lo;
  uint32_t a = lo, b = 0x7FFF;   // largest m with key(m) <= tau
  while (a < b) {
    const uint32_t mid = (a + b + 1) >> 1;
    if (key_of_pattern(mid) <= (uint32_t)tau) a = mid; else b = mid - 1;
  }
  out[3] = a;
}
constexpr int kMaxThresholds = 16;
struct SelectTargets {
  double th[kMaxThresholds];
  int n;
};

__global__ void __launch_bounds__(1024) select_kernel(const uint32_t* hist, const int64_t* seg_off,
                                                      const int64_t* seg_numel, SelectTargets tg, int64_t* select) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long wbase[32];
  const int seg = blockIdx.x;
  // Tensors whose numel is not a multiple of 8 are stored zero-padded up to the next offset: the
  // padding was histogrammed into bin 0 and is taken out again here.
  const int64_t extent = seg_off[seg + 1] - seg_off[seg];
  const int64_t numel = seg_numel ? seg_numel[seg] : extent;
  const unsigned long long pad = (unsigned long long)(extent - numel);
  const uint32_t* h = hist + (size_t)seg * kBins;
  // thread t owns the 32 bins [hi-31, hi], hi = 32767 - 32 t  (descending order)
  const int hi = kBins - 1 - 32 * (int)threadIdx.x;
  unsigned long long local = 0;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) local += h[hi - i];
  if (hi - 31 == 0) local -= pad;
  // exclusive scan of `local` across the block, in thread order
  unsigned long long incl = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = wsum[lane], s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    wbase[lane] = s - v;
  }
  __syncthreads();
  const unsigned long long above0 = wbase[warp] + incl - local;  // #{key > hi}
  for (int j = 0; j < tg.n; ++j) {
    // Python: int(threshold * num_elements) — a double product, truncated (generate_mask.py:69)
    long long k = (long long)(tg.th[j] * (double)numel);
    if (k > numel) k = numel;  // indices[:k] clamps
    int64_t* out = select + ((size_t)seg * tg.n + j) * kSelFields;
    if (k <= 0) {
      if (threadIdx.x == 0) write_select(out, -1, 0, 0);
      continue;
    }
    if ((unsigned long long)k > above0 && (unsigned long long)k <= above0 + local) {
      unsigned long long above = above0;
      for (int i = 0; i < 32; ++i) {
        const unsigned long long c = h[hi - i] - ((hi - i) == 0 ? pad : 0ull);
        if ((unsigned long long)k <= above + c) { write_select(out, hi - i, (int64_t)((unsigned long long)k - above), (int64_t)c); break; }
        above += c;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Global-threshold scope (upstream SalUn's global rank; SURVEY.md App. B #11, north_star "a global threshold"): ONE ranking
// over all tensors.  The per-tensor histograms above already hold everything needed:
//   hist_total          total[b] = sum_s hist[s][b]  (zero padding of every tensor taken out of bin 0) — with the flat
//                       buffer sharded over ranks, the caller all-reduces `total` (256 KB of int64) here;
//   select_global_tau   tau_j = key of the k_j-th largest element overall, k_j = int(th_j * total numel), and the tie budget;
//   select_global_fill  per tensor the same (tau, ties, lo, hi, total) record the per-tensor select writes, with the
//                       global tie budget handed out in flat index order: tensor s takes
//                       clamp(budget - #ties in earlier tensors, 0, #ties in s).
// The mask pass is unchanged: it consumes the per-tensor records.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hist_total_kernel(const uint32_t* hist, const int64_t* seg_off, const int64_t* seg_numel,
                                                         int n_segs, long long* total) {
  const int b = blockIdx.x * 256 + threadIdx.x;  // one bin per thread; a warp reads 128 contiguous bytes per tensor
  long long t = 0;
  for (int s = 0; s < n_segs; ++s) t += hist[(size_t)s * kBins + b];
  if (b == 0 && seg_numel)
    for (int s = 0; s < n_segs; ++s) t -= (seg_off[s + 1] - seg_off[s]) - seg_numel[s];
  total[b] = t;
}

__global__ void __launch_bounds__(1024) select_global_tau_kernel(const long long* total, long long n_total, SelectTargets tg,
                                                                 int64_t* tau_budget) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long wbase[32];
  const int hi = kBins - 1 - 32 * (int)threadIdx.x;  // thread t owns the 32 bins [hi-31, hi], descending
  unsigned long long local = 0;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) local += (unsigned long long)total[hi - i];
  unsigned long long incl = local;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = wsum[lane], s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    wbase[lane] = s - v;
  }
  __syncthreads();
  const unsigned long long above0 = wbase[warp] + incl - local;  // #{key > hi}
  for (int j = 0; j < tg.n; ++j) {
    long long k = (long long)(tg.th[j] * (double)n_total);  // int(threshold * num_elements), generate_mask.py:69
    if (k > n_total) k = n_total;
    if (k <= 0) {
      if (threadIdx.x == 0) { tau_budget[2 * j] = -1; tau_budget[2 * j + 1] = 0; }
      continue;
    }
    if ((unsigned long long)k > above0 && (unsigned long long)k <= above0 + local) {
      unsigned long long above = above0;
      for (int i = 0; i < 32; ++i) {
        const unsigned long long c = (unsigned long long)total[hi - i];
        if ((unsigned long long)k <= above + c) {
          tau_budget[2 * j] = hi - i;
          tau_budget[2 * j + 1] = (int64_t)((unsigned long long)k - above);
          break;
        }
        above += c;
      }
    }
  }
}

// One CTA per threshold: exclusive scan, in tensor order, of the tensors' tie counts against the budget.
__global__ void __launch_bounds__(1024) select_global_fill_kernel(const uint32_t* hist, const int64_t* seg_off,
                                                                  const int64_t* seg_numel, int n_segs, const int64_t* tau_budget,
                                                                  const int64_t* budget_override, int n_th, int64_t* select) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry;
  const int j = blockIdx.x;
  const int64_t tau = tau_budget[2 * j];
  const unsigned long long budget = (unsigned long long)(budget_override ? budget_override[j] : tau_budget[2 * j + 1]);
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_segs; base += 1024) {
    const int s = base + (int)threadIdx.x;
    unsigned long long c = 0;
    if (s < n_segs && tau >= 0) {
      c = hist[(size_t)s * kBins + tau];
      if (tau == 0 && seg_numel) c -= (unsigned long long)((seg_off[s + 1] - seg_off[s]) - seg_numel[s]);
    }
    unsigned long long incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long wb = 0;
    for (int w = 0; w < warp; ++w) wb += wsum[w];
    const unsigned long long excl = carry + wb + incl - c;
    if (s < n_segs) {
      int64_t* out = select + ((size_t)s * n_th + j) * kSelFields;
      if (tau < 0) {
        write_select(out, -1, 0, 0);
      } else {
        const unsigned long long left = budget > excl ? budget - excl : 0ull;
        write_select(out, tau, (int64_t)(left < c ? left : c), (int64_t)c);
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + c;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Mask.  Chunks of 8192 elements never straddle tensors (chunk_prefix[s] = number of chunks before
// tensor s).  Flat order inside a chunk is (iter, thread, lane-in-vector), loads stay coalesced.
// ------------------------------------------------------------------------------------------------
// One streaming pass (2 B read + 1 B written per parameter) with no communication between CTAs:
//   mask_kernel   every chunk writes  key > tau -> 1  and counts its ties (key == tau).  Whether the ties
//                 of a chunk are selected depends on how many ties precede it in the tensor, which the
//                 chunk does not know yet; it SPECULATES from the histogram: ties are taken in chunks
//                 before the point where `budget / total` of the tensor's chunks have passed.
//   tie_scan      exclusive scan of the chunk tie counts per tensor (a few thousand numbers).
//   tie_fix       every chunk checks its speculation against the exact prefix; the chunks that guessed
//                 wrong (those between the guessed and the true boundary) and the one boundary chunk
//                 are rewritten exactly, ties ranked in flat index order.
// The result never depends on the speculation; only the amount of rewriting does (a handful of
// chunks per tensor when ties are spread evenly, every chunk of a tensor in the worst case).
constexpr int kCountShift = 48;  // scratch word = (chunk tie count << 48) | exclusive tie prefix

__device__ __forceinline__ void locate_chunk(int64_t c, const int64_t* seg_off, const int64_t* chunk_prefix, int n_segs,
                                             int& seg, int64_t& first_chunk, int64_t& begin, int64_t& end) {
  int lo = 0, hi = n_segs - 1;  // largest s with chunk_prefix[s] <= c
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_prefix[mid] <= c) lo = mid; else hi = mid - 1;
  }
  seg = lo;
  first_chunk = chunk_prefix[lo];
  begin = seg_off[lo] + (c - first_chunk) * kChunk;
  end = min(begin + (int64_t)kChunk, seg_off[lo + 1]);
}

// Speculation: chunk number `rel` of `n` takes its ties iff rel / n < budget / total.
__device__ __forceinline__ bool speculate_take(int64_t rel, int64_t n, int64_t budget, int64_t total) {
  return (unsigned long long)rel * (unsigned long long)total < (unsigned long long)budget * (unsigned long long)n;
}

// Eight mask elements of one thread (sel01[q]: bit 0 = element 2q, bit 16 = element 2q + 1), at flat element e
// (a multiple of 8).  int8 format: one byte per parameter, the reference's `with_{th}.pt` dtype
// (generate_mask.py:72).  Bit-packed format: bit (e & 7) of byte e >> 3, 1/8 of the traffic for the masked update.
template <bool kPacked>
__device__ __forceinline__ void store_mask8(void* mask, int64_t e, const uint32_t (&sel01)[4]) {
  if constexpr (kPacked) {
    // bits {0, 16} of four words -> bits {0, 2, 4, 6} and {16, 18, 20, 22} of one, then fold the upper half down by 15
    const uint32_t x = sel01[0] + (sel01[1] << 2) + (sel01[2] << 4) + (sel01[3] << 6);
    static_cast<uint8_t*>(mask)[e >> 3] = (uint8_t)((x | (x >> 15)) & 0xFFu);
  } else {
    *reinterpret_cast<uint2*>(static_cast<int8_t*>(mask) + e) =
        make_uint2(__byte_perm(sel01[0], sel01[1], 0x6420), __byte_perm(sel01[2], sel01[3], 0x6420));
  }
}

// Exact rewrite of one chunk given the number of ties before it.  Called by a whole CTA.
template <bool kPacked>
__device__ __noinline__ void mask_ranked_chunk(const __nv_bfloat16* acc, int64_t begin, int64_t end, uint32_t lo2,
                                               uint32_t hi2, unsigned long long prefix, unsigned long long budget,
                                               uint32_t* wcnt /* [kChunkIters][kChunkThreads / 32] */, void* mask) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kChunkThreads / 32;
  uint32_t before[kChunkIters];
  for (int it = 0; it < kChunkIters; ++it) {
    const int64_t e = begin + ((int64_t)it * kChunkThreads + threadIdx.x) * 8;
    uint32_t cnt = 0;
    if (e < end) {
      const uint4 raw = ld_stream(acc + e);
      const uint32_t w[4] = {raw.x & 0x7FFF7FFFu, raw.y & 0x7FFF7FFFu, raw.z & 0x7FFF7FFFu, raw.w & 0x7FFF7FFFu};
      for (int q = 0; q < 4; ++q) cnt += __popc(__vcmpgeu2(w[q], lo2) & ~__vcmpgtu2(w[q], hi2) & 0x00010001u);
    }
    uint32_t incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    before[it] = incl - cnt;
    if (lane == 31) wcnt[it * kWarps + warp] = incl;
  }
  __syncthreads();
  unsigned long long run = prefix;
  for (int it = 0; it < kChunkIters; ++it) {
    unsigned long long rank = run + before[it];
    for (int w = 0; w < kWarps; ++w) {
      if (w < warp) rank += wcnt[it * kWarps + w];
      run += wcnt[it * kWarps + w];
    }
    const int64_t e = begin + ((int64_t)it * kChunkThreads + threadIdx.x) * 8;
    if (e < end) {
      const uint4 raw = ld_stream(acc + e);
      const uint32_t w[4] = {raw.x & 0x7FFF7FFFu, raw.y & 0x7FFF7FFFu, raw.z & 0x7FFF7FFFu, raw.w & 0x7FFF7FFFu};
      uint32_t sel01[4];
      for (int q = 0; q < 4; ++q) {  // element 2q is the low half of word q, 2q + 1 the high half
        const uint32_t g = __vcmpgtu2(w[q], hi2);
        const uint32_t t = __vcmpgeu2(w[q], lo2) & ~g;
        uint32_t kp = 0;
        if (t & 0x0000FFFFu) { if (rank < budget) kp |= 0x0000FFFFu; ++rank; }
        if (t & 0xFFFF0000u) { if (rank < budget) kp |= 0xFFFF0000u; ++rank; }
        sel01[q] = (g | kp) & 0x00010001u;
      }
      store_mask8<kPacked>(mask, e, sel01);
    }
  }
  __syncthreads();  // wcnt may be reused by the caller's next chunk
}

// Pattern comparisons two at a time.  Sign-stripped bf16 bit patterns order like the bf16 values they
// encode, so `pattern > hi` is one HSET2.BF16 per word (the "unordered" forms treat NaN patterns, which
// sit above infinity, as greater).  If a bound itself is a NaN pattern (a threshold that lands among NaN
// keys) the integer SIMD comparison is used instead.
template <bool kFloatCmp>
__device__ __forceinline__ void cmp_word(uint32_t w, uint32_t lo2, uint32_t hi2, uint32_t& g, uint32_t& ge) {
  if constexpr (kFloatCmp) {
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&w);
    g = __hgtu2_mask(x, *reinterpret_cast<const __nv_bfloat162*>(&hi2));
    ge = __hgeu2_mask(x, *reinterpret_cast<const __nv_bfloat162*>(&lo2));
  } else {
    g = __vcmpgtu2(w, hi2);
    ge = __vcmpgeu2(w, lo2);
  }
}

constexpr int kChunksPerCta = 4;

template <bool kFloatCmp, bool kPacked>
__device__ __forceinline__ uint32_t mask_chunk_fast(const __nv_bfloat16* acc, int64_t begin, int64_t end, uint32_t lo2,
                                                    uint32_t hi2, uint32_t take_all, void* mask) {
  uint4 raw[kChunkIters];
#pragma unroll
  for (int it = 0; it < kChunkIters; ++it) {
    const int64_t e = begin + ((int64_t)it * kChunkThreads + threadIdx.x) * 8;
    raw[it] = (e < end) ? ld_stream(acc + e) : make_uint4(0, 0, 0, 0);
  }
  uint32_t packed = 0;  // two 16-bit tie counters
#pragma unroll
  for (int it = 0; it < kChunkIters; ++it) {
    const int64_t e = begin + ((int64_t)it * kChunkThreads + threadIdx.x) * 8;
    if (e < end) {
      const uint32_t w[4] = {raw[it].x & 0x7FFF7FFFu, raw[it].y & 0x7FFF7FFFu, raw[it].z & 0x7FFF7FFFu, raw[it].w & 0x7FFF7FFFu};
      uint32_t sel01[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // per 32-bit word (two elements): g = pattern > hi, tie = lo <= pattern <= hi
        uint32_t g, ge;
        cmp_word<kFloatCmp>(w[q], lo2, hi2, g, ge);
        const uint32_t t01 = ge & ~g & 0x00010001u;
        packed += t01;
        sel01[q] = (g & 0x00010001u) | (t01 & take_all);
      }
      store_mask8<kPacked>(mask, e, sel01);
    }
  }
  return (packed & 0xFFFFu) + (packed >> 16);
}

template <bool kPacked>
__global__ void __launch_bounds__(kChunkThreads) mask_kernel(const __nv_bfloat16* acc, const int64_t* seg_off,
                                                             const int64_t* chunk_prefix, int n_segs, int64_t n_chunks,
                                                             const int64_t* select, int n_th, int th_index,
                                                             unsigned long long* scratch, void* mask) {
  __shared__ uint32_t wsum[kChunksPerCta][kChunkThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * kChunksPerCta;
  int seg;
  int64_t first_chunk, begin, end;
  locate_chunk(c0, seg_off, chunk_prefix, n_segs, seg, first_chunk, begin, end);
  int64_t seg_last = chunk_prefix[seg + 1];  // one past the tensor's last chunk
#pragma unroll 1
  for (int i = 0; i < kChunksPerCta; ++i) {
    const int64_t c = c0 + i;
    if (c >= n_chunks) break;
    if (c >= seg_last) {  // crossed into the next non-empty tensor
      do { ++seg; } while (chunk_prefix[seg + 1] <= c);
      first_chunk = chunk_prefix[seg];
      seg_last = chunk_prefix[seg + 1];
    }
    begin = seg_off[seg] + (c - first_chunk) * kChunk;
    end = min(begin + (int64_t)kChunk, seg_off[seg + 1]);
    const int64_t* sel = select + ((size_t)seg * n_th + th_index) * kSelFields;
    const uint32_t plo = (uint32_t)sel[2], phi = (uint32_t)sel[3];
    const uint32_t lo2 = plo | (plo << 16), hi2 = phi | (phi << 16);
    const uint32_t take_all = speculate_take(c - first_chunk, seg_last - first_chunk, sel[1], sel[4]) ? 0xFFFFFFFFu : 0u;
    const bool float_ok = plo <= 0x7F80u && phi <= 0x7F80u;  // block-uniform
    const uint32_t mine = float_ok ? mask_chunk_fast<true, kPacked>(acc, begin, end, lo2, hi2, take_all, mask)
                                   : mask_chunk_fast<false, kPacked>(acc, begin, end, lo2, hi2, take_all, mask);
    const uint32_t tot = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) wsum[i][warp] = tot;
  }
  __syncthreads();
  if (threadIdx.x < kChunksPerCta && c0 + threadIdx.x < n_chunks) {
    unsigned long long chunk_ties = 0;
    for (int w = 0; w < kChunkThreads / 32; ++w) chunk_ties += wsum[threadIdx.x][w];
    scratch[c0 + threadIdx.x] = chunk_ties << kCountShift;
  }
}

// One CTA per tensor: scratch[c] = (count << 48) | (ties in earlier chunks of the tensor).
__global__ void __launch_bounds__(1024) tie_scan_kernel(const int64_t* chunk_prefix, unsigned long long* scratch) {
  __shared__ unsigned long long carry;
  __shared__ unsigned long long wsum[32];
  const int64_t c0 = chunk_prefix[blockIdx.x], c1 = chunk_prefix[blockIdx.x + 1];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = c0; base < c1; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const unsigned long long word = (i < c1) ? scratch[i] : 0;
    const unsigned long long v = word >> kCountShift;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long wb = 0;
    for (int w = 0; w < warp; ++w) wb += wsum[w];
    const unsigned long long excl = carry + wb + incl - v;
    if (i < c1) scratch[i] = (v << kCountShift) | excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
}

// Thread t of CTA b audits chunk b * 256 + t; the CTA then rewrites the chunks that guessed wrong.
template <bool kPacked>
__global__ void __launch_bounds__(kChunkThreads) tie_fix_kernel(const __nv_bfloat16* acc, const int64_t* seg_off,
                                                                const int64_t* chunk_prefix, int n_segs, int64_t n_chunks,
                                                                const int64_t* select, int n_th, int th_index,
                                                                const unsigned long long* scratch, void* mask) {
  __shared__ uint32_t wcnt[kChunkIters * (kChunkThreads / 32)];
  __shared__ int n_wrong;
  __shared__ int wrong[kChunkThreads];
  if (threadIdx.x == 0) n_wrong = 0;
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * kChunkThreads + threadIdx.x;
  if (c < n_chunks) {
    int seg;
    int64_t first_chunk, begin, end;
    locate_chunk(c, seg_off, chunk_prefix, n_segs, seg, first_chunk, begin, end);
    const int64_t* sel = select + ((size_t)seg * n_th + th_index) * kSelFields;
    const unsigned long long budget = (unsigned long long)sel[1];
    const unsigned long long word = scratch[c];
    const unsigned long long cnt = word >> kCountShift, prefix = word & ((1ull << kCountShift) - 1);
    const bool spec = speculate_take(c - first_chunk, chunk_prefix[seg + 1] - first_chunk, sel[1], sel[4]);
    const bool all = prefix + cnt <= budget, none = prefix >= budget;
    const bool ok = cnt == 0 || (spec ? all : none);
    if (!ok) wrong[atomicAdd(&n_wrong, 1)] = (int)threadIdx.x;
  }
  __syncthreads();
  const int nw = n_wrong;
  for (int i = 0; i < nw; ++i) {  // block-uniform loop
    const int64_t cw = (int64_t)blockIdx.x * kChunkThreads + wrong[i];
    int seg;
    int64_t first_chunk, begin, end;
    locate_chunk(cw, seg_off, chunk_prefix, n_segs, seg, first_chunk, begin, end);
    const int64_t* sel = select + ((size_t)seg * n_th + th_index) * kSelFields;
    const uint32_t plo = (uint32_t)sel[2], phi = (uint32_t)sel[3];
    mask_ranked_chunk<kPacked>(acc, begin, end, plo | (plo << 16), phi | (phi << 16),
                               scratch[cw] & ((1ull << kCountShift) - 1), (unsigned long long)sel[1], wcnt, mask);
  }
}

// ------------------------------------------------------------------------------------------------
// masked AdamW
// ------------------------------------------------------------------------------------------------
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): the eight fp32 moments of a thread's eight parameters are 32
// contiguous bytes, so one instruction per warp covers 1 KB contiguously instead of two half-used 16-byte accesses at a
// 32-byte stride.
__device__ __forceinline__ void ld256(const float* p, float (&x)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float (&x)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]),
               "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7])
               : "memory");
}

constexpr int kAdamVecs = 1;  // 128-bit parameter vectors per thread and iteration (loads issued before any math)

// kPackedMask: `mask` is the bit-packed format (bit i & 7 of byte i >> 3): 22.125 instead of 23 bytes per parameter
template <bool kPackedMask>
__global__ void __launch_bounds__(256) adamw_kernel(__nv_bfloat16* param, const __nv_bfloat16* grad, float* exp_avg,
                                                    float* exp_avg_sq, const int8_t* mask, int64_t n, AdamConsts c) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float gscale = c.grad_scale ? *c.grad_scale : 1.f;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kAdamVecs) {
    uint4 pr[kAdamVecs], gr[kAdamVecs];
    float mm_[kAdamVecs][8], vv_[kAdamVecs][8];
    uint2 mk[kAdamVecs];
#pragma unroll
    for (int u = 0; u < kAdamVecs; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < nvec) {
        pr[u] = reinterpret_cast<const uint4*>(param)[i];
        gr[u] = ld_stream(reinterpret_cast<const uint4*>(grad) + i);
        ld256(exp_avg + (i << 3), mm_[u]);
        ld256(exp_avg_sq + (i << 3), vv_[u]);
        if constexpr (kPackedMask) mk[u] = make_uint2(reinterpret_cast<const uint8_t*>(mask)[i], 0u);
        else mk[u] = mask ? *reinterpret_cast<const uint2*>(mask + (i << 3)) : make_uint2(0x01010101u, 0x01010101u);
      }
    }
#pragma unroll
    for (int u = 0; u < kAdamVecs; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= nvec) break;
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(&pr[u]);
      const __nv_bfloat16* gh = reinterpret_cast<const __nv_bfloat16*>(&gr[u]);
      float (&mm)[8] = mm_[u];
      float (&vv)[8] = vv_[u];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t mb = kPackedMask ? ((mk[u].x >> j) & 1u) : (((j < 4 ? mk[u].x : mk[u].y) >> (8 * (j & 3))) & 0xFFu);
        float g = mb ? __bfloat162float(gh[j]) : 0.f;  // grad *= mask
        if (c.grad_scale) g = __bfloat162float(__float2bfloat16_rn(__fmul_rn(g, gscale)));  // grad.mul_(clip_coef) in bf16
        ph[j] = __float2bfloat16_rn(adam_one(__bfloat162float(ph[j]), g, mm[j], vv[j], c));
      }
      reinterpret_cast<uint4*>(param)[i] = pr[u];
      st256(exp_avg + (i << 3), mm);
      st256(exp_avg_sq + (i << 3), vv);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec << 3; i < n; ++i) {
      const bool on = kPackedMask ? ((reinterpret_cast<const uint8_t*>(mask)[i >> 3] >> (i & 7)) & 1u) : (!mask || mask[i]);
      float g = on ? __bfloat162float(grad[i]) : 0.f;
      if (c.grad_scale) g = __bfloat162float(__float2bfloat16_rn(__fmul_rn(g, gscale)));
      param[i] = __float2bfloat16_rn(adam_one(__bfloat162float(param[i]), g, exp_avg[i], exp_avg_sq[i], c));
    }
  }
}

// int8 mask (one byte per parameter, the on-disk `with_{th}.pt` format) <-> bit-packed mask.  One thread per 8 elements.
__global__ void __launch_bounds__(256) mask_pack_kernel(const int8_t* mask, int64_t n, uint8_t* bits) {
  const int64_t nbytes = (n + 7) >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
    uint32_t b = 0;
    if ((i << 3) + 8 <= n) {
      const uint2 m = *reinterpret_cast<const uint2*>(mask + (i << 3));
#pragma unroll
      for (int j = 0; j < 8; ++j) b |= ((((j < 4 ? m.x : m.y) >> (8 * (j & 3))) & 0xFFu) ? 1u : 0u) << j;
    } else {
      for (int j = 0; (i << 3) + j < n; ++j) b |= (mask[(i << 3) + j] ? 1u : 0u) << j;
    }
    bits[i] = (uint8_t)b;
  }
}

__global__ void __launch_bounds__(256) mask_unpack_kernel(const uint8_t* bits, int64_t n, int8_t* mask) {
  const int64_t nbytes = (n + 7) >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += stride) {
    const uint32_t b = bits[i];
    if ((i << 3) + 8 <= n) {
      uint2 m;
      m.x = (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21);
      m.y = ((b >> 4) & 1u) | (((b >> 4) & 2u) << 7) | (((b >> 4) & 4u) << 14) | (((b >> 4) & 8u) << 21);
      *reinterpret_cast<uint2*>(mask + (i << 3)) = m;
    } else {
      for (int j = 0; (i << 3) + j < n; ++j) mask[(i << 3) + j] = (int8_t)((b >> j) & 1u);
    }
  }
}

// diagnostic override of the lane-private window (CLEAR_HIST_WIN, read once, clamped to what the kernel allocates)
int hist_window_slots() {
  static const int slots = [] {
    const char* e = getenv("CLEAR_HIST_WIN");
    int v = e ? atoi(e) : kWinSlots;
    return v < 0 ? 0 : (v > kWinSlots ? kWinSlots : v);
  }();
  return slots;
}

int stream_grid(int64_t nvec, int threads) {
  const int sms = device_sm_count();
  int64_t blocks = (nvec + threads - 1) / threads;
  const int64_t cap = (int64_t)(sms > 0 ? sms : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace clear

using namespace clear;

template <bool kPacked>
static int saliency_mask_impl(const char* who, const void* acc, const int64_t* seg_offsets_dev, const int64_t* chunk_prefix_dev,
                              int32_t n_segs, int64_t n_chunks, const int64_t* select, int32_t n_th, int32_t th_index,
                              uint64_t* tie_scratch, void* mask, void* stream_) {
  if (!acc || !seg_offsets_dev || !chunk_prefix_dev || !select || !tie_scratch || !mask)
    return fail(CLEAR_ERR_NULL, "%s: null pointer", who);
  if (n_segs <= 0 || n_chunks <= 0 || n_chunks > INT32_MAX || th_index < 0 || th_index >= n_th)
    return fail(CLEAR_ERR_SHAPE, "%s: bad sizes", who);
  if ((reinterpret_cast<uintptr_t>(acc) & 15) || (!kPacked && (reinterpret_cast<uintptr_t>(mask) & 7)) ||
      (reinterpret_cast<uintptr_t>(tie_scratch) & 7))
    return fail(CLEAR_ERR_ALIGN, "%s: acc must be 16-byte, mask and tie_scratch 8-byte aligned", who);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(acc);
  unsigned long long* scratch = reinterpret_cast<unsigned long long*>(tie_scratch);
  mask_kernel<kPacked><<<(unsigned)((n_chunks + kChunksPerCta - 1) / kChunksPerCta), kChunkThreads, 0, stream>>>(
      a, seg_offsets_dev, chunk_prefix_dev, n_segs, n_chunks, select, n_th, th_index, scratch, mask);
  tie_scan_kernel<<<(unsigned)n_segs, 1024, 0, stream>>>(chunk_prefix_dev, scratch);
  tie_fix_kernel<kPacked><<<(unsigned)((n_chunks + kChunkThreads - 1) / kChunkThreads), kChunkThreads, 0, stream>>>(
      a, seg_offsets_dev, chunk_prefix_dev, n_segs, n_chunks, select, n_th, th_index, scratch, mask);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" {

int clear_saliency_accumulate(void* acc, const void* grad, int64_t n, void* stream) {
  if (!acc || !grad) return fail(CLEAR_ERR_NULL, "clear_saliency_accumulate: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_saliency_accumulate: n <= 0");
  if ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(grad)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_saliency_accumulate: pointers must be 16-byte aligned");
  if (((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(grad)) & 31) == 0)
    accumulate_kernel<true><<<stream_grid(n >> 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(acc), static_cast<const __nv_bfloat16*>(grad), n);
  else
    accumulate_kernel<false><<<stream_grid(n >> 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(acc), static_cast<const __nv_bfloat16*>(grad), n);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_saliency_histogram(const void* acc, const int64_t* seg_offsets_dev, int32_t n_segs, int64_t n_total,
                             uint32_t* hist, void* stream) {
  if (!acc || !seg_offsets_dev || !hist) return fail(CLEAR_ERR_NULL, "clear_saliency_histogram: null pointer");
  if (n_segs <= 0 || n_total <= 0 || (n_total & 7)) return fail(CLEAR_ERR_SHAPE, "clear_saliency_histogram: n_total must be a positive multiple of 8");
  if (reinterpret_cast<uintptr_t>(acc) & 15) return fail(CLEAR_ERR_ALIGN, "clear_saliency_histogram: acc must be 16-byte aligned");
  CLEAR_CUDA_OK(cudaFuncSetAttribute(histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHistSmemBytes));  // per device: set every time
  const int sms = device_sm_count();
  int64_t blocks = sms > 0 ? sms : 148;
  int64_t per_block = ((n_total + blocks - 1) / blocks + 4095) / 4096 * 4096;
  blocks = (n_total + per_block - 1) / per_block;
  histogram_kernel<<<(unsigned)blocks, kHistThreads, kHistSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(acc), seg_offsets_dev, n_segs, n_total, per_block, hist,
      hist_window_slots());
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_saliency_select(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev,
                          int32_t n_segs, const double* thresholds_host, int32_t n_th, int64_t* select, void* stream) {
  if (!hist || !seg_offsets_dev || !thresholds_host || !select) return fail(CLEAR_ERR_NULL, "clear_saliency_select: null pointer");
  if (n_segs <= 0 || n_th <= 0 || n_th > kMaxThresholds) return fail(CLEAR_ERR_SHAPE, "clear_saliency_select: 1 <= n_th <= %d", kMaxThresholds);
  SelectTargets tg;
  tg.n = n_th;
  for (int j = 0; j < n_th; ++j) tg.th[j] = thresholds_host[j];
  select_kernel<<<(unsigned)n_segs, 1024, 0, static_cast<cudaStream_t>(stream)>>>(hist, seg_offsets_dev, seg_numel_dev, tg, select);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_saliency_hist_total(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev, int32_t n_segs,
                              int64_t* total, void* stream) {
  if (!hist || !seg_offsets_dev || !total) return fail(CLEAR_ERR_NULL, "clear_saliency_hist_total: null pointer");
  if (n_segs <= 0) return fail(CLEAR_ERR_SHAPE, "clear_saliency_hist_total: n_segs <= 0");
  hist_total_kernel<<<kBins / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, seg_offsets_dev, seg_numel_dev, n_segs,
                                                                               reinterpret_cast<long long*>(total));
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_saliency_select_global_tau(const int64_t* total, int64_t n_total, const double* thresholds_host, int32_t n_th,
                                     int64_t* tau_budget, void* stream) {
  if (!total || !thresholds_host || !tau_budget) return fail(CLEAR_ERR_NULL, "clear_saliency_select_global_tau: null pointer");
  if (n_total <= 0 || n_th <= 0 || n_th > kMaxThresholds)
    return fail(CLEAR_ERR_SHAPE, "clear_saliency_select_global_tau: n_total > 0 and 1 <= n_th <= %d required", kMaxThresholds);
  SelectTargets tg;
  tg.n = n_th;
  for (int j = 0; j < n_th; ++j) tg.th[j] = thresholds_host[j];
  select_global_tau_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(total),
                                                                              (long long)n_total, tg, tau_budget);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_saliency_select_global_fill(const uint32_t* hist, const int64_t* seg_offsets_dev, const int64_t* seg_numel_dev,
                                      int32_t n_segs, const int64_t* tau_budget, const int64_t* budget_override, int32_t n_th,
                                      int64_t* select, void* stream) {
  if (!hist || !seg_offsets_dev || !tau_budget || !select) return fail(CLEAR_ERR_NULL, "clear_saliency_select_global_fill: null pointer");
  if (n_segs <= 0 || n_th <= 0 || n_th > kMaxThresholds) return fail(CLEAR_ERR_SHAPE, "clear_saliency_select_global_fill: bad sizes");
  select_global_fill_kernel<<<(unsigned)n_th, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      hist, seg_offsets_dev, seg_numel_dev, n_segs, tau_budget, budget_override, n_th, select);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int64_t clear_saliency_chunk_elems(void) { return kChunk; }

int clear_saliency_mask(const void* acc, const int64_t* seg_offsets_dev, const int64_t* chunk_prefix_dev,
                        int32_t n_segs, int64_t n_chunks, const int64_t* select, int32_t n_th, int32_t th_index,
                        uint64_t* tie_scratch, int8_t* mask, void* stream) {
  return saliency_mask_impl<false>("clear_saliency_mask", acc, seg_offsets_dev, chunk_prefix_dev, n_segs, n_chunks, select, n_th,
                                   th_index, tie_scratch, mask, stream);
}

int clear_saliency_mask_bits(const void* acc, const int64_t* seg_offsets_dev, const int64_t* chunk_prefix_dev,
                             int32_t n_segs, int64_t n_chunks, const int64_t* select, int32_t n_th, int32_t th_index,
                             uint64_t* tie_scratch, uint8_t* mask_bits, void* stream) {
  return saliency_mask_impl<true>("clear_saliency_mask_bits", acc, seg_offsets_dev, chunk_prefix_dev, n_segs, n_chunks, select,
                                  n_th, th_index, tie_scratch, mask_bits, stream);
}

int clear_mask_pack(const int8_t* mask, int64_t n, uint8_t* mask_bits, void* stream) {
  if (!mask || !mask_bits) return fail(CLEAR_ERR_NULL, "clear_mask_pack: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_mask_pack: n <= 0");
  if (reinterpret_cast<uintptr_t>(mask) & 7) return fail(CLEAR_ERR_ALIGN, "clear_mask_pack: mask must be 8-byte aligned");
  mask_pack_kernel<<<stream_grid((n + 7) >> 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, n, mask_bits);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_mask_unpack(const uint8_t* mask_bits, int64_t n, int8_t* mask, void* stream) {
  if (!mask || !mask_bits) return fail(CLEAR_ERR_NULL, "clear_mask_unpack: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_mask_unpack: n <= 0");
  if (reinterpret_cast<uintptr_t>(mask) & 7) return fail(CLEAR_ERR_ALIGN, "clear_mask_unpack: mask must be 8-byte aligned");
  mask_unpack_kernel<<<stream_grid((n + 7) >> 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask_bits, n, mask);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_masked_adamw(const ClearAdamWArgs* a, void* stream) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_masked_adamw: args is NULL");
  if (!a->param || !a->grad || !a->exp_avg || !a->exp_avg_sq) return fail(CLEAR_ERR_NULL, "clear_masked_adamw: null tensor");
  if (a->n <= 0 || a->step < 1) return fail(CLEAR_ERR_SHAPE, "clear_masked_adamw: n > 0 and step >= 1 required");
  if ((reinterpret_cast<uintptr_t>(a->param) | reinterpret_cast<uintptr_t>(a->grad) | reinterpret_cast<uintptr_t>(a->exp_avg) |
       reinterpret_cast<uintptr_t>(a->exp_avg_sq)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw: tensors must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(a->exp_avg) | reinterpret_cast<uintptr_t>(a->exp_avg_sq)) & 31)
    return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw: the fp32 moments must be 32-byte aligned (256-bit accesses)");
  if (a->mask && (reinterpret_cast<uintptr_t>(a->mask) & 7)) return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw: mask must be 8-byte aligned");
  if (a->mask && a->mask_bits) return fail(CLEAR_ERR_KIND, "clear_masked_adamw: pass mask or mask_bits, not both");
  const AdamConsts c = make_adam_consts(a->lr, a->beta1, a->beta2, a->eps, a->weight_decay, a->step, a->grad_scale_dev);
  if (a->mask_bits)
    adamw_kernel<true><<<stream_grid(a->n >> 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(a->param), static_cast<const __nv_bfloat16*>(a->grad), a->exp_avg, a->exp_avg_sq,
        reinterpret_cast<const int8_t*>(a->mask_bits), a->n, c);
  else
    adamw_kernel<false><<<stream_grid(a->n >> 3, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(a->param), static_cast<const __nv_bfloat16*>(a->grad), a->exp_avg, a->exp_avg_sq,
        a->mask, a->n, c);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
