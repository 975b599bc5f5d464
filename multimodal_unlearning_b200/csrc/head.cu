// Loss head of the CLEAR unlearning step: projection / row-wise loss + dlogits / backward projections.
//
// Reference path: `model(**inputs)` -> lm_head -> logits2probs / kl_div / HF CE / NPO /
// entropy -> autograd (mm/trainer.py:168-449, mm/trainer_utils.py:16-21,73-81).  The reference
// materialises 2-6 fp32 [N,V] tensors per loss; here the only [N,V]-sized storage is ONE bf16
// buffer per projected operand (student, and teacher when the kind needs it): the projection's
// epilogue emits the log-sum-exp statistics, the row pass turns logits into dlogits in place, and
// the backward projections consume that buffer.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"
#include "gemm_sm100.cuh"

namespace clear {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kIgnoreIndex = -100;
// The streaming row kernels keep one or two vectors per thread in flight in registers; vectors this many iterations ahead
// are requested into L2 with prefetch.global.L2 (forward KL, 256 threads: 6 iterations = 1536 vectors ahead, like the NPO
// kernel's 3 iterations of 512 threads; measured on the NPO kernel: 203 -> 191 us per 4096 rows, distance 2: 196, 6: 209;
// forward KL 146 -> 142 us; no gain on the one-operand CE kernel).
constexpr int kRowL2Ahead = 6;

struct RowLossParams {
  __nv_bfloat16* logits;
  const float2* lse_part;
  const void* ref_logits;  // bf16 or fp32 (template parameter kRefF32 of the kernels)
  const float2* ref_lse_part;
  const int64_t* labels;
  const float* row_weight;
  const float* scale_dev;
  float* loss_rows;
  int64_t ld, ref_ld;
  int vocab, n_vtiles;
  float grad_scale, beta;
  int write_grad;
  const int* m_dev;   // optional device-side row count (row compaction): rows >= *m_dev - m_off are not evaluated; those below the
  int m_off;          // next multiple of 64 get zero dlogits (they fall into the last K block of the dW contraction)
  int* status_flags;  // optional: bit 0 <- a CE label outside [0, vocab)
};

// Teacher logits arrive as bf16 (a bf16 checkpoint's lm_head output: the reference's case, SURVEY.md App. A dtype notes) or as
// fp32 (a teacher evaluated under autocast / an fp32 oracle).  One 8-logit vector = 16 or 32 bytes.
template <bool kF32>
struct RefVec {
  uint4 a, b;
};
template <>
struct RefVec<false> {   // bf16: exactly the one 128-bit register quadruple the kernels held before the fp32 variant existed
  uint4 a;
};
template <bool kF32>
__device__ __forceinline__ RefVec<kF32> ld_ref8(const void* row, int v) {
  RefVec<kF32> r;
  if constexpr (kF32) {
    r.a = reinterpret_cast<const uint4*>(row)[2 * v];
    r.b = reinterpret_cast<const uint4*>(row)[2 * v + 1];
  } else {
    r.a = reinterpret_cast<const uint4*>(row)[v];
  }
  return r;
}
template <bool kF32>
__device__ __forceinline__ RefVec<kF32> zero_ref8() {
  RefVec<kF32> r;
  r.a = make_uint4(0, 0, 0, 0);
  if constexpr (kF32) r.b = r.a;
  return r;
}
template <bool kF32>
__device__ __forceinline__ void prefetch_ref8(const void* row, int v) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const char*>(row) + (size_t)v * (kF32 ? 32 : 16)));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Every warp merges the row's per-tile (max, sum-exp) partials redundantly (a few hundred bytes from
// L2) — cheaper than a block barrier.  Returns the row's log-sum-exp in natural-log units.
__device__ __forceinline__ float merge_lse(const float2* part, int n) {
  const int lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, part[i].x);
  m = warp_max(m);
  float s = 0.f;
  for (int i = lane; i < n; i += 32) {
    const float2 p = part[i];
    s += p.y * exp2f((p.x - m) * kLog2e);
  }
  s = warp_sum(s);
  return m + logf(s);
}

// Student and teacher statistics in ONE round trip to L2: every lane fetches its (up to 4 + 4) partials before any
// arithmetic.  n <= 128 tiles (vocabularies up to 32768) take the register path; larger ones fall back to the loops.
__device__ __forceinline__ void merge_lse2(const float2* ps, const float2* pt, int n, float& lse_s, float& lse_t) {
  if (n > 128) {
    lse_s = merge_lse(ps, n);
    lse_t = merge_lse(pt, n);
    return;
  }
  const int lane = threadIdx.x & 31;
  float2 a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = lane + 32 * k;
    a[k] = (i < n) ? ps[i] : make_float2(-INFINITY, 0.f);
    b[k] = (i < n) ? pt[i] : make_float2(-INFINITY, 0.f);
  }
  const float ma = warp_max(fmaxf(fmaxf(a[0].x, a[1].x), fmaxf(a[2].x, a[3].x)));
  const float mb = warp_max(fmaxf(fmaxf(b[0].x, b[1].x), fmaxf(b[2].x, b[3].x)));
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (lane + 32 * k < n) {
      sa += a[k].y * exp2f((a[k].x - ma) * kLog2e);
      sb += b[k].y * exp2f((b[k].x - mb) * kLog2e);
    }
  }
  lse_s = ma + logf(warp_sum(sa));
  lse_t = mb + logf(warp_sum(sb));
}

// One operand, one round trip to L2 (n <= 128 tiles: every lane fetches its up to four partials before any arithmetic).
__device__ __forceinline__ float merge_lse1(const float2* ps, int n) {
  if (n > 128) return merge_lse(ps, n);
  const int lane = threadIdx.x & 31;
  float2 a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = lane + 32 * k;
    a[k] = (i < n) ? ps[i] : make_float2(-INFINITY, 0.f);
  }
  const float ma = warp_max(fmaxf(fmaxf(a[0].x, a[1].x), fmaxf(a[2].x, a[3].x)));
  float sa = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (lane + 32 * k < n) sa += a[k].y * exp2f((a[k].x - ma) * kLog2e);
  return ma + logf(warp_sum(sa));
}

template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // protects `red` against the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < kThreads / 32) ? red[lane] : 0.f;
  return warp_sum(t);  // every thread holds the total
}

// bf16 -> fp32 is a 16-bit shift: one SHF / LOP per element (the cuda_bf16 conversion emits PRMT + SHF)
__device__ __forceinline__ void unpack8(const uint4& raw, float (&x)[8]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[2 * j] = __uint_as_float(w[j] << 16);
    x[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
template <bool kF32>
__device__ __forceinline__ void unpack_ref8(const RefVec<kF32>& raw, float (&x)[8]) {
  if constexpr (kF32) {
    x[0] = __uint_as_float(raw.a.x); x[1] = __uint_as_float(raw.a.y); x[2] = __uint_as_float(raw.a.z); x[3] = __uint_as_float(raw.a.w);
    x[4] = __uint_as_float(raw.b.x); x[5] = __uint_as_float(raw.b.y); x[6] = __uint_as_float(raw.b.z); x[7] = __uint_as_float(raw.b.w);
  } else {
    unpack8(raw.a, x);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&x)[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
  return raw;
}

// two sums behind one pair of barriers
template <int kThreads>
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = make_float2(a, b);
  __syncthreads();
  const float2 t = (lane < kThreads / 32) ? red[lane] : make_float2(0.f, 0.f);
  return make_float2(warp_sum(t.x), warp_sum(t.y));
}

// ---------------------------------------------------------------------------------------------------
// CE: loss_row = lse - z[y]; dz = g (p - e_y).  Streaming, one pass.  (A2, A6c)
// ---------------------------------------------------------------------------------------------------
template <int kThreads>
__global__ void __launch_bounds__(kThreads) rowloss_ce_kernel(const RowLossParams a) {
  const int64_t row = blockIdx.x;
  __nv_bfloat16* z = a.logits + row * a.ld;
  if (a.m_dev) {  // row compaction: only the first *m_dev - m_off rows of this block exist
    const int m = *a.m_dev - a.m_off;
    if (row >= m) {
      if (a.write_grad && row < ((max(m, 0) + 63) & ~63)) {
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int v = threadIdx.x; v < (int)(a.ld >> 3); v += kThreads) reinterpret_cast<uint4*>(z)[v] = zero;
      }
      return;
    }
  }
  const int64_t y = a.labels[row];
  const bool valid = (y != kIgnoreIndex);
  if (valid && (y < 0 || y >= a.vocab)) {  // torch raises a device-side assert here; be as loud as a kernel can:
    if (threadIdx.x == 0) {                // NaN loss row (poisons the loss) and the caller's status flag
      a.loss_rows[row] = __int_as_float(0x7fc00000);
      if (a.status_flags) atomicOr(a.status_flags, 1);
    }
    if (a.write_grad) {
      const uint4 zero = make_uint4(0, 0, 0, 0);
      for (int v = threadIdx.x; v < (int)(a.ld >> 3); v += kThreads) reinterpret_cast<uint4*>(z)[v] = zero;
    }
    return;
  }
  const float w = a.row_weight ? a.row_weight[row] : 1.f;
  if (!a.write_grad) {
    if (threadIdx.x < 32) {
      float loss = 0.f;
      if (valid) loss = (merge_lse(a.lse_part + row * a.n_vtiles, a.n_vtiles) - __bfloat162float(z[y])) * w;
      if (threadIdx.x == 0) a.loss_rows[row] = loss;
    }
    return;
  }
  const int nvec = (int)(a.ld >> 3);
  if (!valid) {  // ignored row: dlogits = 0, nothing to read
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int v = threadIdx.x; v < nvec; v += kThreads) reinterpret_cast<uint4*>(z)[v] = zero;
    if (threadIdx.x == 0) a.loss_rows[row] = 0.f;
    return;
  }
  // the first two vectors of every thread are in flight while the row statistics are merged; one more is requested
  // per iteration (register double-buffering)
  uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
  if ((int)threadIdx.x < nvec) n0 = reinterpret_cast<const uint4*>(z)[threadIdx.x];
  if ((int)threadIdx.x + kThreads < nvec) n1 = reinterpret_cast<const uint4*>(z)[threadIdx.x + kThreads];
  const float lse = merge_lse(a.lse_part + row * a.n_vtiles, a.n_vtiles);
  const float g = a.grad_scale * (a.scale_dev ? *a.scale_dev : 1.f) * w;
  const float nl2 = -lse * kLog2e;
  const int yv = (int)(y >> 3), yj = (int)(y & 7);
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    const uint4 raw = n0;
    n0 = n1;
    if (v + 2 * kThreads < nvec) n1 = reinterpret_cast<const uint4*>(z)[v + 2 * kThreads];
    float x[8];
    unpack8(raw, x);
    if (v == yv) a.loss_rows[row] = (lse - x[yj]) * w;
    const int c0 = v << 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p = ptx::ex2_approx(fmaf(x[j], kLog2e, nl2));
      const float d = (v == yv && j == yj) ? (p - 1.f) : p;
      x[j] = (c0 + j < a.vocab) ? g * d : 0.f;
    }
    reinterpret_cast<uint4*>(z)[v] = pack8(x);
  }
}

// ---------------------------------------------------------------------------------------------------
// forward KL: loss_row = sum q (lq - lp); dz = g (p - q).  Streaming, one pass.  (A4)
// ---------------------------------------------------------------------------------------------------
template <int kThreads, bool kRefF32>
__global__ void __launch_bounds__(kThreads) rowloss_klfwd_kernel(const RowLossParams a) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  __nv_bfloat16* z = a.logits + row * a.ld;
  const char* zr = static_cast<const char*>(a.ref_logits) + row * a.ref_ld * (kRefF32 ? 4 : 2);
  const int nvec = (int)(a.ld >> 3);
  const int nfull = a.vocab >> 3;
  const int ref_nvec = (int)(a.ref_ld >> 3);  // a materialised teacher may be narrower than the staging buffer
  auto ld_ref = [&](int v) { return v < ref_nvec ? ld_ref8<kRefF32>(zr, v) : zero_ref8<kRefF32>(); };
  // first vector pair in flight while the row statistics are merged; register double-buffering afterwards
  uint4 nx = make_uint4(0, 0, 0, 0);
  RefVec<kRefF32> nr = zero_ref8<kRefF32>();
  if ((int)threadIdx.x < nvec) {
    nx = reinterpret_cast<const uint4*>(z)[threadIdx.x];
    nr = ld_ref(threadIdx.x);
  }
  float lse_s, lse_t;
  merge_lse2(a.lse_part + row * a.n_vtiles, a.ref_lse_part + row * a.n_vtiles, a.n_vtiles, lse_s, lse_t);
  const float nls2 = -lse_s * kLog2e, nlt2 = -lse_t * kLog2e, dlse = lse_s - lse_t;
  const float w = a.row_weight ? a.row_weight[row] : 1.f;
  const float g = a.grad_scale * (a.scale_dev ? *a.scale_dev : 1.f) * w;
  float acc = 0.f;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float x[8], r[8];
    unpack8(nx, x);
    unpack_ref8<kRefF32>(nr, r);
    if (v + kThreads < nvec) {
      nx = reinterpret_cast<const uint4*>(z)[v + kThreads];
      nr = ld_ref(v + kThreads);
    }
    if (v + kRowL2Ahead * kThreads < min(nvec, ref_nvec)) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(z) + v + kRowL2Ahead * kThreads));
      prefetch_ref8<kRefF32>(zr, v + kRowL2Ahead * kThreads);
    }
    if (v < nfull) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p = ptx::ex2_approx(fmaf(x[j], kLog2e, nls2)), q = ptx::ex2_approx(fmaf(r[j], kLog2e, nlt2));
        acc = fmaf(q, (r[j] - x[j]) + dlse, acc);  // q (lq - lp)
        x[j] = g * (p - q);
      }
    } else {
      const int c0 = v << 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool in = (c0 + j < a.vocab);
        const float p = ptx::ex2_approx(fmaf(x[j], kLog2e, nls2)), q = ptx::ex2_approx(fmaf(r[j], kLog2e, nlt2));
        if (in) acc = fmaf(q, (r[j] - x[j]) + dlse, acc);
        x[j] = in ? g * (p - q) : 0.f;
      }
    }
    if (a.write_grad) reinterpret_cast<uint4*>(z)[v] = pack8(x);
  }
  acc = block_sum<kThreads>(acc, red);
  if (threadIdx.x == 0) a.loss_rows[row] = acc * w;
}

// ---------------------------------------------------------------------------------------------------
// Two-pass kinds: a row statistic is needed before any gradient element can be formed.
//   KL_REV : K = sum p (lp - lq);            dz = g p ((lp - lq) - K)                  (A5)
//   NPO    : S = sum s, s = t/(1+t), t = (p/(q+1e-8))^beta; loss = sum log(1+t);
//            dz = g (s - p S)                                                          (A6)
//   ENTROPY: E = sum p lp;                   dz = g p (lp - E)                         (A6b)
// The second pass re-reads the row (it was just streamed by this CTA: an L2 hit).  NPO is bound by the XU
// (MUFU) pipe, not by HBM: 7 transcendental ops per logit over the two passes, so everything around them
// is kept to a minimum — exponent arguments are one FFMA (constants pre-scaled by log2 e), the loss is
// accumulated in log2 units, and the vocabulary bound is tested only in the row's last (ragged) vectors.
// ---------------------------------------------------------------------------------------------------
struct RowConsts {
  float nls2, nlt2;  // -lse * log2(e) of student / teacher
  float dlse;        // lse_t - lse_s
  float lse_s;
  float beta;
};

template <int kKind, bool kBetaOne>
struct RowTerm {
  // One logit: p and the auxiliary term pass 2 needs; adds to the row statistic and to the loss (log2 units for NPO).
  __device__ static __forceinline__ void eval(float x, float r, const RowConsts& c, float& p, float& aux, float& stat,
                                              float& loss) {
    p = ptx::ex2_approx(fmaf(x, kLog2e, c.nls2));
    if constexpr (kKind == CLEAR_KIND_KL_REV) {
      aux = (x - r) + c.dlse;  // lp - lq
      stat = fmaf(p, aux, stat);
    } else if constexpr (kKind == CLEAR_KIND_ENTROPY) {
      aux = x - c.lse_s;       // lp
      stat = fmaf(p, aux, stat);  // p == 0 gives 0 (the reference's fp32 softmax would give 0 * -inf = NaN; App. B #10)
    } else {  // NPO
      const float q = ptx::ex2_approx(fmaf(r, kLog2e, c.nlt2));
      const float A = q + 1e-8f;
      if constexpr (kBetaOne) {
        const float invB = ptx::rcp_approx(p + A);
        aux = p * invB;                          // s = t / (1 + t) = p / (p + q + eps)
        loss -= ptx::lg2_approx(A * invB);       // log2(1 + t) = -log2(1 - s)
      } else {
        const float t = ptx::ex2_approx(c.beta * ptx::lg2_approx(p * ptx::rcp_approx(A)));
        const float opt = 1.f + t;
        aux = t * ptx::rcp_approx(opt);
        loss += ptx::lg2_approx(opt);
      }
      stat += aux;
    }
  }
  // gscale = g (KL_REV, ENTROPY) ; for NPO the caller passes g and gS = g * S
  __device__ static __forceinline__ float grad(float p, float aux, float g, float gstat) {
    if constexpr (kKind == CLEAR_KIND_NPO) return fmaf(-p, gstat, g * aux);   // g (s - p S)
    else return (g * p) * (aux - gstat);                                      // gstat = K or E (unscaled)
  }
};

template <int kKind, bool kBetaOne, int kThreads, bool kRefF32>
__global__ void __launch_bounds__(kThreads) rowloss_twopass_kernel(const RowLossParams a) {
  __shared__ float red[32];
  constexpr bool kNeedsRef = (kKind != CLEAR_KIND_ENTROPY);
  const int64_t row = blockIdx.x;
  __nv_bfloat16* z = a.logits + row * a.ld;
  const char* zr = kNeedsRef ? static_cast<const char*>(a.ref_logits) + row * a.ref_ld * (kRefF32 ? 4 : 2) : nullptr;
  RowConsts c;
  float lse_t = 0.f;
  {  // warp 0 merges the student's partials, warp 1 the teacher's (see rowloss_npo1_kernel: these kernels are issue-bound too)
    __shared__ float row_lse[2];
    if (threadIdx.x < (kNeedsRef ? 64 : 32)) {
      const int wq = threadIdx.x >> 5;
      const float l = merge_lse1((wq ? a.ref_lse_part : a.lse_part) + row * a.n_vtiles, a.n_vtiles);
      if ((threadIdx.x & 31) == 0) row_lse[wq] = l;
    }
    __syncthreads();
    c.lse_s = row_lse[0];
    if constexpr (kNeedsRef) lse_t = row_lse[1];
  }
  c.nls2 = -c.lse_s * kLog2e;
  c.nlt2 = -lse_t * kLog2e;
  c.dlse = lse_t - c.lse_s;
  c.beta = a.beta;
  const float w = a.row_weight ? a.row_weight[row] : 1.f;
  const float g = a.grad_scale * (a.scale_dev ? *a.scale_dev : 1.f) * w;
  const int nvec = (int)(a.ld >> 3);
  const int nfull = a.vocab >> 3;  // vectors entirely inside the vocabulary (the teacher row covers them: ref_ld >= vocab)
  const int ref_nvec = kNeedsRef ? (int)(a.ref_ld >> 3) : 0;
  float stat = 0.f, loss = 0.f;
  for (int v = threadIdx.x; v < nfull; v += kThreads) {
    float x[8], r[8];
    unpack8(reinterpret_cast<const uint4*>(z)[v], x);
    if constexpr (kNeedsRef) unpack_ref8<kRefF32>(ld_ref8<kRefF32>(zr, v), r);
    if (v + 3 * kThreads < nfull) {  // later vectors of the row on their way into L2 (kThreads = 512: 1536 vectors ahead)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(z) + v + 3 * kThreads));
      if constexpr (kNeedsRef) prefetch_ref8<kRefF32>(zr, v + 3 * kThreads);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p, aux;
      RowTerm<kKind, kBetaOne>::eval(x[j], kNeedsRef ? r[j] : 0.f, c, p, aux, stat, loss);
    }
  }
  {  // ragged tail: at most one partially valid vector
    const int v = nfull + threadIdx.x;
    if (threadIdx.x == 0 && (a.vocab & 7)) {
      float x[8], r[8];
      unpack8(reinterpret_cast<const uint4*>(z)[v], x);
      if constexpr (kNeedsRef) unpack_ref8<kRefF32>(v < ref_nvec ? ld_ref8<kRefF32>(zr, v) : zero_ref8<kRefF32>(), r);
      for (int j = 0; j < (a.vocab & 7); ++j) {
        float p, aux;
        RowTerm<kKind, kBetaOne>::eval(x[j], kNeedsRef ? r[j] : 0.f, c, p, aux, stat, loss);
      }
    }
  }
  stat = block_sum<kThreads>(stat, red);
  if constexpr (kKind == CLEAR_KIND_NPO) loss = block_sum<kThreads>(loss, red) * kLn2;
  else loss = stat;
  if (threadIdx.x == 0) a.loss_rows[row] = loss * w;
  if (!a.write_grad) return;
  const float gstat = (kKind == CLEAR_KIND_NPO) ? g * stat : stat;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float x[8], r[8];
    if (v < nfull) {
      unpack8(reinterpret_cast<const uint4*>(z)[v], x);
      if constexpr (kNeedsRef) unpack_ref8<kRefF32>(ld_ref8<kRefF32>(zr, v), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p, aux, st = 0.f, ls = 0.f;
        RowTerm<kKind, kBetaOne>::eval(x[j], kNeedsRef ? r[j] : 0.f, c, p, aux, st, ls);
        x[j] = RowTerm<kKind, kBetaOne>::grad(p, aux, g, gstat);
      }
    } else {  // ragged / padding vectors
      unpack8(reinterpret_cast<const uint4*>(z)[v], x);
      if constexpr (kNeedsRef) unpack_ref8<kRefF32>(v < ref_nvec ? ld_ref8<kRefF32>(zr, v) : zero_ref8<kRefF32>(), r);
      const int c0 = v << 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p, aux, st = 0.f, ls = 0.f;
        RowTerm<kKind, kBetaOne>::eval(x[j], kNeedsRef ? r[j] : 0.f, c, p, aux, st, ls);
        x[j] = (c0 + j < a.vocab) ? RowTerm<kKind, kBetaOne>::grad(p, aux, g, gstat) : 0.f;
      }
    }
    reinterpret_cast<uint4*>(z)[v] = pack8(x);
  }
}

// ---------------------------------------------------------------------------------------------------
// NPO, beta = 1 (the shipped recipe: mm/trainer.py:343-353 with loss_beta = 1) — the headline row pass.
// The generic two-pass kernel above spends 7 MUFU operations per logit and is bound by the XU pipe (83 % busy
// in ncu) at a clock the tensor-heavy step has already pulled down.  This version needs 3.5:
//   * one reciprocal serves four logits (1 / (B0 B1 B2 B3) unfolded with six multiplies; B = p + q + 1e-8 lies
//     in [1e-8, 2], so the product of four stays a normal fp32 number);
//   * one logarithm serves four logits (sum_j log(1 + t_j) = -log2 prod_j (1 - s_j); each factor >= 5e-9);
//   * pass 1 parks s = t / (1 + t) in shared memory as fp32 (up to kNpoCacheVecs 8-logit vectors per row: two
//     CTAs of 96 KB per SM), so pass 2 only re-forms p = exp(z - lse) and combines: dz = g (s - p S).
// Vectors beyond the cache (vocabularies above 24576 keep 3/4 of the row cached) recompute s in pass 2.
// ---------------------------------------------------------------------------------------------------
constexpr int kNpoCacheVecs = 3072;  // 3072 x 8 fp32 = 96 KB
#ifndef CLEAR_NPO_L2_AHEAD
#define CLEAR_NPO_L2_AHEAD 3
#endif
constexpr int kL2Ahead = CLEAR_NPO_L2_AHEAD;

struct Npo8 {
  float p[8], s[8];
};

// p, s of eight logits; adds sum(s) to stat and log2 prod(1 - s) to lossl2 (negative of the loss in log2 units)
__device__ __forceinline__ void npo8_eval(const float (&x)[8], const float (&r)[8], float nls2, float nlt2, Npo8& o,
                                          float& stat, float& lossl2, bool want_loss) {
  float A[8], B[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    o.p[j] = ptx::ex2_approx(fmaf(x[j], kLog2e, nls2));
    A[j] = ptx::ex2_approx(fmaf(r[j], kLog2e, nlt2)) + 1e-8f;
    B[j] = o.p[j] + A[j];
  }
#pragma unroll
  for (int h = 0; h < 8; h += 4) {
    const float b01 = B[h] * B[h + 1], b23 = B[h + 2] * B[h + 3];
    const float inv = ptx::rcp_approx(b01 * b23);
    const float i01 = inv * b23, i23 = inv * b01;  // 1 / b01, 1 / b23
    const float i0 = i01 * B[h + 1], i1 = i01 * B[h], i2 = i23 * B[h + 3], i3 = i23 * B[h + 2];
    o.s[h] = o.p[h] * i0;
    o.s[h + 1] = o.p[h + 1] * i1;
    o.s[h + 2] = o.p[h + 2] * i2;
    o.s[h + 3] = o.p[h + 3] * i3;
    stat += (o.s[h] + o.s[h + 1]) + (o.s[h + 2] + o.s[h + 3]);
    // prod_j A_j / B_j = (prod_j A_j) * inv: four multiplies instead of seven (A_j >= 1e-8: the product of four stays normal)
    if (want_loss) lossl2 += ptx::lg2_approx(((A[h] * A[h + 1]) * (A[h + 2] * A[h + 3])) * inv);
  }
}

template <int kThreads, int kMinBlocks, bool kRefF32, int kPf2 = 1>
__global__ void __launch_bounds__(kThreads, kMinBlocks) rowloss_npo1_kernel(const RowLossParams a, const int cache_vecs) {
  extern __shared__ float4 s_cache[];  // [2][cache_vecs]: low / high half of each 8-logit vector (conflict-free float4 rows)
  __shared__ float2 red[32];
  // One CTA per row.  (Measured and rejected in round 2: a persistent grid of two CTAs per SM walking the rows, each pulling
  // its next row's first vectors and statistics into L2 while it finishes the current one — 209 us against 196 us per 4096
  // rows: the hardware's row-by-row CTA scheduling keeps the XU-heavy first pass of some rows beside the store-heavy second
  // pass of others, a lock-step grid does not.)
  const int64_t row = blockIdx.x;
  __nv_bfloat16* z = a.logits + row * a.ld;
  const char* zr = static_cast<const char*>(a.ref_logits) + row * a.ref_ld * (kRefF32 ? 4 : 2);
  const int nvec = (int)(a.ld >> 3);
  const int nfull = a.vocab >> 3;
  const int ref_nvec = (int)(a.ref_ld >> 3);
  const int ncache = a.write_grad ? min(cache_vecs, nfull) : 0;
  // register double-buffering: the next vector pair is in flight while this one is being evaluated (the first pair
  // while the row statistics are merged)
  uint4 nx = make_uint4(0, 0, 0, 0);
  RefVec<kRefF32> nr = zero_ref8<kRefF32>();
  if ((int)threadIdx.x < nfull) {
    nx = reinterpret_cast<const uint4*>(z)[threadIdx.x];
    nr = ld_ref8<kRefF32>(zr, threadIdx.x);
  }
  // The row's two log-sum-exps: warp 0 merges the student's partials, warp 1 the teacher's, everybody else waits at the
  // barrier with its first vectors in flight.  (Every warp used to merge both redundantly, "cheaper than a barrier" — but the
  // kernel is bound by instruction issue (73 % of the issue slots, 3 eligible warps per scheduler and cycle: ncu, round 2), and
  // the redundant merges were 17 % of all instructions executed.)
  __shared__ float row_lse[2];
  if (threadIdx.x < 64) {
    const int wq = threadIdx.x >> 5;
    const float l = merge_lse1((wq ? a.ref_lse_part : a.lse_part) + row * a.n_vtiles, a.n_vtiles);
    if ((threadIdx.x & 31) == 0) row_lse[wq] = l;
  }
  __syncthreads();
  const float lse_s = row_lse[0], lse_t = row_lse[1];
  const float nls2 = -lse_s * kLog2e, nlt2 = -lse_t * kLog2e;
  const float w = a.row_weight ? a.row_weight[row] : 1.f;
  const float g = a.grad_scale * (a.scale_dev ? *a.scale_dev : 1.f) * w;
  float stat = 0.f, lossl2 = 0.f;
  // (Measured and rejected, round 2: the same loop with two register sets that swap roles and running pointers — 11 % fewer
  // instructions, 5 % SLOWER: once the redundant merges were gone the kernel stopped being bound by instruction issue alone.)
  for (int v = threadIdx.x; v < nfull; v += kThreads) {
    const uint4 cx = nx;
    const RefVec<kRefF32> cr = nr;
    if (v + kThreads < nfull) {
      nx = reinterpret_cast<const uint4*>(z)[v + kThreads];
      nr = ld_ref8<kRefF32>(zr, v + kThreads);
    }
    if (kL2Ahead && v + kL2Ahead * kThreads < nfull) {  // pull the vectors of later iterations into L2 (no register cost)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const uint4*>(z) + v + kL2Ahead * kThreads));
      prefetch_ref8<kRefF32>(zr, v + kL2Ahead * kThreads);
    }
    float x[8], r[8];
    unpack8(cx, x);
    unpack_ref8<kRefF32>(cr, r);
    Npo8 o;
    npo8_eval(x, r, nls2, nlt2, o, stat, lossl2, true);
    if (v < ncache) {
      s_cache[v] = make_float4(o.s[0], o.s[1], o.s[2], o.s[3]);
      s_cache[cache_vecs + v] = make_float4(o.s[4], o.s[5], o.s[6], o.s[7]);
    }
  }
  if (threadIdx.x == 0 && (a.vocab & 7)) {  // ragged tail: one partially valid vector, scalar arithmetic
    RowConsts c;
    c.nls2 = nls2; c.nlt2 = nlt2; c.dlse = lse_t - lse_s; c.lse_s = lse_s; c.beta = 1.f;
    float x[8], r[8], l = 0.f;
    unpack8(reinterpret_cast<const uint4*>(z)[nfull], x);
    unpack_ref8<kRefF32>(nfull < ref_nvec ? ld_ref8<kRefF32>(zr, nfull) : zero_ref8<kRefF32>(), r);
    for (int j = 0; j < (a.vocab & 7); ++j) {
      float p, aux;
      RowTerm<CLEAR_KIND_NPO, true>::eval(x[j], r[j], c, p, aux, stat, l);
    }
    lossl2 -= l;
  }
  // pass 2 is light (one MUFU per logit) and reads from L2: three vectors per thread are kept in flight, the first
  // three across the barriers
  uint4 px[kPf2];
#pragma unroll
  for (int k = 0; k < kPf2; ++k) {
    px[k] = make_uint4(0, 0, 0, 0);
    if (a.write_grad && (int)threadIdx.x + k * kThreads < nvec) px[k] = reinterpret_cast<const uint4*>(z)[threadIdx.x + k * kThreads];
  }
  const float2 tot = block_sum2<kThreads>(stat, lossl2, red);
  stat = tot.x;
  lossl2 = tot.y;
  if (threadIdx.x == 0) a.loss_rows[row] = -lossl2 * kLn2 * w;
  if (!a.write_grad) return;
  const float gS = g * stat;
  for (int v = threadIdx.x; v < nvec; v += kThreads) {
    float x[8];
    unpack8(px[0], x);
#pragma unroll
    for (int k = 0; k + 1 < kPf2; ++k) px[k] = px[k + 1];
    if (v + kPf2 * kThreads < nvec) px[kPf2 - 1] = reinterpret_cast<const uint4*>(z)[v + kPf2 * kThreads];
    if (v < ncache) {
      const float4 lo = s_cache[v], hi = s_cache[cache_vecs + v];
      const float s[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(-ptx::ex2_approx(fmaf(x[j], kLog2e, nls2)), gS, g * s[j]);
    } else if (v < nfull) {
      float r[8], st = 0.f, ls = 0.f;
      unpack_ref8<kRefF32>(ld_ref8<kRefF32>(zr, v), r);
      Npo8 o;
      npo8_eval(x, r, nls2, nlt2, o, st, ls, false);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = fmaf(-o.p[j], gS, g * o.s[j]);
    } else {  // ragged / padding vectors
      RowConsts c;
      c.nls2 = nls2; c.nlt2 = nlt2; c.dlse = lse_t - lse_s; c.lse_s = lse_s; c.beta = 1.f;
      float r[8];
      unpack_ref8<kRefF32>(v < ref_nvec ? ld_ref8<kRefF32>(zr, v) : zero_ref8<kRefF32>(), r);
      const int c0 = v << 3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p, aux, st = 0.f, ls = 0.f;
        RowTerm<CLEAR_KIND_NPO, true>::eval(x[j], r[j], c, p, aux, st, ls);
        x[j] = (c0 + j < a.vocab) ? RowTerm<CLEAR_KIND_NPO, true>::grad(p, aux, g, gS) : 0.f;
      }
    }
    reinterpret_cast<uint4*>(z)[v] = pack8(x);
  }
}

// (max, sum-exp) partials of already-materialised logits (teacher logits handed over as a tensor), in
// the [row][256-column tile] layout the projection's epilogue emits.  One warp per tile: a lane holds
// 8 consecutive columns (one 128-bit load).
template <bool kF32>
__global__ void __launch_bounds__(256) lse_partials_kernel(const void* logits, int64_t ld, int vocab,
                                                           int n_vtiles, float2* part) {
  const int64_t row = blockIdx.x;
  const char* z = static_cast<const char*>(logits) + row * ld * (kF32 ? 4 : 2);
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x >> 5; t < n_vtiles; t += 8) {
    const int c0 = t * 256 + lane * 8;
    float x[8];
    if (c0 + 8 <= vocab) {
      unpack_ref8<kF32>(ld_ref8<kF32>(z, c0 >> 3), x);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float val = -INFINITY;
        if (c0 + j < vocab) val = kF32 ? reinterpret_cast<const float*>(z)[c0 + j] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(z)[c0 + j]);
        x[j] = val;
      }
    }
    float m = x[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) m = fmaxf(m, x[j]);
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += ptx::ex2_approx((x[j] - m) * kLog2e);
    s = warp_sum(s);
    if (lane == 0) part[row * n_vtiles + t] = make_float2(m, s);
  }
}

// fp32 [rows, cols] (dense) -> bf16 [rows, ld]: the conversion pass after a split-K dH contraction.
__global__ void __launch_bounds__(256) f32_to_bf16_rows_kernel(const float* src, __nv_bfloat16* dst, int cols, int64_t ld,
                                                               int64_t nvec, const int* m_dev, int m_off) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int vec_per_row = cols >> 3;
  if (m_dev) nvec = min(nvec, (int64_t)max(*m_dev - m_off, 0) * vec_per_row);
  if (i >= nvec) return;
  const int64_t row = i / vec_per_row;
  const int c = (int)(i - row * vec_per_row) << 3;
  const float4 a = *reinterpret_cast<const float4*>(src + row * cols + c);
  const float4 b = *reinterpret_cast<const float4*>(src + row * cols + c + 4);
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  *reinterpret_cast<uint4*>(dst + row * ld + c) = pack8(x);
}

// ---------------------------------------------------------------------------------------------------
// Row compaction for CE-type losses.  HF's causal-LM CE ignores rows whose (shifted) label is -100 — in CLEAR's
// multimodal batches that is the padding, the 576 image positions and the question, i.e. most rows — yet the reference
// projects every position and lets autograd push zero rows through the backward GEMMs.  Ignored rows contribute nothing
// to the loss, to dW, or to their own dH, so the head can run on the valid rows only: same loss, same gradients.
//   compact_rows:  ascending indices of the valid rows + their count (single CTA, block scan; N <= 2^24)
//   rows_gather:   dst[i, :] = src[idx[i], :]        rows_scatter: dst[idx[i], :] = src[i, :]     (row_bytes % 4 == 0)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) compact_rows_kernel(const int64_t* labels, int64_t n, int32_t* row_index, int32_t* count) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int valid = (i < n && labels[i] != kIgnoreIndex) ? 1 : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    const int before = __popc(ballot & ((1u << lane) - 1u));
    if (lane == 0) wsum[warp] = __popc(ballot);
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < warp; ++w) wb += wsum[w];
    if (valid) row_index[carry + wb + before] = (int32_t)i;
    __syncthreads();
    if (threadIdx.x == 1023) carry += wb + __popc(ballot);
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

// n_dev: optional device-side number of indices (<= n_idx).  Gather with pad_to > 0: destination rows [n, round_up(n, pad_to))
// (below n_idx) are zero-filled — they fall into the last K block of the dW contraction.  row_index NULL = identity.
template <bool kScatter>
__global__ void __launch_bounds__(256) rows_move_kernel(const uint8_t* src, int64_t src_ld_bytes, const int32_t* row_index,
                                                        int64_t n_idx, int row_bytes, uint8_t* dst, int64_t dst_ld_bytes,
                                                        const int* n_dev, int pad_to) {
  const int lanes_per_row = (row_bytes & 15) ? ((row_bytes + 3) >> 2) : (row_bytes >> 4);
  const int64_t n = n_dev ? min(n_idx, (int64_t)max(*n_dev, 0)) : n_idx;
  int64_t n_out = n;
  if (!kScatter && pad_to > 0) n_out = min(n_idx, (n + pad_to - 1) / pad_to * pad_to);
  const int64_t total = n_out * lanes_per_row;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / lanes_per_row;
    const int c = (int)(t - i * lanes_per_row);
    if (i >= n) {  // zero padding row (gather only)
      uint8_t* d = dst + i * dst_ld_bytes;
      if (row_bytes & 15) reinterpret_cast<uint32_t*>(d)[c] = 0u;
      else reinterpret_cast<uint4*>(d)[c] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const int64_t r = row_index ? (int64_t)row_index[i] : i;
    const uint8_t* s = src + (kScatter ? i : r) * src_ld_bytes;
    uint8_t* d = dst + (kScatter ? r : i) * dst_ld_bytes;
    if (row_bytes & 15) reinterpret_cast<uint32_t*>(d)[c] = reinterpret_cast<const uint32_t*>(s)[c];
    else reinterpret_cast<uint4*>(d)[c] = reinterpret_cast<const uint4*>(s)[c];
  }
}

// out[i] = outer[inner[i]] for i < *n_dev: the source row of compacted row i when the rows are also remapped (row_index)
__global__ void __launch_bounds__(256) compose_index_kernel(const int32_t* inner, const int32_t* outer, int64_t n_max,
                                                            const int* n_dev, int32_t* out) {
  const int64_t n = n_dev ? min(n_max, (int64_t)max(*n_dev, 0)) : n_max;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) out[i] = outer[inner[i]];
}

// fp32 rows [0, *m_dev - m_off) x cols <- 0 (the split-K scratch of a dH contraction over a device-side row count)
__global__ void __launch_bounds__(256) zero_rows_f32_kernel(float* dst, int cols, int64_t n_rows_max, const int* m_dev, int m_off) {
  const int64_t rows = m_dev ? min(n_rows_max, (int64_t)max(*m_dev - m_off, 0)) : n_rows_max;
  const int64_t nvec = rows * (cols >> 2);
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) reinterpret_cast<float4*>(dst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------
// All-gather half of the fused dW all-reduce: this rank's reduced slab -> every rank's buffer.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

struct PeerPtrs {
  __nv_bfloat16* p[8];
};

__global__ void __launch_bounds__(256) dweight_allgather_kernel(const __nv_bfloat16* local, __nv_bfloat16* mc, PeerPtrs peers,
                                                                int n_peers, int64_t ld, int row0, int vec_per_row,
                                                                int64_t nvec, float scale) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += stride) {
    const int64_t r = i / vec_per_row;
    const int64_t off = (row0 + r) * ld + ((i - r * vec_per_row) << 3);
    uint4 raw = *reinterpret_cast<const uint4*>(local + off);
    if (scale != 1.f) {
      float x[8];
      unpack8(raw, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= scale;
      raw = pack8(x);
    }
    if (mc) {
      multimem_st_v4(mc + off, raw);
    } else {
      for (int p = 0; p < n_peers; ++p) *reinterpret_cast<uint4*>(peers.p[p] + off) = raw;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// small reductions
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) count_valid_kernel(const int64_t* labels, int64_t n, float* inv_count) {
  __shared__ float red[32];
  float c = 0.f;  // exact in fp32 up to 2^24 rows per call
  for (int64_t i = threadIdx.x; i < n; i += 1024) c += (labels[i] != kIgnoreIndex) ? 1.f : 0.f;
  c = block_sum<1024>(c, red);
  if (threadIdx.x == 0) *inv_count = 1.f / c;
}

// Fixed-shape tree (thread-strided partials, warp shuffles, one smem stage): bitwise reproducible.
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const float* rows, int64_t n, float scale,
                                                           const float* scale_dev, float* out, int accumulate) {
  __shared__ float red[32];
  float c = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) c += rows[i];
  c = block_sum<1024>(c, red);
  if (threadIdx.x == 0) {
    const float v = c * scale * (scale_dev ? *scale_dev : 1.f);
    *out = accumulate ? (*out + v) : v;
  }
}

__global__ void __launch_bounds__(256) reduce_sequences_kernel(const float* rows, int64_t rows_per_seq, float* seq) {
  __shared__ float red[32];
  const float* r = rows + (int64_t)blockIdx.x * rows_per_seq;
  float c = 0.f;
  for (int64_t i = threadIdx.x; i < rows_per_seq; i += 256) c += r[i];
  c = block_sum<256>(c, red);
  if (threadIdx.x == 0) seq[blockIdx.x] = c;
}

}  // namespace
}  // namespace clear

using namespace clear;

extern "C" {

int clear_head_vtiles(int vocab) { return (vocab + 255) / 256; }

// diagnostic switch CLEAR_GEMM_GROUP=0: every contraction is its own launch (the round-1 schedule); read once
static int group_level() {
  static const int level = [] { const char* e = getenv("CLEAR_GEMM_GROUP"); return e ? atoi(e) : 2; }();
  return level;   // 0: never; 1: projection pair always, dH + dW only where dH alone starves its last wave (the round-2 default
                  // before the fence-free scheduler ring); 2 (default): dH + dW always
}
static bool group_launches() { return group_level() != 0; }

static int project_desc(const ClearProjectArgs* a, GemmDesc* g) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_head_project: args is NULL");
  if (!a->hidden || !a->weight || !a->logits) return fail(CLEAR_ERR_NULL, "clear_head_project: null tensor");
  if (a->n_rows <= 0 || a->n_rows > INT32_MAX || a->hidden_dim <= 0 || a->vocab <= 0)
    return fail(CLEAR_ERR_SHAPE, "clear_head_project: bad sizes n_rows=%lld hidden=%d vocab=%d", (long long)a->n_rows,
                a->hidden_dim, a->vocab);
  if (a->logits_ld % 64 || a->logits_ld < a->vocab)
    return fail(CLEAR_ERR_SHAPE, "clear_head_project: logits_ld (%lld) must be a multiple of 64 and >= vocab",
                (long long)a->logits_ld);
  g->a = GemmOperand{a->hidden, a->hidden_ld, kMajorK};
  g->b = GemmOperand{a->weight, a->weight_ld, kMajorK};
  g->d = a->logits;
  g->d_is_f32 = 0;
  g->ldd = a->logits_ld;
  g->m = (int)a->n_rows; g->n = a->vocab; g->k = a->hidden_dim;
  g->lse_part = a->lse_part;
  g->dyn_dim = a->n_rows_dev ? 1 : 0;
  return 0;
}

int clear_head_project(const ClearProjectArgs* a, void* stream) {
  GemmDesc g;
  if (int rc = project_desc(a, &g)) return rc;
  return launch_gemm_group(a->cta_group, g, nullptr, static_cast<cudaStream_t>(stream), 1, a->n_rows_dev, a->n_rows_dev_offset, 0);
}

int clear_head_project_pair(const ClearProjectArgs* student, const ClearProjectArgs* teacher, void* stream) {
  GemmDesc g0, g1;
  int rc;
  if ((rc = project_desc(student, &g0)) || (rc = project_desc(teacher, &g1))) return rc;
  if (!group_launches()) {
    if ((rc = clear_head_project(student, stream))) return rc;
    return clear_head_project(teacher, stream);
  }
  if ((student->lse_part == nullptr) != (teacher->lse_part == nullptr) || student->cta_group != teacher->cta_group ||
      student->n_rows_dev != teacher->n_rows_dev || student->n_rows_dev_offset != teacher->n_rows_dev_offset)
    return fail(CLEAR_ERR_KIND, "clear_head_project_pair: both projections must agree in lse_part presence, cta_group and the device-side row count");
  return launch_gemm_group(student->cta_group, g0, &g1, static_cast<cudaStream_t>(stream), 1, student->n_rows_dev,
                           student->n_rows_dev_offset, 0);
}

static int lse_partials_impl(const void* logits, bool f32, int64_t logits_ld, int64_t n_rows, int32_t vocab, float* lse_part,
                             void* stream) {
  if (!logits || !lse_part) return fail(CLEAR_ERR_NULL, "clear_head_lse_partials: null pointer");
  if (n_rows <= 0 || n_rows > INT32_MAX || vocab <= 0 || logits_ld < vocab || logits_ld % 8)
    return fail(CLEAR_ERR_SHAPE, "clear_head_lse_partials: bad sizes");
  if (reinterpret_cast<uintptr_t>(logits) & 15) return fail(CLEAR_ERR_ALIGN, "clear_head_lse_partials: logits must be 16-byte aligned");
  if (f32)
    lse_partials_kernel<true><<<(unsigned)n_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, logits_ld, vocab, clear_head_vtiles(vocab), reinterpret_cast<float2*>(lse_part));
  else
    lse_partials_kernel<false><<<(unsigned)n_rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, logits_ld, vocab, clear_head_vtiles(vocab), reinterpret_cast<float2*>(lse_part));
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_head_lse_partials(const void* logits, int64_t logits_ld, int64_t n_rows, int32_t vocab, float* lse_part,
                            void* stream) {
  return lse_partials_impl(logits, false, logits_ld, n_rows, vocab, lse_part, stream);
}

int clear_head_lse_partials_f32(const void* logits, int64_t logits_ld, int64_t n_rows, int32_t vocab, float* lse_part,
                                void* stream) {
  return lse_partials_impl(logits, true, logits_ld, n_rows, vocab, lse_part, stream);
}

int clear_head_rowloss(const ClearRowLossArgs* a, void* stream_) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_head_rowloss: args is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a->logits || !a->lse_part || !a->loss_rows) return fail(CLEAR_ERR_NULL, "clear_head_rowloss: null tensor");
  if (a->n_rows <= 0 || a->n_rows > INT32_MAX || a->vocab <= 0)
    return fail(CLEAR_ERR_SHAPE, "clear_head_rowloss: bad sizes");
  if (a->logits_ld % 8 || a->logits_ld < a->vocab) return fail(CLEAR_ERR_SHAPE, "clear_head_rowloss: bad logits_ld");
  const bool needs_ref = a->kind == CLEAR_KIND_KL_FWD || a->kind == CLEAR_KIND_KL_REV || a->kind == CLEAR_KIND_NPO;
  if (needs_ref) {
    if (!a->ref_logits || !a->ref_lse_part) return fail(CLEAR_ERR_NULL, "clear_head_rowloss: kind %d needs teacher logits", a->kind);
    if (a->ref_ld < a->vocab || a->ref_ld % 8) return fail(CLEAR_ERR_SHAPE, "clear_head_rowloss: ref_ld must be >= vocab and a multiple of 8");
    if (reinterpret_cast<uintptr_t>(a->ref_logits) & 15) return fail(CLEAR_ERR_ALIGN, "clear_head_rowloss: ref_logits must be 16-byte aligned");
  }
  if (a->kind == CLEAR_KIND_CE && !a->labels) return fail(CLEAR_ERR_NULL, "clear_head_rowloss: CE needs labels");
  if (a->kind == CLEAR_KIND_NPO && !(a->beta > 0.f)) return fail(CLEAR_ERR_SHAPE, "clear_head_rowloss: NPO beta must be > 0");

  RowLossParams p;
  p.logits = static_cast<__nv_bfloat16*>(a->logits);
  p.lse_part = reinterpret_cast<const float2*>(a->lse_part);
  p.ref_logits = a->ref_logits;
  p.ref_lse_part = reinterpret_cast<const float2*>(a->ref_lse_part);
  p.labels = a->labels;
  p.row_weight = a->row_weight;
  p.scale_dev = a->scale_dev;
  p.loss_rows = a->loss_rows;
  p.ld = a->logits_ld;
  p.ref_ld = a->ref_ld;
  p.vocab = a->vocab;
  p.n_vtiles = clear_head_vtiles(a->vocab);
  p.grad_scale = a->grad_scale;
  p.beta = a->beta;
  p.write_grad = a->write_grad;
  p.m_dev = a->n_rows_dev;
  p.m_off = a->n_rows_dev_offset;
  p.status_flags = a->status_flags;
  if (a->n_rows_dev && a->kind != CLEAR_KIND_CE)
    return fail(CLEAR_ERR_KIND, "clear_head_rowloss: a device-side row count (row compaction) exists for CE only");
  const bool rf32 = needs_ref && a->ref_logits_f32;
  const unsigned grid = (unsigned)a->n_rows;
  switch (a->kind) {
    case CLEAR_KIND_CE:
      rowloss_ce_kernel<256><<<grid, 256, 0, stream>>>(p);
      break;
    case CLEAR_KIND_KL_FWD:
      if (rf32) rowloss_klfwd_kernel<256, true><<<grid, 256, 0, stream>>>(p);
      else      rowloss_klfwd_kernel<256, false><<<grid, 256, 0, stream>>>(p);
      break;
    case CLEAR_KIND_KL_REV:
      if (rf32) rowloss_twopass_kernel<CLEAR_KIND_KL_REV, true, 512, true><<<grid, 512, 0, stream>>>(p);
      else      rowloss_twopass_kernel<CLEAR_KIND_KL_REV, true, 512, false><<<grid, 512, 0, stream>>>(p);
      break;
    case CLEAR_KIND_ENTROPY:
      rowloss_twopass_kernel<CLEAR_KIND_ENTROPY, true, 512, false><<<grid, 512, 0, stream>>>(p);
      break;
    case CLEAR_KIND_NPO:
      if (a->beta == 1.0f) {
        // 512 threads x 2 CTAs per SM with 96 KB of s-cache each measured best on B200 (tools/gpu_rowloss_ab.py, interleaved
        // A/B per 4096 rows: 203 us; 1024 x 1 with the whole row cached 241 us; 256 x 4 / 48 KB 211 us; 384 x 3 / 64 KB 208 us;
        // deeper register prefetch in either pass 207-217 us)
        constexpr int kSmem = kNpoCacheVecs * 32;
        const int nfull = a->vocab >> 3;
        const int cache_vecs = a->write_grad ? (nfull < kNpoCacheVecs ? nfull : kNpoCacheVecs) : 0;
        // the attribute is per device: set on every launch (a host-side table lookup, no device work)
        if (rf32) {
          CLEAR_CUDA_OK(cudaFuncSetAttribute(rowloss_npo1_kernel<512, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
          rowloss_npo1_kernel<512, 2, true><<<grid, 512, (size_t)cache_vecs * 32, stream>>>(p, cache_vecs);
        } else {
          CLEAR_CUDA_OK(cudaFuncSetAttribute(rowloss_npo1_kernel<512, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
          rowloss_npo1_kernel<512, 2, false><<<grid, 512, (size_t)cache_vecs * 32, stream>>>(p, cache_vecs);
        }
      } else {
        if (rf32) rowloss_twopass_kernel<CLEAR_KIND_NPO, false, 512, true><<<grid, 512, 0, stream>>>(p);
        else      rowloss_twopass_kernel<CLEAR_KIND_NPO, false, 512, false><<<grid, 512, 0, stream>>>(p);
      }
      break;
    default:
      return fail(CLEAR_ERR_KIND, "clear_head_rowloss: unknown kind %d", a->kind);
  }
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_head_compact_rows(const int64_t* labels, int64_t n_rows, int32_t* row_index, int32_t* count, void* stream) {
  if (!labels || !row_index || !count) return fail(CLEAR_ERR_NULL, "clear_head_compact_rows: null pointer");
  if (n_rows <= 0 || n_rows > (1 << 24)) return fail(CLEAR_ERR_SHAPE, "clear_head_compact_rows: n_rows out of range");
  compact_rows_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(labels, n_rows, row_index, count);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

static int rows_move(bool scatter, const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx,
                     int32_t row_bytes, void* dst, int64_t dst_ld_bytes, void* stream, const int32_t* n_idx_dev = nullptr,
                     int32_t pad_to = 0, bool identity_ok = false) {
  const char* who = scatter ? "clear_rows_scatter" : "clear_rows_gather";
  if (!src || !dst || (!row_index && !identity_ok)) return fail(CLEAR_ERR_NULL, "%s: null pointer", who);
  if (pad_to < 0) return fail(CLEAR_ERR_SHAPE, "%s: pad_to < 0", who);
  if (n_idx < 0 || row_bytes <= 0 || (row_bytes & 3) || src_ld_bytes < row_bytes || dst_ld_bytes < row_bytes)
    return fail(CLEAR_ERR_SHAPE, "%s: row_bytes must be a positive multiple of 4 within both strides", who);
  const int align = (row_bytes & 15) ? 3 : 15;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)src_ld_bytes | (uintptr_t)dst_ld_bytes) & align))
    return fail(CLEAR_ERR_ALIGN, "%s: pointers and strides must be %d-byte aligned", who, align + 1);
  if (n_idx == 0) return 0;
  const int lanes_per_row = (row_bytes & 15) ? ((row_bytes + 3) >> 2) : (row_bytes >> 4);
  int64_t blocks = (n_idx * lanes_per_row + 255) / 256;
  const int64_t cap = (int64_t)(device_sm_count() > 0 ? device_sm_count() : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (scatter)
    rows_move_kernel<true><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(src), src_ld_bytes, row_index, n_idx, row_bytes, static_cast<uint8_t*>(dst), dst_ld_bytes,
        n_idx_dev, 0);
  else
    rows_move_kernel<false><<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(src), src_ld_bytes, row_index, n_idx, row_bytes, static_cast<uint8_t*>(dst), dst_ld_bytes,
        n_idx_dev, pad_to);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_rows_gather_dev(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx_max,
                          const int32_t* n_idx_dev, int32_t pad_to, int32_t row_bytes, void* dst, int64_t dst_ld_bytes, void* stream) {
  return rows_move(false, src, src_ld_bytes, row_index, n_idx_max, row_bytes, dst, dst_ld_bytes, stream, n_idx_dev, pad_to, true);
}

int clear_rows_scatter_dev(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx_max,
                           const int32_t* n_idx_dev, int32_t row_bytes, void* dst, int64_t dst_ld_bytes, void* stream) {
  return rows_move(true, src, src_ld_bytes, row_index, n_idx_max, row_bytes, dst, dst_ld_bytes, stream, n_idx_dev, 0, false);
}

int clear_index_compose(const int32_t* inner, const int32_t* outer, int64_t n_max, const int32_t* n_dev, int32_t* out, void* stream) {
  if (!inner || !outer || !out) return fail(CLEAR_ERR_NULL, "clear_index_compose: null pointer");
  if (n_max <= 0) return fail(CLEAR_ERR_SHAPE, "clear_index_compose: n_max <= 0");
  int64_t blocks = (n_max + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  compose_index_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(inner, outer, n_max, n_dev, out);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_rows_gather(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx, int32_t row_bytes,
                      void* dst, int64_t dst_ld_bytes, void* stream) {
  return rows_move(false, src, src_ld_bytes, row_index, n_idx, row_bytes, dst, dst_ld_bytes, stream);
}

int clear_rows_scatter(const void* src, int64_t src_ld_bytes, const int32_t* row_index, int64_t n_idx, int32_t row_bytes,
                       void* dst, int64_t dst_ld_bytes, void* stream) {
  return rows_move(true, src, src_ld_bytes, row_index, n_idx, row_bytes, dst, dst_ld_bytes, stream);
}

int clear_head_count_valid(const int64_t* labels, int64_t n_rows, float* inv_count, void* stream) {
  if (!labels || !inv_count) return fail(CLEAR_ERR_NULL, "clear_head_count_valid: null pointer");
  if (n_rows <= 0 || n_rows > (1 << 24)) return fail(CLEAR_ERR_SHAPE, "clear_head_count_valid: n_rows out of range");
  count_valid_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(labels, n_rows, inv_count);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_head_reduce_rows(const float* loss_rows, int64_t n_rows, float scale, const float* scale_dev, float* out,
                           int accumulate, void* stream) {
  if (!loss_rows || !out) return fail(CLEAR_ERR_NULL, "clear_head_reduce_rows: null pointer");
  if (n_rows <= 0) return fail(CLEAR_ERR_SHAPE, "clear_head_reduce_rows: n_rows <= 0");
  reduce_rows_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(loss_rows, n_rows, scale, scale_dev, out, accumulate);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_head_reduce_sequences(const float* loss_rows, int64_t n_seqs, int64_t rows_per_seq, float* seq_sums,
                                void* stream) {
  if (!loss_rows || !seq_sums) return fail(CLEAR_ERR_NULL, "clear_head_reduce_sequences: null pointer");
  if (n_seqs <= 0 || rows_per_seq <= 0 || n_seqs > INT32_MAX) return fail(CLEAR_ERR_SHAPE, "clear_head_reduce_sequences: bad sizes");
  reduce_sequences_kernel<<<(unsigned)n_seqs, 256, 0, static_cast<cudaStream_t>(stream)>>>(loss_rows, rows_per_seq, seq_sums);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_head_backward(const ClearBackwardArgs* a, void* stream_) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_head_backward: args is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a->dlogits) return fail(CLEAR_ERR_NULL, "clear_head_backward: dlogits is NULL");
  if (a->n_rows <= 0 || a->n_rows > INT32_MAX || a->hidden_dim <= 0 || a->vocab <= 0)
    return fail(CLEAR_ERR_SHAPE, "clear_head_backward: bad sizes");
  if (a->hidden_dim % 8) return fail(CLEAR_ERR_SHAPE, "clear_head_backward: hidden_dim must be a multiple of 8");
  int rc;
  if (a->n_peers < 0 || a->n_peers > 8) return fail(CLEAR_ERR_SHAPE, "clear_head_backward: 0 <= n_peers <= 8");
  if (a->n_peers > 0 && a->dweight_f32) return fail(CLEAR_ERR_KIND, "clear_head_backward: the peer reduce-scatter accumulates in bf16");
  const bool want_dw = a->dweight || a->n_peers > 0;
  // Both gradients, both bf16: ONE launch — the 501-K-block tiles of dH = dZ W come first, the shorter tiles of dW = dZ^T H
  // fill the SM pairs that finish early, so the two share one pipeline ramp and one tail (dH alone: 6.92 of 7 waves, dW alone:
  // 27.2 of 28) and dH needs no split-K scratch (no zeroing, no conversion pass) when its row count is small (small staging
  // blocks, compacted rows).  At full waves (8192 rows) one launch measured 1.5 % SLOWER than two while every tile boundary
  // cost the MMA thread a fence (round 2, first half); with the fence-free scheduler ring it measures 1.3 % of the NPO step
  // faster (4 of 4 alternating pairs of processes on one box, tools/r02_group_ab2.sh), so it is the default everywhere.
  // CLEAR_GEMM_GROUP=1 restores the old rule, 0 separates everything.
  const bool dh_starved = a->n_rows_dev != nullptr ||
                          gemm_pick_k_splits(a->cta_group, (int)a->n_rows, a->hidden_dim, a->vocab) > 1;
  if (a->dhidden && want_dw && !a->dweight_f32 && (dh_starved || group_level() >= 2) && group_launches()) {
    if (!a->weight || !a->hidden) return fail(CLEAR_ERR_NULL, "clear_head_backward: weight and hidden are required");
    GemmDesc gh, gw;
    gh.a = GemmOperand{a->dlogits, a->dlogits_ld, kMajorK};
    gh.b = GemmOperand{a->weight, a->weight_ld, kMajorMN};
    gh.d = a->dhidden; gh.ldd = a->dhidden_ld;
    gh.m = (int)a->n_rows; gh.n = a->hidden_dim; gh.k = a->vocab;
    gh.dyn_dim = a->n_rows_dev ? 1 : 0;
    gw.a = GemmOperand{a->dlogits, a->dlogits_ld, kMajorMN};
    gw.b = GemmOperand{a->hidden, a->hidden_ld, kMajorMN};
    gw.d = a->dweight ? a->dweight : a->dweight_peers[0]; gw.ldd = a->dweight_ld;
    gw.m = a->vocab; gw.n = a->hidden_dim; gw.k = (int)a->n_rows;
    gw.reduce_add = a->accumulate;
    gw.peer_d = a->n_peers > 0 ? a->dweight_peers : nullptr;
    gw.n_peers = a->n_peers;
    gw.dyn_dim = a->n_rows_dev ? 2 : 0;
    return launch_gemm_group(a->cta_group, gh, &gw, stream, 1, a->n_rows_dev, a->n_rows_dev_offset, 0);
  }
  if (a->dhidden) {
    if (!a->weight) return fail(CLEAR_ERR_NULL, "clear_head_backward: dhidden requested but weight is NULL");
    // dH[n, h] = sum_v dZ[n, v] W[v, h]: A = dZ (K-major, K = vocab), B = W read MN-major.
    GemmOperand A{a->dlogits, a->dlogits_ld, kMajorK};
    GemmOperand B{a->weight, a->weight_ld, kMajorMN};
    GemmDyn dyn;
    dyn.count = a->n_rows_dev;
    dyn.offset = a->n_rows_dev_offset;
    dyn.dim = 1;
    if (a->n_rows_dev && a->dhidden_ws) {
      // Row count known on the device only (row compaction): the number of output tiles, hence the K split that fills the
      // SMs, is chosen inside the kernel; the fp32 scratch is zeroed / converted for the live rows only.
      dyn.splits = 1;
      const int64_t nvec4 = (int64_t)a->n_rows * (a->hidden_dim / 4);
      zero_rows_f32_kernel<<<(unsigned)std::min<int64_t>((nvec4 + 255) / 256, 148 * 8), 256, 0, stream>>>(
          a->dhidden_ws, a->hidden_dim, a->n_rows, a->n_rows_dev, a->n_rows_dev_offset);
      rc = launch_gemm_bf16(a->cta_group, A, B, a->dhidden_ws, 1, a->hidden_dim, (int)a->n_rows, a->hidden_dim, a->vocab,
                            nullptr, /*reduce_add=*/1, /*raster: traffic model*/ -1, stream, 1, nullptr, 0, dyn);
      if (rc) return rc;
      const int64_t nvec = (int64_t)a->n_rows * (a->hidden_dim / 8);
      f32_to_bf16_rows_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, stream>>>(
          a->dhidden_ws, static_cast<__nv_bfloat16*>(a->dhidden), a->hidden_dim, a->dhidden_ld, nvec, a->n_rows_dev,
          a->n_rows_dev_offset);
      CLEAR_CUDA_OK(cudaGetLastError());
    } else {
      const int splits = (a->dhidden_ws && !a->n_rows_dev) ? gemm_pick_k_splits(a->cta_group, (int)a->n_rows, a->hidden_dim, a->vocab) : 1;
      if (splits > 1) {
        const size_t ws_elems = (size_t)a->n_rows * a->hidden_dim;
        CLEAR_CUDA_OK(cudaMemsetAsync(a->dhidden_ws, 0, ws_elems * sizeof(float), stream));
        rc = launch_gemm_bf16(a->cta_group, A, B, a->dhidden_ws, 1, a->hidden_dim, (int)a->n_rows, a->hidden_dim, a->vocab,
                              nullptr, /*reduce_add=*/1, /*raster: traffic model*/ -1, stream, splits);
        if (rc) return rc;
        const int64_t nvec = (int64_t)a->n_rows * (a->hidden_dim / 8);
        f32_to_bf16_rows_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, stream>>>(
            a->dhidden_ws, static_cast<__nv_bfloat16*>(a->dhidden), a->hidden_dim, a->dhidden_ld, nvec, nullptr, 0);
        CLEAR_CUDA_OK(cudaGetLastError());
      } else {
        rc = launch_gemm_bf16(a->cta_group, A, B, a->dhidden, 0, a->dhidden_ld, (int)a->n_rows, a->hidden_dim, a->vocab,
                              nullptr, 0, /*raster: traffic model*/ -1, stream, 1, nullptr, 0, dyn);
        if (rc) return rc;
      }
    }
  }
  if (a->dweight || a->n_peers > 0) {
    if (!a->hidden) return fail(CLEAR_ERR_NULL, "clear_head_backward: dweight requested but hidden is NULL");
    // dW[v, h] = sum_n dZ[n, v] H[n, h]: both operands read MN-major (K = rows).
    GemmOperand A{a->dlogits, a->dlogits_ld, kMajorMN};
    GemmOperand B{a->hidden, a->hidden_ld, kMajorMN};
    void* d = a->dweight ? a->dweight : a->dweight_peers[0];
    GemmDyn dyn;   // K = rows: with a device-side count, the rows of dlogits / hidden up to the next multiple of 64 are zero
    dyn.count = a->n_rows_dev;
    dyn.offset = a->n_rows_dev_offset;
    dyn.dim = 2;
    rc = launch_gemm_bf16(a->cta_group, A, B, d, a->dweight_f32, a->dweight_ld, a->vocab, a->hidden_dim,
                          (int)a->n_rows, nullptr, a->accumulate, /*raster: traffic model*/ -1, stream, 1,
                          a->n_peers > 0 ? a->dweight_peers : nullptr, a->n_peers, dyn);
    if (rc) return rc;
  }
  return 0;
}

int clear_head_dweight_slab(int32_t vocab, int32_t n_peers, int32_t rank, int32_t cta_group, int32_t* row0, int32_t* row1) {
  if (!row0 || !row1) return fail(CLEAR_ERR_NULL, "clear_head_dweight_slab: null pointer");
  if (vocab <= 0 || n_peers <= 0 || n_peers > 8 || rank < 0 || rank >= n_peers) return fail(CLEAR_ERR_SHAPE, "clear_head_dweight_slab: bad sizes");
  int r0, r1;
  gemm_peer_rows(cta_group, vocab, n_peers, rank, &r0, &r1);
  *row0 = r0;
  *row1 = r1;
  return 0;
}

int clear_dweight_allgather(const void* local_buf, void* multicast_buf, void* const* peer_bufs, int32_t n_peers, int64_t ld,
                            int32_t row0, int32_t row1, int32_t hidden_dim, float scale, void* stream) {
  if (!local_buf || (!multicast_buf && !peer_bufs)) return fail(CLEAR_ERR_NULL, "clear_dweight_allgather: null pointer");
  if (n_peers <= 0 || n_peers > 8 || row0 < 0 || row1 < row0 || hidden_dim <= 0 || (hidden_dim & 7) || (ld & 7) || ld < hidden_dim)
    return fail(CLEAR_ERR_SHAPE, "clear_dweight_allgather: bad sizes");
  if ((reinterpret_cast<uintptr_t>(local_buf) | reinterpret_cast<uintptr_t>(multicast_buf)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_dweight_allgather: buffers must be 16-byte aligned");
  if (row1 == row0) return 0;
  PeerPtrs pp = {};
  if (!multicast_buf)
    for (int p = 0; p < n_peers; ++p) {
      if (!peer_bufs[p] || (reinterpret_cast<uintptr_t>(peer_bufs[p]) & 15)) return fail(CLEAR_ERR_ALIGN, "clear_dweight_allgather: peer buffer %d", p);
      pp.p[p] = static_cast<__nv_bfloat16*>(peer_bufs[p]);
    }
  const int vec_per_row = hidden_dim >> 3;
  const int64_t nvec = (int64_t)(row1 - row0) * vec_per_row;
  // a few CTAs are enough to saturate NVLink; keep the SM footprint small so the overlapped compute is not disturbed
  int64_t blocks = (nvec + 255) / 256;
  if (blocks > 32) blocks = 32;
  dweight_allgather_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(local_buf), static_cast<__nv_bfloat16*>(multicast_buf), pp, n_peers, ld, row0, vec_per_row,
      nvec, scale);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
