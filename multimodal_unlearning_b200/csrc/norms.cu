// Whole-model reductions around the update step (HBM-bound, one pass each, bitwise reproducible):
//   * sum of squares of the (masked) gradients  -> global gradient-norm clipping.  The reference trains through
//     HF Trainer, whose default max_grad_norm = 1.0 applies (`max_grad_norm` is commented out at mm/forget.py:151),
//     i.e. torch.nn.utils.clip_grad_norm_ over all parameters before optimizer.step();
//   * L1 norm and L0 "norm" of the parameters, or of (parameters - teacher parameters): the weight regulariser of
//     mm/trainer.py:451-481 / nlp/dataloader.py:461-483, which the reference evaluates on a torch.cat of all
//     parameters (a 14 GB temporary at 7 B).  norm(cat(x_t)) = sum_t norm(x_t) for p = 0, 1, so the reduction runs
//     tensor by tensor (or over a flat buffer) and accumulates into one device scalar.
// Sums are carried in double: per 8-element vector in fp32, then one DADD; block partials are combined by a second
// kernel in a fixed order.
#include <cuda_bf16.h>

#include "common.h"

namespace clear {
namespace {

constexpr int kRedThreads = 256;

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = (lane < kRedThreads / 32) ? red[lane] : 0.0;
  return warp_sum_d(t);
}

// partials[blockIdx.x] = sum over this block's vectors of (g * mask)^2.  Four vectors (and their mask words) per thread
// are in flight before any arithmetic.
constexpr int kSqUnroll = 4;

__device__ __forceinline__ uint32_t half_mask2(uint32_t two_bits) {  // bit 0 -> low half, bit 1 -> high half
  return ((two_bits & 1u) ? 0x0000FFFFu : 0u) | ((two_bits & 2u) ? 0xFFFF0000u : 0u);
}

__global__ void __launch_bounds__(kRedThreads) sqnorm_kernel(const __nv_bfloat16* g, const int8_t* mask,
                                                             const uint8_t* mask_bits, int64_t n, double* partials) {
  __shared__ double red[32];
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * kRedThreads;
  double acc = 0.0;
  for (int64_t i0 = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i0 < nvec; i0 += stride * kSqUnroll) {
    uint4 raw[kSqUnroll];
    uint2 mk[kSqUnroll];
    uint32_t mb[kSqUnroll];
#pragma unroll
    for (int u = 0; u < kSqUnroll; ++u) {
      const int64_t i = i0 + u * stride;
      raw[u] = make_uint4(0, 0, 0, 0);
      mk[u] = make_uint2(0x01010101u, 0x01010101u);
      mb[u] = 0xFFu;
      if (i < nvec) {
        raw[u] = ld_stream16(reinterpret_cast<const uint4*>(g) + i);
        if (mask) mk[u] = *reinterpret_cast<const uint2*>(mask + (i << 3));
        if (mask_bits) mb[u] = mask_bits[i];
      }
    }
#pragma unroll
    for (int u = 0; u < kSqUnroll; ++u) {
      uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
      if (mask) {  // int8 mask: any non-zero byte selects
        const uint32_t lo = __vcmpne4(mk[u].x, 0u), hi = __vcmpne4(mk[u].y, 0u);  // 0xFF per selected element
        w[0] &= __byte_perm(lo, 0, 0x1100);
        w[1] &= __byte_perm(lo, 0, 0x3322);
        w[2] &= __byte_perm(hi, 0, 0x1100);
        w[3] &= __byte_perm(hi, 0, 0x3322);
      } else if (mask_bits) {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] &= half_mask2(mb[u] >> (2 * q));
      }
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = __uint_as_float(w[q] << 16), b = __uint_as_float(w[q] & 0xffff0000u);
        s = fmaf(a, a, s);
        s = fmaf(b, b, s);
      }
      acc += (double)s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec << 3; i < n; ++i) {
      bool on = !mask || mask[i];
      if (mask_bits) on = (mask_bits[i >> 3] >> (i & 7)) & 1u;
      const float x = on ? __bfloat162float(g[i]) : 0.f;
      acc += (double)(x * x);
    }
  }
  acc = block_sum_d(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// partials[2 b] = sum |p - r|, partials[2 b + 1] = #{p != r}; r = 0 when ref is NULL.  The difference is formed like
// torch forms `param - teacher_param` on bf16 tensors: in fp32, rounded to bf16.
__global__ void __launch_bounds__(kRedThreads) l1l0_kernel(const __nv_bfloat16* p, const __nv_bfloat16* ref, int64_t n,
                                                           double* partials) {
  __shared__ double red[32];
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * kRedThreads;
  double l1 = 0.0, l0 = 0.0;
  auto one = [&](__nv_bfloat16 a, __nv_bfloat16 b, float& s, int& c) {
    float d = __bfloat162float(a);
    if (ref) d = __bfloat162float(__float2bfloat16_rn(d - __bfloat162float(b)));
    s += fabsf(d);
    c += (d != 0.f);
  };
  for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < nvec; i += stride) {
    const uint4 ra = ld_stream16(reinterpret_cast<const uint4*>(p) + i);
    uint4 rb = make_uint4(0, 0, 0, 0);
    if (ref) rb = ld_stream16(reinterpret_cast<const uint4*>(ref) + i);
    const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&ra);
    const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&rb);
    float s = 0.f;
    int c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) one(ha[j], hb[j], s, c);
    l1 += (double)s;
    l0 += (double)c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = nvec << 3; i < n; ++i) {
      float s = 0.f;
      int c = 0;
      one(p[i], ref ? ref[i] : __float2bfloat16_rn(0.f), s, c);
      l1 += (double)s;
      l0 += (double)c;
    }
  }
  l1 = block_sum_d(l1, red);
  l0 = block_sum_d(l0, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = l1;
    partials[2 * blockIdx.x + 1] = l0;
  }
}

// out = bf16(coef * sign(p - r)): the gradient of coef * ||p - r||_1 (torch: sign(0) = 0).  coef is a device scalar
// (lambda_1 times the upstream gradient of the loss), so the backward needs no host synchronisation.
__global__ void __launch_bounds__(kRedThreads) l1_grad_kernel(const __nv_bfloat16* p, const __nv_bfloat16* ref, int64_t n,
                                                              const float* coef, __nv_bfloat16* out) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * kRedThreads;
  const float c = *coef;
  const __nv_bfloat16 pos = __float2bfloat16_rn(c), neg = __float2bfloat16_rn(-c), zero = __float2bfloat16_rn(0.f);
  auto one = [&](__nv_bfloat16 a, __nv_bfloat16 b) {
    float d = __bfloat162float(a);
    if (ref) d = __bfloat162float(__float2bfloat16_rn(d - __bfloat162float(b)));
    return d > 0.f ? pos : (d < 0.f ? neg : zero);
  };
  for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < nvec; i += stride) {
    const uint4 ra = ld_stream16(reinterpret_cast<const uint4*>(p) + i);
    uint4 rb = make_uint4(0, 0, 0, 0);
    if (ref) rb = ld_stream16(reinterpret_cast<const uint4*>(ref) + i);
    const __nv_bfloat16* ha = reinterpret_cast<const __nv_bfloat16*>(&ra);
    const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&rb);
    uint4 o;
    __nv_bfloat16* ho = reinterpret_cast<__nv_bfloat16*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) ho[j] = one(ha[j], hb[j]);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nvec << 3; i < n; ++i) out[i] = one(p[i], ref ? ref[i] : zero);
}

// g = bf16(g * coef) in place: torch's `grad.mul_(clip_coef)` on a bf16 gradient (fp32 product, one rounding)
__global__ void __launch_bounds__(kRedThreads) scale_kernel(__nv_bfloat16* g, int64_t n, const float* coef) {
  const int64_t nvec = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * kRedThreads;
  const float c = *coef;
  for (int64_t i = (int64_t)blockIdx.x * kRedThreads + threadIdx.x; i < nvec; i += stride) {
    uint4 raw = reinterpret_cast<const uint4*>(g)[i];
    __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __float2bfloat16_rn(__fmul_rn(__bfloat162float(h[j]), c));
    reinterpret_cast<uint4*>(g)[i] = raw;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = nvec << 3; i < n; ++i) g[i] = __float2bfloat16_rn(__fmul_rn(__bfloat162float(g[i]), c));
}

// out[f] (+)= sum_b partials[b * n_fields + f], one thread per field walking the blocks in order
__global__ void finish_kernel(const double* partials, int n_blocks, int n_fields, double* out, int accumulate) {
  const int f = threadIdx.x;
  if (f >= n_fields) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partials[b * n_fields + f];
  out[f] = accumulate ? out[f] + s : s;
}

__global__ void clip_coef_kernel(const double* sumsq, double max_norm, float* coef) {
  // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
  const double c = max_norm / (sqrt(*sumsq) + 1e-6);
  *coef = (float)(c < 1.0 ? c : 1.0);
}

int red_blocks(int64_t n) {
  const int sms = device_sm_count();
  int64_t b = ((n >> 3) + kRedThreads - 1) / kRedThreads;
  const int64_t cap = (int64_t)(sms > 0 ? sms : 148) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace
}  // namespace clear

using namespace clear;

extern "C" {

int64_t clear_reduce_partials_elems(void) { return (int64_t)(device_sm_count() > 0 ? device_sm_count() : 148) * 8 * 2; }

int clear_grad_sqnorm(const void* grad, const int8_t* mask, const uint8_t* mask_bits, int64_t n, double* partials,
                      double* out, int accumulate, void* stream_) {
  if (!grad || !partials || !out) return fail(CLEAR_ERR_NULL, "clear_grad_sqnorm: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_grad_sqnorm: n <= 0");
  if (reinterpret_cast<uintptr_t>(grad) & 15) return fail(CLEAR_ERR_ALIGN, "clear_grad_sqnorm: grad must be 16-byte aligned");
  if (mask && (reinterpret_cast<uintptr_t>(mask) & 7)) return fail(CLEAR_ERR_ALIGN, "clear_grad_sqnorm: mask must be 8-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int blocks = red_blocks(n);
  if (mask && mask_bits) return fail(CLEAR_ERR_KIND, "clear_grad_sqnorm: pass mask or mask_bits, not both");
  sqnorm_kernel<<<blocks, kRedThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(grad), mask, mask_bits, n, partials);
  finish_kernel<<<1, 32, 0, stream>>>(partials, blocks, 1, out, accumulate);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_param_l1l0(const void* param, const void* ref, int64_t n, double* partials, double* out2, int accumulate,
                     void* stream_) {
  if (!param || !partials || !out2) return fail(CLEAR_ERR_NULL, "clear_param_l1l0: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_param_l1l0: n <= 0");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(ref)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_param_l1l0: tensors must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int blocks = red_blocks(n);
  l1l0_kernel<<<blocks, kRedThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(param),
                                                  static_cast<const __nv_bfloat16*>(ref), n, partials);
  finish_kernel<<<1, 32, 0, stream>>>(partials, blocks, 2, out2, accumulate);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_param_l1_grad(const void* param, const void* ref, int64_t n, const float* coef, void* out, void* stream_) {
  if (!param || !coef || !out) return fail(CLEAR_ERR_NULL, "clear_param_l1_grad: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_param_l1_grad: n <= 0");
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_param_l1_grad: tensors must be 16-byte aligned");
  l1_grad_kernel<<<red_blocks(n), kRedThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(param), static_cast<const __nv_bfloat16*>(ref), n, coef,
      static_cast<__nv_bfloat16*>(out));
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_grad_scale(void* grad, int64_t n, const float* coef, void* stream_) {
  if (!grad || !coef) return fail(CLEAR_ERR_NULL, "clear_grad_scale: null pointer");
  if (n <= 0) return fail(CLEAR_ERR_SHAPE, "clear_grad_scale: n <= 0");
  if (reinterpret_cast<uintptr_t>(grad) & 15) return fail(CLEAR_ERR_ALIGN, "clear_grad_scale: grad must be 16-byte aligned");
  scale_kernel<<<red_blocks(n), kRedThreads, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<__nv_bfloat16*>(grad), n, coef);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

int clear_clip_coef(const double* sumsq, double max_norm, float* coef, void* stream) {
  if (!sumsq || !coef) return fail(CLEAR_ERR_NULL, "clear_clip_coef: null pointer");
  if (!(max_norm > 0.0)) return fail(CLEAR_ERR_SHAPE, "clear_clip_coef: max_norm must be > 0");
  clip_coef_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sumsq, max_norm, coef);
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
