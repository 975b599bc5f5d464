// Masked AdamW with 8-bit block-wise optimiser states.
//
// The reference trains with `optim="adamw_bnb_8bit"` (mm/forget.py:156, nlp/forget.py:166): bitsandbytes' block-wise
// 8-bit AdamW — both moments stored as one byte per parameter, an index into a 256-entry non-uniform ("dynamic")
// codebook, times one fp32 absmax per block of parameters.  The memory footprint of the drop-in depends on it: 2 bytes
// of optimiser state per parameter instead of 8.  bitsandbytes is a third-party dependency that is absent from
// /root/reference and from this image, so this is a B200 design of the same scheme, NOT a bit-compatible clone
// (parity unpinned; the oracle restates exactly what is computed here, oracle/salun_oracle.py::adamw8_step):
//   * block = 256 parameters = the 8 consecutive elements of each of a warp's 32 lanes: the block absmax is a warp
//     shuffle reduction, no shared memory, no second pass;
//   * one fused pass: dequantise m, v (codebook lookup x block absmax) -> masked, optionally clipped gradient -> the AdamW
//     update in fp32 (adam_common.cuh::adam8_one: torch's quantities, divisions spelled as reciprocal multiplies) -> new
//     block absmax -> requantise -> store; 10.2 bytes of HBM traffic per parameter instead of 22.1;
//   * quantisation = the nearest codebook entry of x * (1 / absmax) ROUNDED TO ITS TOP 16 BITS (sign, exponent, 7 mantissa
//     bits: key = (bits + 0x8000) >> 16, round half away from zero — unbiased, and 0.4 % of the value at most, below the
//     1.4 % spacing of the finest codes), which makes the quantiser ONE byte load from a 65536-entry (signed) / 32768-entry
//     (unsigned) table in shared memory: no binary search (8 dependent loads per state and element), no data-dependent
//     loop, no divergence.  A first version searched the exact midpoints from a table lower bound: 58.9 ms per step over
//     6.74 B parameters (LSU- and branch-bound); this one is bound by the exact IEEE division / square root of the update.
#include <cuda_bf16.h>

#include "adam_common.cuh"
#include "common.h"

namespace clear {
namespace {

constexpr int kBlock = 256;       // parameters per absmax block (one warp x 8 elements)
constexpr int kThreads = 512;     // x 2 CTAs per SM (98 KB of tables each): 1024 threads, 48 KB of loads in flight per SM
constexpr int kLut1 = 65536;      // signed codebook: every top-16-bit float pattern
constexpr int kLut2 = 32768;      // unsigned codebook: the non-negative half
constexpr int kSmemBytes = kLut1 + kLut2 + 2 * 256 * 4;   // + code1, code2

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// index of the nearest code value of `a` rounded to its top 16 bits (see the header comment); `mask` = 0xFFFF (signed table
// over all patterns) or 0x7FFF (unsigned table: a >= 0)
__device__ __forceinline__ uint32_t quantise(float a, const uint8_t* lut, uint32_t mask) {
  return lut[((__float_as_uint(a) + 0x8000u) >> 16) & mask];
}

struct Adam8Params {
  __nv_bfloat16* param;
  const __nv_bfloat16* grad;
  uint8_t* s1;
  uint8_t* s2;
  float* absmax1;
  float* absmax2;
  const float* code1;
  const float* code2;
  const uint8_t* lut1;
  const uint8_t* lut2;
  const void* mask;
  int64_t n;
};

template <bool kPackedMask>
__global__ void __launch_bounds__(kThreads, 2) adamw8_kernel(const Adam8Params a, const AdamConsts c, const float inv_bc2_sqrt) {
  extern __shared__ uint8_t smem[];
  uint8_t* lut1 = smem;
  uint8_t* lut2 = smem + kLut1;
  float* code1 = reinterpret_cast<float*>(smem + kLut1 + kLut2);
  float* code2 = code1 + 256;
  for (int i = threadIdx.x; i < kLut1 / 16; i += kThreads) reinterpret_cast<uint4*>(lut1)[i] = reinterpret_cast<const uint4*>(a.lut1)[i];
  for (int i = threadIdx.x; i < kLut2 / 16; i += kThreads) reinterpret_cast<uint4*>(lut2)[i] = reinterpret_cast<const uint4*>(a.lut2)[i];
  if (threadIdx.x < 256) {
    const int i = threadIdx.x;
    code1[i] = a.code1[i];
    code2[i] = a.code2[i];
  }
  __syncthreads();
  const float gscale = c.grad_scale ? *c.grad_scale : 1.f;
  const int64_t nvec = (a.n + 7) >> 3;                      // the last vector may be partial
  const int64_t nvec_warp = (nvec + 31) & ~int64_t(31);     // whole warps iterate together (shuffle reductions)
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec_warp; i += stride) {
    const int64_t e0 = i << 3;
    const int live = (int)max((int64_t)0, min((int64_t)8, a.n - e0));   // elements of this lane's vector inside the tensor
    float p[8], g[8], m[8], v[8];
    uint32_t mk = 0xFFu;
    const int64_t blk = i >> 5;                              // absmax block = this warp-iteration
    if (live == 8) {
      const uint4 pr = *reinterpret_cast<const uint4*>(a.param + e0);
      const uint4 gr = ld_stream16(a.grad + e0);
      const uint2 q1 = *reinterpret_cast<const uint2*>(a.s1 + e0);
      const uint2 q2 = *reinterpret_cast<const uint2*>(a.s2 + e0);
      const __nv_bfloat16* ph = reinterpret_cast<const __nv_bfloat16*>(&pr);
      const __nv_bfloat16* gh = reinterpret_cast<const __nv_bfloat16*>(&gr);
      const float a1 = a.absmax1[blk], a2 = a.absmax2[blk];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = __bfloat162float(ph[j]);
        g[j] = __bfloat162float(gh[j]);
        m[j] = __fmul_rn(code1[((j < 4 ? q1.x : q1.y) >> (8 * (j & 3))) & 0xFFu], a1);
        v[j] = __fmul_rn(code2[((j < 4 ? q2.x : q2.y) >> (8 * (j & 3))) & 0xFFu], a2);
      }
      if (a.mask) {
        if constexpr (kPackedMask) {
          mk = static_cast<const uint8_t*>(a.mask)[i];
        } else {
          const uint2 mm = *reinterpret_cast<const uint2*>(static_cast<const int8_t*>(a.mask) + e0);
          mk = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) mk |= ((((j < 4 ? mm.x : mm.y) >> (8 * (j & 3))) & 0xFFu) ? 1u : 0u) << j;
        }
      }
    } else {
      const float a1 = live ? a.absmax1[blk] : 0.f, a2 = live ? a.absmax2[blk] : 0.f;
      mk = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p[j] = g[j] = m[j] = v[j] = 0.f;
        if (j < live) {
          p[j] = __bfloat162float(a.param[e0 + j]);
          g[j] = __bfloat162float(a.grad[e0 + j]);
          m[j] = __fmul_rn(code1[a.s1[e0 + j]], a1);
          v[j] = __fmul_rn(code2[a.s2[e0 + j]], a2);
          bool on = true;
          if (a.mask) on = kPackedMask ? ((static_cast<const uint8_t*>(a.mask)[(e0 + j) >> 3] >> ((e0 + j) & 7)) & 1u)
                                       : (static_cast<const int8_t*>(a.mask)[e0 + j] != 0);
          mk |= (on ? 1u : 0u) << j;
        }
      }
    }
    float mx1 = 0.f, mx2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < live) {
        float gj = ((mk >> j) & 1u) ? g[j] : 0.f;                                                    // grad *= mask
        if (c.grad_scale) gj = __bfloat162float(__float2bfloat16_rn(__fmul_rn(gj, gscale)));        // grad.mul_(clip_coef) in bf16
        p[j] = adam8_one(p[j], gj, m[j], v[j], c, inv_bc2_sqrt);
        mx1 = fmaxf(mx1, fabsf(m[j]));
        mx2 = fmaxf(mx2, v[j]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
      mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    }
    const float inv1 = (mx1 > 0.f) ? __frcp_rn(mx1) : 0.f;   // one IEEE reciprocal per block and state, then a multiply per element
    const float inv2 = (mx2 > 0.f) ? __frcp_rn(mx2) : 0.f;
    uint32_t q1[8], q2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      q1[j] = quantise(__fmul_rn(m[j], inv1), lut1, 0xFFFFu);
      q2[j] = quantise(__fmul_rn(v[j], inv2), lut2, 0x7FFFu);
    }
    if (live == 8) {
      uint4 pr;
      __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(&pr);
#pragma unroll
      for (int j = 0; j < 8; ++j) ph[j] = __float2bfloat16_rn(p[j]);
      *reinterpret_cast<uint4*>(a.param + e0) = pr;
      *reinterpret_cast<uint2*>(a.s1 + e0) = make_uint2(q1[0] | (q1[1] << 8) | (q1[2] << 16) | (q1[3] << 24),
                                                        q1[4] | (q1[5] << 8) | (q1[6] << 16) | (q1[7] << 24));
      *reinterpret_cast<uint2*>(a.s2 + e0) = make_uint2(q2[0] | (q2[1] << 8) | (q2[2] << 16) | (q2[3] << 24),
                                                        q2[4] | (q2[5] << 8) | (q2[6] << 16) | (q2[7] << 24));
    } else {
      for (int j = 0; j < live; ++j) {
        a.param[e0 + j] = __float2bfloat16_rn(p[j]);
        a.s1[e0 + j] = (uint8_t)q1[j];
        a.s2[e0 + j] = (uint8_t)q2[j];
      }
    }
    if ((threadIdx.x & 31) == 0 && (blk << 8) < a.n) {
      a.absmax1[blk] = mx1;
      a.absmax2[blk] = mx2;
    }
  }
}

}  // namespace
}  // namespace clear

using namespace clear;

extern "C" {

int32_t clear_adamw8_block_elems(void) { return kBlock; }

int clear_masked_adamw_8bit(const ClearAdamW8Args* a, void* stream) {
  if (!a) return fail(CLEAR_ERR_NULL, "clear_masked_adamw_8bit: args is NULL");
  if (!a->param || !a->grad || !a->state1 || !a->state2 || !a->absmax1 || !a->absmax2 || !a->code1 || !a->code2 || !a->qlut1 || !a->qlut2)
    return fail(CLEAR_ERR_NULL, "clear_masked_adamw_8bit: null tensor");
  if (a->n <= 0 || a->step < 1) return fail(CLEAR_ERR_SHAPE, "clear_masked_adamw_8bit: n > 0 and step >= 1 required");
  if ((reinterpret_cast<uintptr_t>(a->param) | reinterpret_cast<uintptr_t>(a->grad) | reinterpret_cast<uintptr_t>(a->qlut1) |
       reinterpret_cast<uintptr_t>(a->qlut2)) & 15)
    return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw_8bit: param, grad and the quantisation tables must be 16-byte aligned");
  if ((reinterpret_cast<uintptr_t>(a->state1) | reinterpret_cast<uintptr_t>(a->state2)) & 7)
    return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw_8bit: the 8-bit states must be 8-byte aligned");
  if (a->mask && (reinterpret_cast<uintptr_t>(a->mask) & 7)) return fail(CLEAR_ERR_ALIGN, "clear_masked_adamw_8bit: mask must be 8-byte aligned");
  if (a->mask && a->mask_bits) return fail(CLEAR_ERR_KIND, "clear_masked_adamw_8bit: pass mask or mask_bits, not both");
  const AdamConsts c = make_adam_consts(a->lr, a->beta1, a->beta2, a->eps, a->weight_decay, a->step, a->grad_scale_dev);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow(a->beta2, (double)a->step)));
  Adam8Params p;
  p.param = static_cast<__nv_bfloat16*>(a->param);
  p.grad = static_cast<const __nv_bfloat16*>(a->grad);
  p.s1 = a->state1; p.s2 = a->state2;
  p.absmax1 = a->absmax1; p.absmax2 = a->absmax2;
  p.code1 = a->code1; p.code2 = a->code2;
  p.lut1 = a->qlut1; p.lut2 = a->qlut2;
  p.mask = a->mask_bits ? static_cast<const void*>(a->mask_bits) : static_cast<const void*>(a->mask);
  p.n = a->n;
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const int64_t nvec = (a->n + 7) >> 3;
  int64_t blocks = (nvec + kThreads - 1) / kThreads;
  if (blocks > (int64_t)sms * 2) blocks = (int64_t)sms * 2;   // persistent: every CTA stages 98 KB of tables once
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  if (a->mask_bits) {
    CLEAR_CUDA_OK(cudaFuncSetAttribute(adamw8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    adamw8_kernel<true><<<(unsigned)blocks, kThreads, kSmemBytes, cs>>>(p, c, inv_bc2_sqrt);
  } else {
    CLEAR_CUDA_OK(cudaFuncSetAttribute(adamw8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    adamw8_kernel<false><<<(unsigned)blocks, kThreads, kSmemBytes, cs>>>(p, c, inv_bc2_sqrt);
  }
  CLEAR_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
