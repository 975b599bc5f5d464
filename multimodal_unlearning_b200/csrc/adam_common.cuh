// Shared by the masked AdamW kernels (fp32 moments: salun.cu; 8-bit block-wise moments: adamw8.cu): the scalar prologue and
// the per-element update of torch.optim.adam._single_tensor_adam, every fp32 operation spelled with an explicit
// round-to-nearest intrinsic so that no FMA contraction can change a bit relative to the oracle.
#pragma once
#include <math.h>

namespace clear {

struct AdamConsts {
  float decay;      // 1 - lr * weight_decay
  float w1;         // 1 - beta1
  float beta2, w2;  // beta2, 1 - beta2
  float bc2_sqrt;   // sqrt(1 - beta2^step)
  float step_size;  // lr / (1 - beta1^step)
  float eps;
  const float* grad_scale;  // optional device scalar: clipping coefficient
};

__device__ __forceinline__ float adam_one(float p, float g, float& m, float& v, const AdamConsts& c) {
  p = __fmul_rn(p, c.decay);                                        // param.mul_(1 - lr * wd)
  m = __fadd_rn(m, __fmul_rn(c.w1, __fsub_rn(g, m)));               // exp_avg.lerp_(grad, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, c.beta2), __fmul_rn(c.w2, __fmul_rn(g, g)));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  // The IEEE square root / division take a slow path (a subroutine call, the whole warp with it) when an
  // operand is zero — which every masked-out parameter is, forever.  A zero operand is swapped for 1 and
  // the exact result (sqrt(+0) = +0, +-0 / d = +-0) selected afterwards: same bits, fast path.
  const float sq = (v == 0.f) ? v : __fsqrt_rn((v == 0.f) ? 1.f : v);
  const float sd = (sq == 0.f) ? sq : __fdiv_rn((sq == 0.f) ? 1.f : sq, c.bc2_sqrt);
  const float denom = __fadd_rn(sd, c.eps);
  const float q = (m == 0.f) ? __fmul_rn(m, 1.f) : __fdiv_rn((m == 0.f) ? 1.f : m, denom);
  return __fsub_rn(p, __fmul_rn(c.step_size, q));   // addcdiv_(exp_avg, denom, -step_size)
}

// The update of the 8-bit-state kernel.  Same quantities, but the two divisions are spelled as multiplications by exactly
// rounded reciprocals (x / bc2_sqrt -> x * fp32(1 / bc2_sqrt), m / denom -> m * (1 / denom)): an IEEE division costs ~10
// instructions with its slow-path check, and this kernel is issue-bound, not HBM-bound (10 bytes per parameter).  The result
// differs from adam_one's in the last bit now and then — far below the 8-bit resolution of the states it is stored in; the
// oracle (oracle/salun_oracle.py::adamw8_step) restates exactly this sequence.
__device__ __forceinline__ float adam8_one(float p, float g, float& m, float& v, const AdamConsts& c, float inv_bc2_sqrt) {
  p = __fmul_rn(p, c.decay);
  m = __fadd_rn(m, __fmul_rn(c.w1, __fsub_rn(g, m)));
  v = __fadd_rn(__fmul_rn(v, c.beta2), __fmul_rn(c.w2, __fmul_rn(g, g)));
  const float sq = (v == 0.f) ? v : __fsqrt_rn((v == 0.f) ? 1.f : v);   // sqrt(+0) takes the slow path: masked-out parameters stay at zero forever
  const float denom = __fadd_rn(__fmul_rn(sq, inv_bc2_sqrt), c.eps);
  return __fsub_rn(p, __fmul_rn(c.step_size, __fmul_rn(m, __frcp_rn(denom))));
}

// Scalar prologue of torch.optim.adam._single_tensor_adam, in double like the Python floats there.
inline AdamConsts make_adam_consts(double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                                   const float* grad_scale_dev) {
  AdamConsts c;
  c.decay = (float)(1.0 - lr * weight_decay);
  c.w1 = (float)(1.0 - beta1);
  c.beta2 = (float)beta2;
  c.w2 = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  c.step_size = (float)(lr / bc1);
  c.bc2_sqrt = (float)sqrt(bc2);
  c.eps = (float)eps;
  c.grad_scale = grad_scale_dev;
  return c;
}

}  // namespace clear
