"""CPU restatement of SalUn's saliency mask (reference: methods/salun/generate_mask.py:17-89) and of the
masked AdamW update.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg, never by the product package.

Parity status: PINNED for the mask.  The reference has no tests; the pins are outputs of the reference
itself: oracle/gen_golden.py runs the unmodified `save_gradient_ratio` on a tiny Llama (CPU) and stores
the per-batch gradients and the ten masks under tests/golden/salun_tiny.npz; tests/test_oracle_golden.py
checks this file against them.  The reference's `torch.sort(descending=True)` is unstable, so elements
EQUAL to the k-th largest value are chosen in an unspecified order; the contract here
is the stable order (lowest flat index first).  The pin is therefore two-fold: bit-exact against the
reference run with `stable=True` added to its one sort call, and — against the unmodified reference —
equal everywhere except on elements whose key equals the threshold key, with equal popcounts.
The masked update is not in the reference (no consumer of the masks exists there): "parity unpinned"
for `adamw_step`; it restates torch.optim.adam._single_tensor_adam (torch 2.11) in explicit fp32 ops
and is cross-checked against torch.optim.AdamW in the tests.

Integer/byte arithmetic is numpy; torch is used only to read bf16.
"""
from __future__ import annotations

import numpy as np
import torch

REFERENCE_THRESHOLDS = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)  # generate_mask.py:52


def accumulate(acc: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """gradients[name] += param.grad.data (generate_mask.py:41-44) in the model dtype (bf16)."""
    return (acc.float() + grad.float()).to(torch.bfloat16)


def saliency_keys(acc: torch.Tensor) -> np.ndarray:
    """abs_().cpu().to(torch.float16) (generate_mask.py:47-49) as uint16 bit patterns.  Non-negative
    fp16 values order like their bit patterns; overflow gives +inf (0x7C00), tiny values flush."""
    half = acc.detach().float().abs().cpu().numpy().astype(np.float16)
    return half.view(np.uint16).reshape(-1) & np.uint16(0x7FFF)


def mask_of_keys(keys: np.ndarray, threshold: float) -> np.ndarray:
    """process_threshold for one tensor (generate_mask.py:61-72) with a STABLE descending sort:
    k = int(threshold * numel) largest keys -> 1, ties by lowest flat index."""
    n = keys.size
    k = int(threshold * n)
    order = np.argsort(np.int32(0x7FFF) - keys.astype(np.int32), kind="stable")
    mask = np.zeros(n, dtype=np.int8)
    mask[order[:k]] = 1
    return mask


def masks(named_acc: dict, thresholds=REFERENCE_THRESHOLDS) -> dict:
    """{threshold: {name: int8 mask with the tensor's shape}} — the ten `with_{th}.pt` dicts."""
    out = {}
    keys = {n: saliency_keys(a) for n, a in named_acc.items()}
    for th in thresholds:
        out[th] = {n: mask_of_keys(keys[n], th).reshape(tuple(named_acc[n].shape)) for n in named_acc}
    return out


def select(keys: np.ndarray, threshold: float) -> tuple[int, int]:
    """(tau, ties): bit pattern of the k-th largest key and how many elements equal to it are taken;
    (-1, 0) when k == 0.  What clear_saliency_select computes from the histogram."""
    n = keys.size
    k = min(int(threshold * n), n)
    if k <= 0:
        return -1, 0
    hist = np.bincount(keys, minlength=32768).astype(np.int64)
    above = 0
    for b in range(32767, -1, -1):
        if k <= above + hist[b]:
            return b, k - above
        above += hist[b]
    raise AssertionError("unreachable")


def adamw_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: np.ndarray, exp_avg_sq: np.ndarray,
               mask: np.ndarray | None, step: int, lr=1e-5, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
    """One masked AdamW step.  param, grad: bf16 torch tensors; moments: fp32 numpy (updated copies are
    returned).  Operation order of torch.optim.adam._single_tensor_adam, every op a separately rounded
    fp32 operation (no fused multiply-add), scalars prepared in double then rounded to fp32."""
    f = np.float32
    p = param.detach().float().cpu().numpy().reshape(-1).astype(f)
    g = grad.detach().float().cpu().numpy().reshape(-1).astype(f)
    if mask is not None:
        g = np.where(mask.reshape(-1) != 0, g, f(0))           # grad *= mask
    m = exp_avg.reshape(-1).astype(f).copy()
    v = exp_avg_sq.reshape(-1).astype(f).copy()
    decay = f(1.0 - lr * weight_decay)
    w1 = f(1.0 - beta1)
    b2 = f(beta2)
    w2 = f(1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = f(lr / bc1)
    bc2_sqrt = f(bc2 ** 0.5)
    p = p * decay                                               # param.mul_(1 - lr * wd)
    m = m + w1 * (g - m)                                        # exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + w2 * (g * g)                                   # mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    denom = np.sqrt(v) / bc2_sqrt + f(eps)                      # (sqrt / bias_correction2_sqrt).add_(eps)
    p = p - step_size * (m / denom)                             # addcdiv_(exp_avg, denom, -step_size)
    new_param = torch.from_numpy(p.astype(f)).to(torch.bfloat16).reshape(param.shape)
    return new_param, m.reshape(exp_avg.shape), v.reshape(exp_avg_sq.shape)


def grad_sqnorm(grads, masks=None) -> float:
    """sum over tensors of sum((g * mask)^2) in double — the square of torch.nn.utils.clip_grad_norm_'s total
    norm (HF Trainer default max_grad_norm = 1.0; the override is commented out at mm/forget.py:151)."""
    total = 0.0
    for i, g in enumerate(grads):
        x = g.detach().float().cpu().numpy().astype(np.float64).reshape(-1)
        if masks is not None and masks[i] is not None:
            x = np.where(np.asarray(masks[i]).reshape(-1) != 0, x, 0.0)
        total += float(np.dot(x, x))
    return total


def clip_coef(sumsq: float, max_norm: float) -> np.float32:
    """clip_coef_clamped of torch.nn.utils.clip_grad_norm_: min(1, max_norm / (total_norm + 1e-6))."""
    return np.float32(min(1.0, max_norm / (np.sqrt(sumsq) + 1e-6)))


def scale_grad(grad: torch.Tensor, coef) -> torch.Tensor:
    """grad.mul_(coef) on a bf16 gradient: fp32 product, one rounding to bf16."""
    g = grad.detach().float().cpu().numpy().astype(np.float32) * np.float32(coef)
    return torch.from_numpy(g.astype(np.float32)).to(torch.bfloat16)


def param_l1l0(params, refs=None) -> tuple[float, float]:
    """(||cat(p - ref)||_1, ||cat(p - ref)||_0) of mm/trainer.py:451-481 / nlp/dataloader.py:461-483; `p - ref`
    is a bf16 tensor in the reference (fp32 subtraction rounded to bf16); sums in double."""
    l1 = l0 = 0.0
    for i, p in enumerate(params):
        d = p.detach().to(torch.bfloat16).cpu()
        if refs is not None:
            d = d - refs[i].detach().to(torch.bfloat16).cpu()
        x = d.float().numpy().astype(np.float64).reshape(-1)
        l1 += float(np.abs(x).sum())
        l0 += float(np.count_nonzero(x))
    return l1, l0
