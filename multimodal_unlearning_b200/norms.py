"""Whole-model reductions around the update step, over the C ABI (csrc/norms.cu).

* `clip_grad_norm_`  — drop-in for `torch.nn.utils.clip_grad_norm_` (2-norm): HF Trainer's default
  `max_grad_norm = 1.0` applies to every reference run (the override is commented out at mm/forget.py:151).
  One streaming pass per tensor accumulates the sum of squares in double on the device; the clipping
  coefficient stays on the device (no host synchronisation) and is either applied by the fused update
  (`MaskedAdamW(max_grad_norm=...)`, salun.py) or, here, by scaling the gradients in place.
* `param_l1l0` / `weight_regulariser` — the L1 / L0 weight regulariser of mm/trainer.py:451-481 and
  nlp/dataloader.py:461-483.  The reference concatenates every trainable parameter (or parameter minus
  teacher parameter) into one 7 B-element temporary per step and takes `torch.norm(p=1)` / `(p=0)` of it;
  the norm of a concatenation is the sum of per-tensor norms for p in {0, 1}, so this walks the tensors
  and accumulates two device doubles.  Backward of the L1 term: lambda_1 * sign(p - p_teacher) (torch:
  sign(0) = 0); the L0 count has zero gradient, exactly as `torch.norm(p=0)` under autograd.

bf16 CUDA tensors only; anything else raises (no CPU or eager fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

import torch

from . import _lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _flat_bf16(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.bfloat16:
        raise RuntimeError(f"{what}: need CUDA bf16 tensors (got {t.dtype} on {t.device}); there is no fallback path")
    t = t.detach()
    if not t.is_contiguous() or t.data_ptr() % 16:
        t = t.contiguous().clone()
    return t.view(-1)


class _Scratch:
    """Per-device partial-sum workspace (doubles), allocated once."""

    _bufs: dict = {}

    @classmethod
    def get(cls, device) -> torch.Tensor:
        key = torch.device(device).index or 0
        buf = cls._bufs.get(key)
        if buf is None:
            with torch.cuda.device(key):
                n = int(_lib.load().clear_reduce_partials_elems())
            buf = torch.empty(n, dtype=torch.float64, device=torch.device("cuda", key))
            cls._bufs[key] = buf
        return buf


def grad_sqnorm(grads: Sequence[torch.Tensor], masks: Sequence[torch.Tensor | None] | None = None) -> torch.Tensor:
    """Device double [1]: sum over tensors of sum((g * mask)^2).  No host synchronisation."""
    grads = [g for g in grads if g is not None and g.numel()]
    if not grads:
        raise RuntimeError("grad_sqnorm: no gradients")
    lib = _lib.load()
    dev = grads[0].device
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    part = _Scratch.get(dev)
    for i, g in enumerate(grads):
        gf = _flat_bf16(g, "grad_sqnorm")
        m = masks[i] if masks is not None else None
        m8 = mb = None
        if m is not None:
            if not m.is_cuda:
                raise RuntimeError("grad_sqnorm: mask must be a CUDA tensor")
            m = m.contiguous()
            if m.dtype == torch.uint8:      # bit-packed (salun.pack_mask / SaliencyState.mask_flat(packed=True))
                if m.numel() != (gf.numel() + 7) // 8:
                    raise RuntimeError("grad_sqnorm: packed mask must hold ceil(numel / 8) bytes")
                mb = m
            elif m.dtype == torch.int8 and m.numel() == gf.numel():
                m8 = m.clone() if m.data_ptr() % 8 else m
            else:
                raise RuntimeError("grad_sqnorm: mask must be int8 of the gradient's size or bit-packed uint8")
        _lib.check(lib.clear_grad_sqnorm(gf.data_ptr(), m8.data_ptr() if m8 is not None else None,
                                         mb.data_ptr() if mb is not None else None, gf.numel(),
                                         part.data_ptr(), out.data_ptr(), 1, _stream()), "clear_grad_sqnorm")
    return out


def clip_coef(sumsq: torch.Tensor, max_norm: float) -> torch.Tensor:
    """Device float [1]: min(1, max_norm / (sqrt(sumsq) + 1e-6)) — torch.nn.utils.clip_grad_norm_'s coefficient."""
    coef = torch.empty(1, dtype=torch.float32, device=sumsq.device)
    _lib.check(_lib.load().clear_clip_coef(sumsq.data_ptr(), float(max_norm), coef.data_ptr(), _stream()), "clear_clip_coef")
    return coef


def clip_grad_norm_(parameters: Iterable[torch.Tensor] | torch.Tensor, max_norm: float,
                    masks: dict | None = None) -> torch.Tensor:
    """Drop-in for torch.nn.utils.clip_grad_norm_(parameters, max_norm) (norm_type 2).  Returns the total
    norm as a 0-dim fp32 device tensor, scales every .grad in place by the clamped coefficient.
    `masks` (param -> int8) restricts the norm to the SalUn-selected elements (the gradient the masked update sees)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    params = [p for p in parameters if p.grad is not None]
    if not params:
        return torch.zeros((), dtype=torch.float32)
    sumsq = grad_sqnorm([p.grad for p in params], [masks.get(p) for p in params] if masks else None)
    coef = clip_coef(sumsq, max_norm)
    lib = _lib.load()
    for p in params:  # grads.mul_(clip_coef_clamped), as torch does (also when the coefficient is 1)
        g = p.grad
        if not g.is_contiguous() or g.dtype != torch.bfloat16 or g.data_ptr() % 16:
            raise RuntimeError("clip_grad_norm_: gradients must be contiguous, 16-byte aligned CUDA bf16 tensors")
        _lib.check(lib.clear_grad_scale(g.data_ptr(), g.numel(), coef.data_ptr(), _stream()), "clear_grad_scale")
    return sumsq.sqrt().to(torch.float32).reshape(())


def param_l1l0(params: Sequence[torch.Tensor], refs: Sequence[torch.Tensor] | None = None) -> torch.Tensor:
    """Device double [2]: (sum |p - ref|, #{p != ref}) over all tensors; `refs` None = distance from zero.
    The difference is rounded to bf16 first, as `param - teacher_param` is in the reference."""
    lib = _lib.load()
    dev = params[0].device
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    part = _Scratch.get(dev)
    for i, p in enumerate(params):
        if p.numel() == 0:
            continue
        pf = _flat_bf16(p, "param_l1l0")
        rf = _flat_bf16(refs[i], "param_l1l0") if refs is not None else None
        if rf is not None and rf.numel() != pf.numel():
            raise RuntimeError("param_l1l0: parameter / reference shape mismatch")
        _lib.check(lib.clear_param_l1l0(pf.data_ptr(), rf.data_ptr() if rf is not None else None, pf.numel(),
                                        part.data_ptr(), out.data_ptr(), 1, _stream()), "clear_param_l1l0")
    return out


class _WeightNorms(torch.autograd.Function):
    """(||cat(p - ref)||_1, ||cat(p - ref)||_0) as an fp32 [2] tensor with the L1 gradient attached."""

    @staticmethod
    def forward(ctx, n_params, *tensors):
        params = tensors[:n_params]
        refs = tensors[n_params:] or None
        ctx.params = params
        ctx.refs = refs
        return param_l1l0(params, refs).to(torch.float32)

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load()
        coef = gout[0:1].to(torch.float32).contiguous()
        grads = []
        for i, p in enumerate(ctx.params):
            if not ctx.needs_input_grad[1 + i] or p.numel() == 0:
                grads.append(None)
                continue
            pf = _flat_bf16(p, "weight_regulariser")
            rf = _flat_bf16(ctx.refs[i], "weight_regulariser") if ctx.refs is not None else None
            g = torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
            _lib.check(lib.clear_param_l1_grad(pf.data_ptr(), rf.data_ptr() if rf is not None else None, pf.numel(),
                                               coef.data_ptr(), g.data_ptr(), _stream()), "clear_param_l1_grad")
            grads.append(g)
        n_refs = len(ctx.refs) if ctx.refs is not None else 0
        return (None, *grads, *([None] * n_refs))


def weight_regulariser(params: Sequence[torch.Tensor], refs: Sequence[torch.Tensor] | None, l1_lambda: float,
                       l0_lambda: float) -> torch.Tensor:
    """l1_lambda * ||cat(p - ref)||_1 + l0_lambda * ||cat(p - ref)||_0 as a 0-dim fp32 tensor (autograd: L1 term)."""
    params = list(params)
    if not params:
        raise RuntimeError("weight_regulariser: no trainable parameters")
    tensors = params + (list(refs) if refs is not None else [])
    norms = _WeightNorms.apply(len(params), *tensors)
    return (l1_lambda or 0.0) * norms[0] + (l0_lambda or 0.0) * norms[1]
