"""SalUn gradient-saliency masks and the masked AdamW update, over the C ABI.

Host-side mirror of methods/salun/generate_mask.py:17-89 (`save_gradient_ratio`):
    :41-44  gradients[name] += param.grad          -> SaliencyState.accumulate        (bf16 += bf16)
    :47-49  abs_().cpu().to(float16)               -> keys = fp16 bit pattern of |acc|, formed in-kernel
    :52-85  per tensor, per threshold: sort desc., top int(th * numel) -> 1 (int8)
                                                   -> SaliencyState.masks (histogram + select + mask)
    :87-89  torch.save(dict[name -> int8 mask])    -> save_gradient_ratio writes the same `with_{th}.pt`
The masked update (`grad *= mask` then AdamW) is MaskedAdamW.

All parameters of a model live in ONE flat bf16 buffer (tensors start at multiples of 8 elements, zero
padded), so that every stage is a single grid over HBM instead of one launch per tensor.
There is no CPU path: the reference's CPU sort is replaced by order statistics over the 32768
non-negative fp16 bit patterns; ties at the threshold value are broken by flat index (stable sort).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Iterator, Sequence

import torch

from . import _lib

REFERENCE_THRESHOLDS = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)  # generate_mask.py:52
N_BINS = 32768


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FlatLayout:
    """Placement of named tensors in a flat buffer: offsets are multiples of 8 elements (16 B of bf16)."""

    def __init__(self, named_shapes: Sequence[tuple[str, Sequence[int]]]):
        self.names = [n for n, _ in named_shapes]
        self.shapes = [tuple(s) for _, s in named_shapes]
        self.numels = [int(torch.Size(s).numel()) for s in self.shapes]
        offs = [0]
        for n in self.numels:
            offs.append(offs[-1] + (n + 7) // 8 * 8)
        self.offsets = offs
        self.total = offs[-1]
        chunk = int(_lib.load().clear_saliency_chunk_elems())
        pref = [0]
        for n in self.numels:
            pref.append(pref[-1] + ((n + 7) // 8 * 8 + chunk - 1) // chunk)
        self.chunk_prefix = pref
        self.n_chunks = pref[-1]

    @classmethod
    def of_model(cls, model: torch.nn.Module) -> "FlatLayout":
        return cls([(n, p.shape) for n, p in model.named_parameters()])

    def view(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        return flat[self.offsets[i]: self.offsets[i] + self.numels[i]].view(self.shapes[i])

    def device_tables(self, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        return (torch.tensor(self.offsets, dtype=torch.int64, device=device),
                torch.tensor(self.numels, dtype=torch.int64, device=device),
                torch.tensor(self.chunk_prefix, dtype=torch.int64, device=device))


class SaliencyState:
    """The `gradients` dict of save_gradient_ratio as one flat bf16 buffer on the GPU."""

    def __init__(self, layout: FlatLayout, device="cuda"):
        self.layout = layout
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SaliencyState lives on a CUDA device: SalUn masks have no CPU path here")
        self.acc = torch.zeros(layout.total, dtype=torch.bfloat16, device=self.device)
        self.seg_off, self.seg_numel, self.chunk_prefix = layout.device_tables(self.device)
        self.launches = 0

    # ---- generate_mask.py:41-44 -------------------------------------------------------------------
    def accumulate_flat(self, grads: torch.Tensor) -> None:
        """acc += grads for a flat bf16 gradient buffer in this layout (one launch)."""
        if grads.dtype != torch.bfloat16 or grads.numel() != self.layout.total or not grads.is_cuda:
            raise RuntimeError("accumulate_flat: need a CUDA bf16 buffer of layout.total elements")
        _lib.check(_lib.load().clear_saliency_accumulate(self.acc.data_ptr(), grads.data_ptr(), grads.numel(), _stream()),
                   "clear_saliency_accumulate")
        self.launches += 1

    def accumulate_model(self, model: torch.nn.Module) -> None:
        """gradients[name] += param.grad for every parameter that has a gradient."""
        lib = _lib.load()
        params = dict(model.named_parameters())
        for i, name in enumerate(self.layout.names):
            g = params[name].grad
            if g is None or self.layout.numels[i] == 0:
                continue
            g = g.detach().to(torch.bfloat16).contiguous().view(-1)
            if g.data_ptr() % 16:
                g = g.clone()
            dst = self.acc[self.layout.offsets[i]:]
            _lib.check(lib.clear_saliency_accumulate(dst.data_ptr(), g.data_ptr(), g.numel(), _stream()),
                       "clear_saliency_accumulate")
            self.launches += 1

    # ---- generate_mask.py:47-85 -------------------------------------------------------------------
    def select(self, thresholds: Sequence[float] = REFERENCE_THRESHOLDS) -> torch.Tensor:
        """Per tensor and threshold: (bit pattern tau of the k-th largest key, ties to take, lo, hi) with
        [lo, hi] the bf16 magnitude patterns whose key is tau, total = #{key == tau}).  int64 [S, n_th, 5]."""
        lib = _lib.load()
        lay = self.layout
        n_segs = len(lay.names)
        hist = torch.zeros((n_segs, N_BINS), dtype=torch.int32, device=self.device)
        _lib.check(lib.clear_saliency_histogram(self.acc.data_ptr(), self.seg_off.data_ptr(), n_segs, lay.total,
                                                hist.data_ptr(), _stream()), "clear_saliency_histogram")
        th = (C.c_double * len(thresholds))(*[float(t) for t in thresholds])
        sel = torch.full((n_segs, len(thresholds), 5), -1, dtype=torch.int64, device=self.device)
        _lib.check(lib.clear_saliency_select(hist.data_ptr(), self.seg_off.data_ptr(), self.seg_numel.data_ptr(), n_segs,
                                             th, len(thresholds), sel.data_ptr(), _stream()), "clear_saliency_select")
        self.launches += 2
        self.hist = hist
        return sel

    def mask_flat(self, sel: torch.Tensor, th_index: int, out: torch.Tensor | None = None,
                  packed: bool = False) -> torch.Tensor:
        """Mask of one threshold, flat in this layout: int8 (one byte per parameter, the reference's format) or,
        with packed=True, uint8 bit-packed (bit i & 7 of byte i >> 3; layout.total / 8 bytes)."""
        lib = _lib.load()
        lay = self.layout
        if out is None:
            out = (torch.empty(lay.total // 8, dtype=torch.uint8, device=self.device) if packed
                   else torch.empty(lay.total, dtype=torch.int8, device=self.device))
        scratch = torch.empty(lay.n_chunks + 1, dtype=torch.int64, device=self.device)
        fn = lib.clear_saliency_mask_bits if packed else lib.clear_saliency_mask
        _lib.check(fn(self.acc.data_ptr(), self.seg_off.data_ptr(), self.chunk_prefix.data_ptr(), len(lay.names),
                      lay.n_chunks, sel.data_ptr(), sel.shape[1], th_index, scratch.data_ptr(), out.data_ptr(),
                      _stream()), "clear_saliency_mask_bits" if packed else "clear_saliency_mask")
        self.launches += 3
        return out

    def masks(self, thresholds: Sequence[float] = REFERENCE_THRESHOLDS) -> Iterator[tuple[float, dict]]:
        """Yields (threshold, {name: int8 mask of the parameter's shape}) — `process_threshold`'s hard_dict."""
        sel = self.select(thresholds)
        for j, th in enumerate(thresholds):
            flat = self.mask_flat(sel, j)
            yield th, {n: self.layout.view(flat, i) for i, n in enumerate(self.layout.names)}


def sharded_masks(state: SaliencyState, thresholds: Sequence[float] = REFERENCE_THRESHOLDS, group=None,
                  reduce_acc: bool = True) -> list[torch.Tensor]:
    """Data-parallel mask generation.  The reference sums SIGNED gradients over all forget batches
    before taking |.| (generate_mask.py:44 then :49), so with the batches dealt out to ranks the accumulators are
    all-reduced (sum, bf16) first.  Its thresholds are per tensor (:54-85): tensors are independent units — rank r
    computes the masks of the tensors `distributed.partition_units` deals to it (no collective on that path) and the
    bit-packed results are exchanged.  Returns one flat bit-packed mask per threshold in `state.layout`, identical on
    every rank (feed `layout`-offset slices to MaskedAdamW, or `unpack_mask` them)."""
    import torch.distributed as dist

    from . import distributed as D

    lay = state.layout
    world = D.world_size()
    if world == 1:
        sel = state.select(thresholds)
        return [state.mask_flat(sel, j, packed=True) for j in range(len(thresholds))]
    rank = dist.get_rank(group)
    if reduce_acc:
        dist.all_reduce(state.acc, op=dist.ReduceOp.SUM, group=group)
    units = D.partition_units(lay.numels, world)
    sub_layouts = [FlatLayout([(lay.names[i], lay.shapes[i]) for i in units[r]]) for r in range(world)]
    mine = sub_layouts[rank]
    sub = SaliencyState(mine, state.device)
    for j, i in enumerate(units[rank]):
        n = lay.numels[i]
        sub.acc[mine.offsets[j]: mine.offsets[j] + n].copy_(state.acc[lay.offsets[i]: lay.offsets[i] + n])
    sel = sub.select(thresholds) if units[rank] else None
    out = []
    for t in range(len(thresholds)):
        full = torch.zeros(lay.total // 8, dtype=torch.uint8, device=state.device)
        for r in range(world):
            lr = sub_layouts[r]
            if lr.total == 0:
                continue
            buf = sub.mask_flat(sel, t, packed=True) if r == rank else torch.empty(lr.total // 8, dtype=torch.uint8, device=state.device)
            dist.broadcast(buf, src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            for j, i in enumerate(units[r]):
                nb = (lay.numels[i] + 7) // 8
                full[lay.offsets[i] // 8: lay.offsets[i] // 8 + nb].copy_(buf[lr.offsets[j] // 8: lr.offsets[j] // 8 + nb])
        out.append(full)
    state.launches += 3 * len(thresholds) + 2
    return out


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """int8 / bool mask (any shape, CUDA) -> bit-packed uint8 [ceil(numel / 8)] (bit i & 7 of byte i >> 3)."""
    if not mask.is_cuda:
        raise RuntimeError("pack_mask: CUDA tensors only")
    m = mask.to(torch.int8).contiguous().view(-1)
    if m.data_ptr() % 8:
        m = m.clone()
    bits = torch.empty((m.numel() + 7) // 8, dtype=torch.uint8, device=m.device)
    _lib.check(_lib.load().clear_mask_pack(m.data_ptr(), m.numel(), bits.data_ptr(), _stream()), "clear_mask_pack")
    return bits


def unpack_mask(bits: torch.Tensor, shape: Sequence[int]) -> torch.Tensor:
    """bit-packed uint8 -> int8 mask of `shape` (the reference's `with_{th}.pt` element format)."""
    n = int(torch.Size(shape).numel())
    if not bits.is_cuda or bits.dtype != torch.uint8 or bits.numel() != (n + 7) // 8:
        raise RuntimeError("unpack_mask: need a CUDA uint8 tensor of ceil(numel / 8) bytes")
    out = torch.empty(n, dtype=torch.int8, device=bits.device)
    _lib.check(_lib.load().clear_mask_unpack(bits.contiguous().data_ptr(), n, out.data_ptr(), _stream()), "clear_mask_unpack")
    return out.view(tuple(shape))


def load_masks(path: str, device="cuda", packed: bool = True) -> dict:
    """Read one of the reference's mask files (`with_{th}.pt` = torch.save(dict name -> int8 tensor of the parameter's
    shape), generate_mask.py:72-89) onto the GPU; packed=True converts each tensor to the bit-packed format that
    MaskedAdamW consumes (1 bit instead of 1 byte per parameter in HBM)."""
    raw = torch.load(path, map_location="cpu")
    out = {}
    for name, m in raw.items():
        m = m.to(device=device, dtype=torch.int8)
        out[name] = pack_mask(m) if packed else m
    return out


def save_gradient_ratio(forget_loader: Iterable, model: torch.nn.Module, criterion, cfg) -> None:
    """Drop-in for methods/salun/generate_mask.py:17 — same arguments, same `with_{th}.pt` files.
    `cfg` needs `.save_dir` (unlearn_lr / weight_decay only fed the reference's unused SGD object)."""
    model.eval()
    state = SaliencyState(FlatLayout.of_model(model), next(model.parameters()).device)
    for batch in forget_loader:
        input_ids, labels, attention_mask = batch[0], batch[1], batch[2]
        device = state.device
        loss = model(input_ids=input_ids.to(device), attention_mask=attention_mask.to(device),
                     labels=labels.to(device)).loss
        model.zero_grad(set_to_none=True)
        loss.backward()
        state.accumulate_model(model)
    os.makedirs(cfg.save_dir, exist_ok=True)
    for th, hard_dict in state.masks(REFERENCE_THRESHOLDS):
        torch.save({k: v.cpu() for k, v in hard_dict.items()}, os.path.join(cfg.save_dir, f"with_{th}.pt"))


class MaskedAdamW(torch.optim.Optimizer):
    """AdamW whose gradient is multiplied by a SalUn mask first (upstream SalUn: `param.grad *= mask[name]`
    before `optimizer.step()`).  bf16 parameters and gradients, fp32 moments, torch's
    `_single_tensor_adam` operation order; one fused pass over 23 bytes per parameter."""

    def __init__(self, params, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, masks: dict | None = None,
                 max_grad_norm: float | None = None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.masks = masks or {}  # param -> int8 tensor
        # Global gradient-norm clipping fused into the update (HF Trainer's default max_grad_norm = 1.0 applies to
        # the reference runs: the override is commented out at mm/forget.py:151).  The norm is taken over the MASKED
        # gradients (what the update consumes); one extra 2-3 B/param pass, the coefficient never leaves the device.
        self.max_grad_norm = max_grad_norm
        self.last_grad_norm = None  # device fp64 [1] after a clipped step

    def set_mask(self, param: torch.Tensor, mask: torch.Tensor | None, packed: bool = True) -> None:
        """mask: int8 / bool of the parameter's shape, or an already bit-packed uint8 [ceil(numel / 8)] tensor.
        Stored bit-packed by default (0.875 GB instead of 7 GB for a 7 B-parameter model)."""
        if mask is None:
            self.masks.pop(param, None)
            return
        mask = mask.to(device=param.device)
        if mask.dtype == torch.uint8 and mask.dim() == 1 and mask.numel() == (param.numel() + 7) // 8 and mask.numel() != param.numel():
            self.masks[param] = mask.contiguous()
        elif mask.numel() != param.numel():
            raise RuntimeError("set_mask: mask must have the parameter's number of elements (or be its bit-packed form)")
        elif packed:
            self.masks[param] = pack_mask(mask)
        else:
            self.masks[param] = mask.to(torch.int8).contiguous()

    def set_masks_from_file(self, path: str, model: torch.nn.Module) -> None:
        """Attach the masks of one of the reference's `with_{th}.pt` files to this optimizer's parameters."""
        named = dict(model.named_parameters())
        for name, bits in load_masks(path, device=next(model.parameters()).device).items():
            if name in named:
                self.masks[named[name]] = bits

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        scale = None
        if self.max_grad_norm is not None:
            from . import norms
            ps = [p for grp in self.param_groups for p in grp["params"] if p.grad is not None]
            if ps:
                sumsq = norms.grad_sqnorm([p.grad for p in ps], [self.masks.get(p) for p in ps])
                scale = norms.clip_coef(sumsq, self.max_grad_norm)
                self.last_grad_norm = sumsq.sqrt()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.bfloat16 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("MaskedAdamW updates contiguous CUDA bf16 parameters")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                    st["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                st["step"] += 1
                g = p.grad.to(torch.bfloat16).contiguous()
                m = self.masks.get(p)
                bits = m if (m is not None and m.dtype == torch.uint8) else None
                if bits is not None:
                    m = None
                a = _lib.AdamWArgs(param=p.data_ptr(), grad=g.data_ptr(), exp_avg=st["exp_avg"].data_ptr(),
                                   exp_avg_sq=st["exp_avg_sq"].data_ptr(), mask=m.data_ptr() if m is not None else None,
                                   mask_bits=bits.data_ptr() if bits is not None else None,
                                   n=p.numel(), lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"],
                                   weight_decay=group["weight_decay"], step=st["step"],
                                   grad_scale_dev=scale.data_ptr() if scale is not None else None)
                _lib.check(lib.clear_masked_adamw(C.byref(a), _stream()), "clear_masked_adamw")
        return loss


def masked_adamw_flat(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                      mask: torch.Tensor | None, step: int, lr=1e-5, betas=(0.9, 0.999), eps=1e-8,
                      weight_decay=0.01, grad_scale: torch.Tensor | None = None) -> None:
    """One launch over a flat parameter buffer (the layout of FlatLayout).  `grad_scale`: optional device fp32
    scalar (norms.clip_coef) applied to the masked gradient first, rounded to bf16 like `grad.mul_(coef)`."""
    bits = mask if (mask is not None and mask.dtype == torch.uint8) else None  # uint8 = bit-packed, int8 = one byte each
    if bits is not None:
        mask = None
    a = _lib.AdamWArgs(param=param.data_ptr(), grad=grad.data_ptr(), exp_avg=exp_avg.data_ptr(),
                       exp_avg_sq=exp_avg_sq.data_ptr(), mask=mask.data_ptr() if mask is not None else None,
                       mask_bits=bits.data_ptr() if bits is not None else None,
                       n=param.numel(), lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay,
                       step=step, grad_scale_dev=grad_scale.data_ptr() if grad_scale is not None else None)
    _lib.check(_lib.load().clear_masked_adamw(C.byref(a), _stream()), "clear_masked_adamw")
