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

    def __init__(self, layout: FlatLayout, device="cuda", acc: torch.Tensor | None = None):
        """acc: optional existing flat bf16 buffer in `layout` (e.g. a slice of another state's accumulator) to work on
        in place instead of allocating a zeroed one."""
        self.layout = layout
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("SaliencyState lives on a CUDA device: SalUn masks have no CPU path here")
        if acc is None:
            acc = torch.zeros(layout.total, dtype=torch.bfloat16, device=self.device)
        elif (acc.dtype != torch.bfloat16 or not acc.is_cuda or acc.numel() != layout.total or not acc.is_contiguous()
              or acc.data_ptr() % 16):
            raise RuntimeError("SaliencyState: acc must be a contiguous, 16-byte aligned CUDA bf16 buffer of layout.total elements")
        self.acc = acc
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
    def histogram(self) -> torch.Tensor:
        """uint32 [S, 32768] (held as int32): per tensor, the number of elements whose fp16(|acc|) bit pattern is the bin."""
        lay = self.layout
        n_segs = len(lay.names)
        hist = torch.zeros((n_segs, N_BINS), dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().clear_saliency_histogram(self.acc.data_ptr(), self.seg_off.data_ptr(), n_segs, lay.total,
                                                        hist.data_ptr(), _stream()), "clear_saliency_histogram")
        self.launches += 1
        self.hist = hist
        return hist

    def hist_total(self, hist: torch.Tensor) -> torch.Tensor:
        """int64 [32768]: the histogram of ALL tensors of this layout (zero padding excluded)."""
        total = torch.empty(N_BINS, dtype=torch.int64, device=self.device)
        _lib.check(_lib.load().clear_saliency_hist_total(hist.data_ptr(), self.seg_off.data_ptr(), self.seg_numel.data_ptr(),
                                                         len(self.layout.names), total.data_ptr(), _stream()),
                   "clear_saliency_hist_total")
        self.launches += 1
        return total

    def select_global(self, hist: torch.Tensor, total: torch.Tensor, n_total: int, thresholds: Sequence[float],
                      rank_budget=None) -> torch.Tensor:
        """Global-threshold records from the (possibly all-reduced) total histogram.  n_total: number of elements the
        ranking runs over (all ranks).  rank_budget: optional callable (tau int64 [n_th], budget int64 [n_th]) -> int64
        [n_th] giving the part of each tie budget that falls to THIS layout (ranks that own consecutive flat slices:
        distributed.tie_budget_for_rank); None = the whole budget."""
        lib = _lib.load()
        n_th = len(thresholds)
        tb = global_tau(total, n_total, thresholds)
        override = None
        if rank_budget is not None:
            override = rank_budget(tb[:, 0], tb[:, 1]).to(torch.int64).contiguous()
        n_segs = len(self.layout.names)
        sel = torch.full((n_segs, n_th, 5), -1, dtype=torch.int64, device=self.device)
        _lib.check(lib.clear_saliency_select_global_fill(hist.data_ptr(), self.seg_off.data_ptr(), self.seg_numel.data_ptr(),
                                                         n_segs, tb.data_ptr(), override.data_ptr() if override is not None else None,
                                                         n_th, sel.data_ptr(), _stream()), "clear_saliency_select_global_fill")
        self.launches += 2
        self.tau_budget = tb
        return sel

    def select(self, thresholds: Sequence[float] = REFERENCE_THRESHOLDS, scope: str = "per_tensor") -> torch.Tensor:
        """Per tensor and threshold: (bit pattern tau of the k-th largest key, ties to take, lo, hi, total) with
        [lo, hi] the bf16 magnitude patterns whose key is tau, total = #{key == tau}.  int64 [S, n_th, 5].
        scope="per_tensor": k = int(th * numel) within each tensor — the shipped reference (generate_mask.py:54-85);
        scope="global": ONE ranking over all tensors, k = int(th * total numel), ties by global flat index (upstream
        SalUn's global rank, of which the reference's `start_index` is the vestige)."""
        lib = _lib.load()
        lay = self.layout
        n_segs = len(lay.names)
        hist = self.histogram()
        if scope == "global":
            return self.select_global(hist, self.hist_total(hist), sum(lay.numels), thresholds)
        if scope != "per_tensor":
            raise ValueError(f"unknown scope {scope!r}: 'per_tensor' or 'global'")
        th = (C.c_double * len(thresholds))(*[float(t) for t in thresholds])
        sel = torch.full((n_segs, len(thresholds), 5), -1, dtype=torch.int64, device=self.device)
        _lib.check(lib.clear_saliency_select(hist.data_ptr(), self.seg_off.data_ptr(), self.seg_numel.data_ptr(), n_segs,
                                             th, len(thresholds), sel.data_ptr(), _stream()), "clear_saliency_select")
        self.launches += 1
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

    def masks(self, thresholds: Sequence[float] = REFERENCE_THRESHOLDS, scope: str = "per_tensor") -> Iterator[tuple[float, dict]]:
        """Yields (threshold, {name: int8 mask of the parameter's shape}) — `process_threshold`'s hard_dict."""
        sel = self.select(thresholds, scope)
        for j, th in enumerate(thresholds):
            flat = self.mask_flat(sel, j)
            yield th, {n: self.layout.view(flat, i) for i, n in enumerate(self.layout.names)}


def slice_layout(layout: FlatLayout, c0: int, c1: int) -> tuple[FlatLayout, int, int]:
    """The mask chunks [c0, c1) of `layout` as a layout of their own: a contiguous range [f0, f1) of the flat buffer whose
    segments are the PIECES of the tensors it intersects (a tensor cut by a slice boundary appears as one piece on either
    side; chunks never straddle tensors, so every piece is a whole number of chunks of one tensor).  Returns
    (piece layout, f0, f1); piece k of tensor `name` starting at element e is named "name@e"."""
    chunk = int(_lib.load().clear_saliency_chunk_elems())
    pieces, f0, f1 = [], None, None
    for i, name in enumerate(layout.names):
        a, b = max(c0, layout.chunk_prefix[i]), min(c1, layout.chunk_prefix[i + 1])
        if a >= b:
            continue
        start = (a - layout.chunk_prefix[i]) * chunk
        stop = min(layout.numels[i], (b - layout.chunk_prefix[i]) * chunk)
        if f0 is None:
            f0 = layout.offsets[i] + start
        f1 = layout.offsets[i] + (stop + 7) // 8 * 8
        pieces.append((f"{name}@{start}", (stop - start,)))
    sub = FlatLayout(pieces)
    if f0 is None:
        return sub, 0, 0
    assert sub.total == f1 - f0, "pieces of a chunk range are contiguous in the flat buffer"
    return sub, f0, f1


def sharded_masks(state: SaliencyState, thresholds: Sequence[float] = REFERENCE_THRESHOLDS, group=None,
                  reduce_acc: bool = True, scope: str = "per_tensor") -> list[torch.Tensor]:
    """Data-parallel mask generation.  The reference sums SIGNED gradients over all forget batches
    before taking |.| (generate_mask.py:44 then :49), so with the batches dealt out to ranks the accumulators are
    all-reduced (sum, bf16) first.  Returns one flat bit-packed mask per threshold in `state.layout`, identical on
    every rank (feed `layout`-offset slices to MaskedAdamW, or `unpack_mask` them).

    scope="per_tensor" (the shipped reference, :54-85): tensors are independent units — rank r computes the masks of the
    tensors `distributed.partition_units` deals to it, no collective on that path; the packed results are merged by ONE
    all-reduce per threshold (the ranks' bytes are disjoint).
    scope="global" (upstream SalUn's global rank): rank r owns a contiguous 1/R slice of the flat buffer (whole mask
    chunks); it histograms its slice, the 32768-bin totals are all-reduced over NCCL (256 KB), every rank derives the same
    threshold key and tie budget, the budget is split over ranks by an exclusive prefix in slice (= flat index) order
    (`distributed.tie_budget_for_rank`), each rank writes the mask of its slice, one all-gather per threshold.

    Measured on 2 x B200 at 6.74 B parameters (profiles/r02_sharded_masks.md): both scopes reproduce the one-GPU masks bit for
    bit, and neither beats a rank that simply computes every mask itself (29 ms for ten thresholds: the masks come out of HBM
    faster than NVLink can deliver them) — under DDP call `state.select(...)` / `mask_flat` on every rank after the
    accumulator's all-reduce; this function is for consumers that shard the parameters (and for the north-star's histogram
    reduction: only the 256 KB histogram and the tie budgets cross ranks to fix the threshold)."""
    import torch.distributed as dist

    from . import distributed as D

    lay = state.layout
    world = D.world_size()
    if world == 1:
        sel = state.select(thresholds, scope)
        return [state.mask_flat(sel, j, packed=True) for j in range(len(thresholds))]
    if scope not in ("per_tensor", "global"):
        raise ValueError(f"unknown scope {scope!r}: 'per_tensor' or 'global'")
    rank = dist.get_rank(group)
    if reduce_acc:
        dist.all_reduce(state.acc, op=dist.ReduceOp.SUM, group=group)
    dev = state.device
    n_th = len(thresholds)

    if scope == "global":
        ranges = D.partition_range(lay.n_chunks, world)
        subs = [slice_layout(lay, *ranges[r]) for r in range(world)]
        mine, f0, f1 = subs[rank]
        sub = hist = None
        local_total = torch.zeros(N_BINS, dtype=torch.int64, device=dev)
        if mine.total:
            sub = SaliencyState(mine, dev, acc=state.acc[f0:f1])
            hist = sub.histogram()
            local_total = sub.hist_total(hist)
        total = D.reduce_histogram_(local_total.clone(), group)          # the NCCL histogram reduction (256 KB)

        def rank_budget(tau, budget):
            local_ties = torch.where(tau >= 0, local_total[tau.clamp(min=0)], torch.zeros_like(tau))
            return D.tie_budget_for_rank(local_ties, budget, group)      # exclusive prefix over ranks = flat index order

        if sub is not None:
            sel = sub.select_global(hist, total, sum(lay.numels), thresholds, rank_budget)
        else:  # a rank without a slice still takes part in the collectives
            tb = global_tau(total, sum(lay.numels), thresholds)
            rank_budget(tb[:, 0], tb[:, 1])
        width = max((b - a) // 8 for _, a, b in subs)
        out = []
        for t in range(n_th):
            send = torch.zeros(width, dtype=torch.uint8, device=dev)
            if sub is not None:
                sub.mask_flat(sel, t, out=send[: mine.total // 8], packed=True)
            recv = torch.empty(world * width, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(recv, send, group=group)
            full = torch.zeros(lay.total // 8, dtype=torch.uint8, device=dev)
            for r, (_, a, b) in enumerate(subs):
                if b > a:
                    full[a // 8: b // 8].copy_(recv[r * width: r * width + (b - a) // 8])
            out.append(full)
        state.launches += (sub.launches if sub is not None else 0)
        return out

    units = D.partition_units(lay.numels, world)
    mine = FlatLayout([(lay.names[i], lay.shapes[i]) for i in units[rank]])
    sub = SaliencyState(mine, dev)
    for j, i in enumerate(units[rank]):
        n = lay.numels[i]
        sub.acc[mine.offsets[j]: mine.offsets[j] + n].copy_(state.acc[lay.offsets[i]: lay.offsets[i] + n])
    sel = sub.select(thresholds) if units[rank] else None
    out = []
    for t in range(n_th):
        full = torch.zeros(lay.total // 8, dtype=torch.uint8, device=dev)
        if units[rank]:
            buf = sub.mask_flat(sel, t, packed=True)
            for j, i in enumerate(units[rank]):
                nb = (lay.numels[i] + 7) // 8
                full[lay.offsets[i] // 8: lay.offsets[i] // 8 + nb].copy_(buf[mine.offsets[j] // 8: mine.offsets[j] // 8 + nb])
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)         # disjoint bytes: the sum is the union
        out.append(full)
    state.launches += sub.launches
    return out


def global_tau(total: torch.Tensor, n_total: int, thresholds: Sequence[float]) -> torch.Tensor:
    """int64 [n_th, 2] = (threshold key tau or -1, tie budget) of the global ranking, from the total histogram."""
    n_th = len(thresholds)
    th = (C.c_double * n_th)(*[float(t) for t in thresholds])
    tb = torch.empty((n_th, 2), dtype=torch.int64, device=total.device)
    _lib.check(_lib.load().clear_saliency_select_global_tau(total.data_ptr(), int(n_total), th, n_th, tb.data_ptr(), _stream()),
               "clear_saliency_select_global_tau")
    return tb


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
    scope = getattr(cfg, "scope", "per_tensor")  # extension: "global" = one ranking over all tensors (upstream SalUn)
    for th, hard_dict in state.masks(REFERENCE_THRESHOLDS, scope):
        torch.save({k: v.cpu() for k, v in hard_dict.items()}, os.path.join(cfg.save_dir, f"with_{th}.pt"))


# --------------------------------------------------------------------------------------------------
# 8-bit block-wise optimiser states (the reference trains with optim="adamw_bnb_8bit",
# mm/forget.py:156 — bitsandbytes, absent here: parity unpinned, see csrc/adamw8.cu)
# --------------------------------------------------------------------------------------------------
ADAMW8_MIN_NUMEL = 4096   # smaller tensors (norm weights of tiny models, biases) keep fp32 moments, as bitsandbytes' min_8bit_size does


def dynamic_codebook(signed: bool) -> torch.Tensor:
    """256 sorted fp32 code values in [-1, 1] (signed: exp_avg) or [0, 1] (unsigned: exp_avg_sq) of the "dynamic" 8-bit data
    type of block-wise 8-bit optimisers: seven decades 1e-6 .. 1, decade d holding 2^d (signed: per sign) / 2^(d+1) (unsigned)
    values at the centres of equal cells of [0.1, 1] * 10^(d-6), plus 0 and 1 — fine resolution near full scale, seven
    orders of magnitude of dynamic range below it."""
    vals = [0.0, 1.0]
    for d in range(7):
        cells = 2 ** d if signed else 2 ** (d + 1)
        edges = torch.linspace(0.1, 1.0, cells + 1, dtype=torch.float32)
        centres = ((edges[:-1] + edges[1:]) / 2.0) * (10.0 ** (d - 6))
        vals += centres.tolist()
        if signed:
            vals += (-centres).tolist()
    code = torch.tensor(sorted(vals), dtype=torch.float32)
    assert code.numel() == 256
    return code


def quantisation_tables(code: torch.Tensor, signed: bool) -> torch.Tensor:
    """uint8 table of the kernel's quantiser: for every 16-bit pattern (sign, exponent, 7 mantissa bits; unsigned codebook:
    the 32768 non-negative ones) the index of the code value nearest to the float `pattern << 16` — the number of fp32
    midpoints of neighbouring code values below it.  The kernel rounds x / absmax to its top 16 bits and looks the index up."""
    mid = ((code[:-1] + code[1:]) * 0.5).to(torch.float32)
    bits = torch.arange(65536 if signed else 32768, dtype=torch.int64) << 16
    x = torch.where(bits >= 0x80000000, bits - (1 << 32), bits).to(torch.int32).view(torch.float32)
    x = torch.nan_to_num(x, nan=0.0)                 # NaN patterns never occur (|x / absmax| <= 1)
    return torch.searchsorted(mid, x.contiguous(), right=False).to(torch.uint8)


class _Adam8Tables:
    """Device copies of the two codebooks and their quantisation tables, one set per device."""

    _cache: dict = {}

    @classmethod
    def get(cls, device):
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in cls._cache:
            c1, c2 = dynamic_codebook(True), dynamic_codebook(False)
            cls._cache[key] = (c1.to(device), c2.to(device), quantisation_tables(c1, True).to(device), quantisation_tables(c2, False).to(device),
                               int(torch.nonzero(c1 == 0)[0]), int(torch.nonzero(c2 == 0)[0]))
        return cls._cache[key]


def masked_adamw8_flat(param: torch.Tensor, grad: torch.Tensor, state1: torch.Tensor, state2: torch.Tensor, absmax1: torch.Tensor,
                       absmax2: torch.Tensor, mask: torch.Tensor | None, step: int, lr=1e-5, betas=(0.9, 0.999), eps=1e-8,
                       weight_decay=0.01, grad_scale: torch.Tensor | None = None) -> None:
    """One launch of the 8-bit-state masked AdamW over a flat bf16 parameter buffer.  state1 / state2: uint8 [n] code indices
    (start: adamw8_init); absmax1 / absmax2: fp32 [ceil(n / 256)]."""
    c1, c2, q1, q2, _, _ = _Adam8Tables.get(param.device)
    bits = mask if (mask is not None and mask.dtype == torch.uint8) else None
    if bits is not None:
        mask = None
    a = _lib.AdamW8Args(param=param.data_ptr(), grad=grad.data_ptr(), state1=state1.data_ptr(), state2=state2.data_ptr(),
                        absmax1=absmax1.data_ptr(), absmax2=absmax2.data_ptr(), code1=c1.data_ptr(), code2=c2.data_ptr(),
                        qlut1=q1.data_ptr(), qlut2=q2.data_ptr(), mask=mask.data_ptr() if mask is not None else None,
                        mask_bits=bits.data_ptr() if bits is not None else None, n=param.numel(), lr=lr, beta1=betas[0], beta2=betas[1],
                        eps=eps, weight_decay=weight_decay, step=step,
                        grad_scale_dev=grad_scale.data_ptr() if grad_scale is not None else None)
    _lib.check(_lib.load().clear_masked_adamw_8bit(C.byref(a), _stream()), "clear_masked_adamw_8bit")


def adamw8_init(n: int, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(state1, state2, absmax1, absmax2) of n parameters at step 0: both moments exactly zero."""
    device = torch.device(device)
    _, _, _, _, z1, z2 = _Adam8Tables.get(device)
    nb = (n + 255) // 256
    return (torch.full((n,), z1, dtype=torch.uint8, device=device), torch.full((n,), z2, dtype=torch.uint8, device=device),
            torch.zeros(nb, dtype=torch.float32, device=device), torch.zeros(nb, dtype=torch.float32, device=device))


class MaskedAdamW(torch.optim.Optimizer):
    """AdamW whose gradient is multiplied by a SalUn mask first (upstream SalUn: `param.grad *= mask[name]`
    before `optimizer.step()`).  bf16 parameters and gradients, fp32 moments, torch's
    `_single_tensor_adam` operation order; one fused pass over 23 bytes per parameter."""

    def __init__(self, params, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, masks: dict | None = None,
                 max_grad_norm: float | None = None, state_bits: int = 32):
        """state_bits=8: block-wise 8-bit moments (2 bytes of optimiser state per parameter instead of 8; the scheme of the
        `adamw_bnb_8bit` optimiser the reference trains with, mm/forget.py:156) for tensors of at least ADAMW8_MIN_NUMEL
        elements; 32 (default): fp32 moments, bit-exact with torch.optim.AdamW's operation order."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if state_bits not in (8, 32):
            raise ValueError("state_bits must be 8 or 32")
        self.state_bits = state_bits
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
                eight = self.state_bits == 8 and p.numel() >= ADAMW8_MIN_NUMEL
                if not st:
                    st["step"] = 0
                    if eight:
                        st["state1"], st["state2"], st["absmax1"], st["absmax2"] = adamw8_init(p.numel(), p.device)
                    else:
                        st["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                        st["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                st["step"] += 1
                g = p.grad.to(torch.bfloat16).contiguous()
                m = self.masks.get(p)
                if eight:
                    masked_adamw8_flat(p.view(-1), g.view(-1), st["state1"], st["state2"], st["absmax1"], st["absmax2"], m,
                                       st["step"], lr=group["lr"], betas=(b1, b2), eps=group["eps"],
                                       weight_decay=group["weight_decay"], grad_scale=scale)
                    continue
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
