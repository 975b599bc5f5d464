"""The fused loss head: lm_head projection + row-wise loss + backward, over the C ABI.

Replaces, for one `model(**split)` call of the reference trainers, everything between "the trunk
produced hidden states" and "autograd filled dH and lm_head.weight.grad":
    nn.Linear lm_head            mm/trainer.py:180,184,188-189,...   (A1)
    HF ForCausalLMLoss           outputs.loss                        (A2)
    logits2probs                 mm/trainer_utils.py:16-21           (A3)
    kl_div(batchmean)            mm/trainer.py:203-208,237-242       (A4, A5)
    NPO ratio                    mm/trainer.py:348-353               (A6)
    entropy                      mm/trainer.py:195-196               (A6b)
    get_batch_loss               mm/trainer_utils.py:73-81           (A6c)
(IDs: SURVEY.md §8a.)  The reference materialises 2-6 fp32 `[B*T, V]` tensors per loss; here rows are
streamed in blocks of `block_rows`: one bf16 `[block_rows, V]` staging buffer per projected operand is
the only vocab-wide storage, the projection's epilogue emits the log-sum-exp statistics, one row pass
turns logits into dlogits in place, and two more tcgen05 contractions consume them.

Everything here runs on the CUDA library; there is no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import KIND_CE, KIND_ENTROPY, KIND_KL_FWD, KIND_KL_REV, KIND_NPO

IGNORE_INDEX = -100
KIND_NAMES = {KIND_CE: "ce", KIND_KL_FWD: "kl_fwd", KIND_KL_REV: "kl_rev", KIND_NPO: "npo", KIND_ENTROPY: "entropy"}
_NEEDS_REF = (KIND_KL_FWD, KIND_KL_REV, KIND_NPO)

# Rows staged per pass.  8192 rows x 32064 bf16 = 525 MB per projected operand, small against 180 GB of HBM3e.
# Measured on B200 under the 1 kW power cap (bench.py --block-rows, final build of round 1, NPO step at N = 8192):
# 8192-row blocks 6.46 ms, 4096-row blocks 6.93 ms, 2048-row blocks 7.43 ms — fewer blocks mean fewer re-reads of W,
# fewer read-modify-write passes over dW and fuller waves of output tiles in the dH contraction.
DEFAULT_BLOCK_ROWS = 8192
# Up to this many row blocks accumulate dW directly in bf16; beyond it an fp32 accumulator is used and converted once.
MAX_BF16_ACCUM_BLOCKS = 4


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class StageTimer:
    """Optional per-stage timing with CUDA events on the launching stream (bench.py's live roofline measurement).
    `with StageTimer() as t: ...; t.summary()` -> {stage: (launches, total ms)}.

    By default the timed invocations run the product path — the one-call C entry — and the LIBRARY records the events
    between its stages (ClearHeadArgs.stage_events): what is timed is what the trainers run.  staged=True drives the stage
    entry points one by one from Python instead (tests compare the two bit for bit)."""

    active: "StageTimer | None" = None
    _pool: list = []           # idle cudaEvent_t handles (creating one costs ~0.1 ms of host time: they are created once and reused)

    def __init__(self, staged: bool = False):
        self.staged = staged
        self.records = []      # (tag, torch event pair) or (tag, raw event pair)
        self._raw = []

    @classmethod
    def reserve(cls, n: int) -> None:
        """Make sure n idle events exist (call before a timed region so that none is created inside it)."""
        lib = _lib.load()
        while len(cls._pool) < n:
            e = C.c_void_p()
            _lib.check(lib.clear_event_create(C.byref(e)), "clear_event_create")
            cls._pool.append(e)

    def __enter__(self):
        StageTimer.active = self
        return self

    def __exit__(self, *exc):
        StageTimer.active = None

    def raw_events(self, n: int):
        """n cudaEvent_t handles of the library's CUDA runtime, owned by this timer."""
        StageTimer.reserve(n)
        out = [StageTimer._pool.pop() for _ in range(n)]
        self._raw.extend(out)
        return out

    def summary(self) -> dict:
        torch.cuda.synchronize()
        lib = _lib.load()
        out = {}
        for tag, a, b in self.records:
            if isinstance(a, C.c_void_p):
                ms = C.c_float()
                _lib.check(lib.clear_event_elapsed_ms(a, b, C.byref(ms)), "clear_event_elapsed_ms")
                dt = ms.value
            else:
                dt = a.elapsed_time(b)
            n, tot = out.get(tag, (0, 0.0))
            out[tag] = (n + 1, tot + dt)
        StageTimer._pool.extend(self._raw)   # back to the pool (never destroyed: a few KB for the life of the process)
        self._raw = []
        self.records = []
        return out


def _timed(tag: str, fn) -> None:
    t = StageTimer.active
    if t is None:
        fn()
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    t.records.append((tag, a, b))


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def shift_labels(labels: torch.Tensor) -> torch.Tensor:
    """Row (b, t) is scored against labels[b, t+1]; the last row of each sequence is ignored
    (HF ForCausalLMLoss: pad one -100 on the right, drop the first).  [B, T] -> [B*T] int64."""
    y = torch.nn.functional.pad(labels, (0, 1), value=IGNORE_INDEX)[..., 1:]
    return y.reshape(-1).contiguous()


def _check_2d_bf16(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the CLEAR loss head has no CPU path")
    if t.dtype != torch.bfloat16:
        raise RuntimeError(f"{name} must be bfloat16, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name} must be a 2-D row-major tensor, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


# --------------------------------------------------------------------------------------------------
# thin wrappers of the three stages (one C call each)
# --------------------------------------------------------------------------------------------------
def vtiles(vocab: int) -> int:
    return _lib.load().clear_head_vtiles(vocab)


def project(hidden: torch.Tensor, weight: torch.Tensor, logits: torch.Tensor, lse_part: torch.Tensor,
            cta_group: int = 0) -> None:
    """logits[:, :V] = hidden @ weight.T (bf16) and per-(row, 256-column tile) (max, sum-exp)."""
    lib = _lib.load()
    n, h = hidden.shape
    v = weight.shape[0]
    a = _lib.ProjectArgs(hidden=hidden.data_ptr(), weight=weight.data_ptr(), logits=logits.data_ptr(),
                         lse_part=lse_part.data_ptr(), hidden_ld=hidden.stride(0), weight_ld=weight.stride(0),
                         logits_ld=logits.stride(0), n_rows=n, hidden_dim=h, vocab=v, cta_group=cta_group)
    _timed("gemm_project", lambda: _lib.check(lib.clear_head_project(C.byref(a), _stream()), "clear_head_project"))


def project_pair(hidden, weight, logits, lse_part, ref_hidden, ref_weight, ref_logits, ref_lse_part, cta_group: int = 0) -> None:
    """Student and teacher projections of the same rows as ONE launch (clear_head_project_pair)."""
    lib = _lib.load()

    def args(hd, w, z, part):
        return _lib.ProjectArgs(hidden=hd.data_ptr(), weight=w.data_ptr(), logits=z.data_ptr(), lse_part=part.data_ptr(),
                                hidden_ld=hd.stride(0), weight_ld=w.stride(0), logits_ld=z.stride(0), n_rows=hd.shape[0],
                                hidden_dim=hd.shape[1], vocab=w.shape[0], cta_group=cta_group)

    a, b = args(hidden, weight, logits, lse_part), args(ref_hidden, ref_weight, ref_logits, ref_lse_part)
    _timed("gemm_project_pair", lambda: _lib.check(lib.clear_head_project_pair(C.byref(a), C.byref(b), _stream()), "clear_head_project_pair"))


def rowloss(kind: int, logits: torch.Tensor, lse_part: torch.Tensor, loss_rows: torch.Tensor, vocab: int, *,
            ref_logits: torch.Tensor | None = None, ref_lse_part: torch.Tensor | None = None,
            labels: torch.Tensor | None = None, row_weight: torch.Tensor | None = None,
            scale_dev: torch.Tensor | None = None, grad_scale: float = 1.0, beta: float = 1.0,
            write_grad: bool = True) -> None:
    lib = _lib.load()
    a = _lib.RowLossArgs(logits=logits.data_ptr(), lse_part=lse_part.data_ptr(), ref_logits=_ptr(ref_logits),
                         ref_lse_part=_ptr(ref_lse_part), labels=_ptr(labels), row_weight=_ptr(row_weight),
                         scale_dev=_ptr(scale_dev), loss_rows=loss_rows.data_ptr(), logits_ld=logits.stride(0),
                         ref_ld=ref_logits.stride(0) if ref_logits is not None else 0, n_rows=logits.shape[0],
                         vocab=vocab, kind=kind, grad_scale=grad_scale, beta=beta, write_grad=int(write_grad),
                         ref_logits_f32=int(ref_logits is not None and ref_logits.dtype == torch.float32))
    _timed("rowloss", lambda: _lib.check(lib.clear_head_rowloss(C.byref(a), _stream()), "clear_head_rowloss"))


def backward(dlogits: torch.Tensor, vocab: int, hidden: torch.Tensor, weight: torch.Tensor,
             dhidden: torch.Tensor | None, dweight: torch.Tensor | None, accumulate: bool = False,
             cta_group: int = 0, dhidden_ws: torch.Tensor | None = None, dweight_peers=None) -> None:
    """dweight_peers: optional list of the data-parallel ranks' accumulation-buffer pointers (distributed.
    FusedGradReducer): the dW contraction reduce-adds each output tile into its owner's buffer (fused reduce-scatter)."""
    lib = _lib.load()
    a = _lib.BackwardArgs(dlogits=dlogits.data_ptr(), hidden=hidden.data_ptr(), weight=weight.data_ptr(),
                          dhidden=_ptr(dhidden), dweight=_ptr(dweight), dlogits_ld=dlogits.stride(0),
                          hidden_ld=hidden.stride(0), weight_ld=weight.stride(0),
                          dhidden_ld=dhidden.stride(0) if dhidden is not None else 0,
                          dweight_ld=dweight.stride(0) if dweight is not None else hidden.shape[1], n_rows=dlogits.shape[0],
                          hidden_dim=hidden.shape[1], vocab=vocab,
                          dweight_f32=int(dweight is not None and dweight.dtype == torch.float32),
                          accumulate=int(accumulate), cta_group=cta_group, dhidden_ws=_ptr(dhidden_ws))
    if dweight_peers:
        if dweight is not None:
            raise RuntimeError("backward: pass dweight or dweight_peers, not both")
        a.n_peers = len(dweight_peers)
        for r, ptr in enumerate(dweight_peers):
            a.dweight_peers[r] = ptr
        dweight = True  # a dW contraction is requested (its output goes to the peers' buffers)
    # dH and dW together are ONE launch when both are bf16 (csrc/head.cu: clear_head_backward)
    tag = "gemm_dweight" if dhidden is None else ("gemm_dhidden" if dweight is None else "gemm_backward")
    _timed(tag, lambda: _lib.check(lib.clear_head_backward(C.byref(a), _stream()), "clear_head_backward"))


def count_valid(labels: torch.Tensor, inv_count: torch.Tensor) -> None:
    _lib.check(_lib.load().clear_head_count_valid(labels.data_ptr(), labels.numel(), inv_count.data_ptr(), _stream()),
               "clear_head_count_valid")


def reduce_rows(loss_rows: torch.Tensor, out: torch.Tensor, scale: float, scale_dev: torch.Tensor | None = None,
                accumulate: bool = False) -> None:
    _lib.check(_lib.load().clear_head_reduce_rows(loss_rows.data_ptr(), loss_rows.numel(), scale, _ptr(scale_dev),
                                                  out.data_ptr(), int(accumulate), _stream()),
               "clear_head_reduce_rows")


def reduce_sequences(loss_rows: torch.Tensor, n_seqs: int, seq_sums: torch.Tensor) -> None:
    _lib.check(_lib.load().clear_head_reduce_sequences(loss_rows.data_ptr(), n_seqs, loss_rows.numel() // n_seqs,
                                                       seq_sums.data_ptr(), _stream()),
               "clear_head_reduce_sequences")


def compact_rows(labels: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """(row_index int32 [N], count int32 [1]) on the device: ascending indices of the rows with a label != -100."""
    n = labels.numel()
    idx = torch.empty(n, dtype=torch.int32, device=labels.device)
    cnt = torch.empty(1, dtype=torch.int32, device=labels.device)
    _lib.check(_lib.load().clear_head_compact_rows(labels.data_ptr(), n, idx.data_ptr(), cnt.data_ptr(), _stream()),
               "clear_head_compact_rows")
    return idx, cnt


def _rows_move(fn_name: str, src: torch.Tensor, idx: torch.Tensor, n_idx: int, dst: torch.Tensor) -> None:
    """src / dst: 1-D or 2-D row-major tensors of one dtype; moves n_idx whole rows."""
    esz = src.element_size()
    width = src.shape[1] if src.dim() == 2 else 1
    s_ld = (src.stride(0) if src.dim() == 2 else 1) * esz
    d_ld = (dst.stride(0) if dst.dim() == 2 else 1) * esz
    fn = getattr(_lib.load(), fn_name)
    _lib.check(fn(src.data_ptr(), s_ld, idx.data_ptr(), n_idx, width * esz, dst.data_ptr(), d_ld, _stream()), fn_name)


def gather_rows(src: torch.Tensor, idx: torch.Tensor, n_idx: int) -> torch.Tensor:
    dst = torch.empty((n_idx,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if n_idx:
        _rows_move("clear_rows_gather", src, idx, n_idx, dst)
    return dst


def scatter_rows(src: torch.Tensor, idx: torch.Tensor, n_idx: int, n_rows: int) -> torch.Tensor:
    """zeros [n_rows, ...] with row idx[i] = src[i]."""
    dst = torch.zeros((n_rows,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if n_idx:
        _rows_move("clear_rows_scatter", src, idx, n_idx, dst)
    return dst


# --------------------------------------------------------------------------------------------------
# the fused forward+backward of one head invocation
# --------------------------------------------------------------------------------------------------
@dataclass
class HeadResult:
    loss: torch.Tensor                 # fp32 [1], already carries `coef` and the reference's normalisation
    seq_loss: torch.Tensor | None      # fp32 [n_seqs]: per-sequence sums of the unscaled row terms
    dhidden: torch.Tensor | None       # bf16 [N, H] = d loss / d hidden
    dweight: torch.Tensor | None       # bf16 [V, H], or the fp32 accumulation buffer that was passed in
    loss_rows: torch.Tensor            # fp32 [N] unscaled row terms
    launches: int                      # kernels of libclear_b200 launched by this call


def _normalisation(kind: int, reduction: str, n_rows: int, vocab: int, beta: float, n_seqs: int | None):
    """(loss factor, gradient factor, per_valid): host factors multiplying the summed row terms / the
    kernel's dterm, and whether 1/#valid (device scalar) also applies.  The two differ only for NPO,
    whose dterm s - p*sum(s) is the derivative of (1/beta) * sum log(1 + t).
    Reference semantics: CE mean over non-ignored rows; KL over ALL rows / N;
    NPO (2/beta) * mean over N*V; entropy raw sum; get_batch_loss per-sequence sums averaged over B."""
    gfac = beta if kind == KIND_NPO else 1.0
    if reduction == "sum":
        return 1.0, gfac, False
    if reduction == "seq_mean":
        if not n_seqs:
            raise ValueError("reduction='seq_mean' needs n_seqs")
        return 1.0 / n_seqs, gfac / n_seqs, False
    if reduction != "reference":
        raise ValueError(f"unknown reduction {reduction!r}")
    if kind == KIND_CE:
        return 1.0, 1.0, True
    if kind in (KIND_KL_FWD, KIND_KL_REV):
        return 1.0 / n_rows, 1.0 / n_rows, False
    if kind == KIND_NPO:
        return 2.0 / (beta * n_rows * vocab), 2.0 / (n_rows * vocab), False
    if kind == KIND_ENTROPY:
        return 1.0, 1.0, False
    raise ValueError(f"unknown kind {kind}")


def head_loss_fwd_bwd(hidden: torch.Tensor, weight: torch.Tensor, kind: int, *, labels: torch.Tensor | None = None,
                      ref_hidden: torch.Tensor | None = None, ref_weight: torch.Tensor | None = None,
                      ref_logits: torch.Tensor | None = None, beta: float = 1.0, coef: float = 1.0,
                      reduction: str = "reference", n_seqs: int | None = None,
                      row_weight: torch.Tensor | None = None, need_dhidden: bool = True, need_dweight: bool = True,
                      dweight_accum: torch.Tensor | None = None, block_rows: int | None = None,
                      cta_group: int = 0, overlap: bool = False, dweight_target=None,
                      compact_ignored: bool = False, row_index: torch.Tensor | None = None,
                      status_flags: torch.Tensor | None = None) -> HeadResult:
    """One head invocation: loss = coef * norm * sum_i term_i, plus d loss/d hidden and d loss/d weight.

    hidden [N, H] bf16, weight [V, H] bf16; labels [N] int64 ALREADY SHIFTED (see shift_labels) for CE;
    the teacher as (ref_hidden [N, H], ref_weight [V, H]) — projected here, block by block, never fully
    materialised — or as materialised ref_logits [N, >=V] bf16.
    dweight_accum: optional fp32 [V, H] buffer; dW is ADDED into it (gradient accumulation across
    splits / micro-batches happens in fp32 inside the contraction's epilogue) and returned as is.
    need_dhidden = need_dweight = False gives the forward-only mode (no dlogits are written).
    dweight_target: a `distributed.FusedGradReducer.begin()` target — dW is not returned; every tile of it is
    reduce-added into the owner rank's buffer by the contraction itself (HeadResult.dweight is None; the reduced,
    averaged gradient comes from `reducer.finish(target)`).
    compact_ignored (CE only): run the head on the rows whose label is not -100 only.  Ignored rows contribute nothing to
    the loss, to dW or to their own dH (their dH rows are returned as zeros), so loss and gradients are those of the full
    evaluation; in CLEAR's batches (padding + 576 image positions + question ignored) this removes most of the work.  The
    valid rows are compacted on the device and their count stays there (every kernel reads it from device memory): no host
    synchronisation.
    row_index: optional int32 [N] row map (`remove_image_tokens`, mm/trainer_utils.py:48-70, as a kernel input): row i of the
    invocation is row row_index[i] of hidden / ref_hidden, which may then have more than N rows; dhidden has hidden's shape,
    rows no index names are zero.  labels / ref_logits / row_weight / loss_rows have N entries.
    ref_logits may be bf16 or fp32.
    status_flags: optional int32 [1] device tensor, OR-ed: bit 0 = a CE label outside [0, V) (where torch would assert).
    overlap: issue the contractions and the row passes on two streams (see the staging comment below).  Off
    by default: under the 1 kW power cap the step is energy-bound and the overlap measured no gain on B200.
    """
    lib = _lib.load()
    hidden = _check_2d_bf16(hidden, "hidden")
    weight = _check_2d_bf16(weight, "weight")
    n_src, h = hidden.shape
    n = n_src
    v = weight.shape[0]
    if weight.shape[1] != h:
        raise RuntimeError(f"weight {tuple(weight.shape)} does not match hidden {tuple(hidden.shape)}")
    dev = hidden.device
    if row_index is not None:
        if not row_index.is_cuda or row_index.dim() != 1:
            raise RuntimeError("row_index must be a 1-D CUDA tensor of row numbers of `hidden`")
        row_index = row_index.to(torch.int32).contiguous()
        n = row_index.numel()
    if n == 0 or n_src == 0:
        raise RuntimeError("head_loss_fwd_bwd: empty batch (N = 0 rows)")
    needs_ref = kind in _NEEDS_REF
    if kind not in KIND_NAMES:
        raise ValueError(f"unknown loss kind {kind}")
    if needs_ref:
        if ref_logits is None and (ref_hidden is None or ref_weight is None):
            raise RuntimeError(f"kind {KIND_NAMES[kind]} needs the frozen teacher: pass ref_hidden+ref_weight or ref_logits")
        if ref_logits is None:
            ref_hidden = _check_2d_bf16(ref_hidden, "ref_hidden")
            ref_weight = _check_2d_bf16(ref_weight, "ref_weight")
            if ref_hidden.shape != hidden.shape or ref_weight.shape != weight.shape:
                raise RuntimeError("teacher shapes differ from the student's")
        else:
            if ref_logits.dtype == torch.float32:
                if not ref_logits.is_cuda or ref_logits.dim() != 2 or ref_logits.stride(1) != 1:
                    raise RuntimeError("ref_logits must be a 2-D row-major CUDA tensor")
            else:
                ref_logits = _check_2d_bf16(ref_logits, "ref_logits")
            if ref_logits.shape[0] != n or ref_logits.shape[1] < v:
                raise RuntimeError("ref_logits must be [N, >= V]")
            if ref_logits.stride(0) % 8 or ref_logits.data_ptr() % 16:
                ref_logits = torch.nn.functional.pad(ref_logits, (0, (-ref_logits.shape[1]) % 8)).contiguous()
    if kind == KIND_CE:
        if labels is None:
            raise RuntimeError("CE needs labels")
        if labels.dtype != torch.int64 or labels.numel() != n or not labels.is_cuda:
            raise RuntimeError("labels must be a CUDA int64 tensor with one (shifted) label per row")
        labels = labels.reshape(-1).contiguous()
    if row_weight is not None:
        row_weight = row_weight.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if row_weight.numel() != n:
            raise RuntimeError("row_weight must have one entry per row")

    compact_ignored = bool(compact_ignored) and kind == KIND_CE
    if status_flags is not None and (status_flags.dtype != torch.int32 or not status_flags.is_cuda or status_flags.numel() < 1):
        raise RuntimeError("status_flags must be a CUDA int32 tensor")
    mapped = compact_ignored or row_index is not None

    write_grad = need_dhidden or need_dweight
    blk = min(n, block_rows or DEFAULT_BLOCK_ROWS)
    n_blocks = (n + blk - 1) // blk
    ld = _round_up(v, 64)
    nt = vtiles(v)
    launches = 0

    host_scale, grad_scale, per_valid = _normalisation(kind, reduction, n, v, beta, n_seqs)
    timer = StageTimer.active
    pipelined = overlap and n_blocks >= 2 and timer is None and not mapped
    one_call = not pipelined and (timer is None or not timer.staged or mapped)   # row maps live in the one-call entry only
    inv_count = None
    if per_valid and not one_call:
        inv_count = torch.empty(1, dtype=torch.float32, device=dev)
        count_valid(labels, inv_count)
    launches += int(per_valid)

    # Staging: two student blocks (block b+1 is projected while block b is still being consumed) and one
    # teacher block.  The row pass is bound by the XU pipe and HBM, the contractions by the tensor pipe, so
    # they are issued on two streams and overlap: row pass of block b runs beside the student projection of
    # block b+1 and beside the backward contractions of block b-1.
    n_stage = 2 if pipelined else (0 if one_call else 1)   # the one-call entry point carves its staging from a workspace
    zs = [torch.empty((blk, ld), dtype=torch.bfloat16, device=dev) for _ in range(n_stage)]
    parts = [torch.empty((blk, nt, 2), dtype=torch.float32, device=dev) for _ in range(n_stage)]
    zr = rpart = None
    if needs_ref and not one_call:
        rpart = torch.empty((blk, nt, 2), dtype=torch.float32, device=dev)
        if ref_logits is None:
            zr = torch.empty((blk, ld), dtype=torch.bfloat16, device=dev)
    loss_rows = torch.empty(n, dtype=torch.float32, device=dev)
    dhidden = torch.empty_like(hidden) if need_dhidden else None     # [n_src, H]: the shape of `hidden`
    dweight = None
    peers = None
    if dweight_target is not None:
        if not need_dweight or dweight_accum is not None:
            raise RuntimeError("dweight_target needs need_dweight=True and excludes dweight_accum")
        if tuple(dweight_target.shape) != tuple(weight.shape):
            raise RuntimeError("dweight_target was created for a different weight shape")
        if cta_group not in (0, getattr(dweight_target, "cta_group", 0)):
            raise RuntimeError("dweight_target: the owners' slabs were cut for the default CTA grouping; pass cta_group=0")
        peers = dweight_target.peer_ptrs
    elif need_dweight:
        if dweight_accum is not None:
            if dweight_accum.dtype != torch.float32 or dweight_accum.shape != weight.shape or not dweight_accum.is_contiguous():
                raise RuntimeError("dweight_accum must be a contiguous fp32 [V, H] tensor")
            dweight = dweight_accum
        elif n_blocks <= MAX_BF16_ACCUM_BLOCKS:
            # bf16 straight from the epilogue; later blocks are added by TMA reduce-add, rounded to bf16 per add —
            # the precision of torch's own bf16 gradient accumulation across micro-batches
            dweight = torch.empty_like(weight)
        else:
            dweight = torch.empty(weight.shape, dtype=torch.float32, device=dev)

    paired_proj = needs_ref and ref_logits is None                      # student + teacher projection: one launch
    dw_f32 = need_dweight and (dweight_accum is not None or (peers is None and dweight is not None and dweight.dtype == torch.float32))
    # dH + dW: one launch whenever both are bf16 (csrc/head.cu: clear_head_backward; CLEAR_GEMM_GROUP changes it for diagnostics)
    paired_bwd = need_dhidden and need_dweight and not dw_f32
    n_bwd = 0 if not write_grad else (1 if paired_bwd else int(need_dhidden) + int(need_dweight))
    launches += n_blocks * (2 + int(needs_ref and not paired_proj) + n_bwd) + 1 + int(bool(n_seqs))
    if compact_ignored:   # compact, gather labels (+ row weights), gather hidden, scatter loss terms (+ dH); a dH contraction on its
        # own (fp32 dW beside it, or no dW) runs over an fp32 scratch: zero + convert per block
        launches += (4 + int(row_weight is not None) + int(row_index is not None)
                     + int(need_dhidden) * (1 + (0 if paired_bwd else 2 * n_blocks)))
    elif row_index is not None:
        launches += 1 + int(needs_ref and ref_logits is None) + int(need_dhidden)
    if one_call:
        # The whole invocation in ONE C call (clear_head_loss_fwd_bwd): the block loop, the normalisations and the
        # workspace carving run in the library; this function only owns the tensors.
        if n_seqs and n % n_seqs:
            raise RuntimeError("n_seqs must divide the number of rows")
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        seq_loss = torch.empty(n_seqs, dtype=torch.float32, device=dev) if n_seqs else None
        red = {"reference": _lib.REDUCE_REFERENCE, "sum": _lib.REDUCE_SUM, "seq_mean": _lib.REDUCE_SEQ_MEAN}[reduction]
        a = _lib.HeadArgs(hidden=hidden.data_ptr(), weight=weight.data_ptr(), labels=_ptr(labels) if kind == KIND_CE else None,
                          ref_hidden=_ptr(ref_hidden) if (needs_ref and ref_logits is None) else None,
                          ref_weight=_ptr(ref_weight) if (needs_ref and ref_logits is None) else None,
                          ref_logits=_ptr(ref_logits) if needs_ref else None, row_weight=_ptr(row_weight),
                          loss=loss.data_ptr(), loss_rows=loss_rows.data_ptr(), seq_loss=_ptr(seq_loss), dhidden=_ptr(dhidden),
                          dweight=_ptr(dweight), n_peers=len(peers) if peers else 0, hidden_ld=hidden.stride(0),
                          weight_ld=weight.stride(0),
                          ref_hidden_ld=ref_hidden.stride(0) if (needs_ref and ref_logits is None) else 0,
                          ref_weight_ld=ref_weight.stride(0) if (needs_ref and ref_logits is None) else 0,
                          ref_logits_ld=ref_logits.stride(0) if (needs_ref and ref_logits is not None) else 0,
                          dhidden_ld=dhidden.stride(0) if dhidden is not None else 0,
                          dweight_ld=dweight.stride(0) if dweight is not None else h, n_rows=n, hidden_dim=h, vocab=v, kind=kind,
                          beta=beta, coef=coef, reduction=red, n_seqs=n_seqs or 0, block_rows=blk,
                          dweight_f32=int(dweight is not None and dweight.dtype == torch.float32),
                          dweight_accumulate=int(dweight_accum is not None), cta_group=cta_group,
                          row_index=_ptr(row_index), n_src_rows=n_src, compact_ignored=int(compact_ignored),
                          ref_logits_f32=int(needs_ref and ref_logits is not None and ref_logits.dtype == torch.float32),
                          status_flags=_ptr(status_flags))
        if peers:
            for r, ptr in enumerate(peers):
                a.dweight_peers[r] = ptr
            # owners zeroed their slabs, all ranks agreed: the library waits on this event right before the dW contraction,
            # so the projections and the row pass do not wait for the previous step's all-gather
            a.dweight_ready_event = dweight_target.ready.cuda_event
        need = C.c_size_t()
        _lib.check(lib.clear_head_loss_workspace_bytes(C.byref(a), C.byref(need)), "clear_head_loss_workspace_bytes")
        ws = torch.empty(max(need.value, 256), dtype=torch.uint8, device=dev)
        if timer is not None:   # the library records events between its stages: 4 per block of rows
            evs = timer.raw_events(4 * n_blocks)
            table = (C.c_void_p * len(evs))(*[e.value for e in evs])
            a.stage_events = C.cast(table, C.c_void_p)
            a.n_stage_events = len(evs)
        _lib.check(lib.clear_head_loss_fwd_bwd(C.byref(a), ws.data_ptr(), need.value, _stream()), "clear_head_loss_fwd_bwd")
        if timer is not None:
            proj_tag = "gemm_project_pair" if (needs_ref and ref_logits is None) else ("gemm_project+ref_lse" if needs_ref else "gemm_project")
            for b in range(n_blocks):
                e = evs[4 * b: 4 * b + 4]
                timer.records.append((proj_tag, e[0], e[1]))
                timer.records.append(("rowloss", e[1], e[2]))
                if write_grad:
                    tag = "gemm_backward" if (need_dhidden and need_dweight) else ("gemm_dhidden" if need_dhidden else "gemm_dweight")
                    timer.records.append((tag, e[2], e[3]))
        if need_dweight and peers is None and dweight_accum is None and dweight.dtype != torch.bfloat16:
            dweight = dweight.to(torch.bfloat16)
        return HeadResult(loss, seq_loss, dhidden, dweight if peers is None else None, loss_rows, launches)

    # fp32 scratch for a split-K dH contraction (used only when the library finds the plain one wave-starved)
    dh_ws = torch.empty((blk, h), dtype=torch.float32, device=dev) if need_dhidden and h % 8 == 0 else None

    def block(b):
        r0, r1 = b * blk, min(n, (b + 1) * blk)
        return r0, r1, zs[b % n_stage][: r1 - r0], parts[b % n_stage][: r1 - r0]

    def do_project(b):
        r0, r1, zb, pb = block(b)
        project(hidden[r0:r1], weight, zb, pb, cta_group)

    def do_teacher(b):
        r0, r1, _, _ = block(b)
        if ref_logits is None:
            project(ref_hidden[r0:r1], ref_weight, zr[: r1 - r0], rpart[: r1 - r0], cta_group)
        else:
            _ref_lse(ref_logits[r0:r1], rpart[: r1 - r0], v)

    def do_project_both(b):   # student + teacher projection: one launch
        r0, r1, zb, pb = block(b)
        project_pair(hidden[r0:r1], weight, zb, pb, ref_hidden[r0:r1], ref_weight, zr[: r1 - r0], rpart[: r1 - r0], cta_group)

    def do_rowloss(b):
        r0, r1, zb, pb = block(b)
        zrb = rpb = None
        if needs_ref:
            rpb = rpart[: r1 - r0]
            zrb = zr[: r1 - r0] if ref_logits is None else ref_logits[r0:r1]
        rowloss(kind, zb, pb, loss_rows[r0:r1], v, ref_logits=zrb, ref_lse_part=rpb,
                labels=labels[r0:r1] if labels is not None and kind == KIND_CE else None,
                row_weight=row_weight[r0:r1] if row_weight is not None else None, scale_dev=inv_count,
                grad_scale=coef * grad_scale, beta=beta, write_grad=write_grad)

    def do_backward(b):
        r0, r1, zb, _ = block(b)
        if peers is not None and b == 0:
            torch.cuda.current_stream().wait_event(dweight_target.ready)  # owners zeroed their slabs, all ranks agreed
        backward(zb, v, hidden[r0:r1], weight, dhidden[r0:r1] if need_dhidden else None,
                 dweight if (need_dweight and peers is None) else None, accumulate=(dweight_accum is not None) or b > 0,
                 cta_group=cta_group, dhidden_ws=dh_ws, dweight_peers=peers)

    if not pipelined:
        for b in range(n_blocks):
            if needs_ref and ref_logits is None:
                do_project_both(b)
            else:
                do_project(b)
                if needs_ref:
                    do_teacher(b)
            do_rowloss(b)
            if write_grad:
                do_backward(b)
    else:
        main = torch.cuda.current_stream()
        gs = _gemm_stream(dev)
        gs.wait_stream(main)
        row_done = [None] * n_blocks
        for b in range(n_blocks + 1):
            if b < n_blocks:
                with torch.cuda.stream(gs):
                    do_project(b)                                   # beside the row pass of block b-1
                    if needs_ref:
                        if b >= 1:
                            gs.wait_event(row_done[b - 1])          # the teacher block is free again
                        do_teacher(b)
                    projected = torch.cuda.Event()
                    projected.record(gs)
                main.wait_event(projected)
                do_rowloss(b)                                       # beside the backward of block b-1
                row_done[b] = torch.cuda.Event()
                row_done[b].record(main)
            if b >= 1 and write_grad:
                with torch.cuda.stream(gs):
                    gs.wait_event(row_done[b - 1])
                    do_backward(b - 1)
        main.wait_stream(gs)

    loss = torch.empty(1, dtype=torch.float32, device=dev)
    reduce_rows(loss_rows, loss, coef * host_scale, inv_count)
    seq_loss = None
    if n_seqs:
        if n % n_seqs:
            raise RuntimeError("n_seqs must divide the number of rows")
        seq_loss = torch.empty(n_seqs, dtype=torch.float32, device=dev)
        reduce_sequences(loss_rows, n_seqs, seq_loss)
    if need_dweight and peers is None and dweight_accum is None and dweight.dtype != torch.bfloat16:
        dweight = dweight.to(torch.bfloat16)
    del lib
    return HeadResult(loss, seq_loss, dhidden, dweight, loss_rows, launches)


_GEMM_STREAMS: dict = {}


def _gemm_stream(dev: torch.device) -> "torch.cuda.Stream":
    """High-priority side stream for the contractions (one per device)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _GEMM_STREAMS:
        _GEMM_STREAMS[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _GEMM_STREAMS[key]


def _ref_lse(ref_logits: torch.Tensor, lse_part: torch.Tensor, vocab: int) -> None:
    """(max, sum-exp) partials of materialised teacher logits (bf16 or fp32), in the layout the projection emits."""
    lib = _lib.load()
    fn = lib.clear_head_lse_partials_f32 if ref_logits.dtype == torch.float32 else lib.clear_head_lse_partials
    _lib.check(fn(ref_logits.data_ptr(), ref_logits.stride(0), ref_logits.shape[0], vocab, lse_part.data_ptr(), _stream()),
               "clear_head_lse_partials")


# --------------------------------------------------------------------------------------------------
# torch.library operator: the PyTorch-visible form of clear_head_loss_fwd_bwd (SURVEY.md §8b "inner operator boundary")
# --------------------------------------------------------------------------------------------------
@torch.library.custom_op("clear_b200::head_loss", mutates_args=())
def head_loss_op(hidden: torch.Tensor, weight: torch.Tensor, kind: int, labels: torch.Tensor | None,
                 ref_hidden: torch.Tensor | None, ref_weight: torch.Tensor | None, ref_logits: torch.Tensor | None,
                 row_weight: torch.Tensor | None, beta: float, coef: float, reduction: str, n_seqs: int, block_rows: int,
                 need_dhidden: bool, need_dweight: bool, compact_ignored: bool,
                 row_index: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(loss [1] fp32, seq_loss [n_seqs] fp32, dhidden [rows of hidden, H] bf16, dweight [V, H] bf16, loss_rows [N] fp32) of
    one head invocation; outputs that were not requested come back as empty tensors.  hidden [N or more, H], weight [V, H]
    bf16 CUDA; labels shifted, flat; row_index int32 [N] optional row map into hidden.  n_seqs = 0 /
    block_rows = 0 mean "none" / "default"."""
    res = head_loss_fwd_bwd(hidden, weight, kind, labels=labels, ref_hidden=ref_hidden, ref_weight=ref_weight,
                            ref_logits=ref_logits, beta=beta, coef=coef, reduction=reduction, n_seqs=n_seqs or None,
                            row_weight=row_weight, need_dhidden=need_dhidden, need_dweight=need_dweight,
                            block_rows=block_rows or None, compact_ignored=compact_ignored, row_index=row_index)
    none = lambda dt: torch.empty(0, dtype=dt, device=hidden.device)
    return (res.loss, res.seq_loss if res.seq_loss is not None else none(torch.float32),
            res.dhidden if res.dhidden is not None else none(torch.bfloat16),
            res.dweight if res.dweight is not None else none(torch.bfloat16), res.loss_rows)


@head_loss_op.register_fake
def _(hidden, weight, kind, labels, ref_hidden, ref_weight, ref_logits, row_weight, beta, coef, reduction, n_seqs, block_rows,
      need_dhidden, need_dweight, compact_ignored, row_index):
    f32 = dict(dtype=torch.float32, device=hidden.device)
    return (torch.empty(1, **f32), torch.empty(n_seqs, **f32),
            torch.empty_like(hidden) if need_dhidden else torch.empty(0, dtype=torch.bfloat16, device=hidden.device),
            torch.empty_like(weight) if need_dweight else torch.empty(0, dtype=torch.bfloat16, device=hidden.device),
            torch.empty(hidden.shape[0] if row_index is None else row_index.shape[0], **f32))


# --------------------------------------------------------------------------------------------------
# autograd boundary
# --------------------------------------------------------------------------------------------------
_DW_REDUCER = None


def set_dweight_reducer(reducer) -> None:
    """Route d loss / d lm_head.weight of every differentiable head invocation through a
    `distributed.FusedGradReducer` (mean over the data-parallel ranks, fused into the contraction).  The parameter
    must then be excluded from DDP's own reducer (`trainer.enable_fused_head_allreduce` does both).  None = off."""
    global _DW_REDUCER
    _DW_REDUCER = reducer


class _FusedHeadLoss(torch.autograd.Function):
    """loss = f(hidden, weight).  Forward runs the fused forward+backward (dlogits exist only inside the
    staging block), backward scales the stored dH / dW by the incoming scalar gradient."""

    @staticmethod
    def forward(ctx, hidden, weight, kind, labels, ref_hidden, ref_weight, ref_logits, beta, coef, reduction, n_seqs,
                row_weight, block_rows, compact_ignored, row_index):
        shape = hidden.shape
        h2 = hidden.reshape(-1, shape[-1])
        cast = h2.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16
        hb = h2.to(torch.bfloat16).contiguous()
        wb = weight.to(torch.bfloat16).contiguous()
        rh = ref_hidden.reshape(-1, shape[-1]).to(torch.bfloat16).contiguous() if ref_hidden is not None else None
        rw = ref_weight.to(torch.bfloat16).contiguous() if ref_weight is not None else None
        rl = None
        if ref_logits is not None:   # fp32 teacher logits are consumed as fp32 (no rounding to bf16), anything else as bf16
            rl = ref_logits.reshape(-1, ref_logits.shape[-1])
            rl = (rl if rl.dtype == torch.float32 else rl.to(torch.bfloat16)).contiguous()
        reducer = _DW_REDUCER if (ctx.needs_input_grad[1] and _DW_REDUCER is not None
                                  and tuple(_DW_REDUCER.shape) == tuple(wb.shape)) else None
        if reducer is None:
            # through the registered operator (torch.ops.clear_b200.head_loss)
            loss, seq_loss, dh, dw, _rows = torch.ops.clear_b200.head_loss(
                hb, wb, kind, labels, rh, rw, rl, row_weight, beta, coef, reduction, n_seqs or 0, block_rows or 0,
                ctx.needs_input_grad[0], ctx.needs_input_grad[1], compact_ignored, row_index)
            res = HeadResult(loss, seq_loss if n_seqs else None, dh if dh.numel() else None, dw if dw.numel() else None, _rows, 0)
        else:
            # the fused data-parallel reducer hands the contraction a set of peer buffers: a Python object, outside the
            # operator schema
            target = reducer.begin()
            res = head_loss_fwd_bwd(hb, wb, kind, labels=labels, ref_hidden=rh, ref_weight=rw, ref_logits=rl, beta=beta,
                                    coef=coef, reduction=reduction, n_seqs=n_seqs, row_weight=row_weight,
                                    need_dhidden=ctx.needs_input_grad[0], need_dweight=ctx.needs_input_grad[1],
                                    block_rows=block_rows, dweight_target=target, compact_ignored=compact_ignored,
                                    row_index=row_index)
        dh = res.dhidden.reshape(shape) if res.dhidden is not None else None
        dw = res.dweight
        ctx.fused = None
        if reducer is not None:
            pin = reducer.can_hold()
            if pin:
                # autograd keeps the reduced gradient where the all-gather wrote it: the buffer is pinned until backward ran;
                # the main stream does not wait for the all-gather here (backward does), so it overlaps what follows
                reducer.hold(target.index)
            done, buf = reducer.finish(target)
            if pin:
                ctx.fused = (reducer, target.index, done)
                dw = buf
            else:
                torch.cuda.current_stream().wait_event(done)
                dw = buf.clone()  # more invocations in flight than buffers to pin: take a private copy
        if cast:
            dh = dh.to(hidden.dtype) if dh is not None else None
            if ctx.fused is None:
                dw = dw.to(weight.dtype) if dw is not None else None
        ctx.weight_dtype = weight.dtype
        ctx.save_for_backward(dh, dw)
        ctx.mark_non_differentiable(*( [res.seq_loss] if res.seq_loss is not None else [] ))
        if res.seq_loss is not None:
            return res.loss.reshape(()), res.seq_loss
        return res.loss.reshape(())

    @staticmethod
    def backward(ctx, g, *unused):
        dh, dw = ctx.saved_tensors
        g = g.to(torch.float32)
        gh = (dh * g.to(dh.dtype)) if dh is not None else None
        if ctx.fused is not None:
            reducer, index, done = ctx.fused
            torch.cuda.current_stream().wait_event(done)        # the all-gather of this invocation's gradient
            gw = (dw * g.to(dw.dtype)).to(ctx.weight_dtype)     # a new tensor: the pinned buffer can rotate again
            reducer.release(index)
            ctx.fused = None
        else:
            gw = (dw * g.to(dw.dtype)) if dw is not None else None
        return (gh, gw) + (None,) * 13


def fused_head_loss(hidden: torch.Tensor, weight: torch.Tensor, kind: int, *, labels: torch.Tensor | None = None,
                    ref_hidden: torch.Tensor | None = None, ref_weight: torch.Tensor | None = None,
                    ref_logits: torch.Tensor | None = None, beta: float = 1.0, coef: float = 1.0,
                    reduction: str = "reference", n_seqs: int | None = None, row_weight: torch.Tensor | None = None,
                    block_rows: int | None = None, compact_ignored: bool = False, row_index: torch.Tensor | None = None):
    """Differentiable loss of one head invocation.  hidden [..., H]; labels already shifted, flat; row_index: optional flat
    int32 row map into hidden.reshape(-1, H) (see head_loss_fwd_bwd).
    Returns a 0-dim fp32 tensor (and the per-sequence sums when n_seqs is given)."""
    return _FusedHeadLoss.apply(hidden, weight, kind, labels, ref_hidden, ref_weight, ref_logits, float(beta),
                                float(coef), reduction, n_seqs, row_weight, block_rows, bool(compact_ignored), row_index)


@torch.no_grad()
def sequence_nll(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, block_rows: int | None = None,
                 row_index: torch.Tensor | None = None):
    """Forward-only `get_batch_loss` (mm/trainer_utils.py:73-81; used by mm/eval.py:64-92): per-sequence
    SUM of shifted token CE.  hidden [B, T', H], labels [B, T] (unshifted); row_index: optional flat int32 [B * T] row map
    into hidden.reshape(B * T', H) (T' > T: `remove_image_tokens`).  Returns fp32 [B]."""
    b, tp, h = hidden.shape
    res = head_loss_fwd_bwd(hidden.reshape(b * tp, h).to(torch.bfloat16).contiguous(),
                            weight.to(torch.bfloat16).contiguous(), KIND_CE, labels=shift_labels(labels),
                            reduction="seq_mean", n_seqs=b, need_dhidden=False, need_dweight=False,
                            block_rows=block_rows, compact_ignored=True, row_index=row_index)
    return res.seq_loss
