"""Data-parallel plumbing: one process per GPU, torch.distributed over NCCL (NVLink 5 / NVSwitch).

The reference has no collective of its own: `accelerate launch` wraps the model in DDP (or DeepSpeed
ZeRO-2), whose reducer averages gradients over ranks (mm/forget.py:23,66-67,161).  The
loss head shards the same way — rows are independent given W, every rank owns its micro-batch and a
replica of W — so the only exchange on the path is the gradient all-reduce of dW; SalUn adds the
reduction of the saliency histogram when the threshold is global, and nothing when it is per tensor
(tensors are independent units and are simply dealt out to the ranks).

Everything here is device-agnostic torch.distributed (the CPU tests run it over gloo with 2 ranks).
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun / accelerate environment variables
    (`WORLD_SIZE`, `RANK`, `LOCAL_RANK`, `MASTER_ADDR`, `MASTER_PORT`; mm/forget.py:23-31)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


class GradReducer:
    """DDP-style gradient averaging of the head's dW, issued on a side stream so that it overlaps whatever
    the compute stream does next (the next head invocation, or the trunk's backward).

        h = reducer.launch(dweight)      # after the head's backward wrote dweight
        ...                              # more compute on the current stream
        reducer.wait(h)                  # before the optimiser (or anything else) reads dweight
    """

    def __init__(self, average: bool = True, group=None):
        self.average = average
        self.group = group
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def launch(self, t: torch.Tensor):
        if world_size() == 1:
            return None
        if self.stream is None:  # CPU / gloo
            dist.all_reduce(t, group=self.group)
            if self.average:
                t.div_(world_size())
            return None
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
            dist.all_reduce(t, op=op, group=self.group)
            t.record_stream(self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return ev

    def wait(self, handle) -> None:
        if handle is not None:
            torch.cuda.current_stream().wait_event(handle)


class FusedGradTarget:
    """One in-flight fused all-reduce: where the dW contraction of this step sends its tiles."""

    def __init__(self, index: int, shape, peer_ptrs, ready, cta_group: int = 0):
        self.index = index
        self.shape = tuple(shape)
        self.peer_ptrs = peer_ptrs   # device pointers of every rank's accumulation buffer, by rank
        self.ready = ready           # CUDA event: owners' slabs are zero on every rank (waited on right before the dW contraction)
        self.cta_group = cta_group   # CTA grouping the owners' slabs were cut for


class FusedGradReducer:
    """The dW gradient all-reduce fused into the contraction that produces it (one box, NVLink 5 / NVSwitch).

    Every rank allocates its accumulation buffers in symmetric memory (torch.distributed._symmetric_memory: the
    plumbing that maps each rank's buffer into every process and creates the NVSwitch multicast mapping).  dW is cut
    into contiguous row slabs of whole output tiles, one owner rank per slab.  Then, per step:

      begin()   no device work: names the buffer of this step and the event after which its slabs are zero everywhere.
      (head)    the projections, the row pass and the dH contraction run without waiting for anything; right before the dW
                contraction the stream waits on that event (ClearHeadArgs.dweight_ready_event); the tcgen05 dW
                contraction's epilogue TMA-reduce-adds every output tile into the OWNER's buffer over NVLink (bf16 adds, like
                DDP's bf16 gradient all-reduce) — the reduce-scatter happens tile by tile while the tensor cores work on
                the next tile; (R-1)/R of the tiles leave the GPU, none needs an extra pass.
      finish()  side stream, behind this rank's dW contraction: barrier (every rank's tiles have landed at their owners); a
                small kernel writes this rank's reduced slab, scaled by 1/world, to EVERY rank's buffer with multimem.st
                through the multicast mapping (the switch replicates; 1/world of dW leaves each GPU once); this rank's slab
                of the NEXT step's buffer is zeroed; barrier (all-gather landed, next slabs zero, everywhere).  Two
                barriers per step; all of it overlaps the next step's projections (the all-gather kernel uses no shared
                memory and at most 32 CTAs, so it co-resides with the persistent contractions instead of displacing them).

    The result is bit-identical on all ranks (each element is accumulated at one place, then broadcast).  The order of
    the bf16 adds at the owner is not fixed, so results can differ in the last bf16 bit between runs; use GradReducer
    (NCCL ring) when run-to-run bitwise reproducibility matters more than overlap.
    Buffers rotate: the gradient returned by finish() stays valid until its buffer is picked as the "next" one, i.e. for
    n_buffers - 2 further steps; hold()/release() pin a buffer for longer (autograd keeps dW until backward)."""

    def __init__(self, shape, device, group=None, average: bool = True, n_buffers: int = 3):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            raise RuntimeError("FusedGradReducer needs an initialised process group with at least 2 ranks")
        if n_buffers < 3:
            raise RuntimeError("FusedGradReducer rotates at least 3 buffers (in use / gradient still read / being zeroed)")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("FusedGradReducer shards over at most 8 ranks (one NVSwitch box)")
        self.shape = tuple(shape)
        vocab, hidden = self.shape
        self.device = torch.device(device)
        self.scale = 1.0 / self.world if average else 1.0
        self.lib = _lib.load()
        self._lib = _lib
        self._C = C
        self.cta_group = 0
        r0, r1 = C.c_int32(), C.c_int32()
        _lib.check(self.lib.clear_head_dweight_slab(vocab, self.world, self.rank, self.cta_group, C.byref(r0), C.byref(r1)),
                   "clear_head_dweight_slab")
        self.row0, self.row1 = r0.value, r1.value
        self.bufs, self.handles = [], []
        for _ in range(n_buffers):
            t = symm.empty(vocab * hidden, dtype=torch.bfloat16, device=self.device)
            self.handles.append(symm.rendezvous(t, self.group))
            self.bufs.append(t.view(vocab, hidden))
        self.stream = torch.cuda.Stream(device=self.device)
        self.held = [0] * n_buffers
        self.next_index = 0
        # once: every slab of every buffer zero on every rank
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for buf in self.bufs:
                if self.row1 > self.row0:
                    buf[self.row0:self.row1].zero_()
            self.handles[0].barrier(channel=0)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready = [ev] * n_buffers

    # ---- buffer lifetime ------------------------------------------------------------------------------
    def can_hold(self) -> bool:
        """True when one more buffer may be pinned and two stay free (the one in use next and the one being zeroed)."""
        return sum(1 for h in self.held if h) + 1 <= len(self.bufs) - 2

    def hold(self, index: int) -> None:
        self.held[index] += 1

    def release(self, index: int) -> None:
        self.held[index] = max(0, self.held[index] - 1)

    def _pick_next(self, k: int) -> int:
        n = len(self.bufs)
        for d in range(1, n):
            c = (k + d) % n
            if not self.held[c]:
                return c
        raise RuntimeError("FusedGradReducer: every other buffer is pinned (hold() without release())")

    # ---- per step ---------------------------------------------------------------------------------------
    def begin(self) -> FusedGradTarget:
        k = self.next_index
        return FusedGradTarget(k, self.shape, [int(p) for p in self.handles[k].buffer_ptrs], self._ready[k], self.cta_group)

    def finish(self, target: FusedGradTarget):
        """After the head invocation.  Returns (event, dW): dW [V, H] bf16 = mean over ranks, valid once the event
        has completed (`torch.cuda.current_stream().wait_event(event)`), until the buffer is reused (see the class docstring)."""
        C, _lib = self._C, self._lib
        k = target.index
        hdl, buf = self.handles[k], self.bufs[k]
        nxt = self._pick_next(k)
        issued = torch.cuda.Event()
        issued.record(torch.cuda.current_stream())              # this rank's contraction (and everything before it) is issued
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(issued)
            hdl.barrier(channel=0)                              # every rank's tiles have landed at their owners
            mc = int(hdl.multicast_ptr or 0)                    # 0: no NVSwitch multicast on this system -> peer stores
            if os.environ.get("CLEAR_NO_MULTICAST") == "1":     # test hook for the peer-store path
                mc = 0
            table = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            _lib.check(self.lib.clear_dweight_allgather(buf.data_ptr(), mc or None, table, self.world, buf.stride(0), self.row0,
                                                        self.row1, self.shape[1], self.scale,
                                                        C.c_void_p(self.stream.cuda_stream)), "clear_dweight_allgather")
            if self.row1 > self.row0:
                self.bufs[nxt][self.row0:self.row1].zero_()     # the next step's slab (whoever read its old gradient ran before `issued`)
            hdl.barrier(channel=0)                              # all-gather landed and next slabs zero, on every rank
            done = torch.cuda.Event()
            done.record(self.stream)
        self._ready[nxt] = done
        self.next_index = nxt
        return done, buf


def partition_units(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Deals independent units (parameter tensors, for the per-tensor SalUn threshold) out to ranks:
    largest first onto the least loaded rank.  Deterministic, identical on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    return [sorted(x) for x in out]


def partition_range(n: int, world: int) -> list[tuple[int, int]]:
    """Cuts [0, n) into `world` contiguous ranges of nearly equal length (the global-threshold SalUn scope: rank r owns
    the mask chunks [lo, hi) of the flat gradient buffer, i.e. a 1/R slice in flat index order).  Identical on every rank."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def reduce_histogram_(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of per-rank saliency histograms (global-threshold scope: every rank histogrammed its slice of
    the flat gradient buffer).  Counts are exact integers, so the reduction order does not matter."""
    if world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def tie_budget_for_rank(local_ties: torch.Tensor, budget: torch.Tensor, group=None) -> torch.Tensor:
    """Stable tie-breaking across ranks that own consecutive slices of one tensor: of the `budget`
    elements equal to the threshold key that are to be selected, lower flat indices win, i.e. lower ranks
    first.  local_ties[j] = this rank's count of such elements for threshold j; returns how many of them
    this rank selects.  One small all-gather."""
    if world_size() == 1:
        return torch.minimum(local_ties, budget)
    gathered = [torch.empty_like(local_ties) for _ in range(world_size())]
    dist.all_gather(gathered, local_ties, group=group)
    before = torch.zeros_like(local_ties)
    for r in range(dist.get_rank()):
        before += gathered[r]
    return torch.clamp(budget - before, min=0).minimum(local_ties)
