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

    def __init__(self, index: int, shape, peer_ptrs, ready):
        self.index = index
        self.shape = tuple(shape)
        self.peer_ptrs = peer_ptrs   # device pointers of every rank's accumulation buffer, by rank
        self.ready = ready           # CUDA event: owners' slabs are zero and every rank has reached this step


class FusedGradReducer:
    """The dW gradient all-reduce fused into the contraction that produces it (one box, NVLink 5 / NVSwitch).

    Every rank allocates its accumulation buffers in symmetric memory (torch.distributed._symmetric_memory: the
    plumbing that maps each rank's buffer into every process and creates the NVSwitch multicast mapping).  dW is cut
    into contiguous row slabs of whole output tiles, one owner rank per slab.  Then, per step:

      begin()   side stream: zero this rank's slab, cross-rank barrier.
      (head)    the tcgen05 dW contraction's epilogue TMA-reduce-adds every output tile into the OWNER's buffer over
                NVLink (bf16 adds, like DDP's bf16 gradient all-reduce) — the reduce-scatter happens tile by tile while
                the tensor cores work on the next tile; 7/8 of the tiles leave the GPU, none needs an extra pass.
      finish()  side stream: barrier (every rank's contraction is done), then a small kernel writes this rank's reduced
                slab, scaled by 1/world, to EVERY rank's buffer with multimem.st through the multicast mapping (the
                switch replicates; 1/world of dW leaves each GPU once), barrier.  It uses no shared memory and at most
                32 CTAs, so it co-resides with the next step's persistent contractions instead of displacing them.

    The result is bit-identical on all ranks (each element is accumulated at one place, then broadcast).  The order of
    the bf16 adds at the owner is not fixed, so results can differ in the last bf16 bit between runs; use GradReducer
    (NCCL ring) when run-to-run bitwise reproducibility matters more than overlap.
    Two buffers alternate, so the gradient of step i stays valid while step i + 1 is being reduced."""

    def __init__(self, shape, device, group=None, average: bool = True, n_buffers: int = 2):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm

        from . import _lib

        if not dist.is_initialized() or dist.get_world_size(group) < 2:
            raise RuntimeError("FusedGradReducer needs an initialised process group with at least 2 ranks")
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
        r0, r1 = C.c_int32(), C.c_int32()
        _lib.check(self.lib.clear_head_dweight_slab(vocab, self.world, self.rank, 0, C.byref(r0), C.byref(r1)), "clear_head_dweight_slab")
        self.row0, self.row1 = r0.value, r1.value
        self.bufs, self.handles = [], []
        for _ in range(n_buffers):
            t = symm.empty(vocab * hidden, dtype=torch.bfloat16, device=self.device)
            self.handles.append(symm.rendezvous(t, self.group))
            self.bufs.append(t.view(vocab, hidden))
        self.stream = torch.cuda.Stream(device=self.device)
        self.count = 0

    def begin(self) -> FusedGradTarget:
        """Call any time before the step's head invocation (the earlier, the more slack the barrier has)."""
        k = self.count % len(self.bufs)
        self.count += 1
        hdl, buf = self.handles[k], self.bufs[k]
        self.stream.wait_stream(torch.cuda.current_stream())   # whoever consumed this buffer's previous gradient is done
        with torch.cuda.stream(self.stream):
            if self.row1 > self.row0:
                buf[self.row0:self.row1].zero_()
            hdl.barrier(channel=0)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return FusedGradTarget(k, self.shape, [int(p) for p in hdl.buffer_ptrs], ready)

    def finish(self, target: FusedGradTarget):
        """After the head invocation.  Returns (event, dW): dW [V, H] bf16 = mean over ranks, valid once the event
        has completed (`torch.cuda.current_stream().wait_event(event)`), until the buffer is reused two steps later."""
        C, _lib = self._C, self._lib
        hdl, buf = self.handles[target.index], self.bufs[target.index]
        self.stream.wait_stream(torch.cuda.current_stream())   # this rank's contraction has been issued
        with torch.cuda.stream(self.stream):
            hdl.barrier(channel=0)                              # every rank's tiles have landed at their owners
            mc = int(hdl.multicast_ptr or 0)                    # 0: no NVSwitch multicast on this system -> peer stores
            if os.environ.get("CLEAR_NO_MULTICAST") == "1":     # test hook for the peer-store path
                mc = 0
            table = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            _lib.check(self.lib.clear_dweight_allgather(buf.data_ptr(), mc or None, table, self.world, buf.stride(0), self.row0,
                                                        self.row1, self.shape[1], self.scale,
                                                        C.c_void_p(self.stream.cuda_stream)), "clear_dweight_allgather")
            hdl.barrier(channel=0)                              # every owner's slab has landed everywhere
            done = torch.cuda.Event()
            done.record(self.stream)
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
