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
