"""Data-parallel plumbing for the hot path (reference: DistributedDataParallel in rnnt/train.py:25-33,67-68).

Utterances shard across ranks with no data-path collective; the only exchange is one NCCL all-reduce of the joint
(and predictor) weight gradients per step, issued on a side stream so it overlaps the rest of the backward.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous equal split of utterance indices (per-rank batch must be equal for mean-loss parity with DDP)."""
    per = n_items // world_size
    if per * world_size != n_items:
        raise ValueError(f"global batch {n_items} must divide by world size {world_size}")
    return rank * per, (rank + 1) * per


def balanced_assignment(cells: Sequence[int], world_size: int):
    """Greedy longest-first assignment of utterances to ranks by lattice cells T_b*(U_b+1) (ragged batches)."""
    order = sorted(range(len(cells)), key=lambda i: -cells[i])
    loads = [0] * world_size
    buckets = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        buckets[r].append(i)
        loads[r] += cells[i]
    return [sorted(b) for b in buckets], loads


class GradAllReducer:
    """Flattens a fixed set of gradients into one bucket and all-reduces it (sum or average) on a side stream."""

    def __init__(self, params: Iterable[torch.nn.Parameter], average: bool = True, group=None):
        self.params = [p for p in params]
        self.average = average
        self.group = group
        self._flat = None
        self._stream = None   # created lazily on the first CUDA all-reduce

    def numel(self) -> int:
        return sum(p.numel() for p in self.params)

    def all_reduce_grads(self, grads: Sequence[torch.Tensor] = None, wait: bool = True):
        grads = [p.grad for p in self.params] if grads is None else list(grads)
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return grads
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
        if grads[0].is_cuda and self._stream is None:
            self._stream = torch.cuda.Stream(device=grads[0].device)
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self._stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            off = 0
            for g in grads:
                self._flat[off:off + g.numel()].copy_(g.reshape(-1))
                off += g.numel()
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self._flat.div_(dist.get_world_size(self.group))
            off = 0
            for g in grads:
                g.copy_(self._flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        if wait:
            self.wait()
        return grads

    def wait(self):
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
