"""Data-parallel plumbing for the hot path (reference: DistributedDataParallel in rnnt/train.py:25-33,67-68).

Utterances shard across ranks with no data-path collective; the only exchange is one NCCL all-reduce of the joint
(and predictor) weight gradients per step, issued on a side stream so it overlaps the rest of the backward.
"""
from __future__ import annotations

import os
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


class WeightGradBucket:
    """One flat fp32 bucket for the step's weight gradients, all-reduced while the backward is still running.

    Layout: [joint_ln.weight (V*H) | joint_ln.bias (V) | extra (e.g. the predictor's gradients)].  Registered with
    `rnnt_b200.functional.set_weight_grad_sink`, the fused backward writes dW / db STRAIGHT into the first two slices
    (no flatten / unflatten copies; autograd hands the slices to `.grad` as they are) and records a CUDA event the
    moment they are final -- before its activation-gradient GEMM is enqueued.  `weight_grads_enqueued` then launches the
    NCCL all-reduce of the joint slice on a side stream gated by that event, so the collective overlaps the dh GEMM
    (DDP's bucket overlap, rnnt/train.py:67-68, for the one bucket this path owns).  `finish()` all-reduces the `extra`
    slice (gradients that only exist after the joint's backward, e.g. the predictor's), applies the averaging and joins
    the side stream."""

    def __init__(self, V: int, H: int, extra: int, device, average: bool = True, group=None):
        self.V, self.H, self.extra = V, H, extra
        self.device = torch.device(device)
        self.flat = torch.zeros(V * H + V + extra, dtype=torch.float32, device=self.device)
        self.average, self.group = average, group
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self.event = None
        if self.stream is not None:
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(self.device))   # materialises the cudaEvent_t handle
        self.pending = False
        # "overlap": joint slice all-reduced as soon as dW/db are final (under the dh GEMM), the rest at finish();
        # "single": one all-reduce of the whole bucket at finish() on the side stream; "inline": same on the compute stream
        self.mode = os.environ.get("RNNT_B200_BUCKET_MODE", "overlap")
        self.time_collectives = False      # bench.py: bracket every collective with CUDA events on the side stream
        self._timing = []
        self._exposed = []

    @property
    def joint_slice(self):
        return self.flat[: self.V * self.H + self.V]

    @property
    def extra_slice(self):
        return self.flat[self.V * self.H + self.V:]

    def _world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _all_reduce(self, t):
        if self.time_collectives and self.stream is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            e1.record(torch.cuda.current_stream(self.device))
            self._timing.append((e0, e1))
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def reset_timing(self):
        self._timing = []
        self._exposed = []

    def exposed_ms(self) -> float:
        """Summed time the compute stream waited in finish() since reset_timing() (synchronises)."""
        torch.cuda.current_stream(self.device).synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self._exposed))

    def collective_ms(self) -> float:
        """Summed device time of the all-reduces since reset_timing() (synchronises)."""
        if self.stream is not None:
            self.stream.synchronize()
        return float(sum(a.elapsed_time(b) for a, b in self._timing))

    # ---- sink protocol used by rnnt_b200.functional._FusedJointLoss.backward
    def accepts(self, V, H, device) -> bool:
        return V == self.V and H == self.H and torch.device(device) == self.device

    def weight_grad_views(self, V, H):
        n = V * H
        ev = self.event.cuda_event if self.event is not None else None
        return self.flat[:n].view(V, H), self.flat[n:n + V], ev

    def weight_grads_enqueued(self):
        """Called right after the backward's kernels were enqueued: dW / db are final once `event` fires."""
        if self._world() == 1 or self.mode != "overlap":
            return
        if self.stream is not None:
            self.stream.wait_event(self.event)
            with torch.cuda.stream(self.stream):
                self._all_reduce(self.joint_slice)
        else:
            self._all_reduce(self.joint_slice)
        self.pending = True

    def finish(self):
        """End of the step: all-reduce what is left (the `extra` slice; the joint slice too if no fused backward fed
        it), average, and make the current stream wait for the result."""
        world = self._world()
        if world == 1:
            return self.flat
        import contextlib
        if self.mode == "none":            # diagnostics only: no collective at all (what rank skew alone costs)
            self.pending = False
            return self.flat
        if self.mode == "inline" and self.stream is not None:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                self.flat.div_(world)
            self.pending = False
            return self.flat
        cur = torch.cuda.current_stream(self.device) if self.stream is not None else None
        if self.stream is not None:
            self.stream.wait_stream(cur)
            ctx = torch.cuda.stream(self.stream)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            rest = self.extra_slice if self.pending else self.flat
            if rest.numel():
                self._all_reduce(rest)
            if self.average:
                self.flat.div_(world)
        if self.stream is not None:
            if self.time_collectives:      # how long the compute stream stalls for the collectives = their EXPOSED time
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(cur)
                cur.wait_stream(self.stream)
                b.record(cur)
                self._exposed.append((a, b))
            else:
                cur.wait_stream(self.stream)
        self.pending = False
        return self.flat


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
