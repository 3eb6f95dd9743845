"""Data-parallel gradient reduction for the one-process-per-GPU layout of the reference (main.py:54-57:
``gpus=N, strategy="ddp"``).  The path is pure data parallelism: the only collective is
the gradient all-reduce.

`torch.nn.parallel.DistributedDataParallel` works unchanged with these modules (every kernel launches on the
current stream and never synchronises).  `FlatGradients` / `allreduce_gradients` are the lean alternative bench.py
uses: every ``.grad`` is a view into ONE flat fp32 buffer (autograd accumulates into it in place), so the reduction is a
single in-place ``ncclAllReduce(avg)`` over NVLink with no bucket copies, issued after backward -- it does not compete
for SMs with the persistent one-CTA-per-SM GEMM / attention kernels, whose static tile schedules stall when NCCL's CTAs
hold some of the SMs they were launched for."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradients:
    """owns one flat gradient buffer per (device, dtype) and points every parameter's ``.grad`` into it"""

    def __init__(self, params: Iterable[torch.nn.Parameter]) -> None:
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.buffers = {}
        groups = {}
        for p in self.params:
            groups.setdefault((p.device, p.dtype), []).append(p)
        self.views = {}
        for key, ps in groups.items():
            total = sum((p.numel() + 3) // 4 * 4 for p in ps)          # 16-byte aligned slices
            buf = torch.zeros(total, device=key[0], dtype=key[1])
            off = 0
            for p in ps:
                self.views[p] = buf[off:off + p.numel()].view_as(p)
                off += (p.numel() + 3) // 4 * 4
            self.buffers[key] = buf

    def zero_(self) -> None:
        """reset the gradients (instead of ``p.grad = None``) and re-attach the views"""
        for buf in self.buffers.values():
            buf.zero_()
        for p in self.params:
            p.grad = self.views[p]

    def allreduce(self, group: Optional[dist.ProcessGroup] = None) -> None:
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        for buf in self.buffers.values():
            dist.all_reduce(buf, op=dist.ReduceOp.AVG if buf.is_cuda else dist.ReduceOp.SUM, group=group)
            if not buf.is_cuda:                                         # gloo has no AVG
                buf.div_(dist.get_world_size(group))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> None:
    """all-reduce (mean) the .grad of every parameter in one flat buffer per (device, dtype).  Stateless variant of
    `FlatGradients` for gradients that already exist as separate tensors (one flatten + one un-flatten copy)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    buckets = {}
    for p in params:
        if p.grad is not None:
            buckets.setdefault((p.grad.device, p.grad.dtype), []).append(p.grad)
    for grads in buckets.values():
        flat = torch._utils._flatten_dense_tensors(grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
        for g, synced in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
            g.copy_(synced)
