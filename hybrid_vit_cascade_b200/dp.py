"""Batch-sharded data parallelism for the backbone: bucketed gradient all-reduce overlapped with backward.

Replaces the reference's ``DistributedDataParallel`` wrapper (train_direct_4gpu.py:146,
train_progressive_4gpu.py:238).  One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch;
gloo in the CPU tests) is the plumbing.  Gradients live in a few flat fp32 buckets (``p.grad`` are views),
filled in reverse registration order -- the order backward produces them; when the last gradient of a
bucket has been accumulated its all-reduce is launched asynchronously, so communication of early
buckets overlaps the remaining backward kernels.  Frozen parameters (``requires_grad=False``) are simply
not bucketed, which replaces ``find_unused_parameters=True``.  The only exchange per step is this
all-reduce; samples never cross ranks.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientBuckets:
    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        params = [p for p in params if p.requires_grad]
        params = list(reversed(params))                     # backward visits the last layers first
        self.buckets: List[torch.Tensor] = []
        self._members: List[List[torch.nn.Parameter]] = []
        self._offsets: List[List[int]] = []
        self._bucket_of = {}
        cur, cur_bytes = [], 0
        for p in params:
            nbytes = (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.reset()

    ALIGN = 64      # elements: every member starts on a 256-byte boundary (kernels read parameters / gradients with 16-byte accesses)

    def _close(self, members):
        offsets, off = [], 0
        for p in members:
            offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        flat = torch.zeros(off, device=members[0].device, dtype=torch.float32)     # the padding stays zero
        idx = len(self.buckets)
        for p, o in zip(members, offsets):
            p.grad = flat[o:o + p.numel()].view_as(p)   # autograd accumulates in place into the bucket
            self._bucket_of[p] = idx
        self.buckets.append(flat)
        self._members.append(members)
        self._offsets.append(offsets)

    def broadcast_parameters(self, params, src: int = 0):
        """One-time replica sync at start-up (what the DDP constructor does)."""
        if self.world > 1:
            for p in params:
                dist.broadcast(p.data, src=src, group=self.group)

    def reset(self):
        """Zero the buckets and re-arm the counters: call where the reference calls optimizer.zero_grad()."""
        for b in self.buckets:
            b.zero_()
        for i, m in enumerate(self._members):
            self._pending[i] = len(m)
        self._works = []

    def _on_grad(self, p):
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self.world > 1:
            if p.grad.data_ptr() < self.buckets[i].data_ptr() or \
                    p.grad.data_ptr() >= self.buckets[i].data_ptr() + self.buckets[i].numel() * 4:
                self._rebind(i)
            op = dist.ReduceOp.SUM
            self._works.append((i, dist.all_reduce(self.buckets[i], op=op, group=self.group, async_op=True)))

    def _rebind(self, i):
        """If something replaced p.grad (e.g. zero_grad(set_to_none=True)), copy back into the bucket views."""
        flat = self.buckets[i]
        for p, off in zip(self._members[i], self._offsets[i]):
            view = flat[off:off + p.numel()].view_as(p)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view

    def finish(self):
        """Wait for the outstanding all-reduces (the current stream waits; the host does not block on NCCL)."""
        for i, w in self._works:
            w.wait()
            if self.average:
                self.buckets[i].mul_(1.0 / self.world)
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()
