"""Batch-sharded data parallelism for the backbone: bucketed gradient all-reduce overlapped with backward.

Replaces the reference's ``DistributedDataParallel`` wrapper (train_direct_4gpu.py:146,
train_progressive_4gpu.py:238).  One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch;
gloo in the CPU tests) is the plumbing.  Gradients live in a few flat fp32 buckets (``p.grad`` are views),
filled in reverse registration order -- the order backward produces them; when the last gradient of a
bucket has been accumulated its all-reduce is launched asynchronously, so communication of early
buckets overlaps the remaining backward kernels.  The only exchange per step is this all-reduce; samples
never cross ranks.

Unused parameters.  Frozen parameters (``requires_grad=False``) are not bucketed.  A TRAINABLE parameter that
takes no part in a step (``ProgressiveCascadeModel(xrays, max_stage=1)`` leaves stages 2-3 trainable but
unused, which is why train_progressive_4gpu.py:238 passes ``find_unused_parameters=True``) never fires its
hook, so its bucket is not complete when backward ends: ``finish()`` reduces every such bucket (the unused
members contribute the zeros ``reset()`` wrote, as DDP does for unused parameters), so the used members that
share the bucket are still averaged and replicas cannot drift.  Complete buckets are reduced in completion order and
the leftovers in index order, so -- as with DDP -- every rank must leave the SAME parameters unused in a step.  ``touched(i)`` tells the optimizer which
members received a gradient (``FlatAdamW`` skips the others like ``torch.optim.AdamW`` skips ``grad is None``).

Buffers.  ``DistributedDataParallel(broadcast_buffers=True)`` (the default the reference trainers run with)
copies rank 0's buffers -- the X-ray encoder's BatchNorm running statistics, models/diagnostic_losses.py:84-94
-- to every rank before each forward; ``broadcast_buffers(module)`` does the same with one flat broadcast per dtype.
"""
import contextlib
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientBuckets:
    ALIGN = 64      # elements: every member starts on a 256-byte boundary (kernels read parameters / gradients with 16-byte accesses)

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 25 << 20,
                 process_group: Optional[dist.ProcessGroup] = None, average: bool = True):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.average = average
        # NCCL averages inside the collective; gloo has no AVG, there the sum is scaled afterwards
        self._avg_in_collective = bool(average and self.world > 1 and dist.get_backend(process_group) == "nccl")
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
        self._touched = [set() for _ in self.buckets]
        self._counted = [set() for _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._works = []
        self._sync = True
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.reset()

    def _close(self, members):
        offsets, off = [], 0
        for p in members:
            offsets.append(off)
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        flat = torch.zeros(off, device=members[0].device, dtype=torch.float32)     # the padding stays zero
        idx = len(self.buckets)
        for k, (p, o) in enumerate(zip(members, offsets)):
            p.grad = flat[o:o + p.numel()].view_as(p)   # autograd accumulates in place into the bucket
            self._bucket_of[p] = (idx, k)
        self.buckets.append(flat)
        self._members.append(members)
        self._offsets.append(offsets)

    # ------------------------------------------------------------------ start-up / per-forward replica sync
    def broadcast_parameters(self, params, src: int = 0):
        """One-time replica sync at start-up (what the DDP constructor does)."""
        if self.world > 1:
            with torch.no_grad():
                for p in params:
                    dist.broadcast(p, src=src, group=self.group)      # in place on p (bumps its version counter)
        _clear_derived_weights()

    def broadcast_buffers(self, module: torch.nn.Module, src: int = 0):
        """DDP's ``broadcast_buffers=True``: rank `src`'s buffers (BatchNorm running_mean / running_var / num_batches_tracked)
        replace every rank's before a forward.  One flat broadcast per dtype."""
        if self.world == 1:
            return
        by_dtype = {}
        for b in module.buffers():
            by_dtype.setdefault(b.dtype, []).append(b)
        with torch.no_grad():
            for bufs in by_dtype.values():
                flat = torch.cat([b.reshape(-1) for b in bufs])
                dist.broadcast(flat, src=src, group=self.group)
                off = 0
                for b in bufs:
                    b.copy_(flat[off:off + b.numel()].view_as(b))
                    off += b.numel()

    # ------------------------------------------------------------------ one step
    def reset(self):
        """Zero the buckets and re-arm the counters: call where the reference calls optimizer.zero_grad()."""
        for b in self.buckets:
            b.zero_()
        for i, m in enumerate(self._members):
            self._pending[i] = len(m)
            self._touched[i].clear()
            self._counted[i].clear()
            self._launched[i] = False
        self._works = []

    @contextlib.contextmanager
    def no_sync(self):
        """Gradient accumulation: backward passes inside this context add into the buckets without arming the all-reduce
        (DDP.no_sync()).  The last micro-batch runs outside it; finish() then reduces the accumulated sums."""
        self._sync = False
        try:
            yield
        finally:
            self._sync = True

    def _on_grad(self, p):
        i, k = self._bucket_of[p]
        self._touched[i].add(k)
        if not self._sync:
            return                          # accumulation pass: the synchronising pass counts this parameter
        if k in self._counted[i]:
            raise RuntimeError("GradientBuckets: a second backward reached a parameter before reset(); wrap the earlier "
                               "micro-batches in no_sync() or call reset() once per step")
        self._counted[i].add(k)
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        self._launched[i] = True
        if self.world == 1:
            return
        self._rebind(i)
        op = dist.ReduceOp.AVG if self._avg_in_collective else dist.ReduceOp.SUM
        self._works.append((i, dist.all_reduce(self.buckets[i], op=op, group=self.group, async_op=True)))

    def _rebind(self, i):
        """If something replaced p.grad (e.g. zero_grad(set_to_none=True)), copy back into the bucket views."""
        flat = self.buckets[i]
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        for p, off in zip(self._members[i], self._offsets[i]):
            if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                continue
            view = flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view

    def finish(self):
        """Reduce the buckets that backward left incomplete (unused trainable parameters: their slots hold zeros), then make the
        current stream wait for every all-reduce (the host does not block on NCCL)."""
        for i in range(len(self.buckets)):
            if not self._launched[i]:
                self._launch(i)
        for i, w in self._works:
            w.wait()
            if self.average and not self._avg_in_collective:
                self.buckets[i].mul_(1.0 / self.world)
        self._works = []

    def touched(self, i):
        """Indices (into the bucket's member list) of the parameters that received a gradient in this step.  If no hook fired at
        all since reset() the gradients were written into the buckets by hand (no backward ran): every member counts."""
        if not any(self._touched):
            return set(range(len(self._members[i])))
        return self._touched[i]

    def all_touched(self):
        return all(len(self.touched(i)) == len(m) for i, m in enumerate(self._members))

    def remove(self):
        for h in self._hooks:
            h.remove()


def _clear_derived_weights():
    """bf16 operand copies are cached per parameter version (ops.w16); a broadcast through ``p.data`` or a load behind autograd's
    back would leave them stale, so every replica sync drops them."""
    from . import ops
    ops.clear_weight_cache()
