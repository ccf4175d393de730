"""Optimizer step on the flat gradient buckets.

The reference trainers call ``torch.nn.utils.clip_grad_norm_(model.parameters(), gradient_clip)`` and ``AdamW.step()``
(train_direct_4gpu.py:72-80; config_direct.json: lr 1e-4, weight_decay 0.01, gradient_clip 1.0).  ``dp.GradientBuckets`` already holds
every gradient as a view into a few flat fp32 buffers; ``FlatAdamW`` lays the parameters and both Adam moments out the same way
(``p.data`` becomes a view, so modules, ``state_dict`` and checkpoints are unaffected) and does the whole step with one
sum-of-squares kernel and one fused clip+AdamW kernel per bucket (``csrc/hvc_optim.cu``).  The step count lives on the device: the
step is CUDA-graph capturable.
"""
import ctypes as C

import torch

from . import _lib, ops
from .kernels import _need_cuda, _ptr, _stream


class FlatAdamW(torch.optim.Optimizer):
    """A ``torch.optim.Optimizer`` (so ``CosineAnnealingLR`` and friends drive its learning rate, train_direct_4gpu.py:165-168) whose
    ``state_dict()`` / ``load_state_dict()`` speak the ``torch.optim.AdamW`` layout: a checkpoint written through it resumes in the
    reference trainer and vice versa (checkpoint.flat_to_torch_state / torch_to_flat_state).  `params`: the parameter list a torch
    optimizer would have been built from, in order (default: the bucketed parameters in registration order; pass
    ``model.parameters()`` when some are frozen so that the state indices match the reference's)."""

    def __init__(self, buckets, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0, params=None):
        self.gb = buckets
        order = list(params) if params is not None else list(reversed([p for m in buckets._members for p in m]))
        super().__init__(order, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._order = order
        self.max_grad_norm = max_grad_norm
        self.params, self.exp_avg, self.exp_avg_sq = [], [], []
        for flat_g, members, offsets in zip(buckets.buckets, buckets._members, buckets._offsets):
            _need_cuda(flat_g)
            flat_p = torch.zeros_like(flat_g)          # (alignment padding between members stays zero through every update)
            with torch.no_grad():
                for p, off in zip(members, offsets):
                    assert p.dtype == torch.float32, "FlatAdamW expects fp32 master parameters"
                    view = flat_p[off:off + p.numel()].view_as(p)
                    view.copy_(p.data)
                    p.data = view                      # same values, new storage: the bucket layout
            self.params.append(flat_p)
            self.exp_avg.append(torch.zeros_like(flat_g))
            self.exp_avg_sq.append(torch.zeros_like(flat_g))
        dev = buckets.buckets[0].device
        self.t = torch.zeros(1, device=dev, dtype=torch.float32)            # step count t (on the device: graph-capturable)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)        # ||g||^2 over all buckets
        ops.clear_weight_cache()

    # hyper-parameters live in param_groups[0] (where torch's LR schedulers write them)
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    betas = property(lambda self: self.param_groups[0]["betas"], lambda self, v: self.param_groups[0].__setitem__("betas", v))
    eps = property(lambda self: self.param_groups[0]["eps"], lambda self, v: self.param_groups[0].__setitem__("eps", v))
    weight_decay = property(lambda self: self.param_groups[0]["weight_decay"], lambda self, v: self.param_groups[0].__setitem__("weight_decay", v))

    def zero_grad(self, set_to_none=False):
        """The gradients ARE the buckets: zero them in place and re-arm the all-reduce counters (never detach p.grad)."""
        self.gb.reset()

    @torch.no_grad()
    def step(self, closure=None):
        """Clip (global L2 norm over all bucketed gradients, as clip_grad_norm_) + AdamW.  Call after GradientBuckets.finish()."""
        assert closure is None, "FlatAdamW does not re-evaluate the model"
        lib, st = _lib.lib(), _stream()
        _lib.check(lib.hvc_adamw_tick(_ptr(self.t), st), "hvc_adamw_tick")
        clip = self.max_grad_norm > 0
        if clip:
            self.sumsq.zero_()
            for g in self.gb.buckets:
                _lib.check(lib.hvc_sumsq_f32(_ptr(g), C.c_int64(g.numel()), _ptr(self.sumsq), st), "hvc_sumsq_f32")
        f = C.c_float

        def tick(pt, gt, mt, vt):
            _lib.check(lib.hvc_adamw_flat(_ptr(pt), _ptr(gt), _ptr(mt), _ptr(vt), C.c_int64(pt.numel()), f(self.lr), f(self.betas[0]),
                                          f(self.betas[1]), f(self.eps), f(self.weight_decay), f(self.max_grad_norm if clip else 0.0),
                                          _ptr(self.sumsq), _ptr(self.t), st), "hvc_adamw_flat")

        for i, (p, g, m, v) in enumerate(zip(self.params, self.gb.buckets, self.exp_avg, self.exp_avg_sq)):
            members, offsets, got = self.gb._members[i], self.gb._offsets[i], self.gb.touched(i)
            if len(got) == len(members):
                tick(p, g, m, v)                       # the usual case: one launch per bucket
                continue
            # a trainable parameter that took no part in this step (its slot holds zeros, not a gradient): torch.optim.AdamW skips
            # parameters whose grad is None -- no weight decay, no moment decay -- so only the members that got a gradient are updated
            for k in sorted(got):
                lo, n = offsets[k], members[k].numel()
                tick(p[lo:lo + n], g[lo:lo + n], m[lo:lo + n], v[lo:lo + n])
        ops.clear_weight_cache()       # the kernels write the parameters behind autograd's version counters

    def grad_norm(self):
        """Global gradient norm of the last step (device tensor; what clip_grad_norm_ returns)."""
        return self.sumsq.sqrt()

    def state_dict(self):
        """The ``torch.optim.AdamW`` layout (per-parameter step / exp_avg / exp_avg_sq, indexed by position in the parameter list)."""
        from .checkpoint import flat_to_torch_state
        return flat_to_torch_state(self, self._order)

    def load_state_dict(self, sd):
        from .checkpoint import torch_to_flat_state
        torch_to_flat_state(self, sd, self._order)
