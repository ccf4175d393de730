"""Drop-in replacements for the reference's ``models/vit_components.py``.

Same class names, constructor arguments, forward signatures, attribute names and ``state_dict``
keys/shapes as the reference, so checkpoints load strictly in both directions; the forward and
backward passes run on the sm_100a kernels of libhvc_sm100a.so (see ops.py).  The ``nn.Linear`` /
``nn.Dropout`` children are kept as parameter/config containers only -- their ``forward`` is never
called on the hot path.
"""
import math

import torch
import torch.nn as nn

import contextlib

from . import kernels as K
from . import ops
from . import ops_fp32

_DROPOUT_POLICY = {"mode": "apply"}
_PRECISION = {"mode": "bf16"}


def set_precision(mode):
    """'bf16' (default): the production path -- bf16 tensor-core operands, fp32 accumulation/statistics/residual stream
    (what torch.autocast(bfloat16) does to the reference).  'fp32': verification mode -- the same kernels fed with
    three-term bf16 splits of the fp32 operands (ops_fp32.py), matching the reference's plain fp32 modules to ~1e-6;
    forward only (call under torch.no_grad()), no dropout, several times slower."""
    assert mode in ("bf16", "fp32")
    _PRECISION["mode"] = mode


@contextlib.contextmanager
def precision(mode):
    """with precision("fp32"), torch.no_grad(): out = model(...)"""
    prev = _PRECISION["mode"]
    set_precision(mode)
    try:
        yield
    finally:
        _PRECISION["mode"] = prev


def _fp32_mode(*drop_modules):
    if _PRECISION["mode"] != "fp32":
        return False
    if torch.is_grad_enabled():
        raise RuntimeError("precision('fp32') is a forward-only verification mode: call the module under torch.no_grad() "
                           "(gradients are produced by the bf16 production path)")
    if any(_p(m) > 0.0 for m in drop_modules):
        raise RuntimeError("precision('fp32') has no dropout: call model.eval() or set_dropout_policy('ignore')")
    return True


def set_dropout_policy(mode):
    """'apply' (default): train-mode nn.Dropout(p) is applied inside the fused kernels (counter-based masks seeded from
    torch's CUDA generator); 'ignore': dropout is skipped even in train mode (outputs equal the reference with dropout
    disabled -- what the parity tests and the benchmark's dropout-off arm use)."""
    assert mode in ("apply", "ignore")
    _DROPOUT_POLICY["mode"] = mode


def _p(drop_module):
    """Effective drop probability of an nn.Dropout child (0 in eval mode or when dropout is switched off)."""
    if _DROPOUT_POLICY["mode"] == "ignore" or not drop_module.training:
        return 0.0
    return float(drop_module.p)


def _drop_cfg(device, p_first, p_second, seed=None, site=0):
    """(seed, site_base, p_first, p_second) for one sub-block, or None when both probabilities are 0."""
    if p_first <= 0.0 and p_second <= 0.0:
        return None
    if seed is None:
        seed = K.new_seed(device)
    return (seed, site, p_first, p_second)


def _check_heads(embed_dim, num_heads):
    assert embed_dim % num_heads == 0
    if embed_dim // num_heads not in (32, 64):
        raise NotImplementedError(
            f"head_dim {embed_dim // num_heads} is not built: the sm_100a attention kernels cover head_dim 64 "
            "(direct_regression / cascade stage 1) and 32 (cascade stages 2-3, the H200 variants)")


class MultiHeadSelfAttention(nn.Module):
    """reference: models/vit_components.py:13-57"""

    def __init__(self, embed_dim, num_heads=8, dropout=0.1):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(embed_dim, embed_dim * 3, bias=False)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.proj_drop = nn.Dropout(dropout)

    def forward(self, x):
        """x: (B, N, C) -> (B, N, C)"""
        _check_heads(self.embed_dim, self.num_heads)
        if _fp32_mode(self.attn_drop, self.proj_drop):
            return ops_fp32.self_attention(x, self.qkv.weight, self.proj.weight, self.proj.bias, self.num_heads).to(x.dtype)
        drop = _drop_cfg(x.device, _p(self.attn_drop), _p(self.proj_drop))
        return ops.SelfAttention.apply(x, self.qkv.weight, self.proj.weight, self.proj.bias, self.num_heads, drop)


class MultiHeadCrossAttention(nn.Module):
    """reference: models/vit_components.py:60-119"""

    def __init__(self, embed_dim, context_dim, num_heads=8, dropout=0.1, store_attention=False):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.store_attention = store_attention
        self.q = nn.Linear(embed_dim, embed_dim, bias=False)
        self.kv = nn.Linear(context_dim, embed_dim * 2, bias=False)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(embed_dim, embed_dim)
        self.proj_drop = nn.Dropout(dropout)
        self.attention_weights = None

    def forward(self, x, context):
        """x: (B, N, C) queries, context: (B, M, context_dim) -> (B, N, C)"""
        _check_heads(self.embed_dim, self.num_heads)
        if _fp32_mode(self.attn_drop, self.proj_drop):
            res = ops_fp32.cross_attention(x, context, self.q.weight, self.kv.weight, self.proj.weight, self.proj.bias,
                                           self.num_heads, self.store_attention)
            if self.store_attention:
                self.attention_weights = res[1].detach()
                return res[0].to(x.dtype)
            return res.to(x.dtype)
        drop = _drop_cfg(x.device, _p(self.attn_drop), _p(self.proj_drop))
        if self.store_attention:
            # diagnostic slow path: the (B,h,N,M) softmax (before dropout) is materialised by an extra GEMM + exp pass
            # (reference :106-108)
            out, probs = ops.CrossAttention.apply(x, context, self.q.weight, self.kv.weight, self.proj.weight,
                                                  self.proj.bias, self.num_heads, True, drop)
            self.attention_weights = probs.detach()
            return out
        return ops.CrossAttention.apply(x, context, self.q.weight, self.kv.weight, self.proj.weight, self.proj.bias,
                                        self.num_heads, False, drop)


class AdaLNModulation(nn.Module):
    """reference: models/vit_components.py:122-149 (zero-initialised linear -> six (B,1,C) chunks)"""

    def __init__(self, embed_dim, cond_dim):
        super().__init__()
        self.linear = nn.Linear(cond_dim, embed_dim * 6, bias=True)
        nn.init.zeros_(self.linear.weight)
        nn.init.zeros_(self.linear.bias)

    def params(self, cond):
        """(B, 6C) fp32 modulation table consumed directly by the fused LayerNorm / epilogue kernels."""
        return ops.AdaLN.apply(cond, self.linear.weight, self.linear.bias)

    def forward(self, x, cond):
        p = self.params(cond).unsqueeze(1)
        return tuple(p.chunk(6, dim=-1))


class SinusoidalTimeEmbedding(nn.Module):
    """reference: models/vit_components.py:152-174 (API surface only: no caller instantiates it; plain torch)"""

    def __init__(self, embed_dim):
        super().__init__()
        self.embed_dim = embed_dim

    def forward(self, t):
        half = self.embed_dim // 2
        k = math.log(10000) / (half - 1)
        freqs = torch.exp(torch.arange(half, device=t.device) * -k)
        ang = t[:, None] * freqs[None, :]
        return torch.cat([ang.sin(), ang.cos()], dim=-1)
