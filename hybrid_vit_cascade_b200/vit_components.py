"""Drop-in replacements for the reference's ``models/vit_components.py``.

Same class names, constructor arguments, forward signatures, attribute names and ``state_dict``
keys/shapes as the reference, so checkpoints load strictly in both directions; the forward and
backward passes run on the sm_100a kernels of libhvc_sm100a.so (see ops.py).  The ``nn.Linear`` /
``nn.Dropout`` children are kept as parameter/config containers only -- their ``forward`` is never
called on the hot path.
"""
import math
import warnings

import torch
import torch.nn as nn

from . import ops

_DROPOUT_POLICY = {"mode": "warn", "warned": False}


def set_dropout_policy(mode):
    """'warn' (default): train-mode dropout p>0 is skipped with a one-time warning; 'error': raise; 'ignore'."""
    assert mode in ("warn", "error", "ignore")
    _DROPOUT_POLICY["mode"] = mode


def _check_dropout(module, p):
    if not module.training or p <= 0.0 or _DROPOUT_POLICY["mode"] == "ignore":
        return
    msg = ("hybrid_vit_cascade_b200: train-mode dropout (p=%g) is not applied by the fused kernels in this round; "
           "outputs equal the reference with dropout disabled" % p)
    if _DROPOUT_POLICY["mode"] == "error":
        raise NotImplementedError(msg)
    if not _DROPOUT_POLICY["warned"]:
        warnings.warn(msg)
        _DROPOUT_POLICY["warned"] = True


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
        _check_dropout(self, self.attn_drop.p)
        return ops.SelfAttention.apply(x, self.qkv.weight, self.proj.weight, self.proj.bias, self.num_heads)


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
        _check_dropout(self, self.attn_drop.p)
        if self.store_attention:
            # diagnostic slow path: the (B,h,N,M) softmax is materialised by an extra GEMM + exp pass (reference :106-108)
            out, probs = ops.CrossAttention.apply(x, context, self.q.weight, self.kv.weight, self.proj.weight,
                                                  self.proj.bias, self.num_heads, True)
            self.attention_weights = probs.detach()
            return out
        return ops.CrossAttention.apply(x, context, self.q.weight, self.kv.weight, self.proj.weight, self.proj.bias,
                                        self.num_heads)


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
