"""Drop-in replacements for the reference's ``models/hybrid_vit_backbone.py``.

``HybridViTBlock3D`` and ``HybridViT3D`` keep the reference constructor/forward signatures, public
attributes (``volume_size``, ``downsampled_size``, ``voxel_embed``, ``pos_embed``, ``blocks``, ``norm``,
``output_proj``) and ``state_dict`` layout; modules are built in the reference's order so the same
``torch.manual_seed`` gives the same initial weights.  Forward/backward run as fused kernel
sequences (ops.py): one for the embedding, three per block, one for the output head.

One extension: ``token_grid`` ("reference" | "conv" | int).  The committed reference cannot run at
128^3 (pos_embed sized for 25^3 tokens, conv stack emits 32^3; hybrid_vit_backbone.py:178-188 vs
:195-204).  "reference" keeps the committed rule (and fails the same way at 128^3), "conv" sizes
the grid from the conv stack (32^3), an int forces target_size (16 = the author's recorded fix).
"""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import kernels as K
from . import ops
from . import ops_fp32
from .vit_components import (AdaLNModulation, MultiHeadCrossAttention, MultiHeadSelfAttention,
                             SinusoidalTimeEmbedding, _check_heads, _drop_cfg, _fp32_mode, _p)  # noqa: F401


class HybridViTBlock3D(nn.Module):
    """reference: models/hybrid_vit_backbone.py:21-143"""

    def __init__(self, voxel_dim: int, num_heads: int = 8, context_dim: int = 512, cond_dim: int = 1024,
                 mlp_ratio: int = 4, dropout: float = 0.1, use_prev_stage: bool = False,
                 return_attention: bool = False):
        super().__init__()
        self.voxel_dim = voxel_dim
        self.use_prev_stage = use_prev_stage
        self.return_attention = return_attention
        adaln_in = cond_dim + (256 if use_prev_stage else 0)
        self.adaln = AdaLNModulation(embed_dim=voxel_dim, cond_dim=adaln_in)
        self.self_attn = MultiHeadSelfAttention(embed_dim=voxel_dim, num_heads=num_heads, dropout=dropout)
        self.cross_attn = MultiHeadCrossAttention(embed_dim=voxel_dim, num_heads=num_heads, context_dim=context_dim,
                                                  dropout=dropout, store_attention=return_attention)
        hidden = int(voxel_dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(voxel_dim, hidden), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden, voxel_dim), nn.Dropout(dropout))
        self.norm1 = nn.LayerNorm(voxel_dim)
        self.norm2 = nn.LayerNorm(voxel_dim)
        self.norm3 = nn.LayerNorm(voxel_dim)

    def _combined_cond(self, cond, prev_stage_embed, batch):
        if not self.use_prev_stage:
            return cond
        if prev_stage_embed is None:
            prev_stage_embed = torch.zeros(batch, 256, device=cond.device, dtype=cond.dtype)
        return torch.cat([cond, prev_stage_embed], dim=-1)

    def _dropouts(self):
        sa, ca = self.self_attn, self.cross_attn
        return (sa.attn_drop, sa.proj_drop, ca.attn_drop, ca.proj_drop, self.mlp[2], self.mlp[4])

    def dropout_probs(self):
        """Effective p of the six nn.Dropout sites of the block (all 0 in eval mode), reference :75-81 and
        vit_components.py:27-29,76-78."""
        sa, ca = self.self_attn, self.cross_attn
        return (_p(sa.attn_drop), _p(sa.proj_drop), _p(ca.attn_drop), _p(ca.proj_drop), _p(self.mlp[2]), _p(self.mlp[4]))

    def _forward_tokens(self, x, ctx16, cond, B, N, M, seed=None, site_base=0):
        """x: fp32 [B*N, C] residual stream; ctx16: bf16 [B*M, Cc]; cond already combined.
        seed/site_base: dropout seed words shared by the whole backbone call and this block's first site id."""
        C = self.voxel_dim
        H = self.self_attn.num_heads
        sa, ca = self.self_attn, self.cross_attn
        ps = self.dropout_probs()
        if seed is None and any(p > 0.0 for p in ps):
            seed = K.new_seed(x.device)
        d_sa = _drop_cfg(x.device, ps[0], ps[1], seed, site_base)
        d_ca = _drop_cfg(x.device, ps[2], ps[3], seed, site_base + 2)
        d_mlp = _drop_cfg(x.device, ps[4], ps[5], seed, site_base + 4)
        mod = self.adaln.params(cond)                                   # (B, 6C)
        x = ops.SelfAttnBranch.apply(x, mod, self.norm1.weight, self.norm1.bias, sa.qkv.weight, sa.proj.weight,
                                     sa.proj.bias, B, N, H, 0, d_sa)
        if ca.store_attention:      # reference :130-133 reads it back from the cross-attention module
            x, probs = ops.CrossAttnBranch.apply(x, ctx16, self.norm2.weight, self.norm2.bias, ca.q.weight, ca.kv.weight,
                                                 ca.proj.weight, ca.proj.bias, B, N, M, H, True, d_ca)
            ca.attention_weights = probs.detach()
        else:
            x = ops.CrossAttnBranch.apply(x, ctx16, self.norm2.weight, self.norm2.bias, ca.q.weight, ca.kv.weight,
                                          ca.proj.weight, ca.proj.bias, B, N, M, H, False, d_ca)
        x = ops.MlpBranch.apply(x, mod, self.norm3.weight, self.norm3.bias, self.mlp[0].weight, self.mlp[0].bias,
                                self.mlp[3].weight, self.mlp[3].bias, B, N, 3 * C, d_mlp)
        return x

    def forward(self, voxel_features: torch.Tensor, xray_context: torch.Tensor, cond: torch.Tensor,
                prev_stage_embed: Optional[torch.Tensor] = None):
        _check_heads(self.voxel_dim, self.self_attn.num_heads)
        B, N, C = voxel_features.shape
        M = xray_context.shape[1]
        cond = self._combined_cond(cond, prev_stage_embed, B)
        x = voxel_features.float().contiguous().view(B * N, C)
        if _fp32_mode(*self._dropouts()):
            out = ops_fp32.block_tokens(self, x, ops_fp32.cast_tokens(xray_context), cond, B, N, M)
            out = out.view(B, N, C).to(voxel_features.dtype)
            return (out, self.cross_attn.attention_weights) if self.return_attention else out
        ctx16 = ops.CastTokens.apply(xray_context)
        x = self._forward_tokens(x, ctx16, cond, B, N, M)
        out = x.view(B, N, C).to(voxel_features.dtype)
        if self.return_attention:   # reference :130-143: (features, cross-attention map (B, heads, N, M))
            return out, self.cross_attn.attention_weights
        return out


class HybridViT3D(nn.Module):
    """reference: models/hybrid_vit_backbone.py:146-274"""

    def __init__(self, volume_size: Tuple[int, int, int] = (64, 64, 64), in_channels: int = 1, voxel_dim: int = 384,
                 depth: int = 6, num_heads: int = 6, context_dim: int = 512, cond_dim: int = 1024,
                 use_prev_stage: bool = False, dropout: float = 0.1, token_grid="reference"):
        super().__init__()
        self.volume_size = volume_size
        self.in_channels = in_channels
        self.voxel_dim = voxel_dim
        self.use_prev_stage = use_prev_stage
        self.token_grid = token_grid

        # token grid rule, reference :174-188
        D, H, W = volume_size
        if isinstance(token_grid, int):
            target = token_grid
        else:
            target = 16 if D <= 64 else (24 if D <= 128 else 32)
        factor = max(max(D // target, H // target, W // target), 1)
        self.downsampled_size = tuple(d // factor for d in volume_size)

        # conv stack, reference :190-210 (same Sequential indices -> same state_dict keys)
        layers, plan = [], []
        cur, remaining = in_channels, factor
        while remaining > 1:
            stride = min(remaining, 2)
            out_dim = voxel_dim // 4 if cur == in_channels else (voxel_dim // 2 if len(layers) < 4 else voxel_dim)
            groups = min(8, out_dim)
            layers += [nn.Conv3d(cur, out_dim, kernel_size=3, stride=stride, padding=1),
                       nn.GroupNorm(groups, out_dim), nn.SiLU()]
            plan.append((cur, out_dim, stride, groups))
            cur = out_dim
            remaining //= stride
        if cur != voxel_dim:
            layers.append(nn.Conv3d(cur, voxel_dim, kernel_size=3, padding=1))
            plan.append((cur, voxel_dim, 1, 0))
        self.voxel_embed = nn.Sequential(*layers)
        self._plan = plan
        if token_grid == "conv":
            dims = list(volume_size)
            for (_, _, s, _) in plan:
                dims = [(d - 1) // s + 1 for d in dims]
            self.downsampled_size = tuple(dims)
        Dd, Hd, Wd = self.downsampled_size

        self.pos_embed = nn.Parameter(torch.randn(1, Dd * Hd * Wd, voxel_dim) * 0.02)
        self.blocks = nn.ModuleList([
            HybridViTBlock3D(voxel_dim=voxel_dim, num_heads=num_heads, context_dim=context_dim, cond_dim=cond_dim,
                             use_prev_stage=use_prev_stage, dropout=dropout)
            for _ in range(depth)])
        self.norm = nn.LayerNorm(voxel_dim)
        self.output_proj = nn.Linear(voxel_dim, 1)

    def _embed_params(self):
        out = []
        for m in self.voxel_embed:
            if isinstance(m, (nn.Conv3d, nn.GroupNorm)):
                out += [m.weight, m.bias]
        return out

    def forward(self, x: torch.Tensor, context: torch.Tensor, cond: torch.Tensor,
                prev_stage_embed: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: (B, C_in, D, H, W), context: (B, M, context_dim), cond: (B, cond_dim) -> (B, 1, D, H, W) fp32"""
        B = x.shape[0]
        D, H, W = self.volume_size
        Dd, Hd, Wd = self.downsampled_size
        N, M = Dd * Hd * Wd, context.shape[1]
        if len(self.blocks):
            _check_heads(self.voxel_dim, self.blocks[0].self_attn.num_heads)
        if not self._plan:
            raise NotImplementedError("in_channels == voxel_dim with no downsampling leaves voxel_embed empty")
        if _fp32_mode(*(m for blk in self.blocks for m in blk._dropouts())):
            return ops_fp32.backbone(self, x, context, cond, prev_stage_embed)
        tok = ops.VoxelEmbed.apply(x, self.pos_embed, self._plan, *self._embed_params())     # fp32 [B*N, C]
        ctx16 = ops.CastTokens.apply(context)
        seed = None
        if any(p > 0.0 for blk in self.blocks for p in blk.dropout_probs()):
            seed = K.new_seed(tok.device)      # one draw per forward; every dropout site of every block derives from it
        for i, blk in enumerate(self.blocks):
            c = blk._combined_cond(cond, prev_stage_embed, B)
            tok = blk._forward_tokens(tok, ctx16, c, B, N, M, seed, 8 * i)
        return ops.OutputHead.apply(tok, self.norm.weight, self.norm.bias, self.output_proj.weight,
                                    self.output_proj.bias, B, (Dd, Hd, Wd), (D, H, W))
