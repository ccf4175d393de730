"""CPU oracle for the Hybrid-ViT-Cascade 3D ViT backbone hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never as the thing shipped.

What it is: a functional restatement (plain torch ops on a ``state_dict``) of
the reference algorithm in

    /root/reference/models/vit_components.py      (a1-a4 of SURVEY.md section 8)
    /root/reference/models/hybrid_vit_backbone.py (a5-a7)

Every function cites the reference lines it follows.  The arithmetic lives in
PyTorch itself (``requirements.txt:4`` -- ``torch>=2.0.0``, unpinned; this image
has torch 2.11.0+cu128), so the oracle uses the same ATen calls in the same
order as the reference and autograd supplies the gradients.

Pinning: the reference ships no golden vectors or known-answer tests for this
path, so the oracle is pinned against outputs of the
reference itself: ``tests/golden/make_golden.py`` imports the real modules from
``/root/reference`` and stores inputs, weights, outputs and gradients under
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them through this
file (bit-exact on the same torch build, 1e-6 otherwise).

Deliberate deviations (both stated in the test output):
  * ``token_grid``: the committed reference cannot run at 128^3 -- ``pos_embed``
    is sized for 25^3 tokens while the conv stack emits 32^3
    (hybrid_vit_backbone.py:178-188 vs :195-204).  ``token_grid="reference"``
    reproduces the committed rule (and its failure); ``"conv"`` sizes the grid
    from what the conv stack emits (primary 128^3 variant); an int forces
    ``target_size`` (16 = the author's recorded hot-fix,
    direct_regression/progressive_cascade/STAGE2_TRAINING_FIXES.md:22-27).
  * ``attn_chunk``: query-chunked attention, mathematically identical to the
    materialised (B,h,N,M) softmax, used only so that N=32768 checks fit in
    memory.  ``attn_chunk=None`` is the literal reference algorithm.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------
# a6: constructor arithmetic (token grid, conv plan) -- hybrid_vit_backbone.py:174-210
# --------------------------------------------------------------------------

@dataclass
class ConvSpec:
    index: int          # position inside nn.Sequential (state_dict key voxel_embed.<index>)
    cin: int
    cout: int
    stride: int
    norm_index: Optional[int]   # GroupNorm position, None for the final plain conv
    groups: int = 0


@dataclass
class BackboneConfig:
    volume_size: Tuple[int, int, int] = (64, 64, 64)
    in_channels: int = 1
    voxel_dim: int = 384
    depth: int = 6
    num_heads: int = 6
    context_dim: int = 512
    cond_dim: int = 1024
    use_prev_stage: bool = False
    token_grid: Union[str, int] = "reference"
    mlp_ratio: int = 4
    # derived
    downsampled_size: Tuple[int, int, int] = field(default=(0, 0, 0))
    convs: List[ConvSpec] = field(default_factory=list)

    def __post_init__(self):
        self.volume_size = tuple(int(v) for v in self.volume_size)
        factor, ds = token_grid_rule(self.volume_size, self.token_grid)
        self.convs = conv_plan(self.in_channels, self.voxel_dim, factor)
        if self.token_grid == "conv":
            ds = conv_output_grid(self.volume_size, self.convs)
        self.downsampled_size = ds

    @property
    def num_tokens(self) -> int:
        d, h, w = self.downsampled_size
        return d * h * w


def token_grid_rule(volume_size: Sequence[int], token_grid: Union[str, int] = "reference"):
    """hybrid_vit_backbone.py:177-188 -- target 16/24/32 by depth, integer factor, floor grid."""
    D, H, W = volume_size
    if isinstance(token_grid, int):
        target = token_grid
    elif D <= 64:
        target = 16
    elif D <= 128:
        target = 24
    else:
        target = 32
    factor = max(D // target, H // target, W // target)
    factor = max(factor, 1)
    return factor, tuple(d // factor for d in volume_size)


def conv_plan(in_channels: int, voxel_dim: int, factor: int) -> List[ConvSpec]:
    """hybrid_vit_backbone.py:190-210 -- the while-loop that builds ``voxel_embed``.

    Note the value comparison ``current_dim == in_channels`` (:197): if
    ``in_channels == voxel_dim // 4`` the second conv also gets ``voxel_dim // 4``.
    ``len(layers) < 4`` counts Sequential entries (3 per strided conv).
    """
    specs: List[ConvSpec] = []
    n_layers = 0
    cur = in_channels
    remaining = factor
    while remaining > 1:
        stride = min(remaining, 2)
        if cur == in_channels:
            cout = voxel_dim // 4
        elif n_layers < 4:
            cout = voxel_dim // 2
        else:
            cout = voxel_dim
        specs.append(ConvSpec(n_layers, cur, cout, stride, n_layers + 1, min(8, cout)))
        n_layers += 3
        cur = cout
        remaining //= stride
    if cur != voxel_dim:
        specs.append(ConvSpec(n_layers, cur, voxel_dim, 1, None))
    return specs


def conv_output_grid(volume_size: Sequence[int], convs: Sequence[ConvSpec]) -> Tuple[int, int, int]:
    """Spatial size after the k3/p1 conv stack (what ``voxel_embed`` really emits)."""
    dims = list(volume_size)
    for c in convs:
        dims = [(d + 2 - 3) // c.stride + 1 for d in dims]
    return tuple(dims)


# --------------------------------------------------------------------------
# a4: SinusoidalTimeEmbedding -- vit_components.py:152-174
# --------------------------------------------------------------------------

def sinusoidal_time_embedding(t: Tensor, embed_dim: int) -> Tensor:
    half = embed_dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -k)
    ang = t[:, None] * freqs[None, :]
    return torch.cat([ang.sin(), ang.cos()], dim=-1)


# --------------------------------------------------------------------------
# attention core shared by a1/a2 -- vit_components.py:46-51 and :103-113
# --------------------------------------------------------------------------

def attention_core(q: Tensor, k: Tensor, v: Tensor, scale: float,
                   attn_chunk: Optional[int] = None, return_probs: bool = False, pmask: Optional[Tensor] = None):
    """softmax(q k^T * scale) v with q,k,v shaped (B, h, N|M, d).

    ``attn_chunk=None`` materialises the full (B,h,N,M) matrix exactly as the
    reference does; otherwise queries are processed ``attn_chunk`` rows at a time
    (each row's softmax is independent, so the result is identical).
    """
    # pmask: attn_drop (vit_components.py:49, :110) as an explicit (B,h,N,M) keep-mask already scaled by 1/(1-p);
    # None = dropout off.  The stored attention map is taken BEFORE dropout (:106-108).
    # A callable pmask is torch's own nn.Dropout applied to the probabilities (oracle.dropout_mask.TorchDropout: the timed baselines,
    # where the masks need not match the kernels').
    if attn_chunk is None:
        attn = (q @ k.transpose(-2, -1)) * scale
        attn = attn.softmax(dim=-1)
        out = (attn if pmask is None else pmask(attn) if callable(pmask) else attn * pmask) @ v
        return (out, attn) if return_probs else (out, None)
    outs = []
    for s in range(0, q.shape[2], attn_chunk):
        a = (q[:, :, s:s + attn_chunk] @ k.transpose(-2, -1)) * scale
        a = a.softmax(dim=-1)
        if pmask is not None:
            a = pmask(a) if callable(pmask) else a * pmask[:, :, s:s + attn_chunk]
        outs.append(a @ v)
    return torch.cat(outs, dim=2), None


def _drop_tokens(drop, site: int, t: Tensor) -> Tensor:
    """proj_drop / MLP dropout on a (B, N, C) tensor: explicit kernel-matching mask, or torch's nn.Dropout (TorchDropout)."""
    if hasattr(drop, "apply_tokens"):
        return drop.apply_tokens(t)
    return t * drop.tokens(site, t.shape[0] * t.shape[1], t.shape[2]).view(t.shape)


# --------------------------------------------------------------------------
# a1: MultiHeadSelfAttention.forward -- vit_components.py:31-57 (dropout off)
# --------------------------------------------------------------------------

def self_attention(x: Tensor, sd: StateDict, pfx: str, num_heads: int,
                   attn_chunk: Optional[int] = None, drop=None, site: int = 0) -> Tensor:
    """drop: None (eval / dropout off) or an oracle.dropout_mask.DropoutOracle giving the masks of attn_drop
    (site) and proj_drop (site + 1), :49 and :55."""
    B, N, C = x.shape
    d = C // num_heads
    qkv = F.linear(x, sd[pfx + "qkv.weight"])                       # :41 (bias=False, :26)
    qkv = qkv.reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)  # :41-42
    q, k, v = qkv[0], qkv[1], qkv[2]
    pmask = drop.attn(site, B, num_heads, N, N) if drop is not None else None
    o, _ = attention_core(q, k, v, d ** -0.5, attn_chunk, pmask=pmask)   # :46-51
    o = o.transpose(1, 2).reshape(B, N, C)                           # :51
    o = F.linear(o, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])  # :54
    if drop is not None:
        o = _drop_tokens(drop, site + 1, o)                          # :55
    return o


# --------------------------------------------------------------------------
# a2: MultiHeadCrossAttention.forward -- vit_components.py:83-119 (dropout off)
# --------------------------------------------------------------------------

def cross_attention(x: Tensor, context: Tensor, sd: StateDict, pfx: str, num_heads: int,
                    attn_chunk: Optional[int] = None, return_probs: bool = False, drop=None, site: int = 0):
    B, N, C = x.shape
    M = context.shape[1]
    d = C // num_heads
    q = F.linear(x, sd[pfx + "q.weight"]).reshape(B, N, num_heads, d).permute(0, 2, 1, 3)   # :95-96
    kv = F.linear(context, sd[pfx + "kv.weight"]).reshape(B, M, 2, num_heads, d)            # :98
    kv = kv.permute(2, 0, 3, 1, 4)                                                          # :99
    k, v = kv[0], kv[1]
    pmask = drop.attn(site, B, num_heads, N, M) if drop is not None else None
    o, probs = attention_core(q, k, v, d ** -0.5, attn_chunk, return_probs, pmask=pmask)    # :103-113
    o = o.transpose(1, 2).reshape(B, N, C)
    o = F.linear(o, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])                         # :116
    if drop is not None:
        o = _drop_tokens(drop, site + 1, o)                                                 # :117
    return (o, probs.detach()) if return_probs else o                                       # :107-108


# --------------------------------------------------------------------------
# a3: AdaLNModulation.forward -- vit_components.py:135-149
# --------------------------------------------------------------------------

def adaln(cond: Tensor, sd: StateDict, pfx: str):
    params = F.linear(cond, sd[pfx + "linear.weight"], sd[pfx + "linear.bias"]).unsqueeze(1)
    return params.chunk(6, dim=-1)   # shift_sa, scale_sa, gate_sa, shift_mlp, scale_mlp, gate_mlp


# --------------------------------------------------------------------------
# a5: HybridViTBlock3D.forward -- hybrid_vit_backbone.py:88-143 (dropout off)
# --------------------------------------------------------------------------

def block(x: Tensor, context: Tensor, cond: Tensor, sd: StateDict, pfx: str, num_heads: int,
          use_prev_stage: bool = False, prev_stage_embed: Optional[Tensor] = None,
          attn_chunk: Optional[int] = None, return_attention: bool = False, drop=None, site_base: int = 0):
    """drop/site_base: train-mode dropout as explicit masks (oracle.dropout_mask.DropoutOracle); sites site_base + 0..5 =
    self-attn probabilities, self-attn proj, cross-attn probabilities, cross-attn proj, MLP activation, MLP output."""
    C = x.shape[-1]
    if use_prev_stage:                                                # :106-114
        if prev_stage_embed is None:
            prev_stage_embed = torch.zeros(x.shape[0], 256, device=x.device, dtype=x.dtype)
        cond = torch.cat([cond, prev_stage_embed], dim=-1)
    shift_sa, scale_sa, gate_sa, shift_mlp, scale_mlp, gate_mlp = adaln(cond, sd, pfx + "adaln.")  # :117

    def ln(t, name):
        return F.layer_norm(t, (C,), sd[pfx + name + ".weight"], sd[pfx + name + ".bias"], 1e-5)

    h = (1 + scale_sa) * ln(x, "norm1") + shift_sa                    # :120-121
    x = x + gate_sa * self_attention(h, sd, pfx + "self_attn.", num_heads, attn_chunk, drop, site_base)   # :122-123
    attn_map = None
    if return_attention:
        ca, attn_map = cross_attention(ln(x, "norm2"), context, sd, pfx + "cross_attn.", num_heads,
                                       attn_chunk, True, drop, site_base + 2)
    else:
        ca = cross_attention(ln(x, "norm2"), context, sd, pfx + "cross_attn.", num_heads, attn_chunk, False, drop,
                             site_base + 2)
    x = x + ca                                                        # :126-128
    h = (1 + scale_mlp) * ln(x, "norm3") + shift_mlp                  # :136-137
    h = F.linear(h, sd[pfx + "mlp.0.weight"], sd[pfx + "mlp.0.bias"]) # :75-81
    h = F.gelu(h)                                                     # nn.GELU() = exact erf
    if drop is not None:
        h = _drop_tokens(drop, site_base + 4, h)                      # mlp.2
    h = F.linear(h, sd[pfx + "mlp.3.weight"], sd[pfx + "mlp.3.bias"])
    if drop is not None:
        h = _drop_tokens(drop, site_base + 5, h)                      # mlp.4
    x = x + gate_mlp * h                                              # :139
    return (x, attn_map) if return_attention else x


# --------------------------------------------------------------------------
# a7: HybridViT3D.forward -- hybrid_vit_backbone.py:233-274
# --------------------------------------------------------------------------

def voxel_embed(x: Tensor, sd: StateDict, pfx: str, cfg: BackboneConfig) -> Tensor:
    """Conv3d(k3,p1)[+GroupNorm+SiLU] stack -- hybrid_vit_backbone.py:195-210,252."""
    for c in cfg.convs:
        x = F.conv3d(x, sd[f"{pfx}{c.index}.weight"], sd[f"{pfx}{c.index}.bias"], stride=c.stride, padding=1)
        if c.norm_index is not None:
            x = F.group_norm(x, c.groups, sd[f"{pfx}{c.norm_index}.weight"], sd[f"{pfx}{c.norm_index}.bias"], 1e-5)
            x = F.silu(x)
    return x


def backbone(x: Tensor, context: Tensor, cond: Tensor, sd: StateDict, cfg: BackboneConfig,
             pfx: str = "", prev_stage_embed: Optional[Tensor] = None,
             attn_chunk: Optional[int] = None, drop=None) -> Tensor:
    B = x.shape[0]
    D, H, W = cfg.volume_size
    Dd, Hd, Wd = cfg.downsampled_size
    x = voxel_embed(x, sd, pfx + "voxel_embed.", cfg)                 # :252
    x = x.flatten(2).transpose(1, 2)                                  # :255
    x = x + sd[pfx + "pos_embed"]                                     # :258 (raises on the 128^3 defect)
    for i in range(cfg.depth):                                        # :261-262
        x = block(x, context, cond, sd, f"{pfx}blocks.{i}.", cfg.num_heads,
                  cfg.use_prev_stage, prev_stage_embed, attn_chunk, drop=drop, site_base=8 * i)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), sd[pfx + "norm.weight"], sd[pfx + "norm.bias"], 1e-5)   # :265
    x = F.linear(x, sd[pfx + "output_proj.weight"], sd[pfx + "output_proj.bias"])    # :266
    x = x.transpose(1, 2).reshape(B, 1, Dd, Hd, Wd)                   # :269
    return F.interpolate(x, size=(D, H, W), mode="trilinear", align_corners=True)    # :272


# --------------------------------------------------------------------------
# weights: same shapes/keys/initial distributions as the reference constructors
# --------------------------------------------------------------------------

def init_state_dict(cfg: BackboneConfig, seed: int = 0, adaln_std: float = 0.02,
                    dtype: torch.dtype = torch.float32) -> StateDict:
    """Random weights with the reference's key names and shapes.

    AdaLN is zero-initialised in the reference (vit_components.py:131-133), which
    zeroes the self-attention and MLP branches; parity tests need those branches
    live, so ``adaln_std`` > 0 re-randomises it ~N(0, adaln_std).
    """
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * b

    sd: StateDict = {}
    C = cfg.voxel_dim
    for c in cfg.convs:
        fan = c.cin * 27
        sd[f"voxel_embed.{c.index}.weight"] = uni((c.cout, c.cin, 3, 3, 3), fan)
        sd[f"voxel_embed.{c.index}.bias"] = uni((c.cout,), fan)
        if c.norm_index is not None:
            sd[f"voxel_embed.{c.norm_index}.weight"] = 1 + 0.1 * torch.randn(c.cout, generator=g, dtype=dtype)
            sd[f"voxel_embed.{c.norm_index}.bias"] = 0.1 * torch.randn(c.cout, generator=g, dtype=dtype)
    sd["pos_embed"] = torch.randn(1, cfg.num_tokens, C, generator=g, dtype=dtype) * 0.02
    cond = cfg.cond_dim + (256 if cfg.use_prev_stage else 0)
    hid = int(C * cfg.mlp_ratio)
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        sd[p + "adaln.linear.weight"] = torch.randn(6 * C, cond, generator=g, dtype=dtype) * adaln_std
        sd[p + "adaln.linear.bias"] = torch.randn(6 * C, generator=g, dtype=dtype) * adaln_std
        sd[p + "self_attn.qkv.weight"] = uni((3 * C, C), C)
        sd[p + "self_attn.proj.weight"] = uni((C, C), C)
        sd[p + "self_attn.proj.bias"] = uni((C,), C)
        sd[p + "cross_attn.q.weight"] = uni((C, C), C)
        sd[p + "cross_attn.kv.weight"] = uni((2 * C, cfg.context_dim), cfg.context_dim)
        sd[p + "cross_attn.proj.weight"] = uni((C, C), C)
        sd[p + "cross_attn.proj.bias"] = uni((C,), C)
        sd[p + "mlp.0.weight"] = uni((hid, C), C)
        sd[p + "mlp.0.bias"] = uni((hid,), C)
        sd[p + "mlp.3.weight"] = uni((C, hid), hid)
        sd[p + "mlp.3.bias"] = uni((C,), hid)
        for n in ("norm1", "norm2", "norm3"):
            sd[p + n + ".weight"] = 1 + 0.1 * torch.randn(C, generator=g, dtype=dtype)
            sd[p + n + ".bias"] = 0.1 * torch.randn(C, generator=g, dtype=dtype)
    sd["norm.weight"] = 1 + 0.1 * torch.randn(C, generator=g, dtype=dtype)
    sd["norm.bias"] = 0.1 * torch.randn(C, generator=g, dtype=dtype)
    sd["output_proj.weight"] = uni((1, C), C)
    sd["output_proj.bias"] = uni((1,), C)
    return sd


# --------------------------------------------------------------------------
# metrics used by every parity test
# --------------------------------------------------------------------------

def rel_fro(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cosine(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    den = a.norm() * b.norm()
    if float(den) == 0.0:
        return 1.0 if float(a.norm()) == float(b.norm()) else 0.0
    return float(a @ b / den)


# --------------------------------------------------------------------------
# algorithmic FLOPs -- forward, per sample
# --------------------------------------------------------------------------

def forward_flops(cfg: BackboneConfig, context_len: int) -> Dict[str, float]:
    N, C, M, Cc = cfg.num_tokens, cfg.voxel_dim, context_len, cfg.context_dim
    cond = cfg.cond_dim + (256 if cfg.use_prev_stage else 0)
    lin = 28 * N * C * C + 4 * M * Cc * C + 2 * cond * 6 * C
    sa = 4 * N * N * C
    ca = 4 * N * M * C
    dims = list(cfg.volume_size)
    emb = 0.0
    for c in cfg.convs:
        dims = [(d + 2 - 3) // c.stride + 1 for d in dims]
        emb += 2 * 27 * c.cin * c.cout * dims[0] * dims[1] * dims[2]
    head = 2 * N * C
    return {"linear": float(lin * cfg.depth), "self_attn": float(sa * cfg.depth),
            "cross_attn": float(ca * cfg.depth), "embed": float(emb), "head": float(head),
            "total": float((lin + sa + ca) * cfg.depth + emb + head)}
