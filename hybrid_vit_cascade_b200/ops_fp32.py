"""fp32 verification mode: the forward pass of the hot path at fp32 accuracy, still on the sm_100a kernels.

The production path (ops.py) computes in bf16 like ``torch.autocast`` does to the reference.  This module is what
``hybrid_vit_cascade_b200.precision("fp32")`` switches the modules to, so they can be checked against the reference's
plain fp32 modules at the 1e-4 bar (BASELINE.json north_star).  Every product still runs on the
tcgen05 GEMM: fp32 operands are split into three bf16 terms and the six significant partial products are summed
along K by one GEMM call (csrc/hvc_fp32.cu); LayerNorm / GroupNorm / AdaLN / head / upsample kernels are fp32 already.
Attention materialises the scores per (batch, head).  Forward only (no autograd graph is recorded), slow on purpose.
"""
import torch

from . import kernels as K
from . import ops
from .ops import _cache_get, _cache_put

A_SIDE, B_SIDE = 0, 1
_W6 = {}
ops._CACHES.append(_W6)


def w6(p, pad_to=None):
    """Six-term B-side operand [out, 6*in] of an fp32 parameter viewed [out, in]; cached like ops.w16."""
    t = _cache_get(_W6, p, pad_to)
    if t is not None:
        return t
    src = p.detach().float().reshape(p.shape[0], -1).contiguous()
    if pad_to is not None and pad_to != src.shape[1]:
        padded = torch.zeros(src.shape[0], pad_to, device=src.device, dtype=torch.float32)
        padded[:, :src.shape[1]] = src
        src = padded
    t = K.split3(src, B_SIDE)
    _cache_put(_W6, p, pad_to, t)
    return t


def clear_weight_cache():
    _W6.clear()


CHAIN = 256   # K' elements per TMEM accumulation chain (the tensor-core accumulator truncates; see hvc_fp32.cu)


def gemm6(a6, b6, out=None, alpha=1.0, b_major=0):
    """fp32-accurate product of two six-term operands: split-K chains of CHAIN elements, fp32 atomic reduction."""
    Kp = a6.shape[1]
    M = a6.shape[0]
    N = b6.shape[0] if b_major == 0 else b6.shape[1]
    if out is None:
        out = torch.zeros(M, N, device=a6.device, dtype=torch.float32)
    else:
        out.zero_()
    return K.gemm(a6, b6, b_major=b_major, epilogue=K.EPI_F32_ATOMIC, k_splits=(Kp + CHAIN - 1) // CHAIN, alpha=alpha, out=out)


def linear(x32, weight, bias=None, pad_to=None, activation=K.ACT_NONE, resid=None, gate=None, gate_ld=0, rows_per_batch=0):
    """y f32 [T, N] = resid + gate * act(x32 [T, K] weight[N, K]^T + bias)."""
    acc = gemm6(K.split3(x32, A_SIDE), w6(weight, pad_to))
    if bias is None and activation == K.ACT_NONE and resid is None and gate is None:
        return acc
    return K.epilogue_f32(acc, bias, activation, resid, gate, gate_ld, rows_per_batch)


def attention(q, k, v, B, H, nq, nk, d, scale, want_probs=False):
    """q f32 view [B*nq, H*d], k/v f32 views [B*nk, H*d] -> o f32 [B*nq, H*d] (and the softmax f32 [B,H,nq,nk])."""
    dev = q.device
    o = torch.empty(B * nq, H * d, device=dev, dtype=torch.float32)
    probs = torch.empty(B, H, nq, nk, device=dev, dtype=torch.float32) if want_probs else None
    s = None if want_probs else torch.empty(nq, nk, device=dev, dtype=torch.float32)
    alpha = scale * 1.4426950408889634
    for b in range(B):
        for h in range(H):
            qh = q[b * nq:(b + 1) * nq, h * d:(h + 1) * d]
            kh = k[b * nk:(b + 1) * nk, h * d:(h + 1) * d]
            vh = v[b * nk:(b + 1) * nk, h * d:(h + 1) * d]
            sb = probs[b, h] if want_probs else s
            gemm6(K.split3(qh, A_SIDE), K.split3(kh, B_SIDE), out=sb, alpha=alpha)
            K.softmax_rows_(sb)
            gemm6(K.split3(sb, A_SIDE), K.split3(vh, B_SIDE, concat_rows=True), out=o[b * nq:(b + 1) * nq, h * d:(h + 1) * d],
                  b_major=1)
    return (o, probs) if want_probs else o


def cast_tokens(x):
    """(B, M, C) any strides / dtype -> f32 [B*M, C] contiguous (a copy is plumbing, not arithmetic)."""
    B, M, C = x.shape
    return x.float().contiguous().view(B * M, C)


# ------------------------------------------------------------------ standalone modules (a1, a2)

def self_attention(x, w_qkv, w_proj, b_proj, H):
    B, N, C = x.shape
    d = C // H
    qkv = linear(cast_tokens(x), w_qkv)
    o = attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, d, d ** -0.5)
    return linear(o, w_proj, b_proj).view(B, N, C)


def cross_attention(x, context, w_q, w_kv, w_proj, b_proj, H, store_probs=False):
    B, N, C = x.shape
    M = context.shape[1]
    d = C // H
    q = linear(cast_tokens(x), w_q)
    kv = linear(cast_tokens(context), w_kv)
    res = attention(q, kv[:, :C], kv[:, C:], B, H, N, M, d, d ** -0.5, want_probs=store_probs)
    o = res[0] if store_probs else res
    out = linear(o, w_proj, b_proj).view(B, N, C)
    return (out, res[1]) if store_probs else out


# ------------------------------------------------------------------ block (a5)

def block_tokens(blk, x, ctx32, cond, B, N, M):
    """x f32 [B*N, C], ctx32 f32 [B*M, Cc], cond combined -> x f32 [B*N, C]   (hybrid_vit_backbone.py:116-139)."""
    C = blk.voxel_dim
    sa, ca = blk.self_attn, blk.cross_attn
    H = sa.num_heads
    d = C // H
    mod = K.adaln_fwd(cond.float().contiguous(), blk.adaln.linear.weight.contiguous(), blk.adaln.linear.bias)
    ld = mod.stride(0)

    def views(off):
        return mod[:, off:off + C], mod[:, off + C:off + 2 * C], mod[:, off + 2 * C:off + 3 * C]

    shift, scale, gate = views(0)
    y, _, _ = K.ln_fwd(x, blk.norm1.weight, blk.norm1.bias, shift, scale, ld, N, out_dtype=torch.float32, save_stats=False)
    qkv = linear(y, sa.qkv.weight)
    o = attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, d, d ** -0.5)
    x = linear(o, sa.proj.weight, sa.proj.bias, resid=x, gate=gate, gate_ld=ld, rows_per_batch=N)

    y, _, _ = K.ln_fwd(x, blk.norm2.weight, blk.norm2.bias, out_dtype=torch.float32, save_stats=False)
    q = linear(y, ca.q.weight)
    kv = linear(ctx32, ca.kv.weight)
    res = attention(q, kv[:, :C], kv[:, C:], B, H, N, M, d, d ** -0.5, want_probs=ca.store_attention)
    if ca.store_attention:
        o, ca.attention_weights = res[0], res[1].detach()
    else:
        o = res
    x = linear(o, ca.proj.weight, ca.proj.bias, resid=x)

    shift, scale, gate = views(3 * C)
    y, _, _ = K.ln_fwd(x, blk.norm3.weight, blk.norm3.bias, shift, scale, ld, N, out_dtype=torch.float32, save_stats=False)
    g = linear(y, blk.mlp[0].weight, blk.mlp[0].bias, activation=K.ACT_GELU)
    return linear(g, blk.mlp[3].weight, blk.mlp[3].bias, resid=x, gate=gate, gate_ld=ld, rows_per_batch=N)


# ------------------------------------------------------------------ backbone (a7)

def voxel_embed(x, pos_embed, plan, params, B):
    """hybrid_vit_backbone.py:252-258 in fp32: im2col (f32) -> split -> GEMM -> GroupNorm+SiLU (f32) ... + pos."""
    _, Cin, D, H, W = x.shape
    xB = 1 if (B > 1 and x.stride(0) == 0) else B
    a = x[:xB].float()
    strides, dims = tuple(a.stride()), (Cin, D, H, W)
    z, pi = None, 0
    for li, (cin, cout, stride, groups) in enumerate(plan):
        _, Dc, Hc, Wc = dims
        weight, bias = params[pi], params[pi + 1]
        cols = K.im2col3d_f32(a, xB, cin, Dc, Hc, Wc, stride, strides)
        Do, Ho, Wo = K.conv_out(Dc, stride), K.conv_out(Hc, stride), K.conv_out(Wc, stride)
        z = linear(cols, weight, bias, pad_to=cols.shape[1])
        V = Do * Ho * Wo
        if groups:
            z, _, _ = K.groupnorm_silu_fwd(z, params[pi + 2], params[pi + 3], xB, V, cout, groups, out_dtype=torch.float32)
            pi += 4
        else:
            pi += 2
        a = z
        strides = (V * cout, 1, Ho * Wo * cout, Wo * cout, cout)
        dims = (cout, Do, Ho, Wo)
    Cout, Dd, Hd, Wd = dims
    n = Dd * Hd * Wd * Cout
    if pos_embed.numel() != n:
        raise RuntimeError(
            f"voxel_embed emits a {Dd}x{Hd}x{Wd} token grid x {Cout} channels but pos_embed has "
            f"{tuple(pos_embed.shape)}: the tensor sizes must match (the committed reference has this "
            "defect at 128^3; construct HybridViT3D(token_grid='conv') or token_grid=16)")
    return K.add_pos(z.view(xB, n), pos_embed.detach().contiguous().view(-1), B).view(B * Dd * Hd * Wd, Cout)


def backbone(model, x, context, cond, prev_stage_embed=None):
    B = x.shape[0]
    D, H, W = model.volume_size
    Dd, Hd, Wd = model.downsampled_size
    N, M = Dd * Hd * Wd, context.shape[1]
    tok = voxel_embed(x, model.pos_embed, model._plan, model._embed_params(), B)
    ctx32 = cast_tokens(context)
    for blk in model.blocks:
        tok = block_tokens(blk, tok, ctx32, blk._combined_cond(cond, prev_stage_embed, B), B, N, M)
    v, _, _ = K.head_fwd(tok, model.norm.weight, model.norm.bias, model.output_proj.weight.reshape(-1), model.output_proj.bias)
    return K.upsample3d_fwd(v, B, (Dd, Hd, Wd), (D, H, W))
