"""Autograd Functions of the hot path: each forward/backward is a fixed sequence of libhvc_sm100a kernels.

Precision policy (matches what ``torch.autocast(bf16)`` does to the reference modules, but keeps more
in fp32): parameters and the residual stream stay fp32; GEMM / attention operands are bf16 with fp32
accumulation; LayerNorm / GroupNorm statistics, softmax statistics, gates and residual adds are fp32.

Gradient layouts are what ``nn.Linear`` expects ((out, in) weights) so optimizers, clipping and
checkpoints are untouched.  Nothing here falls back to torch math: torch only allocates.
"""
import weakref

import torch
from torch.autograd import Function

from . import kernels as K

_SM_COUNT = {}


def _sms(dev):
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    if i not in _SM_COUNT:
        _SM_COUNT[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SM_COUNT[i]


# stride-1 embedding convs with Cin % 64 == 0 run as implicit GEMMs (hvc_conv_taps); False = the patch-matrix path (tests compare the two)
IMPLICIT_EMBED = True
# ... when their patch matrix would be at least this large; below it the matrix stays L2-resident and the patch path has fewer launches
# (64^3 direct regression, one embedded sample: 14 / 28 MB, launch-bound)
IMPLICIT_EMBED_MIN_BYTES = 64 << 20

# ------------------------------------------------------------------ bf16 weight cache
_W16 = {}


def _cache_get(cache, p, pad_to):
    """Entry of a per-parameter cache if it still describes `p`: same object (ids are reused after garbage collection,
    so the entry holds a weak reference), same in-place version, same storage and shape."""
    ent = cache.get((id(p), pad_to))
    if ent is not None and ent[0]() is p and ent[1] == p._version and ent[2] == p.data_ptr() and ent[4] == tuple(p.shape):
        return ent[3]
    return None


def _cache_put(cache, p, pad_to, t):
    if len(cache) > 256:      # drop entries of parameters that no longer exist
        for k in [k for k, e in cache.items() if e[0]() is None]:
            del cache[k]
    cache[(id(p), pad_to)] = (weakref.ref(p), p._version, p.data_ptr(), t, tuple(p.shape))


def w16(p, pad_to=None):
    """bf16 copy of an fp32 parameter viewed [out, in] (zero-padded along `in` to pad_to), cached until
    the parameter is modified in place (optimizer step) or replaced (load_state_dict)."""
    t = _cache_get(_W16, p, pad_to)
    if t is not None:
        return t
    src = p.detach()
    if src.dtype != torch.float32:
        src = src.float()
    src = src.reshape(src.shape[0], -1).contiguous()
    if pad_to is not None and pad_to != src.shape[1]:
        padded = torch.zeros(src.shape[0], pad_to, device=src.device, dtype=torch.float32)
        padded[:, :src.shape[1]] = src
        src = padded
    t = K.cast_bf16(src)
    _cache_put(_W16, p, pad_to, t)
    return t


def w16_taps(p):
    """bf16 [Cout, 27*Cin] copy of a Conv3d weight (Cout, Cin, 3, 3, 3) in tap-major column order (kd, kh, kw, cin): the operand of
    the channels-last patch matrix (hvc_im2col3d_cl).  Cached like w16."""
    t = _cache_get(_W16, p, "taps")
    if t is not None:
        return t
    src = p.detach().float().permute(0, 2, 3, 4, 1).reshape(p.shape[0], -1).contiguous()
    t = K.cast_bf16(src)
    _cache_put(_W16, p, "taps", t)
    return t


def w16_taps_t(p, cout_pad):
    """bf16 [Cin, 27*cout_pad] transposed filter of a Conv3d weight (Cout, Cin, 3, 3, 3): column (kd, kh, kw, co), co zero-padded to
    cout_pad -- the B operand of the implicit-GEMM data gradient (hvc_conv_taps side 1 on the padded output gradient)."""
    key = ("taps_t", cout_pad)
    t = _cache_get(_W16, p, key)
    if t is not None:
        return t
    Cout, Cin = p.shape[0], p.shape[1]
    src = torch.zeros(Cin, 27, cout_pad, device=p.device, dtype=torch.float32)
    src[:, :, :Cout] = p.detach().float().reshape(Cout, Cin, 27).permute(1, 2, 0)
    t = K.cast_bf16(src.view(Cin, 27 * cout_pad))
    _cache_put(_W16, p, key, t)
    return t


# taps of a stride-2 conv grouped by the parity volume they read (K.conv_tap_offsets_s2): tap order, group boundaries, device index
_S2_TAPS = [t for par in range(8) for t in range(27) if ((t // 9 != 1) * 4 + ((t // 3) % 3 != 1) * 2 + (t % 3 != 1)) == par]
_S2_BOUNDS = [sum(1 for t in range(27) if ((t // 9 != 1) * 4 + ((t // 3) % 3 != 1) * 2 + (t % 3 != 1)) < par) for par in range(9)]
_S2_PERM = {}


def w16_taps_t_s2(p, cout_pad):
    """w16_taps_t with the 27 tap blocks reordered parity volume by parity volume, so that the B operand of each of the eight
    data-gradient GEMMs of a stride-2 implicit conv is a column slice [cin, n_taps(par)*cout_pad] of ONE cached matrix."""
    key = ("taps_t_s2", cout_pad)
    t = _cache_get(_W16, p, key)
    if t is not None:
        return t
    dev = p.device
    if dev not in _S2_PERM:
        _S2_PERM[dev] = torch.tensor(_S2_TAPS, device=dev, dtype=torch.long)
    t = w16_taps_t(p, cout_pad).view(p.shape[1], 27, cout_pad).index_select(1, _S2_PERM[dev]).view(p.shape[1], 27 * cout_pad)
    _cache_put(_W16, p, key, t)
    return t


_CACHES = [_W16]        # every per-parameter cache of derived operands registers here (ops_fp32._W6, xray_encoder._W3)


def clear_weight_cache():
    """Drop every cached derived operand: call after parameters were written behind autograd's version counters
    (optim.FlatAdamW does)."""
    for c in _CACHES:
        c.clear()


# ------------------------------------------------------------------ GEMM helpers (no autograd)

def _wgrad(dy16, x16):
    """dW[N,K] = dy[T,N]^T x[T,K], fp32, split over T."""
    T, N = dy16.shape
    Kd = x16.shape[1]
    tiles = ((N + 127) // 128) * ((Kd + 127) // 128)
    kb = (T + 63) // 64
    splits = max(1, min(kb, (4 * _sms(dy16.device)) // max(tiles, 1)))
    return K.gemm(dy16, x16, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=splits)


def _dgrad(dy16, w_16, f32_out=False, **kw):
    """dx[T,K] = dy[T,N] W[N,K] (W as stored)."""
    if f32_out:
        return K.gemm(dy16, w_16, b_major=1, epilogue=K.EPI_F32)
    return K.gemm(dy16, w_16, b_major=1, **kw)


def _drops(drop):
    """drop = None | (seed, site_base, p_first, p_second) -> the two kernel-level dropout sites of a sub-block
    (attention probabilities / projection output, or MLP activation / MLP output)."""
    if drop is None:
        return None, None
    seed, base, pa, pb = drop
    return (K.Drop(seed, base, pa) if pa > 0 else None), (K.Drop(seed, base + 1, pb) if pb > 0 else None)


def _mod_views(mod, off, C):
    return mod[:, off:off + C], mod[:, off + C:off + 2 * C], mod[:, off + 2 * C:off + 3 * C]


# ------------------------------------------------------------------ casts

class CastTokens(Function):
    """(B, M, C) f32|bf16 with any strides -> bf16 [B*M, C]; backward returns the f32/bf16 gradient as (B, M, C)."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        ctx.in_dtype = x.dtype
        return K.cast_tokens(x)

    @staticmethod
    def backward(ctx, g):
        B, M, C = ctx.shape
        return g.to(ctx.in_dtype).view(B, M, C)


# ------------------------------------------------------------------ AdaLN modulation linear (a3)

class AdaLN(Function):
    """params[B, 6C] = cond W^T + b  (vit_components.py:144), fp32."""

    @staticmethod
    def forward(ctx, cond, weight, bias):
        cond = cond.float().contiguous()
        ctx.save_for_backward(cond, weight)
        return K.adaln_fwd(cond, weight.contiguous(), bias)

    @staticmethod
    def backward(ctx, g):
        cond, weight = ctx.saved_tensors
        need_w = ctx.needs_input_grad[1]
        dW, db, dcond = K.adaln_bwd(g.contiguous(), cond, weight.contiguous(), need_w=need_w,
                                    need_cond=ctx.needs_input_grad[0])
        return dcond, dW, db


# ------------------------------------------------------------------ self-attention sub-block (a1 inside a5)

class SelfAttnBranch(Function):
    """x + gate_sa * proj(attn(qkv((1+scale_sa) * LN1(x) + shift_sa)))   hybrid_vit_backbone.py:120-123."""

    @staticmethod
    def forward(ctx, x, mod, ln_w, ln_b, w_qkv, w_proj, b_proj, B, N, H, off, drop=None):
        T, C = x.shape
        d = C // H
        shift, scale, gate = _mod_views(mod, off, C)
        d_attn, d_proj = _drops(drop)
        y, mean, rstd = K.ln_fwd(x, ln_w, ln_b, shift, scale, mod.stride(0), N)
        qkv = K.gemm(y, w16(w_qkv))
        o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, d, d ** -0.5, drop=d_attn)
        branch = torch.empty(T, C, device=x.device, dtype=torch.bfloat16)
        out = K.gemm(o, w16(w_proj), epilogue=K.EPI_RESIDUAL, bias=b_proj, resid=x, gate=gate, gate_ld=mod.stride(0),
                     rows_per_batch=N, out2=branch, drop=d_proj)
        ctx.save_for_backward(x, mod, ln_w, ln_b, w_qkv, w_proj, mean, rstd, y, qkv, o, lse, branch)
        ctx.dims = (B, N, H, off)
        ctx.drop = drop
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mod, ln_w, ln_b, w_qkv, w_proj, mean, rstd, y, qkv, o, lse, branch = ctx.saved_tensors
        B, N, H, off = ctx.dims
        T, C = x.shape
        d = C // H
        dout = dout.contiguous()
        shift, scale, gate = _mod_views(mod, off, C)
        d_attn, d_proj = _drops(ctx.drop)
        dbranch, dgate, db_proj = K.resid_bwd(dout, B, N, branch=branch, gate=gate, gate_ld=mod.stride(0), drop=d_proj)
        dw_proj = _wgrad(dbranch, o)
        d_o = _dgrad(dbranch, w16(w_proj))
        dqkv = torch.empty(T, 3 * C, device=x.device, dtype=torch.bfloat16)
        K.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, d_o, B, H, N, N, d, d ** -0.5,
                   dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], drop=d_attn)
        dw_qkv = _wgrad(dqkv, y)
        dy = _dgrad(dqkv, w16(w_qkv))
        r = K.ln_bwd(dy, x, mean, rstd, ln_w, ln_b, B, N, scale=scale, mod_ld=mod.stride(0), dx_in=dout, want_mod=True)
        dmod = torch.zeros_like(mod)
        dmod[:, off:off + C] = r["dmod"][:, 0]
        dmod[:, off + C:off + 2 * C] = r["dmod"][:, 1]
        dmod[:, off + 2 * C:off + 3 * C] = dgate
        return r["dx"], dmod, r["dw"], r["db"], dw_qkv, dw_proj, db_proj, None, None, None, None, None


# ------------------------------------------------------------------ cross-attention sub-block (a2 inside a5)

class CrossAttnBranch(Function):
    """x + proj(attn(q(LN2(x)), kv(context)))   hybrid_vit_backbone.py:126-128."""

    @staticmethod
    def forward(ctx, x, ctx16, ln_w, ln_b, w_q, w_kv, w_proj, b_proj, B, N, M, H, store_probs=False, drop=None):
        T, C = x.shape
        d = C // H
        d_attn, d_proj = _drops(drop)
        y, mean, rstd = K.ln_fwd(x, ln_w, ln_b)
        q = K.gemm(y, w16(w_q))
        kv = K.gemm(ctx16, w16(w_kv))
        res = K.attn_fwd(q, kv[:, :C], kv[:, C:], B, H, N, M, d, d ** -0.5, want_probs=store_probs, drop=d_attn)
        o, lse = res[0], res[1]
        out = K.gemm(o, w16(w_proj), epilogue=K.EPI_RESIDUAL, bias=b_proj, resid=x, drop=d_proj)
        ctx.save_for_backward(x, ctx16, ln_w, ln_b, w_q, w_kv, w_proj, mean, rstd, y, q, kv, o, lse)
        ctx.dims = (B, N, M, H)
        ctx.drop = drop
        if store_probs:      # attention map (B, H, N, M): a detached diagnostic output (vit_components.py:106-108)
            ctx.mark_non_differentiable(res[2])
            return out, res[2]
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        x, ctx16, ln_w, ln_b, w_q, w_kv, w_proj, mean, rstd, y, q, kv, o, lse = ctx.saved_tensors
        B, N, M, H = ctx.dims
        T, C = x.shape
        d = C // H
        dout = dout.contiguous()
        d_attn, d_proj = _drops(ctx.drop)
        dbranch, _, db_proj = K.resid_bwd(dout, B, N, drop=d_proj)
        dw_proj = _wgrad(dbranch, o)
        d_o = _dgrad(dbranch, w16(w_proj))
        dq = torch.empty(T, C, device=x.device, dtype=torch.bfloat16)
        dkv = torch.empty(B * M, 2 * C, device=x.device, dtype=torch.bfloat16)
        K.attn_bwd(q, kv[:, :C], kv[:, C:], o, lse, d_o, B, H, N, M, d, d ** -0.5, dq, dkv[:, :C], dkv[:, C:], drop=d_attn)
        dw_q = _wgrad(dq, y)
        dy = _dgrad(dq, w16(w_q))
        dw_kv = _wgrad(dkv, ctx16)
        dctx = _dgrad(dkv, w16(w_kv), f32_out=True) if ctx.needs_input_grad[1] else None
        r = K.ln_bwd(dy, x, mean, rstd, ln_w, ln_b, B, N, dx_in=dout)
        return r["dx"], dctx, r["dw"], r["db"], dw_q, dw_kv, dw_proj, db_proj, None, None, None, None, None, None


# ------------------------------------------------------------------ MLP sub-block (inside a5)

class MlpBranch(Function):
    """x + gate_mlp * W2 gelu(W1 ((1+scale_mlp) * LN3(x) + shift_mlp) + b1) + b2   hybrid_vit_backbone.py:136-139."""

    @staticmethod
    def forward(ctx, x, mod, ln_w, ln_b, w1, b1, w2, b2, B, N, off, drop=None):
        T, C = x.shape
        shift, scale, gate = _mod_views(mod, off, C)
        d_act, d_out = _drops(drop)
        y, mean, rstd = K.ln_fwd(x, ln_w, ln_b, shift, scale, mod.stride(0), N)
        h = torch.empty(T, w1.shape[0], device=x.device, dtype=torch.bfloat16)
        g = K.gemm(y, w16(w1), bias=b1, activation=K.ACT_GELU, out2=h, drop=d_act)     # g = Dropout(GELU(h)), :77-78
        branch = torch.empty(T, C, device=x.device, dtype=torch.bfloat16)
        out = K.gemm(g, w16(w2), epilogue=K.EPI_RESIDUAL, bias=b2, resid=x, gate=gate, gate_ld=mod.stride(0),
                     rows_per_batch=N, out2=branch, drop=d_out)
        ctx.save_for_backward(x, mod, ln_w, ln_b, w1, w2, mean, rstd, y, h, g, branch)
        ctx.dims = (B, N, off)
        ctx.drop = drop
        return out

    @staticmethod
    def backward(ctx, dout):
        x, mod, ln_w, ln_b, w1, w2, mean, rstd, y, h, g, branch = ctx.saved_tensors
        B, N, off = ctx.dims
        T, C = x.shape
        dout = dout.contiguous()
        shift, scale, gate = _mod_views(mod, off, C)
        d_act, d_out = _drops(ctx.drop)
        dbranch, dgate, db2 = K.resid_bwd(dout, B, N, branch=branch, gate=gate, gate_ld=mod.stride(0), drop=d_out)
        dw2 = _wgrad(dbranch, g)
        dh = _dgrad(dbranch, w16(w2), activation=K.ACT_GELU_GRAD, aux=h, drop=d_act)
        db1 = K.colsum_bf16(dh)
        dw1 = _wgrad(dh, y)
        dy = _dgrad(dh, w16(w1))
        r = K.ln_bwd(dy, x, mean, rstd, ln_w, ln_b, B, N, scale=scale, mod_ld=mod.stride(0), dx_in=dout, want_mod=True)
        dmod = torch.zeros_like(mod)
        dmod[:, off:off + C] = r["dmod"][:, 0]
        dmod[:, off + C:off + 2 * C] = r["dmod"][:, 1]
        dmod[:, off + 2 * C:off + 3 * C] = dgate
        return r["dx"], dmod, r["dw"], r["db"], dw1, db1, dw2, db2, None, None, None, None


# ------------------------------------------------------------------ standalone attention modules (a1, a2)

def _proj_out_grad(dout, B, N, C, d_proj):
    """Gradient entering the output projection of a standalone attention module: bf16 [B*N, C] (through proj_drop's
    mask when it was applied) and the bias gradient."""
    if d_proj is None:
        dy16 = K.cast_tokens(dout)
        return dy16, K.colsum_bf16(dy16)
    dy16, _, db = K.resid_bwd(dout.float().contiguous().view(B * N, C), B, N, drop=d_proj)
    return dy16, db


class SelfAttention(Function):
    """proj(attn(qkv(x)))  -- MultiHeadSelfAttention.forward, vit_components.py:31-57 (dropout off)."""

    @staticmethod
    def forward(ctx, x, w_qkv, w_proj, b_proj, H, drop=None):
        B, N, C = x.shape
        d = C // H
        d_attn, d_proj = _drops(drop)
        x16 = K.cast_tokens(x)
        qkv = K.gemm(x16, w16(w_qkv))
        o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, d, d ** -0.5, drop=d_attn)
        f32 = x.dtype == torch.float32
        out = K.gemm(o, w16(w_proj), bias=b_proj, epilogue=K.EPI_F32 if f32 else K.EPI_BF16, drop=d_proj)
        ctx.save_for_backward(x16, w_qkv, w_proj, qkv, o, lse)
        ctx.dims = (B, N, C, H, x.dtype)
        ctx.drop = drop
        return out.view(B, N, C)

    @staticmethod
    def backward(ctx, dout):
        x16, w_qkv, w_proj, qkv, o, lse = ctx.saved_tensors
        B, N, C, H, dt = ctx.dims
        d = C // H
        d_attn, d_proj = _drops(ctx.drop)
        dy16, db_proj = _proj_out_grad(dout, B, N, C, d_proj)
        dw_proj = _wgrad(dy16, o)
        d_o = _dgrad(dy16, w16(w_proj))
        dqkv = torch.empty(B * N, 3 * C, device=dout.device, dtype=torch.bfloat16)
        K.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, d_o, B, H, N, N, d, d ** -0.5,
                   dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], drop=d_attn)
        dw_qkv = _wgrad(dqkv, x16)
        dx = _dgrad(dqkv, w16(w_qkv), f32_out=(dt == torch.float32)).view(B, N, C)
        return dx, dw_qkv, dw_proj, db_proj, None, None


class CrossAttention(Function):
    """proj(attn(q(x), kv(context)))  -- MultiHeadCrossAttention.forward, vit_components.py:83-119 (dropout off)."""

    @staticmethod
    def forward(ctx, x, context, w_q, w_kv, w_proj, b_proj, H, store_probs=False, drop=None):
        B, N, C = x.shape
        M, Cc = context.shape[1], context.shape[2]
        d = C // H
        d_attn, d_proj = _drops(drop)
        x16 = K.cast_tokens(x)
        c16 = K.cast_tokens(context)
        q = K.gemm(x16, w16(w_q))
        kv = K.gemm(c16, w16(w_kv))
        res = K.attn_fwd(q, kv[:, :C], kv[:, C:], B, H, N, M, d, d ** -0.5, want_probs=store_probs, drop=d_attn)
        o, lse = res[0], res[1]
        f32 = x.dtype == torch.float32
        out = K.gemm(o, w16(w_proj), bias=b_proj, epilogue=K.EPI_F32 if f32 else K.EPI_BF16, drop=d_proj)
        ctx.save_for_backward(x16, c16, w_q, w_kv, w_proj, q, kv, o, lse)
        ctx.dims = (B, N, M, C, Cc, H, x.dtype, context.dtype)
        ctx.drop = drop
        if store_probs:
            ctx.mark_non_differentiable(res[2])
            return out.view(B, N, C), res[2]
        return out.view(B, N, C)

    @staticmethod
    def backward(ctx, dout, *unused):
        x16, c16, w_q, w_kv, w_proj, q, kv, o, lse = ctx.saved_tensors
        B, N, M, C, Cc, H, dt, cdt = ctx.dims
        d = C // H
        d_attn, d_proj = _drops(ctx.drop)
        dy16, db_proj = _proj_out_grad(dout, B, N, C, d_proj)
        dw_proj = _wgrad(dy16, o)
        d_o = _dgrad(dy16, w16(w_proj))
        dq = torch.empty(B * N, C, device=dout.device, dtype=torch.bfloat16)
        dkv = torch.empty(B * M, 2 * C, device=dout.device, dtype=torch.bfloat16)
        K.attn_bwd(q, kv[:, :C], kv[:, C:], o, lse, d_o, B, H, N, M, d, d ** -0.5, dq, dkv[:, :C], dkv[:, C:], drop=d_attn)
        dw_q = _wgrad(dq, x16)
        dw_kv = _wgrad(dkv, c16)
        dx = _dgrad(dq, w16(w_q), f32_out=(dt == torch.float32)).view(B, N, C)
        dctx = None
        if ctx.needs_input_grad[1]:
            dctx = _dgrad(dkv, w16(w_kv), f32_out=(cdt == torch.float32)).view(B, M, Cc)
        return dx, dctx, dw_q, dw_kv, dw_proj, db_proj, None, None, None


# ------------------------------------------------------------------ voxel embedding + positional encoding (a7 head of forward)

class VoxelEmbed(Function):
    """tokens[B*N, C] = flatten(voxel_embed(x)) + pos_embed   hybrid_vit_backbone.py:252-258.

    plan: list of (cin, cout, stride, groups|0); params: for each conv (weight, bias[, gn_weight, gn_bias]).
    A batch-expanded input (stride(0) == 0, model_direct.py:75) is embedded once and broadcast.
    """

    @staticmethod
    def forward(ctx, x, pos_embed, plan, *params):
        B, Cin, D, H, W = x.shape
        xB = 1 if (B > 1 and x.stride(0) == 0) else B
        xin = x[:xB]
        if xin.dtype not in (torch.float32, torch.bfloat16):
            xin = xin.float()
        saved, geoms = [], []
        a, strides, dims = xin, tuple(xin.stride()), (Cin, D, H, W)
        z = None
        pi = 0
        for li, (cin, cout, stride, groups) in enumerate(plan):
            _, Dc, Hc, Wc = dims
            weight, bias = params[pi], params[pi + 1]
            # channels-last inputs (every layer after the first, and the stage wrappers' 32-channel volume) use the tap-major patch
            # matrix: whole 8-channel runs per access in im2col and col2im (hvc_im2col3d_cl).  A stride-1 layer whose input has a
            # multiple of 64 channels (the last conv of every stack in the reference's configurations) needs no patch matrix at all:
            # implicit GEMM on the zero-padded volume (hvc_conv_taps)
            tm = cin % 8 == 0 and strides[1] == 1 and all(s % 8 == 0 for s in strides[:1] + strides[2:]) and (li > 0 or xB == B)
            Do, Ho, Wo = K.conv_out(Dc, stride), K.conv_out(Hc, stride), K.conv_out(Wc, stride)
            implicit = IMPLICIT_EMBED and tm and cin % 64 == 0 and li > 0 and cout % 8 == 0 and \
                (stride == 1 or (Dc % 2 == 0 and Hc % 2 == 0 and Wc % 2 == 0)) and \
                xB * Do * Ho * Wo * 27 * cin * 2 >= IMPLICIT_EMBED_MIN_BYTES
            if implicit and stride == 1:
                cols = K.pad3d_cl(a.view(xB, Dc, Hc, Wc, cin), xB, Dc, Hc, Wc, cin, cin)
                zp = K.gemm(cols.view(-1, cin), w16_taps(weight), bias=bias, epilogue=K.EPI_F32, taps=(1, cin, K.conv_tap_offsets(Hc, Wc)))
                z = K.unpad3d_cl(zp, xB, Dc, Hc, Wc, cout).view(-1, cout)
                del zp
            elif implicit:
                # stride 2: eight parity volumes (low-side padded, stacked along the rows); the output lives on the same padded
                # half-resolution grid, so every tap is again a row shift (include/hvc.h, hvc_conv_taps)
                rows_p = xB * (Do + 1) * (Ho + 1) * (Wo + 1)
                cols = K.s2d_pad_cl(a.view(xB, Dc, Hc, Wc, cin), xB, Dc, Hc, Wc, cin)
                offs = [par * rows_p + sh for par, sh in K.conv_tap_offsets_s2(rows_p, Ho + 1, Wo + 1)]
                zp = K.gemm(cols.view(-1, cin), w16_taps(weight), bias=bias, epilogue=K.EPI_F32, taps=(1, cin, offs), m_rows=rows_p)
                z = K.unpad3d_cl(zp, xB, Do, Ho, Wo, cout, pad_hi=0).view(-1, cout)
                del zp
            else:
                cols = K.im2col3d(a, xB, cin, Dc, Hc, Wc, stride, strides, tap_major=tm)
                z = K.gemm(cols, w16_taps(weight) if tm else w16(weight, pad_to=cols.shape[1]), bias=bias, epilogue=K.EPI_F32)   # [xB*V, cout] channels-last
            tm = (2 if stride == 1 else 3) if implicit else int(tm)
            V = Do * Ho * Wo
            geoms.append((cin, Dc, Hc, Wc, stride, strides, V, tm))
            if groups:
                gw, gb = params[pi + 2], params[pi + 3]
                last = li == len(plan) - 1          # a stack that ends in GN+SiLU feeds the fp32 token stream
                act, mean, rstd = K.groupnorm_silu_fwd(z, gw, gb, xB, V, cout, groups,
                                                       out_dtype=torch.float32 if last else torch.bfloat16)
                saved += [cols, z, mean, rstd]
                a = act
                if last:
                    z = act
                pi += 4
            else:
                saved += [cols]
                a = None
                pi += 2
            strides = (V * cout, 1, Ho * Wo * cout, Wo * cout, cout)
            dims = (cout, Do, Ho, Wo)
        Cout, Dd, Hd, Wd = dims
        n = Dd * Hd * Wd * Cout
        if pos_embed.numel() != n:
            raise RuntimeError(
                f"voxel_embed emits a {Dd}x{Hd}x{Wd} token grid x {Cout} channels but pos_embed has "
                f"{tuple(pos_embed.shape)}: the tensor sizes must match (the committed reference has this "
                "defect at 128^3; construct HybridViT3D(token_grid='conv') or token_grid=16)")
        tokens = K.add_pos(z.view(xB, n), pos_embed.contiguous().view(-1), B)
        ctx.save_for_backward(*saved, *params)
        ctx.meta = (plan, geoms, xB, B, tuple(x.shape), len(saved), n, Cout)
        ctx.in_strides = tuple(xin.stride())
        return tokens.view(B * Dd * Hd * Wd, Cout)

    @staticmethod
    def backward(ctx, dtok):
        plan, geoms, xB, B, xshape, nsaved, n, Cout = ctx.meta
        tensors = ctx.saved_tensors          # read once: torch.utils.checkpoint allows a single unpack per tensor
        saved, params = tensors[:nsaved], tensors[nsaved:]
        dtok = dtok.contiguous().view(B, n)
        dpos = K.batch_sum(dtok)
        dz = dpos.view(1, n) if xB != B else dtok
        dz = dz.reshape(-1, Cout)
        grads = [None] * len(params)
        # walk the stack backwards
        si, pi = nsaved, len(params)
        dx = None
        for li in range(len(plan) - 1, -1, -1):
            cin, cout, stride, groups = plan[li]
            cin_, Dc, Hc, Wc, stride_, strides, V, tm = geoms[li]
            if groups:
                cols, z, mean, rstd = saved[si - 4:si]
                si -= 4
                weight, bias, gw, gb = params[pi - 4:pi]
                pi -= 4
                dz, dgw, dgb = K.groupnorm_silu_bwd(dz.contiguous(), z, gw, gb, mean, rstd, xB, V, cout, groups)
                grads[pi + 2], grads[pi + 3] = dgw, dgb
            else:
                cols = saved[si - 1]
                si -= 1
                weight, bias = params[pi - 2:pi]
                pi -= 2
            dz16 = K.cast_bf16(dz.contiguous())
            grads[pi + 1] = K.colsum_bf16(dz16)
            need_dx = li > 0 or ctx.needs_input_grad[0]
            if tm >= 2:       # implicit GEMM: padded output gradient (channels padded to whole 64-wide k-blocks), no patch matrices
                cp = (cout + 63) // 64 * 64
                tiles = ((cp + 127) // 128) * ((27 * cin + 127) // 128)
                if tm == 2:
                    dzp = K.pad3d_cl(dz16.view(xB, Dc, Hc, Wc, cout), xB, Dc, Hc, Wc, cout, cp).view(-1, cp)
                    offs = K.conv_tap_offsets(Hc, Wc)
                else:
                    Do, Ho, Wo = Dc // 2, Hc // 2, Wc // 2
                    dzp = K.pad3d_cl(dz16.view(xB, Do, Ho, Wo, cout), xB, Do, Ho, Wo, cout, cp, pad_hi=0).view(-1, cp)
                    rows_p = dzp.shape[0]
                    tt = K.conv_tap_offsets_s2(rows_p, Ho + 1, Wo + 1)
                    offs = [par * rows_p + sh for par, sh in tt]
                splits = max(1, min(dzp.shape[0] // 64, (16 * _sms(dtok.device)) // tiles))
                dwp = K.gemm(dzp, cols.view(-1, cin), a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=splits,
                             taps=(2, cin, offs))[:cout]
                grads[pi] = dwp.view(cout, 3, 3, 3, cin).permute(0, 4, 1, 2, 3).contiguous()
                if tm == 2:
                    dxp = K.gemm(dzp, w16_taps_t(weight, cp), epilogue=K.EPI_F32, taps=(1, cp, K.conv_tap_offsets(Hc, Wc, -1)))
                    dz = K.unpad3d_cl(dxp, xB, Dc, Hc, Wc, cin).view(-1, cin)
                else:
                    # data gradient, one GEMM per parity volume over the taps that read it: dX_par[r] = sum_t dZ[r - shift_t] W_t
                    wt = w16_taps_t_s2(weight, cp)                 # tap blocks grouped by parity volume: column slices, no copies
                    dxp = torch.empty(8 * rows_p, cin, device=dtok.device, dtype=torch.float32)
                    for par in range(8):
                        ts = _S2_TAPS[_S2_BOUNDS[par]:_S2_BOUNDS[par + 1]]
                        K.gemm(dzp, wt[:, _S2_BOUNDS[par] * cp:_S2_BOUNDS[par + 1] * cp], epilogue=K.EPI_F32,
                               taps=(1, cp, [-tt[t][1] for t in ts]), out=dxp[par * rows_p:(par + 1) * rows_p])
                    dz = K.d2s_unpad_cl(dxp, xB, Dc, Hc, Wc, cin).view(-1, cin)
                del dxp, dzp
                continue
            dwp = _wgrad(dz16, cols)                                    # [cout, Kp]
            grads[pi] = dwp.view(cout, 3, 3, 3, cin).permute(0, 4, 1, 2, 3).contiguous() if tm else dwp[:, :cin * 27].reshape(weight.shape)
            if need_dx:
                dcols = _dgrad(dz16, w16_taps(weight) if tm else w16(weight, pad_to=cols.shape[1]))
                if li > 0:
                    prev_c = plan[li - 1][1]
                    d_act = torch.empty(xB * Dc * Hc * Wc, prev_c, device=dtok.device, dtype=torch.float32)
                    K.col2im3d(dcols, xB, cin, Dc, Hc, Wc, stride, d_act, strides, tap_major=bool(tm))
                    dz = d_act
                else:
                    # same memory layout as the forward input (e.g. the channels-last view a stage wrapper hands over)
                    dx1 = torch.empty_strided((xB,) + tuple(xshape[1:]), ctx.in_strides, device=dtok.device, dtype=torch.float32) \
                        if xB == B else torch.empty((xB,) + tuple(xshape[1:]), device=dtok.device, dtype=torch.float32)
                    K.col2im3d(dcols, xB, cin, Dc, Hc, Wc, stride, dx1, tuple(dx1.stride()), tap_major=bool(tm))
                    if xB != B:
                        dx = torch.zeros(xshape, device=dtok.device, dtype=torch.float32)
                        dx[0] = dx1[0]
                    else:
                        dx = dx1
        return (dx, dpos.view(1, -1, Cout), None) + tuple(grads)


# ------------------------------------------------------------------ output head (a7 tail of forward)

class OutputHead(Function):
    """upsample(reshape(output_proj(LayerNorm(tokens))))   hybrid_vit_backbone.py:265-272."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, wo, bo, B, grid, size):
        v, mean, rstd = K.head_fwd(x, ln_w, ln_b, wo.reshape(-1), bo)
        out = K.upsample3d_fwd(v, B, grid, size)
        ctx.save_for_backward(x, ln_w, ln_b, wo, mean, rstd)
        ctx.meta = (B, grid, size)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, ln_w, ln_b, wo, mean, rstd = ctx.saved_tensors
        B, grid, size = ctx.meta
        dv = K.upsample3d_bwd(dout.float().contiguous(), B, grid, size)
        N = grid[0] * grid[1] * grid[2]
        r = K.ln_bwd(dv, x, mean, rstd, ln_w, ln_b, B, N, mult_vec=wo.reshape(-1).contiguous(), head=True)
        return r["dx"], r["dw"], r["db"], r["dvec"].view_as(wo), r["dscalar"], None, None, None
