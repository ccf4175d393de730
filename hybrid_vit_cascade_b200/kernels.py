"""Thin tensor-level wrappers over the C ABI (no autograd here; see ops.py).

torch is plumbing only: it owns the device memory and the stream; every computation below is a
kernel of libhvc_sm100a.so.  All wrappers require CUDA tensors and raise otherwise.
"""
import ctypes as C

import torch

from . import _lib

EPI_BF16, EPI_RESIDUAL, EPI_F32_ATOMIC, EPI_F32 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2


# Optional per-call device timing (bench.py's live roofline measurement): when set to a dict, the wrappers listed
# in _timed() bracket their launches with CUDA events on the current stream (the stream the kernels run on).
_PROF = None


def set_profiler(store):
    global _PROF
    _PROF = store


class _timed:
    def __init__(self, name, flops):
        self.name, self.flops = name, flops

    def __enter__(self):
        if _PROF is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _PROF is not None:
            self.e1.record()
            _PROF.setdefault(self.name, []).append((self.e0, self.e1, self.flops))
        return False


class Drop:
    """One dropout site: (seed, site, p).  seed = int32[2] device tensor drawn from torch's CUDA generator (new_seed),
    site separates the sites sharing it, p the drop probability.  None / p == 0 means no dropout."""
    __slots__ = ("seed", "site", "p")

    def __init__(self, seed, site, p):
        self.seed, self.site, self.p = seed, int(site), float(p)


def new_seed(device):
    """Two random 32-bit words on the device, taken from torch's CUDA generator: no host sync, and
    torch.utils.checkpoint (which restores the generator state before recomputing) replays the same words."""
    return torch.randint(-2 ** 31, 2 ** 31 - 1, (2,), dtype=torch.int32, device=device)


def _set_drop(args, drop):
    if drop is not None and drop.p > 0.0:
        assert drop.seed.is_cuda and drop.seed.dtype == torch.int32 and drop.seed.numel() == 2
        args.drop.seed, args.drop.site, args.drop.p = drop.seed.data_ptr(), drop.site, drop.p


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    """Every tensor on one CUDA device, and that device is the CURRENT one: the kernels are launched on the current device's
    stream (_stream()) and per-device state (shared-memory opt-in, SM count) is keyed by cudaGetDevice().  A tensor of another
    GPU is an error, not a silent launch on the wrong device -- wrap the call in ``torch.cuda.device(t.device)``."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.HvcError("hybrid_vit_cascade_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise _lib.HvcError(f"hybrid_vit_cascade_b200: operands on different devices ({dev} and {t.device})")
    cur = torch.cuda.current_device()
    if dev.index is not None and dev.index != cur:
        raise _lib.HvcError(f"hybrid_vit_cascade_b200: operands live on cuda:{dev.index} but the current device is cuda:{cur}; "
                            f"run the module under torch.cuda.device({dev.index}) (or torch.cuda.set_device)")
    _lib.require_device(cur)


def _row_major_2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} "
                         f"stride {t.stride()}")
    return t.stride(0)


def gemm(a, b, *, a_major=0, b_major=0, out=None, out_dtype=torch.bfloat16, epilogue=EPI_BF16,
         activation=ACT_NONE, bias=None, out2=None, resid=None, gate=None, gate_ld=0, rows_per_batch=0,
         aux=None, alpha=1.0, k_splits=1, drop=None, taps=None, m_rows=None):
    """D[M,N] = alpha * sum_k A(m,k) B(n,k) with a fused epilogue (see include/hvc.h).

    a: bf16, stored [M,K] (a_major=0) or [K,M] (a_major=1); b: bf16, stored [N,K] or [K,N].
    taps = (side, cin, offsets): implicit 3x3x3 convolution, the operand on `side` (1 = a, 2 = b) is the zero-padded channels-last
    volume [rows, cin] and stands for its len(offsets)*cin-wide patch matrix; offsets = the row shift of every tap (hvc_conv_taps in
    include/hvc.h; conv_tap_offsets() builds them).  m_rows: M when `a` (side 1) has more rows than the output (stacked parity volumes).
    """
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lda, ldb = _row_major_2d(a, "a"), _row_major_2d(b, "b")
    M, K = (a.shape if a_major == 0 else (a.shape[1], a.shape[0]))
    N, Kb = (b.shape if b_major == 0 else (b.shape[1], b.shape[0]))
    tap_rows = 0
    if taps is not None and taps[0] == 1:
        assert a_major == 0 and K == taps[1]
        K = len(taps[2]) * K
        if m_rows is not None:
            tap_rows, M = M, m_rows
    if taps is not None and taps[0] == 2:
        assert b_major == 1 and N == taps[1]
        N = len(taps[2]) * N
        tap_rows, Kb = Kb, K
    assert K == Kb, (a.shape, b.shape, a_major, b_major)
    if out is None:
        if epilogue == EPI_BF16:
            out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
        elif epilogue == EPI_F32_ATOMIC:
            out = torch.zeros(M, N, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    args = _lib.GemmArgs()
    args.size = C.sizeof(_lib.GemmArgs)
    args.M, args.N, args.K = M, N, K
    args.A, args.lda, args.a_major = a.data_ptr(), lda, a_major
    args.B, args.ldb, args.b_major = b.data_ptr(), ldb, b_major
    args.epilogue, args.activation = epilogue, activation
    args.out, args.ldo = out.data_ptr(), _row_major_2d(out, "out")
    if out2 is not None:
        args.out2, args.ldo2 = out2.data_ptr(), _row_major_2d(out2, "out2")
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
        args.bias = bias.data_ptr()
    if resid is not None:
        assert resid.dtype == torch.float32
        args.resid, args.ldr = resid.data_ptr(), _row_major_2d(resid, "resid")
    if gate is not None:
        assert gate.dtype == torch.float32
        args.gate, args.gate_ld, args.rows_per_batch = gate.data_ptr(), gate_ld, rows_per_batch
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        args.aux, args.ldaux = aux.data_ptr(), _row_major_2d(aux, "aux")
    args.alpha = alpha
    args.k_splits = k_splits
    _set_drop(args, drop)
    if taps is not None:
        args.taps.side, args.taps.cin, args.taps.n_taps, args.taps.rows = taps[0], taps[1], len(taps[2]), tap_rows
        for i, o in enumerate(taps[2]):
            args.taps.offsets[i] = int(o)
    with _timed("gemm", 2.0 * M * N * K):
        _lib.check(_lib.lib().hvc_gemm(C.byref(args), _stream()), "hvc_gemm")
    return out


def _pad128(n):
    return (n + 127) // 128 * 128


def _attn_args(q, k, v, B, H, nq, nk, d, scale):
    args = _lib.AttnArgs()
    args.size = C.sizeof(_lib.AttnArgs)
    args.batch, args.heads, args.nq, args.nk, args.head_dim = B, H, nq, nk, d
    args.q, args.ldq = q.data_ptr(), _row_major_2d(q, "q")
    args.k, args.ldk = k.data_ptr(), _row_major_2d(k, "k")
    args.v, args.ldv = v.data_ptr(), _row_major_2d(v, "v")
    args.scale = scale
    return args


def attn_fwd(q, k, v, B, H, nq, nk, d, scale, want_probs=False, drop=None):
    """q: bf16 view [B*nq, H*d] (may be a column slice of a packed projection output), k/v: [B*nk, H*d].

    Returns (o bf16 [B*nq, H*d], lse2 f32 [B, H, pad128(nq)]) and, with want_probs (the store_attention slow
    path), the materialised softmax f32 [B, H, nq, nk] as a third element.
    """
    _need_cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.shape == (B * nq, H * d) and k.shape == (B * nk, H * d) and v.shape == (B * nk, H * d)
    o = torch.empty(B * nq, H * d, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, _pad128(nq), device=q.device, dtype=torch.float32)
    args = _attn_args(q, k, v, B, H, nq, nk, d, scale)
    args.o, args.ldo = o.data_ptr(), o.stride(0)
    args.lse = lse.data_ptr()
    _set_drop(args, drop)
    probs = None
    if want_probs:
        probs = torch.empty(B, H, nq, nk, device=q.device, dtype=torch.float32)
        args.probs = probs.data_ptr()
    with _timed("attn_fwd", 4.0 * B * H * nq * nk * d):
        _lib.check(_lib.lib().hvc_attn_fwd(C.byref(args), _stream()), "hvc_attn_fwd")
    if want_probs:
        return o, lse, probs
    return o, lse


def attn_bwd(q, k, v, o, lse, d_o, B, H, nq, nk, d, scale, dq, dk, dv, drop=None):
    """Backward of attn_fwd.  dq/dk/dv: bf16 views [B*n, H*d] (column slices of packed gradient buffers)."""
    _need_cuda(q, k, v, o, d_o)
    assert d_o.dtype == torch.bfloat16 and o.dtype == torch.bfloat16
    nq_pad = _pad128(nq)
    delta = torch.empty(3, B, H, nq_pad, device=q.device, dtype=torch.float32)
    dq_accum = torch.zeros(B, H, nq_pad, d, device=q.device, dtype=torch.float32)
    args = _attn_args(q, k, v, B, H, nq, nk, d, scale)
    args.o, args.ldo = o.data_ptr(), _row_major_2d(o, "o")
    args.lse = lse.data_ptr()
    args.d_o, args.lddo = d_o.data_ptr(), _row_major_2d(d_o, "d_o")
    args.dq, args.lddq = dq.data_ptr(), _row_major_2d(dq, "dq")
    args.dk, args.lddk = dk.data_ptr(), _row_major_2d(dk, "dk")
    args.dv, args.lddv = dv.data_ptr(), _row_major_2d(dv, "dv")
    args.delta = delta.data_ptr()
    args.dq_accum = dq_accum.data_ptr()
    _set_drop(args, drop)
    with _timed("attn_bwd", 10.0 * B * H * nq * nk * d):
        _lib.check(_lib.lib().hvc_attn_bwd(C.byref(args), _stream()), "hvc_attn_bwd")
    return dq, dk, dv


# ------------------------------------------------------------------ LayerNorm / residual / casts / AdaLN

def ln_fwd(x, w, b, shift=None, scale=None, mod_ld=0, rows_per_batch=0, out_dtype=torch.bfloat16, save_stats=True):
    """x f32 [T,C] -> (y [T,C] bf16|f32, mean [T], rstd [T]).  shift/scale: f32 views with unit inner stride."""
    _need_cuda(x, w, b)
    assert x.dtype == torch.float32
    T, Cc = x.shape
    y = torch.empty(T, Cc, device=x.device, dtype=out_dtype)
    mean = torch.empty(T, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(T, device=x.device, dtype=torch.float32) if save_stats else None
    a = _lib.LnArgs()
    a.size = C.sizeof(_lib.LnArgs)
    a.T, a.C = T, Cc
    a.x, a.ldx = x.data_ptr(), _row_major_2d(x, "x")
    a.w, a.b = w.data_ptr(), b.data_ptr()
    if scale is not None:
        a.shift, a.scale, a.mod_ld, a.rows_per_batch = shift.data_ptr(), scale.data_ptr(), mod_ld, rows_per_batch
    a.y, a.ldy, a.y_is_bf16 = y.data_ptr(), y.stride(0), int(out_dtype == torch.bfloat16)
    a.mean, a.rstd = _ptr(mean), _ptr(rstd)
    _lib.check(_lib.lib().hvc_ln_fwd(C.byref(a), _stream()), "hvc_ln_fwd")
    return y, mean, rstd


def ln_bwd(dz, x, mean, rstd, w, b, batch, rows_per_batch, scale=None, mod_ld=0, mult_vec=None, dx_in=None,
           want_mod=False, head=False):
    """Returns dict(dx, dw, db[, dshift, dscale][, dvec, dscalar]).  dz: bf16/f32 [T,C] or f32 [T] (head)."""
    _need_cuda(dz, x)
    T, Cc = x.shape
    dev = x.device
    a = _lib.LnBwdArgs()
    a.size = C.sizeof(_lib.LnBwdArgs)
    a.batch, a.rows_per_batch, a.C = batch, rows_per_batch, Cc
    if dz.dim() == 1:
        assert dz.dtype == torch.float32 and dz.is_contiguous()
        a.dz_row = dz.data_ptr()
    elif dz.dtype == torch.bfloat16:
        a.dz_bf16, a.lddz = dz.data_ptr(), _row_major_2d(dz, "dz")
    else:
        assert dz.dtype == torch.float32
        a.dz_f32, a.lddz = dz.data_ptr(), _row_major_2d(dz, "dz")
    a.x, a.ldx = x.data_ptr(), _row_major_2d(x, "x")
    a.mean, a.rstd, a.w, a.b = mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), b.data_ptr()
    if scale is not None:
        a.scale, a.mod_ld = scale.data_ptr(), mod_ld
    if mult_vec is not None:
        a.mult_vec = mult_vec.data_ptr()
    if dx_in is not None:
        a.dx_in, a.lddx_in = dx_in.data_ptr(), _row_major_2d(dx_in, "dx_in")
    out = {"dx": torch.empty(T, Cc, device=dev, dtype=torch.float32),
           "dw": torch.empty(Cc, device=dev, dtype=torch.float32),
           "db": torch.empty(Cc, device=dev, dtype=torch.float32)}
    a.dx, a.lddx = out["dx"].data_ptr(), Cc
    a.dw, a.db = out["dw"].data_ptr(), out["db"].data_ptr()
    if want_mod:
        out["dmod"] = torch.empty(batch, 2, Cc, device=dev, dtype=torch.float32)   # [:,0]=dshift [:,1]=dscale
        a.dshift, a.dscale, a.dmod_ld = out["dmod"][:, 0].data_ptr(), out["dmod"][:, 1].data_ptr(), 2 * Cc
    if head:
        out["dvec"] = torch.empty(Cc, device=dev, dtype=torch.float32)
        out["dscalar"] = torch.empty(1, device=dev, dtype=torch.float32)
        a.dvec, a.dscalar = out["dvec"].data_ptr(), out["dscalar"].data_ptr()
    scratch = torch.empty(2, batch, Cc, device=dev, dtype=torch.float32)
    a.S1, a.S2 = scratch[0].data_ptr(), scratch[1].data_ptr()
    _lib.check(_lib.lib().hvc_ln_bwd(C.byref(a), _stream()), "hvc_ln_bwd")
    return out


def resid_bwd(dout, batch, rows_per_batch, branch=None, gate=None, gate_ld=0, want_dbias=True, drop=None):
    """dout f32 [T,C] -> (dbranch bf16 [T,C], dgate f32 [batch,C] | None, dbias f32 [C] | None)."""
    _need_cuda(dout)
    T, Cc = dout.shape
    dev = dout.device
    a = _lib.ResidBwdArgs()
    a.size = C.sizeof(_lib.ResidBwdArgs)
    a.batch, a.rows_per_batch, a.C = batch, rows_per_batch, Cc
    a.dout, a.lddout = dout.data_ptr(), _row_major_2d(dout, "dout")
    dbranch = torch.empty(T, Cc, device=dev, dtype=torch.bfloat16)
    a.dbranch, a.lddbranch = dbranch.data_ptr(), Cc
    dgate = None
    if gate is not None:
        a.gate, a.gate_ld = gate.data_ptr(), gate_ld
        if branch is not None:
            a.branch, a.ldbranch = branch.data_ptr(), _row_major_2d(branch, "branch")
            dgate = torch.empty(batch, Cc, device=dev, dtype=torch.float32)
            a.dgate = dgate.data_ptr()
    dbias = torch.empty(Cc, device=dev, dtype=torch.float32) if want_dbias else None
    a.dbias = _ptr(dbias)
    d1 = torch.empty(batch, Cc, device=dev, dtype=torch.float32)
    a.D1 = d1.data_ptr()
    _set_drop(a, drop)
    _lib.check(_lib.lib().hvc_resid_bwd(C.byref(a), _stream()), "hvc_resid_bwd")
    return dbranch, dgate, dbias


def colsum_bf16(x):
    _need_cuda(x)
    assert x.dtype == torch.bfloat16
    out = torch.empty(x.shape[1], device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_colsum_bf16(_ptr(x), C.c_int64(_row_major_2d(x, "x")), x.shape[0], x.shape[1], _ptr(out),
                                          _stream()), "hvc_colsum_bf16")
    return out


def cast_bf16(x):
    """Contiguous f32 -> bf16."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().hvc_cast_bf16(_ptr(x), _ptr(y), C.c_int64(x.numel()), _stream()), "hvc_cast_bf16")
    return y


def cast_tokens(x):
    """x (B, M, C) f32|bf16 with arbitrary strides -> bf16 [B*M, C] contiguous."""
    _need_cuda(x)
    assert x.dim() == 3 and x.dtype in (torch.float32, torch.bfloat16)
    B, M, Cc = x.shape
    y = torch.empty(B * M, Cc, device=x.device, dtype=torch.bfloat16)
    sb, sm, sc = x.stride()
    _lib.check(_lib.lib().hvc_cast_tokens(_ptr(x), int(x.dtype == torch.bfloat16), C.c_int64(sb), C.c_int64(sm),
                                          C.c_int64(sc), _ptr(y), B, M, Cc, _stream()), "hvc_cast_tokens")
    return y


def adaln_fwd(cond, W, bias):
    _need_cuda(cond, W)
    assert cond.dtype == W.dtype == torch.float32 and W.is_contiguous() and cond.stride(1) == 1
    B, K = cond.shape
    J = W.shape[0]
    out = torch.empty(B, J, device=cond.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_adaln_fwd(_ptr(cond), C.c_int64(cond.stride(0)), _ptr(W), _ptr(bias), _ptr(out), B, K, J,
                                        _stream()), "hvc_adaln_fwd")
    return out


def adaln_bwd(dparams, cond, W, need_w=True, need_cond=True):
    _need_cuda(dparams, cond, W)
    assert dparams.dtype == torch.float32 and dparams.is_contiguous() and cond.stride(1) == 1
    B, K = cond.shape
    J = W.shape[0]
    dW = torch.empty(J, K, device=cond.device, dtype=torch.float32) if need_w else None
    db = torch.empty(J, device=cond.device, dtype=torch.float32) if need_w else None
    dcond = torch.empty(B, K, device=cond.device, dtype=torch.float32) if need_cond else None
    _lib.check(_lib.lib().hvc_adaln_bwd(_ptr(dparams), _ptr(cond), C.c_int64(cond.stride(0)), _ptr(W), _ptr(dW), _ptr(db),
                                        _ptr(dcond), B, K, J, _stream()), "hvc_adaln_bwd")
    return dW, db, dcond


# ------------------------------------------------------------------ embed (conv3d / groupnorm) and head

def _geom(B, Cin, D, H, W, stride, strides):
    g = _lib.Conv3dGeom()
    g.B, g.Cin, g.D, g.H, g.W, g.stride = B, Cin, D, H, W, stride
    g.sb, g.sc, g.sd, g.sh, g.sw = strides
    return g


def conv_out(n, stride):
    return (n - 1) // stride + 1


def im2col3d(x, B, Cin, D, H, W, stride, strides, tap_major=False):
    """x: f32|bf16 tensor holding the conv input with element strides (sb, sc, sd, sh, sw).
    Returns bf16 [B*Do*Ho*Wo, Kp], column cin*27 + tap, or tap*Cin + cin with tap_major (channels-last x, Cin % 8 == 0)."""
    _need_cuda(x)
    Do, Ho, Wo = conv_out(D, stride), conv_out(H, stride), conv_out(W, stride)
    Kp = (Cin * 27 + 7) // 8 * 8
    cols = torch.empty(B * Do * Ho * Wo, Kp, device=x.device, dtype=torch.bfloat16)
    g = _geom(B, Cin, D, H, W, stride, strides)
    fn, name = (_lib.lib().hvc_im2col3d_cl, "hvc_im2col3d_cl") if tap_major else (_lib.lib().hvc_im2col3d, "hvc_im2col3d")
    _lib.check(fn(_ptr(x), int(x.dtype == torch.bfloat16), C.byref(g), _ptr(cols), _stream()), name)
    return cols


def col2im3d(dcols, B, Cin, D, H, W, stride, out, strides, tap_major=False):
    """Adjoint of im2col3d into `out` (f32, pre-allocated, every element written)."""
    _need_cuda(dcols, out)
    assert dcols.dtype == torch.bfloat16 and dcols.is_contiguous() and out.dtype == torch.float32
    g = _geom(B, Cin, D, H, W, stride, strides)
    fn, name = (_lib.lib().hvc_col2im3d_cl, "hvc_col2im3d_cl") if tap_major else (_lib.lib().hvc_col2im3d, "hvc_col2im3d")
    _lib.check(fn(_ptr(dcols), C.byref(g), _ptr(out), _stream()), name)
    return out


def groupnorm_silu_fwd(x, w, b, B, V, Cc, groups, out_dtype=torch.bfloat16):
    """x f32 [B*V, C] channels-last -> (y bf16|f32 [B*V, C], mean [B,G], rstd [B,G])."""
    _need_cuda(x, w, b)
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(B * V, Cc, device=x.device, dtype=out_dtype)
    mean = torch.empty(B, groups, device=x.device, dtype=torch.float32)
    rstd = torch.empty(B, groups, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cc, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_groupnorm_silu_fwd(_ptr(x), _ptr(w), _ptr(b), B, V, Cc, groups, _ptr(y),
                                                 int(out_dtype == torch.bfloat16), _ptr(mean), _ptr(rstd),
                                                 _ptr(scratch), _stream()), "hvc_groupnorm_silu_fwd")
    return y, mean, rstd


def groupnorm_silu_bwd(dy, x, w, b, mean, rstd, B, V, Cc, groups):
    _need_cuda(dy, x)
    assert dy.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous()
    dx = torch.empty(B * V, Cc, device=x.device, dtype=torch.float32)
    dw = torch.empty(Cc, device=x.device, dtype=torch.float32)
    db = torch.empty(Cc, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cc + 2 * B * groups, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_groupnorm_silu_bwd(_ptr(dy), _ptr(x), _ptr(w), _ptr(b), _ptr(mean), _ptr(rstd), B, V, Cc, groups,
                                                 _ptr(dx), _ptr(dw), _ptr(db), _ptr(scratch), _stream()), "hvc_groupnorm_silu_bwd")
    return dx, dw, db


def add_pos(x, pos, B):
    """x f32 [xB, n] (xB divides into B by broadcast), pos f32 [n] -> out f32 [B, n]."""
    _need_cuda(x, pos)
    assert x.dtype == pos.dtype == torch.float32 and x.is_contiguous() and pos.is_contiguous()
    xB, n = x.shape
    out = torch.empty(B, n, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_add_pos(_ptr(x), xB, _ptr(pos), _ptr(out), B, C.c_int64(n), _stream()), "hvc_add_pos")
    return out


def batch_sum(x):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    B, n = x.shape
    out = torch.empty(n, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_batch_sum(_ptr(x), _ptr(out), B, C.c_int64(n), _stream()), "hvc_batch_sum")
    return out


def head_fwd(x, w, b, wo, bo):
    """x f32 [T,C] -> (v f32 [T], mean, rstd)."""
    _need_cuda(x)
    T, Cc = x.shape
    v = torch.empty(T, device=x.device, dtype=torch.float32)
    mean = torch.empty(T, device=x.device, dtype=torch.float32)
    rstd = torch.empty(T, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_head_fwd(_ptr(x), C.c_int64(_row_major_2d(x, "x")), _ptr(w), _ptr(b), _ptr(wo), _ptr(bo), _ptr(v),
                                       _ptr(mean), _ptr(rstd), T, Cc, _stream()), "hvc_head_fwd")
    return v, mean, rstd


def upsample3d_fwd(v, B, grid, size):
    _need_cuda(v)
    assert v.dtype == torch.float32 and v.is_contiguous()
    out = torch.empty(B, 1, *size, device=v.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_upsample3d_fwd(_ptr(v), _ptr(out), B, *grid, *size, _stream()), "hvc_upsample3d_fwd")
    return out


def upsample3d_bwd(dout, B, grid, size):
    _need_cuda(dout)
    assert dout.dtype == torch.float32 and dout.is_contiguous()
    dv = torch.empty(B * grid[0] * grid[1] * grid[2], device=dout.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_upsample3d_bwd(_ptr(dout), _ptr(dv), B, *grid, *size, _stream()), "hvc_upsample3d_bwd")
    return dv


# ------------------------------------------------------------------ fp32 verification mode (hvc_fp32.cu)

def split3(x, pattern, concat_rows=False):
    """x f32 2-D (unit inner stride) -> the six-term bf16 operand of hvc_fp32.cu: [R, 6C] or, with concat_rows, [6R, C]."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and pattern in (0, 1, 2, 3)
    ldx = _row_major_2d(x, "x")
    R, Cc = x.shape
    nb = 6 if pattern < 2 else 3
    out = torch.empty((nb * R, Cc) if concat_rows else (R, nb * Cc), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().hvc_split3(_ptr(x), C.c_int64(ldx), R, Cc, _ptr(out), C.c_int64(out.stride(0)), pattern,
                                     int(concat_rows), _stream()), "hvc_split3")
    return out


def softmax_rows_(s, lse2=None):
    """In place row softmax of f32 s [R, M] holding log2-domain scaled scores."""
    _need_cuda(s)
    assert s.dtype == torch.float32 and s.dim() == 2 and s.stride(1) == 1
    _lib.check(_lib.lib().hvc_softmax_rows(_ptr(s), C.c_int64(s.stride(0)), s.shape[0], s.shape[1], _ptr(lse2), _stream()),
               "hvc_softmax_rows")
    return s


def im2col3d_f32(x, B, Cin, D, H, W, stride, strides):
    """im2col3d with an f32 patch matrix (x must be f32)."""
    _need_cuda(x)
    assert x.dtype == torch.float32
    Do, Ho, Wo = conv_out(D, stride), conv_out(H, stride), conv_out(W, stride)
    Kp = (Cin * 27 + 7) // 8 * 8
    cols = torch.empty(B * Do * Ho * Wo, Kp, device=x.device, dtype=torch.float32)
    g = _geom(B, Cin, D, H, W, stride, strides)
    _lib.check(_lib.lib().hvc_im2col3d_f32(_ptr(x), C.byref(g), _ptr(cols), _stream()), "hvc_im2col3d_f32")
    return cols


def epilogue_f32(acc, bias=None, activation=ACT_NONE, resid=None, gate=None, gate_ld=0, rows_per_batch=0, out=None):
    """out = resid + gate * act(acc + bias) on f32 [T, N] (in place on acc unless `out` is given)."""
    _need_cuda(acc)
    assert acc.dtype == torch.float32
    T, N = acc.shape
    out = acc if out is None else out
    _lib.check(_lib.lib().hvc_epilogue_f32(_ptr(acc), C.c_int64(_row_major_2d(acc, "acc")), T, N, _ptr(bias), activation,
                                           _ptr(resid), C.c_int64(_row_major_2d(resid, "resid") if resid is not None else 0),
                                           _ptr(gate), C.c_int64(gate_ld), rows_per_batch, _ptr(out),
                                           C.c_int64(_row_major_2d(out, "out")), _stream()), "hvc_epilogue_f32")
    return out


# ------------------------------------------------------------------ X-ray encoder (hvc_encoder.cu)

def _geom2d(N, Cin, H, W, k, stride, pad, strides):
    g = _lib.Conv2dGeom()
    g.N, g.Cin, g.H, g.W, g.k, g.stride, g.pad = N, Cin, H, W, k, stride, pad
    g.sn, g.sc, g.sh, g.sw = strides
    return g


def conv2d_out(n, k, stride, pad):
    return (n + 2 * pad - k) // stride + 1


def im2col2d(x, N, Cin, H, W, k, stride, pad, strides, out_dtype=torch.bfloat16):
    """x: f32|bf16 tensor holding the conv input with element strides (sn, sc, sh, sw) -> bf16|f32 [N*Ho*Wo, Kp]."""
    _need_cuda(x)
    assert x.dtype in (torch.float32, torch.bfloat16)
    Ho, Wo = conv2d_out(H, k, stride, pad), conv2d_out(W, k, stride, pad)
    Kp = (Cin * k * k + 7) // 8 * 8
    cols = torch.empty(N * Ho * Wo, Kp, device=x.device, dtype=out_dtype)
    g = _geom2d(N, Cin, H, W, k, stride, pad, strides)
    _lib.check(_lib.lib().hvc_im2col2d(_ptr(x), int(x.dtype == torch.bfloat16), C.byref(g), _ptr(cols), int(out_dtype == torch.float32),
                                       _stream()), "hvc_im2col2d")
    return cols


def im2col2d_split(x, N, Cin, H, W, k, stride, pad, strides):
    """f32 conv input -> bf16 [N*Ho*Wo, 3*Kp] = [c0 | c1 | c0] (two-term split fused into the gather)."""
    _need_cuda(x)
    assert x.dtype == torch.float32
    Ho, Wo = conv2d_out(H, k, stride, pad), conv2d_out(W, k, stride, pad)
    Kp = (Cin * k * k + 7) // 8 * 8
    out = torch.empty(N * Ho * Wo, 3 * Kp, device=x.device, dtype=torch.bfloat16)
    g = _geom2d(N, Cin, H, W, k, stride, pad, strides)
    _lib.check(_lib.lib().hvc_im2col2d_split(_ptr(x), C.byref(g), _ptr(out), _stream()), "hvc_im2col2d_split")
    return out


def col2im2d(dcols, N, Cin, H, W, k, stride, pad, out, strides):
    _need_cuda(dcols, out)
    assert dcols.dtype == torch.bfloat16 and dcols.is_contiguous() and out.dtype == torch.float32
    g = _geom2d(N, Cin, H, W, k, stride, pad, strides)
    _lib.check(_lib.lib().hvc_col2im2d(_ptr(dcols), C.byref(g), _ptr(out), _stream()), "hvc_col2im2d")
    return out


ACT_SILU, ACT_RELU, ACT_GELU_ERF = 0, 1, 2


def norm_act_fwd(x, w, b, B, V, Cc, groups, activation, out_dtype=torch.bfloat16, mean=None, rstd=None):
    """x f32 [B*V, C] channels-last -> (y, mean [B,G], rstd [B,G]); mean/rstd given = eval-mode BatchNorm statistics."""
    _need_cuda(x, w, b)
    assert x.dtype == torch.float32 and x.is_contiguous()
    given = mean is not None
    y = torch.empty(B * V, Cc, device=x.device, dtype=out_dtype)
    if not given:
        mean = torch.empty(B, groups, device=x.device, dtype=torch.float32)
        rstd = torch.empty(B, groups, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cc, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_norm_act_fwd(_ptr(x), _ptr(w), _ptr(b), B, V, Cc, groups, activation, int(given), _ptr(y),
                                           int(out_dtype == torch.bfloat16), _ptr(mean), _ptr(rstd), _ptr(scratch), _stream()),
               "hvc_norm_act_fwd")
    return y, mean, rstd


def norm_act_bwd(dy, x, w, b, mean, rstd, B, V, Cc, groups, activation, stats_frozen=False):
    _need_cuda(dy, x)
    assert dy.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous()
    dx = torch.empty(B * V, Cc, device=x.device, dtype=torch.float32)
    dw = torch.empty(Cc, device=x.device, dtype=torch.float32)
    db = torch.empty(Cc, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2 * B * Cc + 2 * B * groups, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_norm_act_bwd(_ptr(dy), _ptr(x), _ptr(w), _ptr(b), _ptr(mean), _ptr(rstd), B, V, Cc, groups, activation,
                                           int(stats_frozen), _ptr(dx), _ptr(dw), _ptr(db), _ptr(scratch), _stream()), "hvc_norm_act_bwd")
    return dx, dw, db


def maxpool2d_fwd(x, N, H, W, Cc, k, stride, pad):
    """x f32 [N*H*W, C] channels-last -> (y f32 [N*Ho*Wo, C], arg u8)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    Ho, Wo = conv2d_out(H, k, stride, pad), conv2d_out(W, k, stride, pad)
    y = torch.empty(N * Ho * Wo, Cc, device=x.device, dtype=torch.float32)
    arg = torch.empty(N * Ho * Wo, Cc, device=x.device, dtype=torch.uint8)
    _lib.check(_lib.lib().hvc_maxpool2d_fwd(_ptr(x), _ptr(y), _ptr(arg), N, H, W, Cc, k, stride, pad, _stream()), "hvc_maxpool2d_fwd")
    return y, arg


def maxpool2d_bwd(dy, arg, N, H, W, Cc, k, stride, pad):
    _need_cuda(dy, arg)
    assert dy.dtype == torch.float32 and dy.is_contiguous()
    dx = torch.empty(N * H * W, Cc, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_maxpool2d_bwd(_ptr(dy), _ptr(arg), _ptr(dx), N, H, W, Cc, k, stride, pad, _stream()), "hvc_maxpool2d_bwd")
    return dx


def view_mean_fwd(x, B, V, P, Cc, want_pooled=True):
    """x f32 [B*V*P, C] -> (feat f32 [B*P, C], pooled f32 [B, C] | None)."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    feat = torch.empty(B * P, Cc, device=x.device, dtype=torch.float32)
    pooled = torch.empty(B, Cc, device=x.device, dtype=torch.float32) if want_pooled else None
    _lib.check(_lib.lib().hvc_view_mean_fwd(_ptr(x), _ptr(feat), _ptr(pooled), B, V, P, Cc, _stream()), "hvc_view_mean_fwd")
    return feat, pooled


def view_mean_bwd(dfeat, dpooled, B, V, P, Cc):
    t = dfeat if dfeat is not None else dpooled
    _need_cuda(t)
    dx = torch.empty(B * V * P, Cc, device=t.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_view_mean_bwd(_ptr(dfeat), _ptr(dpooled), _ptr(dx), B, V, P, Cc, _stream()), "hvc_view_mean_bwd")
    return dx


def silu(x, dy=None):
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and (dy is None or (dy.dtype == torch.float32 and dy.is_contiguous()))
    out = torch.empty_like(x)
    _lib.check(_lib.lib().hvc_silu(_ptr(x), _ptr(dy), _ptr(out), C.c_int64(x.numel()), _stream()), "hvc_silu")
    return out


# ------------------------------------------------------------------ direct-regression loss (hvc_loss.cu)

def ssim_l1_fwd(pred, target, window=11):
    """pred/target f32 (B, 1, D, H, W) contiguous -> (sums f64 [2] = {sum SSIM, sum |pred - target|}, filtered f32 [5n])."""
    _need_cuda(pred, target)
    assert pred.dtype == target.dtype == torch.float32 and pred.is_contiguous() and target.is_contiguous() and pred.shape == target.shape
    B, D, H, W = pred.shape[0] * pred.shape[1], pred.shape[2], pred.shape[3], pred.shape[4]
    n = pred.numel()
    filtered = torch.empty(5 * n, device=pred.device, dtype=torch.float32)
    scratch = torch.empty(10 * n, device=pred.device, dtype=torch.float32)
    sums = torch.empty(2, device=pred.device, dtype=torch.float64)
    _lib.check(_lib.lib().hvc_ssim_l1_fwd(_ptr(pred), _ptr(target), B, D, H, W, window, _ptr(filtered), _ptr(scratch), _ptr(sums), _stream()),
               "hvc_ssim_l1_fwd")
    return sums, filtered


def ssim_l1_bwd(pred, target, filtered, c_ssim, c_l1, window=11, upstream=None):
    _need_cuda(pred, target, filtered)
    assert upstream is None or (upstream.is_cuda and upstream.dtype == torch.float32 and upstream.numel() == 1)
    B, D, H, W = pred.shape[0] * pred.shape[1], pred.shape[2], pred.shape[3], pred.shape[4]
    n = pred.numel()
    scratch = torch.empty(9 * n, device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred)
    _lib.check(_lib.lib().hvc_ssim_l1_bwd(_ptr(pred), _ptr(target), _ptr(filtered), B, D, H, W, window, C.c_float(c_ssim), C.c_float(c_l1),
                                          _ptr(upstream), _ptr(scratch), _ptr(dpred), _stream()), "hvc_ssim_l1_bwd")
    return dpred


# ------------------------------------------------------------------ cascade stage wrappers

def interp3d_fwd(v, B, grid, size, align_corners):
    """v f32 [B, Di,Hi,Wi] contiguous -> f32 [B, Do,Ho,Wo] trilinear (either corner convention)."""
    _need_cuda(v)
    assert v.dtype == torch.float32 and v.is_contiguous()
    out = torch.empty(B, *size, device=v.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_interp3d_fwd(_ptr(v), _ptr(out), B, *grid, *size, int(align_corners), _stream()), "hvc_interp3d_fwd")
    return out


def interp3d_bwd(dout, B, grid, size, align_corners):
    _need_cuda(dout)
    assert dout.dtype == torch.float32 and dout.is_contiguous()
    dv = torch.empty(B, *grid, device=dout.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_interp3d_bwd(_ptr(dout), _ptr(dv), B, *grid, *size, int(align_corners), _stream()), "hvc_interp3d_bwd")
    return dv


def conv_tap_offsets(H, W, sign=1):
    """Row shifts of the 27 taps on a stride-1 padded volume (B, D+2, H+2, W+2, C): (kd-1)*(H+2)*(W+2) + (kh-1)*(W+2) + (kw-1)."""
    sd, sh = (H + 2) * (W + 2), W + 2
    return [sign * ((kd - 1) * sd + (kh - 1) * sh + (kw - 1)) for kd in range(3) for kh in range(3) for kw in range(3)]


def conv_tap_offsets_s2(rows_per_volume, Hp, Wp):
    """Stride-2 conv on the eight stacked parity volumes (s2d_pad_cl): per tap (parity volume index, shift inside it)."""
    out = []
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                par = (kd != 1) * 4 + (kh != 1) * 2 + (kw != 1)
                out.append((par, -(kd == 0) * Hp * Wp - (kh == 0) * Wp - (kw == 0)))
    return out


def chan_dot_fwd(y, w, bias):
    """Conv3d(C -> 1, kernel 1) on channels-last y f32 [M, C]: out[m] = bias + sum_c y[m,c] w[c]."""
    _need_cuda(y, w)
    assert y.dtype == torch.float32 and y.is_contiguous() and w.dtype == torch.float32 and w.is_contiguous() and w.numel() == y.shape[1]
    M, Cc = y.shape
    out = torch.empty(M, device=y.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_chan_dot_fwd(_ptr(y), _ptr(w), _ptr(bias), _ptr(out), C.c_int64(M), Cc, _stream()), "hvc_chan_dot_fwd")
    return out


def chan_dot_bwd(dout, y, w):
    """-> (dy f32 [M, C], dw f32 [C], db f32 [1])"""
    _need_cuda(dout, y, w)
    assert dout.dtype == torch.float32 and dout.is_contiguous() and y.is_contiguous() and dout.numel() == y.shape[0]
    M, Cc = y.shape
    dy = torch.empty_like(y)
    dwb = torch.zeros(Cc + 1, device=y.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_chan_dot_bwd(_ptr(dout), _ptr(y), _ptr(w), _ptr(dy), _ptr(dwb), C.c_void_p(dwb.data_ptr() + 4 * Cc),
                                           C.c_int64(M), Cc, _stream()), "hvc_chan_dot_bwd")
    return dy, dwb[:Cc], dwb[Cc:]


def pad3d_cl(src, B, D, H, W, Cs, Cp, pad_hi=1):
    """src (B, D, H, W, Cs) f32|bf16 dense channels-last -> zero-padded bf16 (B, D+1+pad_hi, H+1+pad_hi, W+1+pad_hi, Cp)."""
    _need_cuda(src)
    assert src.is_contiguous() and src.numel() == B * D * H * W * Cs and src.dtype in (torch.float32, torch.bfloat16)
    dst = torch.empty(B, D + 1 + pad_hi, H + 1 + pad_hi, W + 1 + pad_hi, Cp, device=src.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().hvc_pad3d_cl(_ptr(src), int(src.dtype == torch.bfloat16), _ptr(dst), B, D, H, W, Cs, Cp, pad_hi, _stream()),
               "hvc_pad3d_cl")
    return dst


def unpad3d_cl(src, B, D, H, W, Cc, out=None, pad_hi=1):
    """src f32 (B, D+1+pad_hi, H+1+pad_hi, W+1+pad_hi, C) contiguous -> f32 (B, D, H, W, C) dense."""
    _need_cuda(src)
    assert src.is_contiguous() and src.dtype == torch.float32
    assert src.numel() == B * (D + 1 + pad_hi) * (H + 1 + pad_hi) * (W + 1 + pad_hi) * Cc
    if out is None:
        out = torch.empty(B, D, H, W, Cc, device=src.device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == B * D * H * W * Cc
    _lib.check(_lib.lib().hvc_unpad3d_cl(_ptr(src), _ptr(out), B, D, H, W, Cc, pad_hi, _stream()), "hvc_unpad3d_cl")
    return out


def s2d_pad_cl(src, B, D, H, W, Cc):
    """src (B, D, H, W, C) f32|bf16 dense (even sizes) -> bf16 (8, B, D/2+1, H/2+1, W/2+1, C): the eight parity volumes, low-side padded."""
    _need_cuda(src)
    assert src.is_contiguous() and src.numel() == B * D * H * W * Cc and src.dtype in (torch.float32, torch.bfloat16)
    dst = torch.empty(8, B, D // 2 + 1, H // 2 + 1, W // 2 + 1, Cc, device=src.device, dtype=torch.bfloat16)
    _lib.check(_lib.lib().hvc_s2d_pad_cl(_ptr(src), int(src.dtype == torch.bfloat16), _ptr(dst), B, D, H, W, Cc, _stream()), "hvc_s2d_pad_cl")
    return dst


def d2s_unpad_cl(src, B, D, H, W, Cc):
    """src f32 (8, B, D/2+1, H/2+1, W/2+1, C) contiguous -> f32 (B, D, H, W, C) dense."""
    _need_cuda(src)
    assert src.is_contiguous() and src.dtype == torch.float32 and src.numel() == 8 * B * (D // 2 + 1) * (H // 2 + 1) * (W // 2 + 1) * Cc
    out = torch.empty(B, D, H, W, Cc, device=src.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_d2s_unpad_cl(_ptr(src), _ptr(out), B, D, H, W, Cc, _stream()), "hvc_d2s_unpad_cl")
    return out


# ------------------------------------------------------------------ stage 2-3 loss terms (hvc_loss_multiscale.cu)

def _vol_dims(x):
    """(B, D, H, W) of a contiguous f32 volume tensor (B, 1, D, H, W) / (B, D, H, W): leading dims fold into the batch."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() >= 3
    D, H, W = x.shape[-3:]
    return x.numel() // (D * H * W), D, H, W


def tv_sums(x, eps):
    """f64[3] = sum sqrt(diff^2 + eps) along D, H, W (TotalVariationLoss, loss_multiscale.py:162-170)."""
    _need_cuda(x)
    sums = torch.empty(3, device=x.device, dtype=torch.float64)
    _lib.check(_lib.lib().hvc_tv_fwd(_ptr(x), *_vol_dims(x), C.c_float(eps), _ptr(sums), _stream()), "hvc_tv_fwd")
    return sums


def tv_finalize(sums_pred, sums_target, dims):
    """-> (loss f32[1], coef f32[1] = d loss / d tv_pred)."""
    _need_cuda(sums_pred)
    loss = torch.empty(1, device=sums_pred.device, dtype=torch.float32)
    coef = torch.empty(1, device=sums_pred.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_tv_finalize(_ptr(sums_pred), _ptr(sums_target), *dims, _ptr(loss), _ptr(coef), _stream()), "hvc_tv_finalize")
    return loss, coef


def tv_bwd(x, eps, coef, upstream=None):
    _need_cuda(x, coef)
    dx = torch.empty_like(x)
    _lib.check(_lib.lib().hvc_tv_bwd(_ptr(x), *_vol_dims(x), C.c_float(eps), _ptr(coef), _ptr(upstream), _ptr(dx), _stream()), "hvc_tv_bwd")
    return dx


def freq_l1_sums(spec_pred, spec_target):
    """spectra complex64 [..., D, H, W] contiguous -> f64[2] = (low, high) sums of | |Fp| - |Ft| | (FrequencyLoss, :206-234)."""
    _need_cuda(spec_pred, spec_target)
    assert spec_pred.dtype == torch.complex64 and spec_target.dtype == torch.complex64 and spec_pred.is_contiguous() and spec_target.is_contiguous()
    D, H, W = spec_pred.shape[-3:]
    B = spec_pred.numel() // (D * H * W)
    sums = torch.empty(2, device=spec_pred.device, dtype=torch.float64)
    _lib.check(_lib.lib().hvc_freq_l1_fwd(_ptr(spec_pred), _ptr(spec_target), B, D, H, W, _ptr(sums), _stream()), "hvc_freq_l1_fwd")
    return sums


def freq_l1_bwd(spec_pred, spec_target, c_low, c_high, upstream=None):
    _need_cuda(spec_pred, spec_target)
    D, H, W = spec_pred.shape[-3:]
    B = spec_pred.numel() // (D * H * W)
    g = torch.empty_like(spec_pred)
    _lib.check(_lib.lib().hvc_freq_l1_bwd(_ptr(spec_pred), _ptr(spec_target), B, D, H, W, C.c_float(c_low), C.c_float(c_high), _ptr(upstream),
                                          _ptr(g), _stream()), "hvc_freq_l1_bwd")
    return g


def proj_mean_fwd(vol):
    """vol f32 (B, 1, D, H, W) -> ap (B, H, W) = mean over D, lat (B, D, H) = mean over W (DRRReprojectionLoss.generate_drr, :249-265)."""
    _need_cuda(vol)
    B, D, H, W = _vol_dims(vol)
    ap = torch.empty(B, H, W, device=vol.device, dtype=torch.float32)
    lat = torch.empty(B, D, H, device=vol.device, dtype=torch.float32)
    _lib.check(_lib.lib().hvc_proj_mean_fwd(_ptr(vol), B, D, H, W, _ptr(ap), _ptr(lat), _stream()), "hvc_proj_mean_fwd")
    return ap, lat


def proj_mean_bwd(dap, dlat, shape):
    _need_cuda(dap, dlat)
    dvol = torch.empty(shape, device=dap.device, dtype=torch.float32)
    B, D, H, W = _vol_dims(dvol)
    _lib.check(_lib.lib().hvc_proj_mean_bwd(_ptr(dap), _ptr(dlat), B, D, H, W, _ptr(dvol), _stream()), "hvc_proj_mean_bwd")
    return dvol


def l1_sum(a, b):
    _need_cuda(a, b)
    assert a.dtype == b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous() and a.numel() == b.numel()
    s = torch.empty(1, device=a.device, dtype=torch.float64)
    _lib.check(_lib.lib().hvc_l1_fwd(_ptr(a), _ptr(b), C.c_int64(a.numel()), _ptr(s), _stream()), "hvc_l1_fwd")
    return s


def l1_bwd(a, b, c, upstream=None):
    _need_cuda(a, b)
    da = torch.empty_like(a)
    _lib.check(_lib.lib().hvc_l1_bwd(_ptr(a), _ptr(b), C.c_int64(a.numel()), C.c_float(c), _ptr(upstream), _ptr(da), _stream()), "hvc_l1_bwd")
    return da
