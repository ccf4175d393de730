"""Thin tensor-level wrappers over the C ABI (no autograd here; see ops.py).

torch is plumbing only: it owns the device memory and the stream; every computation below is a
kernel of libhvc_sm100a.so.  All wrappers require CUDA tensors and raise otherwise.
"""
import ctypes as C

import torch

from . import _lib

EPI_BF16, EPI_RESIDUAL, EPI_F32_ATOMIC, EPI_F32 = 0, 1, 2, 3
ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.HvcError("hybrid_vit_cascade_b200 kernels need CUDA tensors (there is no CPU fallback)")
    dev = next(t for t in ts if t is not None).device
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())


def _row_major_2d(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} "
                         f"stride {t.stride()}")
    return t.stride(0)


def gemm(a, b, *, a_major=0, b_major=0, out=None, out_dtype=torch.bfloat16, epilogue=EPI_BF16,
         activation=ACT_NONE, bias=None, out2=None, resid=None, gate=None, gate_ld=0, rows_per_batch=0,
         aux=None, alpha=1.0, k_splits=1):
    """D[M,N] = alpha * sum_k A(m,k) B(n,k) with a fused epilogue (see include/hvc.h).

    a: bf16, stored [M,K] (a_major=0) or [K,M] (a_major=1); b: bf16, stored [N,K] or [K,N].
    """
    _need_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    lda, ldb = _row_major_2d(a, "a"), _row_major_2d(b, "b")
    M, K = (a.shape if a_major == 0 else (a.shape[1], a.shape[0]))
    N, Kb = (b.shape if b_major == 0 else (b.shape[1], b.shape[0]))
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


def attn_fwd(q, k, v, B, H, nq, nk, d, scale):
    """q: bf16 view [B*nq, H*d] (may be a column slice of a packed projection output), k/v: [B*nk, H*d].

    Returns (o bf16 [B*nq, H*d], lse2 f32 [B, H, pad128(nq)]).
    """
    _need_cuda(q, k, v)
    assert q.dtype == k.dtype == v.dtype == torch.bfloat16
    assert q.shape == (B * nq, H * d) and k.shape == (B * nk, H * d) and v.shape == (B * nk, H * d)
    o = torch.empty(B * nq, H * d, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, _pad128(nq), device=q.device, dtype=torch.float32)
    args = _attn_args(q, k, v, B, H, nq, nk, d, scale)
    args.o, args.ldo = o.data_ptr(), o.stride(0)
    args.lse = lse.data_ptr()
    _lib.check(_lib.lib().hvc_attn_fwd(C.byref(args), _stream()), "hvc_attn_fwd")
    return o, lse
