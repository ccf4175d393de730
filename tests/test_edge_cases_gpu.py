"""Edge cases of the C ABI on the GPU: smallest and ragged shapes, and the error channel (no exception or crash crosses the
boundary: bad arguments come back as HvcError with the library's message)."""
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def _attn_ref(q, k, v, H, d):
    B, N, M = q.shape[0], q.shape[1], k.shape[1]
    qh = q.float().view(B, N, H, d).transpose(1, 2)
    kh = k.float().view(B, M, H, d).transpose(1, 2)
    vh = v.float().view(B, M, H, d).transpose(1, 2)
    o = ((qh @ kh.transpose(-1, -2)) * d ** -0.5).softmax(-1) @ vh
    return o.transpose(1, 2).reshape(B, N, H * d)


@pytest.mark.parametrize("B,H,N,M,d", [(1, 1, 1, 1, 64), (1, 1, 1, 1, 32), (2, 3, 1, 130, 32), (1, 2, 129, 1, 64), (3, 1, 127, 257, 64),
                                       (1, 4, 256, 128, 32), (2, 2, 255, 8, 64)])
def test_attention_smallest_and_ragged_shapes(B, H, N, M, d):
    from hybrid_vit_cascade_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N + M)
    C = H * d
    q = torch.randn(B, N, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    k = torch.randn(B, M, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    v = torch.randn(B, M, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    r = torch.randn(B, N, C, device="cuda", generator=g).bfloat16()
    ref = _attn_ref(q, k, v, H, d)
    (ref * r.float()).sum().backward()
    q2, k2, v2 = (t.detach().reshape(-1, C) for t in (q, k, v))
    o, lse = K.attn_fwd(q2, k2, v2, B, H, N, M, d, d ** -0.5)
    assert O.max_rel(o.view(B, N, C), ref) <= 2e-2
    dq, dk, dv = (torch.empty_like(t) for t in (q2, k2, v2))
    K.attn_bwd(q2, k2, v2, o, lse, r.reshape(-1, C), B, H, N, M, d, d ** -0.5, dq, dk, dv)
    for a, b, name in ((dq, q.grad, "dq"), (dk, k.grad, "dk"), (dv, v.grad, "dv")):
        b = b.reshape(-1, C).float()
        if float(b.abs().max()) < 1e-6:           # single key: softmax is constant, dq = dk = 0
            assert float(a.float().abs().max()) < 1e-2, name
        else:
            assert O.cosine(a, b) >= 0.999, (name, O.cosine(a, b))


@pytest.mark.parametrize("M,N,Kd", [(1, 8, 8), (1, 1, 8), (7, 9, 16), (129, 130, 72), (128, 128, 8)])
def test_gemm_smallest_and_ragged_shapes(M, N, Kd):
    from hybrid_vit_cascade_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(M + N + Kd)
    a = torch.randn(M, Kd, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, Kd, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ b.float().t() + bias
    out = K.gemm(a, b, bias=bias, epilogue=K.EPI_F32)
    assert O.max_rel(out, ref) < 1e-5
    out16 = K.gemm(a, b, bias=bias)
    assert O.max_rel(out16, ref) < 1e-2
    resid = torch.randn(M, N, device="cuda", generator=g)
    outr = K.gemm(a, b, bias=bias, epilogue=K.EPI_RESIDUAL, resid=resid)
    assert O.max_rel(outr, ref + resid) < 1e-5


def test_layernorm_odd_token_counts():
    from hybrid_vit_cascade_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(4)
    for T, C in [(1, 64), (3, 256), (130, 96), (257, 1024)]:
        x = torch.randn(T, C, device="cuda", generator=g)
        w, b = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
        y, mean, rstd = K.ln_fwd(x, w, b, out_dtype=torch.float32)
        ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
        assert O.max_rel(y, ref) < 1e-5


def test_errors_come_back_as_exceptions_not_crashes():
    from hybrid_vit_cascade_b200 import _lib, kernels as K
    x = torch.randn(64, 48, device="cuda").bfloat16()
    with pytest.raises(_lib.HvcError, match="head_dim"):
        K.attn_fwd(x, x, x, 1, 1, 64, 64, 48, 48 ** -0.5)                       # unsupported head_dim
    a = torch.randn(16, 12, device="cuda").bfloat16()                          # K = 12: row pitch 24 bytes, not TMA-addressable
    with pytest.raises(_lib.HvcError, match="16"):
        K.gemm(a, a)
    with pytest.raises(_lib.HvcError, match="CUDA tensors"):
        K.gemm(a.cpu(), a.cpu())                                               # no CPU fallback
    with pytest.raises(ValueError):
        K.gemm(a.t(), a)                                                       # non-unit inner stride is refused by the wrapper
    import hybrid_vit_cascade_b200 as hvc
    with pytest.raises(NotImplementedError, match="head_dim"):
        hvc.MultiHeadSelfAttention(96, num_heads=2).cuda()(torch.randn(1, 8, 96, device="cuda"))
    # the library is still usable after the failures
    out = K.gemm(x, x)
    assert out.shape == (64, 64)
