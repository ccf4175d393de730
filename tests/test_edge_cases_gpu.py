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


@pytest.mark.parametrize("N", [8, 32, 40, 64, 72])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_narrow_n_all_operand_layouts(N, a_major, b_major):
    """N <= 64 problems issue a narrower tcgen05.mma (32 or 64 columns instead of 128): every operand-layout combination, ragged N,
    several M tiles and k-blocks."""
    from hybrid_vit_cascade_b200 import kernels as K
    M, Kd = 384 + 8, 200
    g = torch.Generator(device="cuda").manual_seed(N * 7 + a_major * 2 + b_major)
    a = torch.randn(M, Kd, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, Kd, device="cuda", generator=g).bfloat16()
    ref = a.float() @ b.float().t()
    out = K.gemm(a.t().contiguous() if a_major else a, b.t().contiguous() if b_major else b, a_major=a_major, b_major=b_major,
                 epilogue=K.EPI_F32)
    assert out.shape == (M, N) and O.max_rel(out, ref) < 1e-5


@pytest.mark.parametrize("Cin,Cout", [(64, 32), (64, 64), (128, 96)])
def test_gemm_implicit_conv_taps_match_conv3d(Cin, Cout):
    """hvc_conv_taps: the GEMM reads a zero-padded channels-last volume with per-tap row shifts instead of a patch matrix.  Side 1 =
    forward conv (and data gradient with negated strides + transposed filter), side 2 = weight gradient; against fp64 conv3d on the
    bf16-rounded operands."""
    import torch.nn.functional as F
    from hybrid_vit_cascade_b200 import kernels as K
    B, D, H, W = 2, 5, 6, 7
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, D, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) * (27 * Cin) ** -0.5).bfloat16()
    dz = torch.randn(B, Cout, D, H, W, device="cuda", generator=g).bfloat16()
    xd, wd, dzd = x.double().requires_grad_(True), w.double().requires_grad_(True), dz.double()
    ref = F.conv3d(xd, wd, padding=1)
    (ref * dzd).sum().backward()
    xp = K.pad3d_cl(x.permute(0, 2, 3, 4, 1).contiguous(), B, D, H, W, Cin, Cin)
    w_taps = w.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous()
    zp = K.gemm(xp.view(-1, Cin), w_taps, epilogue=K.EPI_F32, taps=(1, Cin, K.conv_tap_offsets(H, W)))
    z = K.unpad3d_cl(zp, B, D, H, W, Cout).permute(0, 4, 1, 2, 3)
    assert O.max_rel(z, ref) < 1e-5
    Cp = (Cout + 63) // 64 * 64
    dzp = K.pad3d_cl(dz.permute(0, 2, 3, 4, 1).contiguous(), B, D, H, W, Cout, Cp).view(-1, Cp)
    dw = K.gemm(dzp, xp.view(-1, Cin), a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=3, taps=(2, Cin, K.conv_tap_offsets(H, W)))[:Cout]
    assert O.max_rel(dw.view(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3), wd.grad) < 1e-5
    wt = torch.zeros(Cin, 27, Cp, device="cuda", dtype=torch.bfloat16)
    wt[:, :, :Cout] = w.reshape(Cout, Cin, 27).permute(1, 2, 0)
    dxp = K.gemm(dzp, wt.view(Cin, 27 * Cp), epilogue=K.EPI_F32, taps=(1, Cp, K.conv_tap_offsets(H, W, -1)))
    assert O.max_rel(K.unpad3d_cl(dxp, B, D, H, W, Cin).permute(0, 4, 1, 2, 3), xd.grad) < 1e-5


@pytest.mark.parametrize("Cin,Cout", [(64, 128), (128, 72)])
def test_gemm_implicit_conv_stride2_matches_conv3d(Cin, Cout):
    """Stride-2 Conv3d(k3, p1) as an implicit GEMM: the input split into eight parity volumes (hvc_s2d_pad_cl), every tap a row shift
    into one of them; forward, weight gradient (taps on B) and data gradient (one GEMM per parity volume) against fp64 conv3d."""
    import torch.nn.functional as F
    from hybrid_vit_cascade_b200 import kernels as K
    B, D, H, W = 2, 6, 4, 8
    g = torch.Generator(device="cuda").manual_seed(3 * Cin + Cout)
    x = torch.randn(B, Cin, D, H, W, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, 3, device="cuda", generator=g) * (27 * Cin) ** -0.5).bfloat16()
    Do, Ho, Wo = D // 2, H // 2, W // 2
    dz = torch.randn(B, Cout, Do, Ho, Wo, device="cuda", generator=g).bfloat16()
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, stride=2, padding=1)
    (ref * dz.double()).sum().backward()
    xs = K.s2d_pad_cl(x.permute(0, 2, 3, 4, 1).contiguous(), B, D, H, W, Cin)
    rows_p = B * (Do + 1) * (Ho + 1) * (Wo + 1)
    tt = K.conv_tap_offsets_s2(rows_p, Ho + 1, Wo + 1)
    offs = [par * rows_p + sh for par, sh in tt]
    w_taps = w.permute(0, 2, 3, 4, 1).reshape(Cout, 27 * Cin).contiguous()
    zp = K.gemm(xs.view(-1, Cin), w_taps, epilogue=K.EPI_F32, taps=(1, Cin, offs), m_rows=rows_p)
    z = K.unpad3d_cl(zp, B, Do, Ho, Wo, Cout, pad_hi=0).permute(0, 4, 1, 2, 3)
    assert O.max_rel(z, ref) < 1e-5
    Cp = (Cout + 63) // 64 * 64
    dzp = K.pad3d_cl(dz.permute(0, 2, 3, 4, 1).contiguous(), B, Do, Ho, Wo, Cout, Cp, pad_hi=0).view(-1, Cp)
    dw = K.gemm(dzp, xs.view(-1, Cin), a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=2, taps=(2, Cin, offs))[:Cout]
    assert O.max_rel(dw.view(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3), wd.grad) < 1e-5
    wt = torch.zeros(Cin, 27, Cp, device="cuda", dtype=torch.bfloat16)
    wt[:, :, :Cout] = w.reshape(Cout, Cin, 27).permute(1, 2, 0)
    dxs = torch.empty(8 * rows_p, Cin, device="cuda", dtype=torch.float32)
    for par in range(8):
        ts = [t for t, (pp, _) in enumerate(tt) if pp == par]
        K.gemm(dzp, wt[:, ts].reshape(Cin, len(ts) * Cp), epilogue=K.EPI_F32, taps=(1, Cp, [-tt[t][1] for t in ts]),
               out=dxs[par * rows_p:(par + 1) * rows_p])
    dx = K.d2s_unpad_cl(dxs, B, D, H, W, Cin).permute(0, 4, 1, 2, 3)
    assert O.max_rel(dx, xd.grad) < 1e-5


def test_voxel_embed_implicit_gemm_matches_patch_matrix_path():
    """HybridViT3D's conv stack at the direct-regression widths (1 -> 64 s2 -> 128 s2 -> 256 s1): the layers with Cin % 64 == 0 run as
    implicit GEMMs (stride 2 through the parity split); with ops.IMPLICIT_EMBED = False they take the patch-matrix path.  Same
    tokens, same gradients."""
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200 import ops
    torch.manual_seed(5)
    m = hvc.HybridViT3D(volume_size=(32, 32, 32), in_channels=1, voxel_dim=256, depth=1, num_heads=4, context_dim=64, token_grid=8).cuda()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaln.linear" in n:
                p.normal_(0.0, 0.02)
    assert [(c[0], c[1], c[2]) for c in m._plan] == [(1, 64, 2), (64, 128, 2), (128, 256, 1)]
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randn(2, 1, 32, 32, 32, device="cuda", generator=g)
    ctx = torch.randn(2, 16, 64, device="cuda", generator=g)
    cond = torch.randn(2, 1024, device="cuda", generator=g)
    r = torch.randn(2, 1, 32, 32, 32, device="cuda", generator=g)
    hvc.set_dropout_policy("ignore")
    res = []
    min_bytes = ops.IMPLICIT_EMBED_MIN_BYTES
    ops.IMPLICIT_EMBED_MIN_BYTES = 0          # (the size rule would keep this small volume on the patch path)
    try:
        for flag in (True, False):
            ops.IMPLICIT_EMBED = flag
            m.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            y = m(xi, ctx, cond)
            (y * r).sum().backward()
            res.append((y.detach(), xi.grad, {n: p.grad.clone() for n, p in m.named_parameters() if n.startswith("voxel_embed")}))
    finally:
        ops.IMPLICIT_EMBED = True
        ops.IMPLICIT_EMBED_MIN_BYTES = min_bytes
        hvc.set_dropout_policy("apply")
    (y1, gx1, gp1), (y0, gx0, gp0) = res
    assert O.max_rel(y1, y0) <= 5e-3 and O.cosine(gx1, gx0) >= 0.9999
    for n in gp0:
        assert O.cosine(gp1[n], gp0[n]) >= 0.9999, (n, O.cosine(gp1[n], gp0[n]))


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


def test_flat_adamw_matches_torch_clip_and_adamw():
    """optim.FlatAdamW (one sum-of-squares + one fused clip+AdamW kernel per gradient bucket) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the same gradients, several steps, clipping active and inactive; the
    modules keep working on the re-pointed parameters (derived bf16 operands are refreshed)."""
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(3)
    shapes = [(257, 33), (64,), (5, 7, 3), (1,), (1024, 130)]
    mine = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    gb = GradientBuckets(mine, bucket_bytes=64 << 10)           # several buckets
    assert len(gb.buckets) >= 2
    opt = hvc.FlatAdamW(gb, lr=1e-2, weight_decay=0.05, max_grad_norm=1.0)
    topt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.05)
    for p, r in zip(mine, ref):
        assert torch.equal(p.data, r.data)                       # re-pointing kept the values
    g = torch.Generator(device="cuda").manual_seed(4)
    for step in range(6):
        scale = 1e-3 if step % 2 else 3.0                        # below / above the clipping threshold
        gb.reset()
        for p, r in zip(mine, ref):
            gr = torch.randn(p.shape, device="cuda", generator=g) * scale
            p.grad.copy_(gr)
            r.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        topt.step()
        opt.step()
        assert abs(float(opt.grad_norm()) - float(total)) <= 1e-5 * float(total)
        for p, r in zip(mine, ref):
            assert O.max_rel(p.data, r.data) <= 2e-6, (step, tuple(p.shape), O.max_rel(p.data, r.data))
    # a module stepping through FlatAdamW: the cached bf16 weights must follow the update
    lin = hvc.MultiHeadSelfAttention(64, num_heads=2, dropout=0.0).cuda()
    x = torch.randn(2, 128, 64, device="cuda", generator=g)
    gb2 = GradientBuckets(list(lin.parameters()))
    opt2 = hvc.FlatAdamW(gb2, lr=0.05, weight_decay=0.0)
    y0 = lin(x).detach().clone()
    gb2.reset()
    lin(x).square().mean().backward()
    opt2.step()
    y1 = lin(x).detach()
    assert O.max_rel(y1, y0) > 1e-2                               # the forward sees the new weights


def test_flat_adamw_skips_unused_parameters_like_torch_adamw():
    """A trainable parameter that takes no part in a step (ProgressiveCascadeModel(xrays, max_stage=1) leaves stages 2-3 unused,
    train_progressive_4gpu.py:238) has grad None under torch: AdamW neither decays it nor touches its moments.  Its bucket slot holds
    zeros here; FlatAdamW must skip it the same way (GradientBuckets.touched)."""
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(5)
    used = torch.nn.Linear(40, 24).cuda()
    unused = torch.nn.Linear(24, 8).cuda()
    ref_used, ref_unused = torch.nn.Linear(40, 24).cuda(), torch.nn.Linear(24, 8).cuda()
    ref_used.load_state_dict(used.state_dict())
    ref_unused.load_state_dict(unused.state_dict())
    mine = list(used.parameters()) + list(unused.parameters())
    ref = list(ref_used.parameters()) + list(ref_unused.parameters())
    gb = GradientBuckets(mine)                                     # ONE bucket holds used and unused members
    assert len(gb.buckets) == 1
    opt = hvc.FlatAdamW(gb, lr=1e-2, weight_decay=0.1)
    topt = torch.optim.AdamW(ref, lr=1e-2, weight_decay=0.1)
    x = torch.randn(16, 40, device="cuda")
    w_unused0 = unused.weight.detach().clone()
    for step in range(3):
        gb.reset()
        topt.zero_grad(set_to_none=True)
        used(x).square().mean().backward()
        ref_used(x).square().mean().backward()
        gb.finish()
        assert gb.touched(0) == {2, 3} and not gb.all_touched()   # reverse registration order: the unused pair comes first
        opt.step()
        topt.step()
    for p, r in zip(mine, ref):
        assert O.max_rel(p.data, r.data) <= 1e-5              # three lr = 1e-2 steps on |w| ~ 0.1 weights (fast-math rsqrt in the kernel)
    assert torch.equal(unused.weight.detach(), w_unused0)         # no weight decay on a parameter that got no gradient
