"""GPU parity at the configuration bench.py publishes: 32^3 = 32768 volume tokens.

The attention kernels at this length run 256 query tiles of fp32 `cp.reduce.async.bulk` accumulation per key tile and address
packed [B*N, 3C] buffers at offsets up to 2e8 elements; the smaller fixtures never reach either.  Three levels:

  * hvc_attn_fwd / hvc_attn_bwd at N = M = 32768, d = 64 (4 heads) and d = 32 (8 heads), against a query-chunked fp32
    restatement of vit_components.py:46-51 and its analytic gradient, evaluated on the GPU (TF32 off).  The restatement is itself
    checked against torch autograd at a small size first.
  * the same at batch 8 (the bench's batch): every (batch, head) slice against the fp32 restatement.
  * HybridViT3D(token_grid="conv") at 128^3 (direct_regression, 4 heads) and at 256^3 with 32 input channels (cascade stage 3,
    8 heads), depth 1, batch 1: forward and EVERY gradient against oracle.vit_oracle.backbone(attn_chunk=2048) in fp32 on the GPU
    (hybrid_vit_backbone.py:233-274), plus a batch-8 run whose last sample is compared with the oracle run on that sample.

Tolerances (north_star): forward max|a-b|/max|b| <= 2e-2, gradient cosine >= 0.999.
"""
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-2
COS_TOL = 0.999
N32K = 32768


class _NoTF32:
    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.old
        return False


def attention_fp32_chunked(q, k, v, do, scale, chunk=2048):
    """softmax(q k^T scale) v and its gradients for ONE (batch, head): q [N,d], k/v [M,d], do [N,d], all fp32.
    vit_components.py:46-51 with the queries taken `chunk` rows at a time (rows of a softmax are independent), and the
    gradient written out: dP = dO V^T, dS = P o (dP - rowsum(P o dP)), dQ = dS K scale, dK = dS^T Q scale, dV = P^T dO."""
    o = torch.empty_like(q)
    dq = torch.empty_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    for s in range(0, q.shape[0], chunk):
        qs, dos = q[s:s + chunk], do[s:s + chunk]
        p = ((qs @ k.t()) * scale).softmax(-1)
        o[s:s + chunk] = p @ v
        dv += p.t() @ dos
        dp = dos @ v.t()
        ds = p * (dp - (p * dp).sum(-1, keepdim=True))
        dq[s:s + chunk] = (ds @ k) * scale
        dk += (ds.t() @ qs) * scale
    return o, dq, dk, dv


def test_chunked_fp32_restatement_matches_autograd():
    """The checker of this file against torch autograd on the literal reference expression (small size)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    N, M, d = 700, 900, 32
    q, k, v, do = (torch.randn(n, d, device="cuda", generator=g) for n in (N, M, M, N))
    with _NoTF32():
        qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
        ref = ((qa @ ka.transpose(-2, -1)) * d ** -0.5).softmax(dim=-1) @ va
        ref.backward(do)
        o, dq, dk, dv = attention_fp32_chunked(q, k, v, do, d ** -0.5, chunk=256)
    for a, b in ((o, ref), (dq, qa.grad), (dk, ka.grad), (dv, va.grad)):
        assert O.max_rel(a, b) < 1e-5


def _attn_case(B, H, d, seed, check_slices):
    from hybrid_vit_cascade_b200 import kernels as K
    g = torch.Generator(device="cuda").manual_seed(seed)
    N, C = N32K, H * d
    # the packed projection output the modules hand to the kernels: [B*N, 3C], feature = which*C + head*d + j (vit_components.py:41)
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    do = torch.randn(B * N, C, device="cuda", generator=g).bfloat16()
    scale = d ** -0.5
    o, lse2 = K.attn_fwd(q, k, v, B, H, N, N, d, scale)
    dqkv = torch.empty_like(qkv)
    K.attn_bwd(q, k, v, o, lse2, do, B, H, N, N, d, scale, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:])
    torch.cuda.synchronize()
    worst_fwd, worst_cos = 0.0, 1.0
    with _NoTF32():
        for b, h in check_slices:
            rows, cols = slice(b * N, (b + 1) * N), slice(h * d, (h + 1) * d)
            qf, kf, vf, dof = (t[rows, cols].float() for t in (q, k, v, do))
            o_ref, dq_ref, dk_ref, dv_ref = attention_fp32_chunked(qf, kf, vf, dof, scale)
            worst_fwd = max(worst_fwd, O.max_rel(o[rows, cols], o_ref))
            lse_ref = torch.cat([torch.logsumexp((qf[s:s + 4096] @ kf.t()) * scale, -1) for s in range(0, N, 4096)])
            assert float((lse2[b, h, :N] - lse_ref * 1.4426950408889634).abs().max()) < 1e-2
            for w, ref in enumerate((dq_ref, dk_ref, dv_ref)):
                got = dqkv[rows, w * C + h * d: w * C + (h + 1) * d]
                cs = O.cosine(got, ref)
                worst_cos = min(worst_cos, cs)
                assert cs >= COS_TOL, f"(b={b}, h={h}) {'qkv'[w]} gradient cosine {cs:.5f}"
                assert O.max_rel(got, ref) <= 5e-2, f"(b={b}, h={h}) d{'qkv'[w]} max-rel"
    assert worst_fwd <= FWD_TOL, worst_fwd
    return worst_fwd, worst_cos


@pytest.mark.parametrize("H,d", [(4, 64), (8, 32)])
def test_attention_forward_backward_32768_tokens(H, d):
    """hvc_attn_fwd + hvc_attn_bwd at the full 32768-token length, every head, batch 1."""
    _attn_case(1, H, d, seed=21 + d, check_slices=[(0, h) for h in range(H)])


def test_attention_forward_backward_32768_tokens_batch8():
    """The bench's launch shape: B = 8, 4 heads, d = 64 -- packed-buffer offsets up to 8*32768*768 = 2.0e8 elements.  Every
    (batch, head) slice of the first, a middle and the last sample against the fp32 restatement."""
    _attn_case(8, 4, 64, seed=33, check_slices=[(b, h) for b in (0, 3, 7) for h in range(4)])


def test_attention_backward_32768_tokens_batch8_head_dim_32():
    """Stage 2/3 shape at batch 8: 8 heads of d = 32; first and last sample."""
    _attn_case(8, 8, 32, seed=35, check_slices=[(b, h) for b in (0, 7) for h in (0, 3, 7)])


def _oracle_run(cfg, sd, x, ctx, cond, r, attn_chunk=2048):
    with _NoTF32():
        sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
        xs = [t.detach().clone().requires_grad_(True) for t in (x, ctx, cond)]
        y = O.backbone(xs[0], xs[1], xs[2], sd, cfg, attn_chunk=attn_chunk)
        (y * r).sum().backward()
        out = y.detach(), {k: v.grad for k, v in sd.items()}, [t.grad for t in xs]
    del y
    torch.cuda.empty_cache()
    return out


def _check_grads(named, ref, what):
    flat_a, flat_b = [], []
    for k, g in named.items():
        assert g is not None, f"{what}: no gradient for {k}"
        r = ref[k].float()
        if float(r.abs().max()) == 0.0:
            assert float(g.abs().max()) < 1e-6, k
            continue
        cs = O.cosine(g, r)
        assert cs >= COS_TOL, f"{what}: grad cosine {cs:.5f} for {k}"
        flat_a.append(g.flatten().float())
        flat_b.append(r.flatten())
    cs = O.cosine(torch.cat(flat_a), torch.cat(flat_b))
    assert cs >= COS_TOL, f"{what}: global grad cosine {cs:.5f}"


HEADLINE = {
    # bench.py's default workload: config_direct.json with volume_size 128^3 (model_direct.py:44-53), conv-stack token grid
    "direct128": dict(kw=dict(volume_size=(128, 128, 128), in_channels=1, voxel_dim=256, depth=1, num_heads=4, context_dim=512,
                              cond_dim=1024), grid="conv", ctx_tokens=4096),
    # cascade stage 3's refiner ViT (model_progressive.py:247-256): 256^3, 32 input channels, 8 heads (d = 32), 4096 context tokens
    "stage3": dict(kw=dict(volume_size=(256, 256, 256), in_channels=32, voxel_dim=256, depth=1, num_heads=8, context_dim=512,
                           cond_dim=1024), grid="reference", ctx_tokens=4096),
}


@pytest.mark.parametrize("name", ["direct128", "stage3"])
def test_backbone_32768_tokens_forward_backward_vs_oracle(name):
    """HybridViT3D at the headline token count, depth 1, batch 1: forward and every gradient against the chunked fp32 oracle."""
    import hybrid_vit_cascade_b200 as hvc
    c = HEADLINE[name]
    kw = c["kw"]
    cfg = O.BackboneConfig(token_grid=c["grid"], **kw)
    assert cfg.num_tokens == N32K
    sd = {k: v.cuda() for k, v in O.init_state_dict(cfg, seed=7).items()}
    m = hvc.HybridViT3D(token_grid=c["grid"], **kw).cuda().eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator(device="cuda").manual_seed(17)
    vol = kw["volume_size"]
    x = torch.randn(1, kw["in_channels"], *vol, device="cuda", generator=g) * 0.5
    ctx = torch.randn(1, c["ctx_tokens"], 512, device="cuda", generator=g)
    cond = torch.randn(1, 1024, device="cuda", generator=g)
    r = torch.randn(1, 1, *vol, device="cuda", generator=g)
    y_ref, pg_ref, ig_ref = _oracle_run(cfg, sd, x, ctx, cond, r)
    xs = [t.clone().requires_grad_(True) for t in (x, ctx, cond)]
    y = m(*xs)
    err = O.max_rel(y, y_ref)
    assert err <= FWD_TOL, err
    (y * r).sum().backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    grads.update(x=xs[0].grad, ctx=xs[1].grad, cond=xs[2].grad)
    _check_grads(grads, dict(pg_ref, x=ig_ref[0], ctx=ig_ref[1], cond=ig_ref[2]), name)


def test_backbone_32768_tokens_batch8_last_sample_vs_oracle():
    """The bench's batch: B = 8 at 128^3 / 32768 tokens.  The last sample (largest offsets) of the batch-8 run against the oracle
    run on that sample alone: forward, and the gradients that are per-sample (input volume, context, cond)."""
    import hybrid_vit_cascade_b200 as hvc
    c = HEADLINE["direct128"]
    kw = c["kw"]
    cfg = O.BackboneConfig(token_grid="conv", **kw)
    sd = {k: v.cuda() for k, v in O.init_state_dict(cfg, seed=8).items()}
    m = hvc.HybridViT3D(token_grid="conv", **kw).cuda().eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator(device="cuda").manual_seed(19)
    B, vol = 8, kw["volume_size"]
    x = torch.randn(B, 1, *vol, device="cuda", generator=g) * 0.5
    ctx = torch.randn(B, 4096, 512, device="cuda", generator=g)
    cond = torch.randn(B, 1024, device="cuda", generator=g)
    r = torch.randn(B, 1, *vol, device="cuda", generator=g)
    xs = [t.clone().requires_grad_(True) for t in (x, ctx, cond)]
    y = m(*xs)
    (y * r).sum().backward()
    y = y.detach()
    gx, gctx, gcond = (t.grad for t in xs)
    m.zero_grad(set_to_none=True)
    torch.cuda.empty_cache()
    for b in (7,):
        y_ref, _, ig_ref = _oracle_run(cfg, sd, x[b:b + 1], ctx[b:b + 1], cond[b:b + 1], r[b:b + 1])
        assert O.max_rel(y[b:b + 1], y_ref) <= FWD_TOL
        for name, got, ref in (("x", gx, ig_ref[0]), ("ctx", gctx, ig_ref[1]), ("cond", gcond, ig_ref[2])):
            cs = O.cosine(got[b:b + 1], ref)
            assert cs >= COS_TOL, f"sample {b}: grad cosine {cs:.5f} for {name}"


@pytest.mark.parametrize("C,heads", [(384, 6), (512, 16)])
def test_backbone_shape_envelope_vs_oracle(C, heads):
    """The rest of the shape envelope: the default constructor's voxel_dim=384 / 6 heads (hybrid_vit_backbone.py:157-167, d = 64)
    and the H200 variant's 512 / 16 heads (model_progressive_h200.py:53-55, d = 32): forward + every gradient vs the fp32 oracle."""
    import hybrid_vit_cascade_b200 as hvc
    volume = (64, 64, 64)
    kw = dict(volume_size=volume, in_channels=1, voxel_dim=C, depth=2, num_heads=heads, context_dim=512, cond_dim=1024)
    cfg = O.BackboneConfig(**kw)
    sd = {k: v.cuda() for k, v in O.init_state_dict(cfg, seed=C).items()}
    m = hvc.HybridViT3D(**kw).cuda().eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator(device="cuda").manual_seed(23)
    B, M = 2, 1024
    x = torch.randn(B, 1, *volume, device="cuda", generator=g) * 0.5
    ctx = torch.randn(B, M, 512, device="cuda", generator=g)
    cond = torch.randn(B, 1024, device="cuda", generator=g)
    r = torch.randn(B, 1, *volume, device="cuda", generator=g)
    y_ref, pg_ref, ig_ref = _oracle_run(cfg, sd, x, ctx, cond, r, attn_chunk=1024)
    xs = [t.clone().requires_grad_(True) for t in (x, ctx, cond)]
    y = m(*xs)
    err = O.max_rel(y, y_ref)
    assert err <= FWD_TOL, err
    (y * r).sum().backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    grads.update(x=xs[0].grad, ctx=xs[1].grad, cond=xs[2].grad)
    _check_grads(grads, dict(pg_ref, x=ig_ref[0], ctx=ig_ref[1], cond=ig_ref[2]), f"C{C}")
