"""Train-mode dropout (SURVEY.md finding 4, section 8(c) "stochastic mode"): the fused kernels regenerate their masks
from (seed, site, row, col); oracle/dropout_mask.py restates the generator, so with the SAME mask the CUDA path must
match the reference arithmetic exactly (bf16 tolerances), forward and backward, plus the statistical properties."""
import pytest
import torch

from oracle import dropout_mask as DM
from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-2
COS_TOL = 0.999


def _seed(words=(0x1234567, -0x3456789)):
    return torch.tensor(words, dtype=torch.int32, device="cuda")


@pytest.mark.parametrize("d,H,nq,nk,p", [(64, 2, 200, 328, 0.1), (32, 4, 384, 130, 0.1), (64, 1, 256, 256, 0.5)])
def test_attention_dropout_same_mask_same_numbers(d, H, nq, nk, p):
    from hybrid_vit_cascade_b200 import kernels as K
    B, C = 2, H * d
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    k = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    v = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    d_o = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    seed = _seed()
    site = 21
    drop = K.Drop(seed, site, p)
    o, lse2 = K.attn_fwd(q, k, v, B, H, nq, nk, d, d ** -0.5, drop=drop)
    dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, d ** -0.5, dq, dk, dv, drop=drop)
    # reference with the oracle's restatement of the mask
    pmask = DM.DropoutOracle(seed.tolist(), p, device="cuda").attn(site, B, H, nq, nk)
    keep_rate = float((pmask > 0).float().mean())
    assert abs(keep_rate - (1 - p)) < 0.01
    qf, kf, vf = (t.float().view(B, -1, H, d).permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v))
    ro, _ = O.attention_core(qf, kf, vf, d ** -0.5, pmask=pmask)
    ro.backward(d_o.float().view(B, nq, H, d).permute(0, 2, 1, 3))
    back = lambda t, n: t.permute(0, 2, 1, 3).reshape(B * n, C)
    assert O.max_rel(o.float(), back(ro.detach(), nq)) <= FWD_TOL
    for mine, ref, n in ((dq, qf.grad, nq), (dk, kf.grad, nk), (dv, vf.grad, nk)):
        r = back(ref, n)
        assert O.cosine(mine.float(), r) >= COS_TOL
        assert O.max_rel(mine.float(), r) <= 3e-2
    # the log-sum-exp is the undropped one (softmax normalises before nn.Dropout)
    o0, lse0 = K.attn_fwd(q, k, v, B, H, nq, nk, d, d ** -0.5)
    # (not bit-equal: the dropout instantiation evaluates every exp2 on the MUFU, the plain one 6 of 16 pairs as a polynomial)
    assert float((lse0[:, :, :nq] - lse2[:, :, :nq]).abs().max()) <= 1e-4
    assert O.max_rel(o.float(), o0.float()) > 0.05      # and the mask really was applied


def test_epilogue_dropout_keep_rate_and_site_independence():
    """GEMM epilogue dropout on a constant matrix: zero fraction = p, kept values = c/(1-p), different sites and seeds
    give independent masks, the same (seed, site) the same mask."""
    from hybrid_vit_cascade_b200 import kernels as K
    T, Kd, N, p = 4096, 64, 256, 0.1
    a = torch.ones(T, Kd, device="cuda", dtype=torch.bfloat16)
    b = torch.full((N, Kd), 1.0 / Kd, device="cuda", dtype=torch.bfloat16)

    def run(seed, site):
        return K.gemm(a, b, epilogue=K.EPI_F32, drop=K.Drop(seed, site, p))

    s1, s2 = _seed(), _seed((99, 100))
    y = run(s1, 3)
    zero = (y == 0)
    assert abs(float(zero.float().mean()) - p) < 2e-3
    assert float((y[~zero] - 1 / (1 - p)).abs().max()) < 1e-2
    assert torch.equal(y, run(s1, 3))
    want = DM.keep_mask(s1.tolist(), 3, torch.arange(T, device="cuda"), torch.arange(N, device="cuda"), p)
    assert torch.equal(~zero, want)
    for other in (run(s1, 4), run(s2, 3)):
        z2 = (other == 0)
        both = float((zero & z2).float().mean())
        assert abs(both - p * p) < 1.5e-3        # independent masks overlap with probability p^2
    # per-row and per-column drop counts are binomial-like (no stripes)
    assert float(zero.float().mean(0).std()) < 2.0 * (p * (1 - p) / T) ** 0.5
    assert float(zero.float().mean(1).std()) < 2.0 * (p * (1 - p) / N) ** 0.5


def _small_model(heads, dropout):
    import hybrid_vit_cascade_b200 as hvc
    torch.manual_seed(0)
    kw = dict(volume_size=(32, 32, 32), in_channels=2, voxel_dim=64 * heads if heads == 1 else 64, depth=2,
              num_heads=heads, context_dim=32, cond_dim=64)
    m = hvc.HybridViT3D(dropout=dropout, **kw).cuda()
    with torch.no_grad():
        for n, prm in m.named_parameters():
            if "adaln.linear" in n:
                prm.normal_(0, 0.02)
    return m, kw


@pytest.mark.parametrize("heads", [1, 2])
def test_backbone_train_mode_matches_oracle_with_the_same_masks(heads, monkeypatch):
    """HybridViT3D.train() with dropout 0.1 (the reference default, hybrid_vit_backbone.py:166): all 6 sites x 2 blocks."""
    from hybrid_vit_cascade_b200 import kernels as K
    p = 0.1
    m, kw = _small_model(heads, p)
    m.train()
    seed = _seed((424242, -77))
    monkeypatch.setattr(K, "new_seed", lambda device: seed)
    g = torch.Generator(device="cuda").manual_seed(3)
    B, M = 2, 72
    x = torch.randn(B, 2, 32, 32, 32, device="cuda", generator=g) * 0.5
    ctx = torch.randn(B, M, 32, device="cuda", generator=g)
    cond = torch.randn(B, 64, device="cuda", generator=g)
    r = torch.randn(B, 1, 32, 32, 32, device="cuda", generator=g)
    xs = [t.clone().requires_grad_(True) for t in (x, ctx, cond)]
    y = m(*xs)
    (y * r).sum().backward()
    cfg = O.BackboneConfig(**kw)
    sd = {k_: v_.detach().clone().requires_grad_(True) for k_, v_ in m.state_dict().items()}
    xr = [t.clone().requires_grad_(True) for t in (x, ctx, cond)]
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y_ref = O.backbone(xr[0], xr[1], xr[2], sd, cfg, drop=DM.DropoutOracle(seed.tolist(), p, device="cuda"))
        (y_ref * r).sum().backward()
        y_nodrop = O.backbone(x, ctx, cond, {k_: v_.detach() for k_, v_ in sd.items()}, cfg)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert O.max_rel(y, y_ref) <= FWD_TOL
    assert O.max_rel(y_ref, y_nodrop) > 2 * FWD_TOL          # dropout changes the output by much more than the tolerance
    flat_a, flat_b = [], []
    for name, prm in m.named_parameters():
        ga, gr = prm.grad, sd[name].grad
        assert O.cosine(ga, gr) >= COS_TOL, name
        flat_a.append(ga.flatten()); flat_b.append(gr.flatten())
    for a_, b_ in zip(xs, xr):
        assert O.cosine(a_.grad, b_.grad) >= COS_TOL
    assert O.cosine(torch.cat(flat_a), torch.cat(flat_b)) >= COS_TOL


def test_dropout_replay_eval_and_checkpoint():
    """Same CUDA generator state -> same masks (what torch.utils.checkpoint relies on); eval() turns dropout off;
    the seed really comes from torch's generator."""
    from torch.utils.checkpoint import checkpoint
    m, kw = _small_model(2, 0.1)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(2, 2, 32, 32, 32, device="cuda", generator=g) * 0.5).requires_grad_(True)
    ctx = torch.randn(2, 40, 32, device="cuda", generator=g)
    cond = torch.randn(2, 64, device="cuda", generator=g)
    m.train()
    torch.manual_seed(11)
    y1 = m(x, ctx, cond)
    y1.square().sum().backward()
    g1 = {k_: p_.grad.clone() for k_, p_ in m.named_parameters()}
    torch.manual_seed(11)
    y2 = m(x, ctx, cond)
    same = 2e-3      # not bit-equal: the GroupNorm statistics of the embed are reduced with float atomics (bf16 flips)
    assert O.max_rel(y1, y2) < same
    y3 = m(x, ctx, cond)                       # generator advanced: a different mask
    assert O.max_rel(y3, y1) > 1e-2
    m.zero_grad(set_to_none=True)
    torch.manual_seed(11)
    y4 = checkpoint(m, x, ctx, cond, use_reentrant=False)
    y4.square().sum().backward()
    assert O.max_rel(y4, y1) < same
    for k_, p_ in m.named_parameters():
        assert O.cosine(p_.grad, g1[k_]) > 0.9999, k_
    m.eval()
    e1, e2 = m(x, ctx, cond), m(x, ctx, cond)
    assert O.max_rel(e1, e2) < same
    # unbiasedness: the mean over many masks approaches the eval output
    m.train()
    acc = torch.zeros_like(e1)
    n = 24
    with torch.no_grad():
        for _ in range(n):
            acc += m(x, ctx, cond)
    single = O.max_rel(y1, e1)
    assert O.max_rel(acc / n, e1) < 0.45 * single


def test_mask_statistics_at_the_bench_shape_from_the_kernels():
    """correlation checks on GPU-GENERATED masks over the index ranges of the benchmark (32 (batch, head) slices x
    32768 queries = 2^20 rows; 32768 key columns): the GEMM epilogue writes 0 where it drops, so a constant product exposes the mask.
    Keep rate, lag correlations along both axes, across row blocks of 32768 (the same query in the next (batch, head) slice), across
    sites and across seeds."""
    from hybrid_vit_cascade_b200 import kernels as K
    p, Kd = 0.1, 64

    def mask(T, N, seed, site):
        a = torch.ones(T, Kd, device="cuda", dtype=torch.bfloat16)
        b = torch.full((N, Kd), 1.0 / Kd, device="cuda", dtype=torch.bfloat16)
        y = K.gemm(a, b, epilogue=K.EPI_F32, drop=K.Drop(seed, site, p))
        return (y != 0)

    s1, s2 = _seed(), _seed((0x1234568, -0x3456789))            # seeds one bit apart
    # (a) the whole column range of one key axis: 4096 rows x 32768 columns
    m = mask(4096, 32768, s1, 0).float() - (1 - p)
    noise = 5 * p * (1 - p) / m.numel() ** 0.5
    assert abs(float(m.mean())) < 4 * (p * (1 - p) / m.numel()) ** 0.5
    for lag in (1, 2, 3, 64, 127, 128, 129, 4096, 16384):
        assert abs(float((m[:, :-lag] * m[:, lag:]).mean())) < noise, ("col lag", lag)
    for lag in (1, 2, 128, 2048):
        assert abs(float((m[:-lag] * m[lag:]).mean())) < noise, ("row lag", lag)
    for other in (mask(4096, 32768, s1, 8), mask(4096, 32768, s1, 1), mask(4096, 32768, s2, 0)):      # next block's site, next site, next seed
        assert abs(float((m * (other.float() - (1 - p))).mean())) < noise
    del m, other
    # (b) the whole row range: 2^20 rows (b, h, q) x 256 columns; rows 32768 apart are the same query of the next (batch, head) slice
    r = mask(1 << 20, 256, s1, 0).float() - (1 - p)
    noise = 5 * p * (1 - p) / r.numel() ** 0.5
    assert abs(float(r.mean())) < 4 * (p * (1 - p) / r.numel()) ** 0.5
    for lag in (1, 32768, 65536, 4 * 32768):
        assert abs(float((r[:-lag] * r[lag:]).mean())) < noise * 1.1, ("row-block lag", lag)
    spread = float(r.mean(1).std())                               # per-row keep rates scatter like a binomial over 256 columns
    assert 0.9 < spread / (p * (1 - p) / 256) ** 0.5 < 1.1
