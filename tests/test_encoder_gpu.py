"""GPU parity of the X-ray encoder and the direct-regression model through the drop-in modules -> C ABI,
against golden fixtures produced by the real reference (tests/golden/make_golden_encoder.py).  Same bars as the backbone:
bf16 forward max|a-b|/max|b| <= 2e-2, gradient cosine >= 0.999 per tensor and globally."""
import os

import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-2
COS_TOL = 0.999
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "encoder.pt"), weights_only=False)


def _check_grads(named, ref, what):
    fa, fb = [], []
    for k, g in named.items():
        r = ref[k].float().cuda()
        if g is None:                          # parameter not on this path (e.g. the to_stage2 branch in a stage-1 call)
            assert float(r.abs().max()) == 0.0, f"{what}: no gradient for {k}"
            continue
        if k.endswith(("encoder.0.bias", "encoder.4.bias", "encoder.8.bias")):
            # a conv bias in front of a train-mode BatchNorm has a mathematically zero gradient (the batch mean removes it):
            # the reference holds fp32 rounding noise there, this path bf16 noise -- both must be negligible next to the
            # gradient of the same conv's weight
            wk = k[:-4] + "weight"
            scale = float(ref[wk].float().abs().max())
            assert float(r.abs().max()) < 1e-2 * scale and float(g.abs().max()) < 1e-2 * scale, (k, float(g.abs().max()), scale)
            continue
        if float(r.abs().max()) < 1e-5:      # exact zeros, and the conv biases in front of a BatchNorm (mathematically zero gradient)
            assert float(g.abs().max()) < 1e-2, (k, float(g.abs().max()))
            continue
        cs = O.cosine(g, r)
        assert cs >= COS_TOL, f"{what}: grad cosine {cs:.5f} for {k}"
        fa.append(g.flatten().float())
        fb.append(r.flatten())
    assert O.cosine(torch.cat(fa), torch.cat(fb)) >= COS_TOL


def test_xray_encoder_train_mode_golden():
    import hybrid_vit_cascade_b200 as hvc
    c = _gold()["encoder"]
    m = hvc.XrayConditioningModule(img_size=64, in_channels=1, embed_dim=64, num_views=2, time_embed_dim=32, cond_dim=96).cuda().train()
    m.load_state_dict(c["sd"], strict=True)
    xr = c["xrays"].cuda().requires_grad_(True)
    ctx, cond, feats = m(xr, c["t"].cuda())
    assert feats.shape == c["feats"].shape and ctx.shape == c["ctx"].shape
    for a, b, n in ((ctx, c["ctx"], "ctx"), (cond, c["cond"], "cond"), (feats, c["feats"], "feats")):
        assert O.max_rel(a, b) <= FWD_TOL, (n, O.max_rel(a, b))
    # BatchNorm buffers were updated like nn.BatchNorm2d does (momentum 0.1, unbiased variance)
    sd = m.state_dict()
    for k, v in c["sd_after"].items():
        if "num_batches" in k:
            assert int(sd[k]) == int(v), k
        else:
            assert O.max_rel(sd[k], v) <= 1e-2, (k, O.max_rel(sd[k], v))
    loss = sum((o * r.cuda()).sum() for o, r in zip((ctx, cond, feats), c["r"]))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    _check_grads(grads, c["pgrad"], "xray_encoder")
    assert O.cosine(xr.grad, c["xgrad"].cuda()) >= COS_TOL
    # the feature map is a free view of the token-major buffer the cross-attention reads
    tok = feats.flatten(2).transpose(1, 2)
    assert tok.stride(2) == 1 and tok.stride(1) == feats.shape[1]


def test_xray_encoder_eval_mode_and_one_view_golden():
    import hybrid_vit_cascade_b200 as hvc
    c = _gold()["encoder"]
    m = hvc.XrayConditioningModule(img_size=64, in_channels=1, embed_dim=64, num_views=2, time_embed_dim=32, cond_dim=96).cuda().eval()
    m.load_state_dict(dict(c["sd"], **c["sd_after"]), strict=True)
    with torch.no_grad():
        ctx, cond, feats = m(c["xrays"].cuda(), c["t"].cuda())
    for a, b in ((ctx, c["eval_ctx"]), (cond, c["eval_cond"]), (feats, c["eval_feats"])):
        assert O.max_rel(a, b) <= FWD_TOL
    c1 = _gold()["encoder_one_view"]
    m1 = hvc.XrayConditioningModule(img_size=32, in_channels=1, embed_dim=32, num_views=1, time_embed_dim=16, cond_dim=48).cuda().train()
    m1.load_state_dict(c1["sd"], strict=True)
    a, b, f = m1(c1["xrays"].cuda(), c1["t"].cuda())
    assert O.max_rel(a, c1["ctx"]) <= FWD_TOL and O.max_rel(b, c1["cond"]) <= FWD_TOL and O.max_rel(f, c1["feats"]) <= FWD_TOL


def _model_vs_fixture(cls, c, what):
    import hybrid_vit_cascade_b200 as hvc
    from conftest import rebuild_from_seed
    m = rebuild_from_seed(cls, c).cuda().train()
    hvc.set_dropout_policy("ignore")                  # the fixture was produced with nn.Dropout switched off
    try:
        y = m(c["xrays"].cuda())
        assert y.shape == c["y"].shape
        err = O.max_rel(y, c["y"])
        assert err <= FWD_TOL, err
        (y * c["r"].cuda()).sum().backward()
        grads = {k: p.grad for k, p in m.named_parameters()}
        _check_grads(grads, c["pgrad"], what)
    finally:
        hvc.set_dropout_policy("apply")


def test_direct_ct_regression_golden():
    """model_direct.py end to end: X-rays -> encoder -> context / cond -> 3D ViT -> volume, forward and every gradient."""
    import hybrid_vit_cascade_b200 as hvc
    _model_vs_fixture(hvc.DirectCTRegression, _gold()["direct"], "direct")


def test_cascade_stage1_golden():
    """Stage1Base64 (model_progressive.py:86-150): MultiScaleXrayEncoder stage-1 branch (two Conv2d s2 + GroupNorm + GELU) -> ViT."""
    import hybrid_vit_cascade_b200 as hvc
    _model_vs_fixture(hvc.Stage1Base64, _gold()["stage1"], "stage1")


def test_multi_scale_xray_encoder_golden():
    """MultiScaleXrayEncoder: features / cond / context of all three stage branches, gradients through the stage-1 branch."""
    import hybrid_vit_cascade_b200 as hvc
    from conftest import rebuild_from_seed
    c = _gold()["multiscale"]
    sd0 = {k: v.clone() for k, v in rebuild_from_seed(hvc.MultiScaleXrayEncoder, c, img_size=128, in_channels=1, base_dim=64,
                                                      num_views=2).state_dict().items()}
    for stage in (3, 2, 1):
        m = hvc.MultiScaleXrayEncoder(img_size=128, in_channels=1, base_dim=64, num_views=2).cuda().train()
        m.load_state_dict(sd0, strict=True)
        f, cond, ctx = m(c["xrays"].cuda(), stage=stage)
        ref = c[f"stage{stage}"]
        assert f.shape == ref["feats"].shape
        assert O.max_rel(f, ref["feats"]) <= FWD_TOL and O.max_rel(cond, ref["cond"]) <= FWD_TOL and O.max_rel(ctx, ref["ctx"]) <= FWD_TOL
    ((f * c["r1"].cuda()).sum() + 0.01 * cond.sum()).backward()
    grads = {k: p.grad for k, p in m.named_parameters() if not k.startswith("to_stage2")}     # the stage-2 branch is not on this path
    _check_grads(grads, c["pgrad1"], "multiscale stage 1")


def test_direct_ct_regression_config_direct_shapes_run():
    """config_direct.json sizes (512^2 X-rays, 64^3 volume, C = 256, 4 heads, 512-channel context): one training step runs and
    produces finite gradients for every parameter."""
    import hybrid_vit_cascade_b200 as hvc
    torch.manual_seed(0)
    m = hvc.DirectCTRegression(volume_size=(64, 64, 64), xray_img_size=512, voxel_dim=256, vit_depth=2, num_heads=4,
                               xray_feature_dim=512).cuda().train()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaln.linear" in n:
                p.normal_(0, 0.02)
    xr = torch.rand(2, 2, 1, 512, 512, device="cuda") * 2 - 1
    y = m(xr)
    assert y.shape == (2, 1, 64, 64, 64) and bool(torch.isfinite(y).all())
    y.abs().mean().backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), n


def test_direct_regression_loss_golden():
    """DirectRegressionLoss (L1 + 0.5 (1 - SSIM3D), model_direct.py:88-131) against values and gradients from the real reference."""
    import hybrid_vit_cascade_b200 as hvc
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "direct_loss.pt"), weights_only=False)
    crit = hvc.DirectRegressionLoss(1.0, 0.5)
    for name, c in gold.items():
        pred = c["pred"].cuda().requires_grad_(True)
        res = crit(pred, c["target"].cuda())
        for k, ref in (("total_loss", c["total"]), ("l1_loss", c["l1"]), ("ssim_loss", c["ssim"])):
            assert abs(float(res[k].detach()) - float(ref)) < 2e-5, (name, k, float(res[k].detach()), float(ref))
        (3.0 * res["total_loss"]).backward()                  # a non-unit upstream gradient, applied on the device
        assert O.max_rel(pred.grad, 3.0 * c["dpred"]) < 1e-4, (name, O.max_rel(pred.grad, 3.0 * c["dpred"]))
        assert O.cosine(pred.grad, c["dpred"].cuda()) > 0.99999
    # properties at the benchmark size: identical volumes -> SSIM = 1, loss = 0
    v = torch.rand(2, 1, 64, 64, 64, device="cuda") * 2 - 1
    res = crit(v, v.clone())
    assert abs(float(res["total_loss"].detach())) < 1e-5


def test_cascade_stage2_refiner_golden():
    """Stage2Refiner128 (model_progressive.py:153-215): trilinear x2 (align_corners=False) -> Conv3d(1->32)+GroupNorm+GELU -> refiner ViT
    (32 input channels, heads of 32) -> base + residual_weight * refinement; forward, parameter and input gradients."""
    import hybrid_vit_cascade_b200 as hvc
    from conftest import rebuild_from_seed
    c = _gold()["stage2"]
    m = rebuild_from_seed(hvc.Stage2Refiner128, c).cuda().train()
    hvc.set_dropout_policy("ignore")
    try:
        v64, feats, cond = (c[k].cuda().requires_grad_(True) for k in ("volume_64", "feats", "cond"))
        y = m(v64, feats, cond)
        assert y.shape == c["y"].shape
        err = O.max_rel(y, c["y"])
        assert err <= FWD_TOL, err
        (y * c["r"].cuda()).sum().backward()
        grads = {k: p.grad for k, p in m.named_parameters()}
        _check_grads(grads, c["pgrad"], "stage2")
        for a, b, n in ((v64.grad, c["vgrad"], "volume_64"), (feats.grad, c["fgrad"], "feats"), (cond.grad, c["cgrad"], "cond")):
            assert O.cosine(a, b.cuda()) >= COS_TOL, (n, O.cosine(a, b.cuda()))
    finally:
        hvc.set_dropout_policy("apply")
