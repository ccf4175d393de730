"""Golden fixtures for the X-ray encoder and the direct-regression model from the REAL reference (authoring container only).

    python tests/golden/make_golden_encoder.py     ->  tests/golden/encoder.pt

Imports XrayConditioningModule (models/diagnostic_losses.py) and DirectCTRegression (direct_regression/model_direct.py) from
/root/reference, runs them on seeded inputs and stores weights, inputs, outputs, updated BatchNorm buffers and all gradients.
BatchNorm is in train mode (what the trainers run); nn.Dropout is switched off (p = 0) and AdaLN re-randomised, as for the
backbone fixtures.
"""
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "direct_regression"))

from models.diagnostic_losses import XrayConditioningModule  # noqa: E402
from model_direct import DirectCTRegression  # noqa: E402
sys.path.insert(0, os.path.join(REF, "direct_regression", "progressive_cascade"))
from model_progressive import MultiScaleXrayEncoder, Stage1Base64, Stage2Refiner128, Stage3Refiner256  # noqa: E402


def grads(mod, outs_and_rs, leaves):
    params = list(mod.parameters())
    names = [n for n, _ in mod.named_parameters()]
    loss = sum((o * r).sum() for o, r in outs_and_rs)
    gs = torch.autograd.grad(loss, params + leaves, allow_unused=True)
    pg = {n: (g if g is not None else torch.zeros_like(p)) for n, g, p in zip(names, gs, params)}
    return pg, list(gs[len(params):])


def randomise_adaln(m, seed):
    ga = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaln.linear" in n:
                p.copy_(torch.randn(p.shape, generator=ga) * 0.02)


def weight_checksum(m):
    return {k: float(v.double().sum()) for k, v in m.state_dict().items() if v.is_floating_point()}


def main():
    g = torch.Generator().manual_seed(4321)
    out = {}
    torch.manual_seed(0)
    enc = XrayConditioningModule(img_size=64, in_channels=1, embed_dim=64, num_views=2, time_embed_dim=32, cond_dim=96,
                                 share_view_weights=False).train()
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
    xr = (torch.rand(2, 2, 1, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    t = torch.randn(2, 32, generator=g)
    ctx, cond, feats = enc(xr, t)
    rs = [torch.randn(o.shape, generator=g) for o in (ctx, cond, feats)]
    pg, ig = grads(enc, list(zip((ctx, cond, feats), rs)), [xr])
    sd1 = {k: v.clone() for k, v in enc.state_dict().items() if "running" in k or "num_batches" in k}
    enc.eval()
    with torch.no_grad():
        ectx, econd, efeats = enc(xr, t)
    out["encoder"] = dict(sd=sd0, sd_after=sd1, xrays=xr.detach(), t=t, ctx=ctx.detach(), cond=cond.detach(), feats=feats.detach(),
                          r=rs, pgrad=pg, xgrad=ig[0], eval_ctx=ectx, eval_cond=econd, eval_feats=efeats)

    # one view (the else branch, diagnostic_losses.py:127)
    torch.manual_seed(1)
    enc1 = XrayConditioningModule(img_size=32, in_channels=1, embed_dim=32, num_views=1, time_embed_dim=16, cond_dim=48).train()
    sd0 = {k: v.clone() for k, v in enc1.state_dict().items()}
    xr1 = torch.rand(3, 1, 1, 32, 32, generator=g) * 2 - 1
    t1 = torch.randn(3, 16, generator=g)
    c1, d1, f1 = enc1(xr1, t1)
    out["encoder_one_view"] = dict(sd=sd0, xrays=xr1, t=t1, ctx=c1.detach(), cond=d1.detach(), feats=f1.detach())

    # DirectCTRegression, small
    torch.manual_seed(2)
    kw = dict(volume_size=(32, 32, 32), xray_img_size=64, voxel_dim=64, vit_depth=1, num_heads=1, xray_feature_dim=64)
    m = DirectCTRegression(**kw).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    randomise_adaln(m, 102)
    wsum = weight_checksum(m)
    xr = torch.rand(2, 2, 1, 64, 64, generator=g) * 2 - 1
    y = m(xr)
    r = torch.randn(y.shape, generator=g)
    pg, _ = grads(m, [(y, r)], [])
    # weights are not stored: the drop-in modules are built in the reference's order, so torch.manual_seed(seed) + the same
    # AdaLN re-randomisation reproduces them bit for bit (checked through `wsum`); gradients are compared by cosine: bf16 storage
    out["direct"] = dict(kwargs=kw, seed=2, adaln_seed=102, wsum=wsum, xrays=xr, y=y.detach(), r=r,
                         pgrad={k: v.bfloat16() for k, v in pg.items()})
    # MultiScaleXrayEncoder (cascade): the three stage branches, gradients through stage 1
    torch.manual_seed(3)
    ms = MultiScaleXrayEncoder(img_size=128, in_channels=1, base_dim=64, num_views=2).train()
    sd0 = {k: v.clone() for k, v in ms.state_dict().items()}
    ms_wsum = weight_checksum(ms)
    xr = torch.rand(2, 2, 1, 128, 128, generator=g) * 2 - 1
    res = {}
    for stage in (1, 2, 3):
        ms.load_state_dict(sd0)
        f, c, x = ms(xr, stage=stage)
        res[stage] = dict(feats=f.detach().clone(), cond=c.detach().clone(), ctx=x.detach().clone())
        if stage == 1:
            r = torch.randn(f.shape, generator=g)
            pg, _ = grads(ms, [(f, r), (c, torch.ones_like(c) * 0.01)], [])
            res["r1"], res["pgrad1"] = r, {k: v.bfloat16() for k, v in pg.items()}
    out["multiscale"] = dict(seed=3, wsum=ms_wsum, xrays=xr, **{f"stage{k}": v for k, v in res.items() if isinstance(k, int)},
                             r1=res["r1"], pgrad1=res["pgrad1"])

    # Stage1Base64 (cascade stage 1), small
    torch.manual_seed(4)
    kw = dict(volume_size=(32, 32, 32), xray_img_size=128, voxel_dim=64, vit_depth=1, num_heads=1, xray_feature_dim=64)
    m = Stage1Base64(**kw).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    randomise_adaln(m, 104)
    wsum = weight_checksum(m)
    xr = torch.rand(2, 2, 1, 128, 128, generator=g) * 2 - 1
    y = m(xr)
    r = torch.randn(y.shape, generator=g)
    pg, _ = grads(m, [(y, r)], [])
    out["stage1"] = dict(kwargs=kw, seed=4, adaln_seed=104, wsum=wsum, xrays=xr, y=y.detach(), r=r,
                         pgrad={k: v.bfloat16() for k, v in pg.items()})
    # Stage2Refiner128 (cascade stage 2), small: 16^3 -> 32^3, 8 x 8 context tokens, two heads of 32
    torch.manual_seed(5)
    kw = dict(volume_size=(32, 32, 32), voxel_dim=64, vit_depth=1, num_heads=2, xray_feature_dim=64)
    m = Stage2Refiner128(**kw).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    randomise_adaln(m, 105)
    wsum = weight_checksum(m)
    v64 = (torch.rand(2, 1, 16, 16, 16, generator=g) * 2 - 1).requires_grad_(True)
    feats = torch.randn(2, 64, 8, 8, generator=g).requires_grad_(True)
    cond = torch.randn(2, 1024, generator=g).requires_grad_(True)
    y = m(v64, feats, cond)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads(m, [(y, r)], [v64, feats, cond])
    out["stage2"] = dict(kwargs=kw, seed=5, adaln_seed=105, wsum=wsum, volume_64=v64.detach(), feats=feats.detach(), cond=cond.detach(),
                         y=y.detach(), r=r, pgrad={k: v.bfloat16() for k, v in pg.items()}, vgrad=ig[0], fgrad=ig[1], cgrad=ig[2])
    # Stage3Refiner256 (cascade stage 3), small: 16^3 -> 32^3 with the detail branch, 8 x 8 context tokens, two heads of 32
    torch.manual_seed(6)
    kw = dict(volume_size=(32, 32, 32), voxel_dim=64, vit_depth=1, num_heads=2, xray_feature_dim=64)
    m = Stage3Refiner256(**kw).train()                 # train(): the torch.utils.checkpoint branch (:286-293) is the one that runs
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    randomise_adaln(m, 106)
    wsum = weight_checksum(m)
    v128 = (torch.rand(2, 1, 16, 16, 16, generator=g) * 2 - 1).requires_grad_(True)
    feats = torch.randn(2, 64, 8, 8, generator=g).requires_grad_(True)
    cond = torch.randn(2, 1024, generator=g).requires_grad_(True)
    y = m(v128, feats, cond)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads(m, [(y, r)], [v128, feats, cond])
    out["stage3"] = dict(kwargs=kw, seed=6, adaln_seed=106, wsum=wsum, volume_128=v128.detach(), feats=feats.detach(), cond=cond.detach(),
                         y=y.detach(), r=r, pgrad={k: v.bfloat16() for k, v in pg.items()}, vgrad=ig[0], fgrad=ig[1], cgrad=ig[2])
    torch.save(out, os.path.join(HERE, "encoder.pt"))
    print({k: list(v.keys()) for k, v in out.items()})


if __name__ == "__main__":
    main()
