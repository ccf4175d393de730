"""Round-2 golden fixtures from the REAL reference (authoring container only; /root/reference is read at generation time, never by a test).

    python tests/golden/make_golden_r02.py   ->  tests/golden/r02_losses.pt, tests/golden/r02_checkpoint_layout.json, tests/golden/r02_views.pt

  * r02_losses.pt: SSIMLoss, TotalVariationLoss (with and without a target), FrequencyLoss, DRRReprojectionLoss, Stage1Loss of
    direct_regression/progressive_cascade/loss_multiscale.py on seeded non-cubic volumes: values and d loss / d pred.
  * r02_checkpoint_direct.pt: a checkpoint written by the reference trainer's own save lines (train_direct_4gpu.py:277-297) for a small
    DirectCTRegression after one AdamW step on the CPU, plus the eval-mode output of the saved model on a recorded input.
  * r02_vgg.pt: TriPlanarVGGLoss, Stage2Loss, Stage3Loss (with X-rays) and MultiScaleLoss(stage=2) with seeded stand-in VGG16 weights.
  * r02_views.pt: XrayConditioningModule constructed with num_views=1 (its default) fed a TWO-view input (diagnostic_losses.py:118-125
    decides on the input's view count).
"""
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "direct_regression"))
sys.path.insert(0, os.path.join(REF, "direct_regression", "progressive_cascade"))

from loss_multiscale import DRRReprojectionLoss, FrequencyLoss, SSIMLoss, Stage1Loss, TotalVariationLoss, compute_psnr, compute_ssim_metric  # noqa: E402
from model_direct import DirectCTRegression, DirectRegressionLoss  # noqa: E402
from models.diagnostic_losses import XrayConditioningModule  # noqa: E402


def val_and_grad(fn, pred):
    p = pred.clone().requires_grad_(True)
    v = fn(p)
    v = v["total_loss"] if isinstance(v, dict) else v
    g, = torch.autograd.grad(v, p)
    return float(v), g


def losses():
    g = torch.Generator().manual_seed(2024)
    B, D, H, W = 2, 16, 20, 24
    pred = torch.rand(B, 1, D, H, W, generator=g) * 2 - 1
    target = (pred + 0.3 * torch.randn(B, 1, D, H, W, generator=g)).clamp(-1, 1)
    xrays = torch.rand(B, 2, 1, 48, 48, generator=g) * 2 - 1
    out = dict(pred=pred, target=target, xrays=xrays, img_size=48)
    out["ssim"] = val_and_grad(lambda p: SSIMLoss()(p, target), pred)
    out["tv_pred_only"] = val_and_grad(lambda p: TotalVariationLoss()(p), pred)
    out["tv_vs_target"] = val_and_grad(lambda p: TotalVariationLoss()(p, target), pred)
    smooth = torch.nn.functional.avg_pool3d(pred, 3, 1, 1)                                   # tv(pred) < tv(target): the other sign of the L1
    out["smooth"] = smooth
    out["tv_vs_target_smooth"] = val_and_grad(lambda p: TotalVariationLoss()(p, target), smooth)
    out["freq"] = val_and_grad(lambda p: FrequencyLoss(high_freq_weight=2.0)(p, target), pred)
    out["drr"] = val_and_grad(lambda p: DRRReprojectionLoss(img_size=48)(p, xrays), pred)
    out["stage1"] = val_and_grad(lambda p: Stage1Loss()(p, target), pred)
    out["psnr"] = compute_psnr(pred, target)
    out["ssim_metric"] = compute_ssim_metric(pred, target)
    torch.save(out, os.path.join(HERE, "r02_losses.pt"))


def checkpoint():
    torch.manual_seed(11)
    cfg = {"model": dict(volume_size=[32, 32, 32], xray_img_size=64, voxel_dim=64, vit_depth=1, num_heads=1, xray_feature_dim=64),
           "training": dict(learning_rate=1e-3, weight_decay=0.01, num_epochs=10, gradient_clip=1.0),
           "checkpoints": dict(save_dir="checkpoints_direct", save_every=5)}
    m = DirectCTRegression(volume_size=tuple(cfg["model"]["volume_size"]), **{k: v for k, v in cfg["model"].items() if k != "volume_size"})
    ga = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaln.linear" in n:
                p.copy_(torch.randn(p.shape, generator=ga) * 0.02)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    opt = torch.optim.AdamW(m.parameters(), lr=cfg["training"]["learning_rate"], weight_decay=cfg["training"]["weight_decay"])   # train_direct_4gpu.py:159-163
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=cfg["training"]["num_epochs"])                                # :165-168
    g = torch.Generator().manual_seed(13)
    xr = torch.rand(2, 2, 1, 64, 64, generator=g) * 2 - 1
    tgt = torch.rand(2, 1, 32, 32, 32, generator=g) * 2 - 1
    crit = DirectRegressionLoss(1.0, 0.5)
    m.train()
    loss = crit(m(xr), tgt)["total_loss"]
    opt.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(m.parameters(), cfg["training"]["gradient_clip"])
    opt.step()
    sched.step()
    ddp_like = torch.nn.Module()
    ddp_like.module = m                                                                      # the trainer saves model.module.state_dict() (:280)
    ckpt = {"epoch": 3, "model_state_dict": ddp_like.module.state_dict(), "optimizer_state_dict": opt.state_dict(),
            "scheduler_state_dict": sched.state_dict(), "val_psnr": 21.5, "best_psnr": 21.5, "config": cfg}                     # :277-287
    # Only the LAYOUT is recorded (a few KB): the tests rebuild a checkpoint of this exact structure with the oracle port + torch's own
    # AdamW / CosineAnnealingLR (the classes the reference trainer uses) and check it against this table.
    import json
    osd = ckpt["optimizer_state_dict"]
    layout = {
        "checkpoint_keys": sorted(ckpt.keys()),
        "model_state_dict": {k: [list(v.shape), str(v.dtype)] for k, v in ckpt["model_state_dict"].items()},
        "optimizer_param_group_keys": sorted(osd["param_groups"][0].keys()),
        "optimizer_params_index": osd["param_groups"][0]["params"],
        "optimizer_state": {str(i): {k: [list(v.shape), str(v.dtype)] for k, v in st.items()} for i, st in osd["state"].items()},
        "parameter_order": [n for n, _ in m.named_parameters()],
        "scheduler_state_keys": sorted(ckpt["scheduler_state_dict"].keys()),
        "config": cfg,
        "torch": torch.__version__,
    }
    with open(os.path.join(HERE, "r02_checkpoint_layout.json"), "w") as f:
        json.dump(layout, f, indent=1)


def views():
    torch.manual_seed(21)
    enc = XrayConditioningModule(img_size=32, in_channels=1, embed_dim=32, time_embed_dim=16, cond_dim=48).train()    # num_views defaults to 1
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
    g = torch.Generator().manual_seed(22)
    xr = torch.rand(3, 2, 1, 32, 32, generator=g) * 2 - 1
    t = torch.randn(3, 16, generator=g)
    c, d, f = enc(xr, t)
    torch.save(dict(sd=sd0, xrays=xr, t=t, ctx=c.detach(), cond=d.detach(), feats=f.detach()), os.path.join(HERE, "r02_views.pt"))


def vgg():
    """TriPlanarVGGLoss / Stage2Loss / Stage3Loss of the reference, with torchvision's vgg16 constructor pointed at seeded stand-in
    weights (oracle.loss_oracle.vgg16_features_state(5)) instead of the ImageNet download -- the only change; the classes are the
    reference's own.  Inputs: pred / target / xrays of r02_losses.pt."""
    import torchvision.models as tvm
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.loss_oracle import vgg16_features_state
    import loss_multiscale as LM
    real_vgg16 = tvm.vgg16

    def seeded_vgg16(*a, **k):
        m = real_vgg16(weights=None)
        m.load_state_dict(vgg16_features_state(5), strict=False)
        return m

    tvm.vgg16 = seeded_vgg16
    try:
        base = torch.load(os.path.join(HERE, "r02_losses.pt"))
        pred, target, xrays = base["pred"], base["target"], base["xrays"]
        out = {"vgg_seed": 5}
        out["vgg"] = val_and_grad(lambda p: LM.TriPlanarVGGLoss()(p, target), pred)
        out["stage2"] = val_and_grad(lambda p: LM.Stage2Loss()(p, target), pred)
        s3 = LM.Stage3Loss()
        s3.drr_loss = LM.DRRReprojectionLoss(img_size=48)
        out["stage3"] = val_and_grad(lambda p: s3(p, target, xrays), pred)
        ms = LM.MultiScaleLoss()
        out["multiscale_stage2_parts"] = {k: float(v) for k, v in ms(pred, target, stage=2).items()}
        torch.save(out, os.path.join(HERE, "r02_vgg.pt"))
    finally:
        tvm.vgg16 = real_vgg16


if __name__ == "__main__":
    losses()
    checkpoint()
    views()
    vgg()
    for f in ("r02_losses.pt", "r02_checkpoint_layout.json", "r02_views.pt", "r02_vgg.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
