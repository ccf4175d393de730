"""CPU restatement of the X-ray encoder in front of the backbone and of the direct-regression model (test infrastructure:
only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this).

    XrayConditioningModule.forward   models/diagnostic_losses.py:107-138
    DirectCTRegression.forward       direct_regression/model_direct.py:59-85
    compute_ssim_loss / DirectRegressionLoss                                  direct_regression/model_direct.py:88-131
    MultiScaleXrayEncoder, Stage1Base64, Stage2Refiner128, Stage3Refiner256, ProgressiveCascadeModel
                                     direct_regression/progressive_cascade/model_progressive.py:57-83, 125-150, 193-215, 273-315, 371-402

Functional over a state_dict, like oracle/vit_oracle.py; autograd gives the gradients.  Pinned against outputs of the real
reference modules run in the authoring container (tests/golden/make_golden_encoder.py -> tests/golden/encoder.pt,
tests/golden/make_golden_loss.py -> direct_loss.pt; replayed by tests/test_oracle_golden.py).
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from . import vit_oracle as V

StateDict = Dict[str, Tensor]


def _conv_bn_relu(x: Tensor, sd: StateDict, pfx: str, ci: int, bi: int, stride: int, pad: int, training: bool, momentum: float = 0.1,
                  new_stats: Optional[dict] = None) -> Tensor:
    """nn.Conv2d -> nn.BatchNorm2d -> nn.ReLU (diagnostic_losses.py:81-93).  Train mode normalises with batch statistics; the
    updated running buffers are returned through `new_stats` instead of being written in place."""
    x = F.conv2d(x, sd[f"{pfx}{ci}.weight"], sd[f"{pfx}{ci}.bias"], stride=stride, padding=pad)
    rm, rv = sd[f"{pfx}{bi}.running_mean"], sd[f"{pfx}{bi}.running_var"]
    if training:
        rm2, rv2 = rm.detach().clone(), rv.detach().clone()
        x = F.batch_norm(x, rm2, rv2, sd[f"{pfx}{bi}.weight"], sd[f"{pfx}{bi}.bias"], True, momentum, 1e-5)
        if new_stats is not None:
            new_stats[f"{pfx}{bi}.running_mean"], new_stats[f"{pfx}{bi}.running_var"] = rm2, rv2
    else:
        x = F.batch_norm(x, rm, rv, sd[f"{pfx}{bi}.weight"], sd[f"{pfx}{bi}.bias"], False, momentum, 1e-5)
    return F.relu(x)


def xray_conditioning(xrays: Tensor, t: Tensor, sd: StateDict, pfx: str = "", training: bool = True, new_stats: Optional[dict] = None):
    """diagnostic_losses.py:107-138 -> (xray_context, time_xray_cond, xray_features_2d)."""
    B, num_views = xrays.shape[0], xrays.shape[1]
    e = pfx + "encoder."

    def encoder(x):                                                            # :81-93
        x = _conv_bn_relu(x, sd, e, 0, 1, 2, 3, training, new_stats=new_stats)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        x = _conv_bn_relu(x, sd, e, 4, 5, 1, 1, training, new_stats=new_stats)
        x = F.max_pool2d(x, kernel_size=2, stride=2)
        return _conv_bn_relu(x, sd, e, 8, 9, 1, 1, training, new_stats=new_stats)

    if num_views > 1:                                                          # :118-125
        feats = encoder(xrays.reshape(B * num_views, *xrays.shape[2:]))
        feats = feats.view(B, num_views, *feats.shape[1:]).mean(dim=1)
    else:
        feats = encoder(xrays[:, 0])                                           # :127
    ctx = F.linear(feats.mean(dim=[-2, -1]), sd[pfx + "to_cond.weight"], sd[pfx + "to_cond.bias"])        # :130-131
    h = F.silu(F.linear(t, sd[pfx + "time_mlp.0.weight"], sd[pfx + "time_mlp.0.bias"]))                  # :98-102, :134
    time_embed = F.linear(h, sd[pfx + "time_mlp.2.weight"], sd[pfx + "time_mlp.2.bias"])
    return ctx, time_embed + ctx, feats                                        # :135-137


def direct_ct_regression(xrays: Tensor, sd: StateDict, cfg: V.BackboneConfig, training: bool = True, new_stats: Optional[dict] = None,
                         attn_chunk: Optional[int] = None) -> Tensor:
    """model_direct.py:59-85 (dropout off in the backbone)."""
    B = xrays.shape[0]
    dummy_t = torch.zeros(B, 256, device=xrays.device, dtype=xrays.dtype)      # :70
    _, cond, feats = xray_conditioning(xrays, dummy_t, sd, "xray_encoder.", training, new_stats)   # :73
    x = sd["initial_volume"].expand(B, -1, -1, -1, -1)                          # :76
    bsd = {k[len("vit_backbone."):]: v for k, v in sd.items() if k.startswith("vit_backbone.")}
    return V.backbone(x, feats.flatten(2).transpose(1, 2), cond, bsd, cfg, attn_chunk=attn_chunk)   # :79-84


def ssim_loss(pred: Tensor, target: Tensor, window_size: int = 11) -> Tensor:
    """compute_ssim_loss, model_direct.py:88-107."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    pad = window_size // 2

    def box(x):
        return F.avg_pool3d(x, window_size, stride=1, padding=pad)

    mu_p, mu_t = box(pred), box(target)
    s_pp = box(pred ** 2) - mu_p ** 2
    s_tt = box(target ** 2) - mu_t ** 2
    s_pt = box(pred * target) - mu_p * mu_t
    ssim = ((2 * mu_p * mu_t + C1) * (2 * s_pt + C2)) / ((mu_p ** 2 + mu_t ** 2 + C1) * (s_pp + s_tt + C2))
    return 1 - ssim.mean()


def direct_regression_loss(pred: Tensor, target: Tensor, l1_weight: float = 1.0, ssim_weight: float = 0.5):
    """DirectRegressionLoss.forward, model_direct.py:118-131."""
    l1 = F.l1_loss(pred, target)
    ss = ssim_loss(pred, target)
    return {"total_loss": l1_weight * l1 + ssim_weight * ss, "l1_loss": l1, "ssim_loss": ss}


def multi_scale_xray_encoder(xrays: Tensor, sd: StateDict, pfx: str = "", stage: int = 1, training: bool = True,
                             new_stats: Optional[dict] = None):
    """MultiScaleXrayEncoder.forward, progressive_cascade/model_progressive.py:57-83 -> (features, time_xray_cond, xray_context)."""
    B = xrays.shape[0]
    dummy_t = torch.zeros(B, 256, device=xrays.device, dtype=xrays.dtype)                        # :70
    ctx, cond, feats = xray_conditioning(xrays, dummy_t, sd, pfx + "xray_encoder.", training, new_stats)   # :73

    def down(x, branch, ci, gi):                                                                  # Conv2d(s2) -> GroupNorm(32) -> GELU, :38-52
        x = F.conv2d(x, sd[f"{pfx}{branch}.{ci}.weight"], sd[f"{pfx}{branch}.{ci}.bias"], stride=2, padding=1)
        x = F.group_norm(x, 32, sd[f"{pfx}{branch}.{gi}.weight"], sd[f"{pfx}{branch}.{gi}.bias"], 1e-5)
        return F.gelu(x)

    if stage == 1:                                                                                # :76-78
        feats = down(down(feats, "to_stage1", 0, 1), "to_stage1", 3, 4)
    elif stage == 2:                                                                              # :79-81
        feats = down(feats, "to_stage2", 0, 1)
    return feats, cond, ctx


def stage1_base64(xrays: Tensor, sd: StateDict, cfg: V.BackboneConfig, training: bool = True, attn_chunk: Optional[int] = None) -> Tensor:
    """Stage1Base64.forward, progressive_cascade/model_progressive.py:125-150 (dropout off in the backbone)."""
    B = xrays.shape[0]
    feats, cond, _ = multi_scale_xray_encoder(xrays, sd, "xray_encoder.", 1, training)
    x = sd["initial_volume"].expand(B, -1, -1, -1, -1)
    bsd = {k[len("vit_backbone."):]: v for k, v in sd.items() if k.startswith("vit_backbone.")}
    return V.backbone(x, feats.flatten(2).transpose(1, 2), cond, bsd, cfg, attn_chunk=attn_chunk)


def stage2_refiner128(volume_64: Tensor, xray_features_2d: Tensor, cond: Tensor, sd: StateDict, cfg: V.BackboneConfig,
                      attn_chunk: Optional[int] = None) -> Tensor:
    """Stage2Refiner128.forward, progressive_cascade/model_progressive.py:193-215 (dropout off in the refiner ViT)."""
    x = F.interpolate(volume_64, scale_factor=2, mode="trilinear", align_corners=False)            # nn.Upsample, :169
    x = F.conv3d(x, sd["upsample_from_64.1.weight"], sd["upsample_from_64.1.bias"], padding=1)     # :170
    x = F.gelu(F.group_norm(x, 8, sd["upsample_from_64.2.weight"], sd["upsample_from_64.2.bias"], 1e-5))   # :171-172
    bsd = {k[len("vit_refiner."):]: v for k, v in sd.items() if k.startswith("vit_refiner.")}
    refinement = V.backbone(x, xray_features_2d.flatten(2).transpose(1, 2), cond, bsd, cfg, attn_chunk=attn_chunk)   # :203-208
    up = F.interpolate(volume_64, size=tuple(cfg.volume_size), mode="trilinear", align_corners=False)     # :211-212
    return up + sd["residual_weight"] * refinement                                                 # :213


def stage3_refiner256(volume_128: Tensor, xray_features_2d: Tensor, cond: Tensor, sd: StateDict, cfg: V.BackboneConfig,
                      attn_chunk: Optional[int] = None) -> Tensor:
    """Stage3Refiner256.forward, progressive_cascade/model_progressive.py:273-307 (dropout off in the refiner ViT; gradient
    checkpointing, :286-293, changes memory only)."""
    x = F.interpolate(volume_128, scale_factor=2, mode="trilinear", align_corners=False)           # nn.Upsample, :239
    x = F.conv3d(x, sd["upsample_from_128.1.weight"], sd["upsample_from_128.1.bias"], padding=1)   # :240
    x = F.gelu(F.group_norm(x, 8, sd["upsample_from_128.2.weight"], sd["upsample_from_128.2.bias"], 1e-5))   # :241-242
    bsd = {k[len("vit_refiner."):]: v for k, v in sd.items() if k.startswith("vit_refiner.")}
    refinement = V.backbone(x, xray_features_2d.flatten(2).transpose(1, 2), cond, bsd, cfg, attn_chunk=attn_chunk)   # :309-315
    up = F.interpolate(volume_128, size=tuple(cfg.volume_size), mode="trilinear", align_corners=False)    # :296-297
    d = F.conv3d(up, sd["detail_enhancer.0.weight"], sd["detail_enhancer.0.bias"], padding=1)      # :260
    d = F.gelu(F.group_norm(d, 16, sd["detail_enhancer.1.weight"], sd["detail_enhancer.1.bias"], 1e-5))    # :261-262
    d = F.conv3d(d, sd["detail_enhancer.3.weight"], sd["detail_enhancer.3.bias"], padding=1)       # :263
    d = F.gelu(F.group_norm(d, 8, sd["detail_enhancer.4.weight"], sd["detail_enhancer.4.bias"], 1e-5))     # :264-265
    d = F.conv3d(d, sd["detail_enhancer.6.weight"], sd["detail_enhancer.6.bias"])                  # :266
    return up + sd["residual_weight"] * refinement + sd["detail_weight"] * d                       # :303-305


def progressive_cascade(xrays: Tensor, sd: StateDict, cfgs: Dict[int, V.BackboneConfig], max_stage: int = 3, training: bool = True,
                        attn_chunk: Optional[int] = None) -> Dict[str, Tensor]:
    """ProgressiveCascadeModel.forward(return_intermediate=True), progressive_cascade/model_progressive.py:371-407.  cfgs[k] is the
    backbone configuration of stage k's ViT.  Stage 1 runs its own encoder (:135), stages 2 and 3 the shared one, once per stage
    on the same X-rays (:386,:394) -- in training mode every one of those calls updates that encoder's BatchNorm statistics, which
    does not change any output of this forward."""
    def sub(p):
        return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    out = {"stage1": stage1_base64(xrays, sub("stage1."), cfgs[1], training=training, attn_chunk=attn_chunk)}
    if max_stage >= 2:
        f2, c2, _ = multi_scale_xray_encoder(xrays, sd, "xray_encoder.", stage=2, training=training)
        out["stage2"] = stage2_refiner128(out["stage1"], f2, c2, sub("stage2."), cfgs[2], attn_chunk=attn_chunk)
    if max_stage >= 3:
        f3, c3, _ = multi_scale_xray_encoder(xrays, sd, "xray_encoder.", stage=3, training=training)
        out["stage3"] = stage3_refiner256(out["stage2"], f3, c3, sub("stage3."), cfgs[3], attn_chunk=attn_chunk)
    return out
