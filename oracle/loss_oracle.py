"""CPU restatement of the stage 2-3 loss terms of the progressive cascade (test infrastructure only: see oracle/vit_oracle.py).

    SSIMLoss                 direct_regression/progressive_cascade/loss_multiscale.py:18-51
    TotalVariationLoss       :140-188
    FrequencyLoss            :191-236
    DRRReprojectionLoss      :239-293
    TriPlanarVGGLoss         :54-137   (the VGG16 `features` weights are an ARGUMENT here: the reference downloads the ImageNet ones in
                             its constructor, which cannot happen offline; tests use seeded stand-in weights, `vgg16_features_state`)
    Stage1Loss / Stage2Loss / Stage3Loss   :296-432   (the stage totals take the VGG term's value as an argument)

Plain torch ops in the reference's order; autograd gives the gradients.  Pinned against outputs of the real reference classes
(tests/golden/make_golden_r02.py -> tests/golden/r02_losses.pt, replayed by tests/test_oracle_golden_r02.py).
"""
import torch
import torch.nn.functional as F


def ssim_loss(pred, target, window_size=11):
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    w = min(window_size, pred.shape[2], pred.shape[3], pred.shape[4])                       # :35

    def box(x):
        return F.avg_pool3d(x, w, stride=1, padding=w // 2)

    mu_p, mu_t = box(pred), box(target)
    s_pp, s_tt, s_pt = box(pred ** 2) - mu_p ** 2, box(target ** 2) - mu_t ** 2, box(pred * target) - mu_p * mu_t
    ssim = ((2 * mu_p * mu_t + C1) * (2 * s_pt + C2)) / ((mu_p ** 2 + mu_t ** 2 + C1) * (s_pp + s_tt + C2))
    return 1 - ssim.mean()                                                                   # :51


def _tv(v, eps):
    v = v.float()
    dd = torch.abs(v[:, :, 1:] - v[:, :, :-1])                                               # :162-164
    dh = torch.abs(v[:, :, :, 1:] - v[:, :, :, :-1])
    dw = torch.abs(v[:, :, :, :, 1:] - v[:, :, :, :, :-1])
    tv = (torch.sqrt(dd.pow(2) + eps).mean() + torch.sqrt(dh.pow(2) + eps).mean() + torch.sqrt(dw.pow(2) + eps).mean()) / 3   # :167-169
    return torch.clamp(tv, 0, 100)                                                           # :172


def total_variation_loss(pred, target=None, eps=1e-8):
    tv_p = _tv(pred, eps)
    if target is None:
        return tv_p                                                                          # :188
    return F.l1_loss(tv_p, _tv(target, eps))                                                 # :186


def frequency_loss(pred, target, high_freq_weight=2.0):
    pf = torch.fft.fftn(pred, dim=(-3, -2, -1))                                              # :206-207
    tf = torch.fft.fftn(target, dim=(-3, -2, -1))
    pm, tm = torch.abs(pf), torch.abs(tf)                                                    # :210-211
    D, H, W = pred.shape[-3:]
    radius = min(D, H, W) // 4                                                               # :216
    dd, hh, ww = torch.meshgrid(torch.arange(D, device=pred.device).float() - D // 2, torch.arange(H, device=pred.device).float() - H // 2,
                                torch.arange(W, device=pred.device).float() - W // 2, indexing="ij")   # :219-223
    mask = (torch.sqrt(dd ** 2 + hh ** 2 + ww ** 2) > radius).float()[None, None]            # :224-228 (on the UNSHIFTED spectrum)
    low = F.l1_loss(pm * (1 - mask), tm * (1 - mask))                                        # :231
    high = F.l1_loss(pm * mask, tm * mask)                                                   # :232
    return low + high_freq_weight * high                                                     # :234


def generate_drr(vol, view_angle, img_size):
    drr = torch.mean(vol, dim=2) if view_angle == 0 else torch.mean(vol, dim=4)              # :260-265
    return F.interpolate(drr, size=(img_size, img_size), mode="bilinear", align_corners=False)   # :268-269


def drr_reprojection_loss(pred, input_xrays, img_size=512):
    ap, lat = generate_drr(pred, 0, img_size), generate_drr(pred, 90, img_size)             # :281-282
    return (F.l1_loss(ap, input_xrays[:, 0]) + F.l1_loss(lat, input_xrays[:, 1])) / 2        # :285-293


# torchvision vgg16().features[:16]: (index, in, out) of its convolutions (3x3, padding 1, each followed by ReLU); max-pool 2x2 at 4 and 9
VGG16_CONVS = ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256))
VGG16_POOLS = (4, 9)


def vgg16_features_state(seed=0):
    """Seeded stand-in for the ImageNet weights, keyed like torchvision's vgg16().state_dict() ('features.N.weight' / '.bias')."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for idx, cin, cout in VGG16_CONVS:
        sd[f"features.{idx}.weight"] = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
        sd[f"features.{idx}.bias"] = torch.randn(cout, generator=g) * 0.05
    return sd


def _vgg_prefix(x, sd, stop):
    """vgg.features[:stop](x), from the input every time as the reference does (:84-92)."""
    convs = {i for i, _, _ in VGG16_CONVS}
    for i in range(stop):
        if i in convs:
            x = F.relu(F.conv2d(x, sd[f"features.{i}.weight"].to(x), sd[f"features.{i}.bias"].to(x), padding=1))
        elif i in VGG16_POOLS:
            x = F.max_pool2d(x, 2, 2)
    return x


def triplanar_vgg_loss(pred, target, vgg_state):
    D, H, W = pred.shape[2:]
    planes = lambda v: (v[:, :, D // 2, :, :], v[:, :, :, H // 2, :], v[:, :, :, :, W // 2])          # :101-114 axial / sagittal / coronal
    total = 0.0
    for ps, ts in zip(planes(pred), planes(target)):
        ps, ts = ((ps + 1) / 2).repeat(1, 3, 1, 1), ((ts + 1) / 2).repeat(1, 3, 1, 1)                    # :121-126
        for stop in (4, 9, 16):                                                                        # relu1_2, relu2_2, relu3_3 (:72-76), weights 1
            total = total + F.l1_loss(_vgg_prefix(ps, vgg_state, stop), _vgg_prefix(ts, vgg_state, stop))   # :133-135
    return total / 3                                                                                   # :137


def stage1_loss(pred, target, l1_weight=1.0, ssim_weight=0.5):
    l1, ss = F.l1_loss(pred, target), ssim_loss(pred, target)
    return {"total_loss": l1_weight * l1 + ssim_weight * ss, "l1_loss": l1, "ssim_loss": ss}   # :315-324


def stage2_loss(pred, target, vgg=0.0, l1_weight=1.0, ssim_weight=0.5, vgg_weight=0.1, tv_weight=0.02, freq_weight=0.05):
    l1, ss = F.l1_loss(pred, target), ssim_loss(pred, target)                                # :353-357
    tv, fr = total_variation_loss(pred, target), frequency_loss(pred, target)
    total = l1_weight * l1 + ssim_weight * ss + vgg_weight * vgg + tv_weight * tv + freq_weight * fr   # :359-363
    return {"total_loss": total, "l1_loss": l1, "ssim_loss": ss, "vgg_loss": vgg, "tv_loss": tv, "freq_loss": fr}


def stage3_loss(pred, target, input_xrays=None, vgg=0.0, l1_weight=1.0, ssim_weight=0.5, vgg_weight=0.1, tv_weight=0.03, freq_weight=0.07,
                drr_weight=0.3, img_size=512):
    out = stage2_loss(pred, target, vgg, l1_weight, ssim_weight, vgg_weight, tv_weight, freq_weight)   # :404-423
    if input_xrays is not None:                                                              # :426-430
        drr = drr_reprojection_loss(pred, input_xrays, img_size)
        out["total_loss"] = out["total_loss"] + drr_weight * drr
        out["drr_loss"] = drr
    return out


def psnr(pred, target):
    mse = torch.mean((pred - target) ** 2)                                                   # :495-500
    return float("inf") if mse == 0 else float(20 * torch.log10(2.0 / torch.sqrt(mse)))
