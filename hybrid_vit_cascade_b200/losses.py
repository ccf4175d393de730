"""Drop-in replacement for the reference's ``DirectRegressionLoss`` / ``compute_ssim_loss``
(direct_regression/model_direct.py:88-131): l1_weight * L1 + ssim_weight * (1 - mean SSIM3D over an 11^3 box window), computed
by the kernels of csrc/hvc_loss.cu (separable box filters on the stacked volumes, one fused pointwise pass each way).
Returns the same dict (``total_loss``, ``l1_loss``, ``ssim_loss``); ``total_loss`` carries the gradient w.r.t. ``pred``
(the trainer backpropagates only it, train_direct_4gpu.py), the two components are detached values.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import kernels as K


class _DirectLoss(Function):
    @staticmethod
    def forward(ctx, pred, target, l1_weight, ssim_weight, window):
        pred32 = pred.float().contiguous()
        target32 = target.detach().float().contiguous()
        sums, filtered = K.ssim_l1_fwd(pred32, target32, window)
        n = pred32.numel()
        mean = (sums / n).float()
        ssim_loss = 1.0 - mean[0]
        l1 = mean[1]
        ctx.save_for_backward(pred32, target32, filtered)
        ctx.cfg = (l1_weight, ssim_weight, window, n, pred.dtype)
        ctx.mark_non_differentiable(l1, ssim_loss)
        return l1_weight * l1 + ssim_weight * ssim_loss, l1, ssim_loss

    @staticmethod
    def backward(ctx, g_total, g_l1, g_ssim):
        pred32, target32, filtered = ctx.saved_tensors
        l1_weight, ssim_weight, window, n, dtype = ctx.cfg
        up = None if g_total is None else g_total.detach().float().reshape(1).contiguous()     # stays on the device: no host sync
        dpred = K.ssim_l1_bwd(pred32, target32, filtered, -ssim_weight / n, l1_weight / n, window, upstream=up)
        return dpred.to(dtype), None, None, None, None


def compute_ssim_loss(pred, target, window_size=11):
    """reference: model_direct.py:88-107"""
    return _DirectLoss.apply(pred, target, 0.0, 1.0, window_size)[0]


class DirectRegressionLoss(nn.Module):
    """reference: model_direct.py:110-131"""

    def __init__(self, l1_weight=1.0, ssim_weight=0.5):
        super().__init__()
        self.l1_weight = l1_weight
        self.ssim_weight = ssim_weight

    def forward(self, pred, target):
        total, l1, ssim_loss = _DirectLoss.apply(pred, target, float(self.l1_weight), float(self.ssim_weight), 11)
        return {"total_loss": total, "l1_loss": l1, "ssim_loss": ssim_loss}


# ----------------------------------------------------------------------------------------------------------------------------
# Stage 2-3 loss terms of the progressive cascade: direct_regression/progressive_cascade/loss_multiscale.py.
# Same class names, constructor arguments, forward signatures and returned dict keys; the arithmetic runs in csrc/hvc_loss.cu and
# csrc/hvc_loss_multiscale.cu (fp32, double accumulators; inputs are cast to fp32 like the reference's `.float()` in the TV term).
# ----------------------------------------------------------------------------------------------------------------------------
def _up(g):
    return None if g is None else g.detach().float().reshape(1).contiguous()


class SSIMLoss(nn.Module):
    """reference: loss_multiscale.py:18-51 (window = min(window_size, D, H, W))"""

    def __init__(self, window_size=11, channel=1):
        super().__init__()
        self.window_size = window_size
        self.channel = channel

    def forward(self, pred, target):
        window = min(self.window_size, pred.shape[2], pred.shape[3], pred.shape[4])
        return _DirectLoss.apply(pred, target, 0.0, 1.0, window)[0]


class _TV(Function):
    @staticmethod
    def forward(ctx, pred, target, eps):
        p32 = pred.float().contiguous()
        dims = K._vol_dims(p32)
        sums_t = None
        if target is not None:
            sums_t = K.tv_sums(target.detach().float().contiguous(), eps)
        loss, coef = K.tv_finalize(K.tv_sums(p32, eps), sums_t, dims)
        ctx.save_for_backward(p32, coef)
        ctx.cfg = (eps, pred.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        p32, coef = ctx.saved_tensors
        eps, dtype = ctx.cfg
        return K.tv_bwd(p32, eps, coef, _up(g)).to(dtype), None, None


class TotalVariationLoss(nn.Module):
    """reference: loss_multiscale.py:140-188 -- clamp(tv(pred), 0, 100), or |tv(pred) - tv(target)| when a target is given."""

    def __init__(self, eps=1e-8):
        super().__init__()
        self.eps = eps

    def forward(self, pred_volume, target_volume=None):
        return _TV.apply(pred_volume, target_volume, float(self.eps))


class _Freq(Function):
    @staticmethod
    def forward(ctx, pred, target, high_freq_weight):
        # the transform is the library FFT (cuFFT), exactly the call the reference makes (:206-207); magnitudes, mask, sums: hvc_freq_l1_*
        sp = torch.fft.fftn(pred.float(), dim=(-3, -2, -1)).contiguous()
        st = torch.fft.fftn(target.detach().float(), dim=(-3, -2, -1)).contiguous()
        sums = K.freq_l1_sums(sp, st)
        n = sp.numel()
        ctx.save_for_backward(sp, st)
        ctx.cfg = (high_freq_weight, n, pred.dtype)
        return ((sums[0] + high_freq_weight * sums[1]) / n).float()

    @staticmethod
    def backward(ctx, g):
        sp, st = ctx.saved_tensors
        hfw, n, dtype = ctx.cfg
        gs = K.freq_l1_bwd(sp, st, 1.0 / n, hfw / n, _up(g))
        # adjoint of the unnormalised forward DFT restricted to real inputs: Re(sum_k G_k e^{+i theta}) = real(ifftn(G)) * N
        return torch.fft.ifftn(gs, dim=(-3, -2, -1), norm="forward").real.to(dtype), None, None


class FrequencyLoss(nn.Module):
    """reference: loss_multiscale.py:191-236 (L1 between FFT magnitudes; low band + high_freq_weight * high band, the mask taken on the
    unshifted spectrum as the reference builds it)."""

    def __init__(self, high_freq_weight=2.0):
        super().__init__()
        self.high_freq_weight = high_freq_weight

    def forward(self, pred_volume, target_volume):
        return _Freq.apply(pred_volume, target_volume, float(self.high_freq_weight))


class _DRR(Function):
    @staticmethod
    def forward(ctx, pred, xray_ap, xray_lat, img_size):
        p32 = pred.float().contiguous()
        B, D, H, W = K._vol_dims(p32)
        ap, lat = K.proj_mean_fwd(p32)
        size = (1, img_size, img_size)
        drr_ap = K.interp3d_fwd(ap, B, (1, H, W), size, False)          # F.interpolate(mode="bilinear", align_corners=False), :267-269
        drr_lat = K.interp3d_fwd(lat, B, (1, D, H), size, False)
        xa, xl = xray_ap.detach().float().contiguous(), xray_lat.detach().float().contiguous()
        n = drr_ap.numel()
        s = (K.l1_sum(drr_ap, xa) + K.l1_sum(drr_lat, xl)) / (2.0 * n)
        ctx.save_for_backward(drr_ap, drr_lat, xa, xl)
        ctx.cfg = (tuple(p32.shape), (B, D, H, W), img_size, n, pred.dtype)
        return s.float().reshape(())

    @staticmethod
    def backward(ctx, g):
        drr_ap, drr_lat, xa, xl = ctx.saved_tensors
        shape, (B, D, H, W), img_size, n, dtype = ctx.cfg
        up = _up(g)
        size = (1, img_size, img_size)
        dap = K.interp3d_bwd(K.l1_bwd(drr_ap, xa, 0.5 / n, up), B, (1, H, W), size, False)
        dlat = K.interp3d_bwd(K.l1_bwd(drr_lat, xl, 0.5 / n, up), B, (1, D, H), size, False)
        return K.proj_mean_bwd(dap, dlat, shape).to(dtype), None, None, None


class DRRReprojectionLoss(nn.Module):
    """reference: loss_multiscale.py:239-293 -- mean-intensity projections of the predicted volume along depth (AP) and width (lateral),
    resized to the X-ray resolution, L1 against the two input X-rays."""

    def __init__(self, img_size=512):
        super().__init__()
        self.img_size = img_size

    def generate_drr(self, ct_volume, view_angle=0):
        """(B, 1, D, H, W) -> (B, 1, img_size, img_size); no gradient (the loss path is the fused Function)."""
        with torch.no_grad():
            v = ct_volume.float().contiguous()
            B, D, H, W = K._vol_dims(v)
            ap, lat = K.proj_mean_fwd(v)
            src, grid = (ap, (1, H, W)) if view_angle == 0 else (lat, (1, D, H))
            return K.interp3d_fwd(src, B, grid, (1, self.img_size, self.img_size), False).view(B, 1, self.img_size, self.img_size)

    def forward(self, pred_volume, input_xrays):
        return _DRR.apply(pred_volume, input_xrays[:, 0], input_xrays[:, 1], int(self.img_size))


class TriPlanarVGGLoss(nn.Module):
    """reference: loss_multiscale.py:54-137 -- L1 between VGG16 relu1_2 / relu2_2 / relu3_3 features of the three central slices
    (axial, sagittal, coronal) of prediction and target, each slice mapped to [0, 1] and replicated to RGB; mean over the planes.

    The reference constructor downloads torchvision's ImageNet VGG16; here the weights are an argument, so the loss also works on a
    machine without network access: `weights` = a torchvision VGG module, its state_dict ('features.N.weight' / '.bias', N in
    0 2 5 7 10 12 14), or a path to one saved with torch.save.  With weights=None the torchvision ImageNet weights are looked up exactly
    as the reference does (local hub cache or download) and the failure, if any, is raised.

    Same values, less work: the three taps come from ONE pass (the reference runs features[:4], [:9] and [:16] from the input each,
    recomputing the shared prefix), the RGB replication is folded into the first convolution (its weights summed over the input
    channels), all same-shaped slices go through the network as one batch, and the target features are computed without autograd.
    The convolutions are cuDNN's (library calls outside the hot path)."""

    _CONVS = ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256))
    _TAPS = {2: 0, 7: 1, 14: 2}          # feature index of the convolution whose ReLU output is relu1_2 / relu2_2 / relu3_3
    _POOL_BEFORE = (5, 10)               # max-pool 2x2 (features[4], features[9]) sits in front of these convolutions

    def __init__(self, weights=None, layer_weights=(1.0, 1.0, 1.0)):
        super().__init__()
        if weights is None:
            from torchvision.models import vgg16, VGG16_Weights
            weights = vgg16(weights=VGG16_Weights.IMAGENET1K_V1)
        if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__"):
            weights = torch.load(weights, map_location="cpu", weights_only=True)
        if isinstance(weights, nn.Module):
            weights = weights.state_dict()
        for idx, cin, cout in self._CONVS:
            w, b = weights[f"features.{idx}.weight"].detach().float(), weights[f"features.{idx}.bias"].detach().float()
            assert w.shape == (cout, cin, 3, 3), f"features.{idx}.weight: {tuple(w.shape)}"
            if idx == 0:
                w = w.sum(1, keepdim=True)            # the same grey slice in all three input channels
            self.register_buffer(f"w{idx}", w.contiguous(), persistent=False)      # frozen (:78-80); not part of any checkpoint
            self.register_buffer(f"b{idx}", b.contiguous(), persistent=False)
        self.layer_weights = list(layer_weights)

    def extract_features(self, x):
        """x: (n, 1, h, w) in [0, 1] -> [relu1_2, relu2_2, relu3_3]"""
        feats = []
        for idx, _, _ in self._CONVS:
            if idx in self._POOL_BEFORE:
                x = torch.nn.functional.max_pool2d(x, 2, 2)
            x = torch.relu(torch.nn.functional.conv2d(x, getattr(self, f"w{idx}").to(x.dtype), getattr(self, f"b{idx}").to(x.dtype), padding=1))
            if idx in self._TAPS:
                feats.append(x)
        return feats

    def forward(self, pred_volume, target_volume):
        D, H, W = pred_volume.shape[2:]
        planes = lambda v: (v[:, :, D // 2, :, :], v[:, :, :, H // 2, :], v[:, :, :, :, W // 2])
        groups = {}                                    # slice shape -> (pred slices, target slices)
        for ps, ts in zip(planes(pred_volume), planes(target_volume)):
            g = groups.setdefault(tuple(ps.shape[2:]), ([], []))
            g[0].append(ps)
            g[1].append(ts)
        total = 0.0
        for ps, ts in groups.values():
            n = len(ps)
            fp = self.extract_features((torch.cat(ps, 0) + 1) / 2)
            with torch.no_grad():
                ft = self.extract_features((torch.cat(ts, 0).detach() + 1) / 2)
            for a, b, w in zip(fp, ft, self.layer_weights):
                total = total + (w * n) * torch.nn.functional.l1_loss(a, b)      # equal-sized slices: sum of the planes' means = n * mean of the batch
        return total / 3


class Stage1Loss(nn.Module):
    """reference: loss_multiscale.py:296-324 (L1 + SSIM; one fused pass)"""

    def __init__(self, l1_weight=1.0, ssim_weight=0.5):
        super().__init__()
        self.l1_weight = l1_weight
        self.ssim_weight = ssim_weight
        self.ssim_loss = SSIMLoss()

    def forward(self, pred, target):
        window = min(11, pred.shape[2], pred.shape[3], pred.shape[4])
        total, l1, ssim_loss = _DirectLoss.apply(pred, target, float(self.l1_weight), float(self.ssim_weight), window)
        return {"total_loss": total, "l1_loss": l1, "ssim_loss": ssim_loss}


class _StageLossBase(nn.Module):
    """Stage 2 / 3: L1 + SSIM (fused) + VGG + TV(pred, target) + frequency [+ DRR].  The reference builds TriPlanarVGGLoss (:54-137) from
    downloaded ImageNet weights inside its constructor (so Stage2Loss / MultiScaleLoss cannot be constructed offline); here the term is
    passed in as `vgg_loss=TriPlanarVGGLoss(weights)`.  Without one it is reported as 0 and leaves the total unchanged."""

    def __init__(self, l1_weight, ssim_weight, vgg_weight, tv_weight, freq_weight, vgg_loss=None):
        super().__init__()
        self.l1_weight, self.ssim_weight, self.vgg_weight = l1_weight, ssim_weight, vgg_weight
        self.tv_weight, self.freq_weight = tv_weight, freq_weight
        self.ssim_loss = SSIMLoss()
        self.vgg_loss = vgg_loss
        self.tv_loss = TotalVariationLoss()
        self.freq_loss = FrequencyLoss(high_freq_weight=2.0)

    def _common(self, pred, target):
        window = min(11, pred.shape[2], pred.shape[3], pred.shape[4])
        base, l1, ssim_loss = _DirectLoss.apply(pred, target, float(self.l1_weight), float(self.ssim_weight), window)
        tv = self.tv_loss(pred, target)
        fr = self.freq_loss(pred, target)
        vgg = self.vgg_loss(pred, target) if self.vgg_loss is not None else torch.zeros((), device=pred.device)
        total = base + self.vgg_weight * vgg + self.tv_weight * tv + self.freq_weight * fr
        return total, {"l1_loss": l1, "ssim_loss": ssim_loss, "vgg_loss": vgg, "tv_loss": tv, "freq_loss": fr}


class Stage2Loss(_StageLossBase):
    """reference: loss_multiscale.py:327-374"""

    def __init__(self, l1_weight=1.0, ssim_weight=0.5, vgg_weight=0.1, tv_weight=0.02, freq_weight=0.05, vgg_loss=None):
        super().__init__(l1_weight, ssim_weight, vgg_weight, tv_weight, freq_weight, vgg_loss)

    def forward(self, pred, target):
        total, parts = self._common(pred, target)
        return dict(total_loss=total, **parts)


class Stage3Loss(_StageLossBase):
    """reference: loss_multiscale.py:377-432 (+ drr_weight * DRR reprojection when the input X-rays are given)"""

    def __init__(self, l1_weight=1.0, ssim_weight=0.5, vgg_weight=0.1, tv_weight=0.03, freq_weight=0.07, drr_weight=0.3, vgg_loss=None):
        super().__init__(l1_weight, ssim_weight, vgg_weight, tv_weight, freq_weight, vgg_loss)
        self.drr_weight = drr_weight
        self.drr_loss = DRRReprojectionLoss()

    def forward(self, pred, target, input_xrays=None):
        total, parts = self._common(pred, target)
        out = dict(total_loss=total, **parts)
        if input_xrays is not None:
            drr = self.drr_loss(pred, input_xrays)
            out["total_loss"] = total + self.drr_weight * drr
            out["drr_loss"] = drr
        return out


class MultiScaleLoss(nn.Module):
    """reference: loss_multiscale.py:435-490 -- picks the stage's loss; same config dict and defaults."""

    def __init__(self, config=None, vgg_loss=None):
        super().__init__()
        if config is None:
            config = {"stage1": {"l1": 1.0, "ssim": 0.5},
                      "stage2": {"l1": 1.0, "ssim": 0.5, "vgg": 0.1, "tv": 0.02, "freq": 0.05},
                      "stage3": {"l1": 1.0, "ssim": 0.5, "vgg": 0.1, "tv": 0.03, "freq": 0.07, "drr": 0.3}}
        self.stage1_loss = Stage1Loss(l1_weight=config["stage1"]["l1"], ssim_weight=config["stage1"]["ssim"])
        self.stage2_loss = Stage2Loss(l1_weight=config["stage2"]["l1"], ssim_weight=config["stage2"]["ssim"], vgg_weight=config["stage2"]["vgg"],
                                      tv_weight=config["stage2"].get("tv", 0.02), freq_weight=config["stage2"].get("freq", 0.05), vgg_loss=vgg_loss)
        self.stage3_loss = Stage3Loss(l1_weight=config["stage3"]["l1"], ssim_weight=config["stage3"]["ssim"], vgg_weight=config["stage3"]["vgg"],
                                      tv_weight=config["stage3"].get("tv", 0.03), freq_weight=config["stage3"].get("freq", 0.07),
                                      drr_weight=config["stage3"]["drr"], vgg_loss=vgg_loss)

    def forward(self, pred, target, stage=1, input_xrays=None):
        if stage == 1:
            return self.stage1_loss(pred, target)
        if stage == 2:
            return self.stage2_loss(pred, target)
        if stage == 3:
            return self.stage3_loss(pred, target, input_xrays)
        raise ValueError(f"Invalid stage: {stage}. Must be 1, 2, or 3.")


def compute_psnr(pred, target):
    """reference: loss_multiscale.py:493-500 (data range 2.0); one sum-of-squares kernel, one value read back (it is a metric)."""
    import ctypes as C
    from . import _lib
    d = (pred.detach().float() - target.detach().float()).contiguous()
    K._need_cuda(d)
    acc = torch.zeros(1, device=d.device, dtype=torch.float64)
    _lib.check(_lib.lib().hvc_sumsq_f32(K._ptr(d), C.c_int64(d.numel()), K._ptr(acc), K._stream()), "hvc_sumsq_f32")
    mse = float(acc.item()) / d.numel()
    if mse == 0:
        return float("inf")
    import math
    return 20.0 * math.log10(2.0 / math.sqrt(mse))


def compute_ssim_metric(pred, target):
    """reference: loss_multiscale.py:503-525"""
    window = min(11, pred.shape[2], pred.shape[3], pred.shape[4])
    with torch.no_grad():
        return 1.0 - float(_DirectLoss.apply(pred, target, 0.0, 1.0, window)[0].item())
