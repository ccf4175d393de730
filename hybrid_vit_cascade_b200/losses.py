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
