"""hybrid_vit_cascade_b200 -- B200-native (sm_100a) drop-in for the 3D ViT backbone hot path of
kanadm12/Hybrid-ViT-Cascade (models/vit_components.py, models/hybrid_vit_backbone.py).

    from hybrid_vit_cascade_b200 import HybridViT3D            # same ctor / forward / state_dict

Host code is PyTorch (memory, streams, autograd glue); every computation on the path is a kernel of
``libhvc_sm100a.so`` (C ABI: include/hvc.h).  There is no CPU or eager fallback: without the built
library or off sm_100 the modules raise.
"""
from .vit_components import (AdaLNModulation, MultiHeadCrossAttention, MultiHeadSelfAttention,  # noqa: F401
                             SinusoidalTimeEmbedding, precision, set_dropout_policy, set_precision)
from .hybrid_vit_backbone import HybridViT3D, HybridViTBlock3D  # noqa: F401
from .xray_encoder import (DirectCTRegression, MultiScaleXrayEncoder, Stage1Base64, Stage2Refiner128, Stage3Refiner256, ProgressiveCascadeModel,  # noqa: F401
                           XrayConditioningModule)
from .losses import (DirectRegressionLoss, DRRReprojectionLoss, FrequencyLoss, MultiScaleLoss, SSIMLoss, Stage1Loss, Stage2Loss, Stage3Loss, TriPlanarVGGLoss,  # noqa: F401
                     TotalVariationLoss, compute_psnr, compute_ssim_loss, compute_ssim_metric)
from . import checkpoint  # noqa: F401
from .optim import FlatAdamW  # noqa: F401

__all__ = ["AdaLNModulation", "MultiHeadCrossAttention", "MultiHeadSelfAttention", "SinusoidalTimeEmbedding",
           "HybridViTBlock3D", "HybridViT3D", "XrayConditioningModule", "MultiScaleXrayEncoder", "DirectCTRegression", "Stage1Base64", "Stage2Refiner128", "Stage3Refiner256", "ProgressiveCascadeModel", "DirectRegressionLoss", "compute_ssim_loss", "SSIMLoss", "TotalVariationLoss", "FrequencyLoss", "DRRReprojectionLoss", "TriPlanarVGGLoss", "Stage1Loss", "Stage2Loss", "Stage3Loss", "MultiScaleLoss", "compute_psnr", "compute_ssim_metric", "checkpoint", "FlatAdamW", "set_dropout_policy", "set_precision",
           "precision"]
