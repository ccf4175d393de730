"""Drop-in replacement for the reference's ``XrayConditioningModule`` (models/diagnostic_losses.py:68-138) and
``DirectCTRegression`` (direct_regression/model_direct.py:15-85): the producer of ``context`` / ``cond`` in front of the
3D ViT backbone and the model that wires both together.

Same class names, constructor arguments, forward signatures and ``state_dict`` keys (``encoder.{0,1,4,5,8,9}.*`` incl. the
BatchNorm running buffers, ``time_mlp.{0,2}.*``, ``to_cond.*``; ``xray_encoder.*``, ``vit_backbone.*``, ``initial_volume``).
The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` children are parameter containers; forward and backward run on
libhvc_sm100a kernels: Conv2d = im2col + tcgen05 GEMM on channels-last activations (forward with a two-term operand split for
near-fp32 products, see ConvBnRelu), BatchNorm2d+ReLU = hvc_norm_act,
max-pool / view mean / pooling = hvc_encoder.cu, the small linears = the fp32 skinny-GEMM kernels.  The last stage emits the
(B, H'W', C) token layout the backbone's cross-attention reads, so ``features.flatten(2).transpose(1, 2)`` is a free view.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import kernels as K
from . import ops
from .hybrid_vit_backbone import HybridViT3D


_W3 = {}
ops._CACHES.append(_W3)


def _w3(p, pad_to):
    """[w0 | w0 | w1] two-term B-side operand of a conv weight viewed [Cout, Cin*k*k] (zero-padded to pad_to); cached like ops.w16."""
    t = ops._cache_get(_W3, p, pad_to)
    if t is None:
        src = p.detach().float().reshape(p.shape[0], -1)
        padded = torch.zeros(src.shape[0], pad_to, device=src.device, dtype=torch.float32)
        padded[:, :src.shape[1]] = src
        t = K.split3(padded, 3)
        ops._cache_put(_W3, p, pad_to, t)
    return t


class ConvBnRelu(Function):
    """relu(batch_norm(conv2d(x)))  -- diagnostic_losses.py:81-93 (one of the three stages).

    x: the conv input as a CUDA tensor addressed through `strides` = (sn, sc, sh, sw) element strides; returns channels-last
    [N*Ho*Wo, Cout] (bf16, or f32 for the last stage).  Train mode normalises with the batch statistics and updates the
    running buffers in place (momentum, unbiased variance) like nn.BatchNorm2d; eval mode uses the running buffers.
    """

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, bn_w, bn_b, run_mean, run_var, geom, training, momentum, out_f32):
        N, Cin, H, W, k, stride, pad, strides = geom
        Cout = conv_w.shape[0]
        Ho, Wo = K.conv2d_out(H, k, stride, pad), K.conv2d_out(W, k, stride, pad)
        M = N * Ho * Wo
        # Near-fp32 product on the bf16 tensor cores: the gather writes [c0 | c1 | c0] (two-term split of the f32 activation), the
        # weight is [w0 | w0 | w1], and ONE GEMM with K' = 3 Kp sums c0 w0 + c1 w0 + c0 w1 (~2^-16).  ReLU and max-pool are
        # discontinuous: with plain bf16 products ~0.2 % of their masks differ from the fp32 reference, enough to pull the gradient
        # cosine of the early layers to 0.99.  Block 0 of the patch matrix is the plain bf16 operand of the backward GEMMs.
        Kp = (Cin * k * k + 7) // 8 * 8
        cols3 = K.im2col2d_split(x, N, Cin, H, W, k, stride, pad, strides)                             # [M, 3 Kp] bf16
        z = K.gemm(cols3, _w3(conv_w, Kp), bias=conv_b, epilogue=K.EPI_F32)                            # [M, Cout] f32
        cols = cols3[:, :Kp]
        out_dtype = torch.float32 if out_f32 else torch.bfloat16
        if training:
            y, mean, rstd = K.norm_act_fwd(z, bn_w, bn_b, 1, M, Cout, Cout, K.ACT_RELU, out_dtype)
            with torch.no_grad():       # running buffers (C-length vectors): momentum update with the unbiased variance
                var = rstd.view(-1).pow(-2) - 1e-5
                run_mean.mul_(1.0 - momentum).add_(mean.view(-1), alpha=momentum)
                run_var.mul_(1.0 - momentum).add_(var * (M / max(M - 1, 1)), alpha=momentum)
        else:
            mean = run_mean.detach().float().view(1, Cout).contiguous()
            rstd = (run_var.detach().float() + 1e-5).rsqrt().view(1, Cout).contiguous()
            y, _, _ = K.norm_act_fwd(z, bn_w, bn_b, 1, M, Cout, Cout, K.ACT_RELU, out_dtype, mean=mean, rstd=rstd)
        ctx.save_for_backward(cols, z, mean, rstd, conv_w, bn_w, bn_b)
        ctx.meta = (geom, training, M, Cout, tuple(x.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        cols, z, mean, rstd, conv_w, bn_w, bn_b = ctx.saved_tensors
        (N, Cin, H, W, k, stride, pad, strides), training, M, Cout, x_shape = ctx.meta
        dy = dy.float().contiguous()
        dz, dbn_w, dbn_b = K.norm_act_bwd(dy, z, bn_w, bn_b, mean, rstd, 1, M, Cout, Cout, K.ACT_RELU, stats_frozen=not training)
        dz16 = K.cast_bf16(dz)
        dconv_b = K.colsum_bf16(dz16)
        dconv_w = ops._wgrad(dz16, cols)[:, :Cin * k * k].reshape(conv_w.shape)
        dx = None
        if ctx.needs_input_grad[0]:
            dcols = ops._dgrad(dz16, _w3(conv_w, cols.shape[1])[:, :cols.shape[1]])      # block 0 of [w0 | w0 | w1] = bf16(w)
            dx = torch.empty(x_shape, device=dy.device, dtype=torch.float32)     # same (contiguous) layout as the forward input
            K.col2im2d(dcols, N, Cin, H, W, k, stride, pad, dx, strides)
        return dx, dconv_w, dconv_b, dbn_w, dbn_b, None, None, None, None, None, None


class MaxPool(Function):
    """nn.MaxPool2d(k, stride, pad) on channels-last f32 [N*H*W, C]  -- diagnostic_losses.py:84,89."""

    @staticmethod
    def forward(ctx, x, N, H, W, C, k, stride, pad):
        y, arg = K.maxpool2d_fwd(x, N, H, W, C, k, stride, pad)
        ctx.save_for_backward(arg)
        ctx.meta = (N, H, W, C, k, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        arg, = ctx.saved_tensors
        N, H, W, C, k, stride, pad = ctx.meta
        return K.maxpool2d_bwd(dy.float().contiguous(), arg, N, H, W, C, k, stride, pad), None, None, None, None, None, None, None


class ViewMeanPool(Function):
    """features = mean over views; pooled = mean over pixels  -- diagnostic_losses.py:120-130."""

    @staticmethod
    def forward(ctx, x, B, V, P, C):
        feat, pooled = K.view_mean_fwd(x, B, V, P, C)
        ctx.meta = (B, V, P, C)
        return feat, pooled

    @staticmethod
    def backward(ctx, dfeat, dpooled):
        B, V, P, C = ctx.meta
        dfeat = None if dfeat is None else dfeat.float().contiguous()
        dpooled = None if dpooled is None else dpooled.float().contiguous()
        return K.view_mean_bwd(dfeat, dpooled, B, V, P, C), None, None, None, None


class Silu(Function):
    """nn.SiLU of the time MLP  -- diagnostic_losses.py:100."""

    @staticmethod
    def forward(ctx, x):
        x = x.float().contiguous()
        ctx.save_for_backward(x)
        return K.silu(x)

    @staticmethod
    def backward(ctx, dy):
        x, = ctx.saved_tensors
        return K.silu(x, dy.float().contiguous())


class XrayConditioningModule(nn.Module):
    """reference: models/diagnostic_losses.py:68-138"""

    def __init__(self, img_size: int = 512, in_channels: int = 1, embed_dim: int = 256, num_views: int = 1,
                 time_embed_dim: int = 256, cond_dim: int = 1024, share_view_weights: bool = True):
        super().__init__()
        self.num_views = num_views
        self.embed_dim = embed_dim
        self.cond_dim = cond_dim
        self.encoder = nn.Sequential(
            nn.Conv2d(in_channels, 64, kernel_size=7, stride=2, padding=3), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            nn.Conv2d(64, 128, kernel_size=3, padding=1), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.MaxPool2d(kernel_size=2, stride=2),
            nn.Conv2d(128, embed_dim, kernel_size=3, padding=1), nn.BatchNorm2d(embed_dim), nn.ReLU(inplace=True),
        )
        self.time_mlp = nn.Sequential(nn.Linear(time_embed_dim, time_embed_dim * 2), nn.SiLU(),
                                      nn.Linear(time_embed_dim * 2, cond_dim))
        self.to_cond = nn.Linear(embed_dim, cond_dim)

    def _stage(self, x, conv, bn, geom, out_f32):
        training = bn.training or not bn.track_running_stats
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        y = ConvBnRelu.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, geom, training,
                             momentum, out_f32)
        if training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        return y

    def forward(self, xrays: torch.Tensor, t: torch.Tensor):
        """xrays: (B, num_views, C, H, W); t: (B, time_embed_dim) -> (xray_context (B, cond_dim), time_xray_cond (B, cond_dim),
        xray_features_2d (B, embed_dim, H/8, W/8))"""
        B, V = xrays.shape[0], xrays.shape[1]
        if V == 1:                             # diagnostic_losses.py:120-128: the INPUT's view count decides, not self.num_views
            x = xrays[:, 0]
        else:
            x = xrays.reshape(B * V, *xrays.shape[2:])
        x = x.float().contiguous()
        N, Cin, H, W = x.shape
        enc = self.encoder
        y = self._stage(x, enc[0], enc[1], (N, Cin, H, W, 7, 2, 3, tuple(x.stride())), True)     # f32: the max-pool needs unrounded values
        H1, W1 = K.conv2d_out(H, 7, 2, 3), K.conv2d_out(W, 7, 2, 3)
        y = MaxPool.apply(y, N, H1, W1, 64, 3, 2, 1)
        H2, W2 = K.conv2d_out(H1, 3, 2, 1), K.conv2d_out(W1, 3, 2, 1)
        y = self._stage(y, enc[4], enc[5], (N, 64, H2, W2, 3, 1, 1, (H2 * W2 * 64, 1, W2 * 64, 64)), True)
        y = MaxPool.apply(y, N, H2, W2, 128, 2, 2, 0)
        H3, W3 = H2 // 2, W2 // 2
        y = self._stage(y, enc[8], enc[9], (N, 128, H3, W3, 3, 1, 1, (H3 * W3 * 128, 1, W3 * 128, 128)), True)   # f32 [N*P, E]
        E, P = self.embed_dim, H3 * W3
        feat, pooled = ViewMeanPool.apply(y, B, V, P, E)                       # (B*P, E), (B, E)
        features = feat.view(B, H3, W3, E).permute(0, 3, 1, 2)                # (B, E, H', W') view of the token-major buffer
        xray_context = ops.AdaLN.apply(pooled, self.to_cond.weight, self.to_cond.bias)
        h = ops.AdaLN.apply(t, self.time_mlp[0].weight, self.time_mlp[0].bias)
        time_embed = ops.AdaLN.apply(Silu.apply(h), self.time_mlp[2].weight, self.time_mlp[2].bias)
        return xray_context, time_embed + xray_context, features


class DirectCTRegression(nn.Module):
    """reference: direct_regression/model_direct.py:15-85 (X-rays -> CT volume, no diffusion)"""

    def __init__(self, volume_size=(64, 64, 64), xray_img_size=512, voxel_dim=256, vit_depth=4, num_heads=4,
                 xray_feature_dim=512, token_grid="reference"):
        super().__init__()
        self.volume_size = volume_size
        self.xray_encoder = XrayConditioningModule(img_size=xray_img_size, in_channels=1, embed_dim=xray_feature_dim,
                                                   num_views=2, time_embed_dim=256, cond_dim=1024, share_view_weights=False)
        self.vit_backbone = HybridViT3D(volume_size=volume_size, in_channels=1, voxel_dim=voxel_dim, depth=vit_depth,
                                        num_heads=num_heads, context_dim=xray_feature_dim, cond_dim=1024,
                                        use_prev_stage=False, token_grid=token_grid)
        D, H, W = volume_size
        self.initial_volume = nn.Parameter(torch.randn(1, 1, D, H, W) * 0.01)

    def forward(self, xrays):
        """xrays: (B, num_views, 1, H, W) -> predicted volume (B, 1, D, H, W)"""
        batch_size = xrays.shape[0]
        dummy_t = torch.zeros(batch_size, 256, device=xrays.device)
        xray_context, time_xray_cond, xray_features_2d = self.xray_encoder(xrays, dummy_t)
        x = self.initial_volume.expand(batch_size, -1, -1, -1, -1)
        return self.vit_backbone(x=x, context=xray_features_2d.flatten(2).transpose(1, 2), cond=time_xray_cond,
                                 prev_stage_embed=None)


class ConvGnGelu(Function):
    """gelu(group_norm(conv2d(x)))  -- one step of MultiScaleXrayEncoder.to_stage1 / to_stage2
    (direct_regression/progressive_cascade/model_progressive.py:38-52): Conv2d(k3, stride 2, pad 1) + GroupNorm(32) + GELU.
    x: f32 feature map addressed through `strides`; returns channels-last f32 [N*Ho*Wo, Cout].  All three pieces are smooth, so the
    convolution is a plain bf16 tensor-core GEMM."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, gn_w, gn_b, geom, groups):
        N, Cin, H, W, k, stride, pad, strides = geom
        Cout = conv_w.shape[0]
        Ho, Wo = K.conv2d_out(H, k, stride, pad), K.conv2d_out(W, k, stride, pad)
        cols = K.im2col2d(x, N, Cin, H, W, k, stride, pad, strides)
        z = K.gemm(cols, ops.w16(conv_w, pad_to=cols.shape[1]), bias=conv_b, epilogue=K.EPI_F32)
        y, mean, rstd = K.norm_act_fwd(z, gn_w, gn_b, N, Ho * Wo, Cout, groups, K.ACT_GELU_ERF, torch.float32)
        ctx.save_for_backward(cols, z, mean, rstd, conv_w, gn_w, gn_b)
        ctx.meta = (geom, groups, Ho * Wo, Cout, tuple(x.shape), tuple(x.stride()))
        return y

    @staticmethod
    def backward(ctx, dy):
        cols, z, mean, rstd, conv_w, gn_w, gn_b = ctx.saved_tensors
        (N, Cin, H, W, k, stride, pad, strides), groups, V, Cout, x_shape, x_strides = ctx.meta
        dz, dgn_w, dgn_b = K.norm_act_bwd(dy.float().contiguous(), z, gn_w, gn_b, mean, rstd, N, V, Cout, groups, K.ACT_GELU_ERF)
        dz16 = K.cast_bf16(dz)
        dconv_b = K.colsum_bf16(dz16)
        dconv_w = ops._wgrad(dz16, cols)[:, :Cin * k * k].reshape(conv_w.shape)
        dx = None
        if ctx.needs_input_grad[0]:
            dcols = ops._dgrad(dz16, ops.w16(conv_w, pad_to=cols.shape[1]))
            dx = torch.empty_strided(x_shape, x_strides, device=dy.device, dtype=torch.float32)   # same memory layout as the input view
            K.col2im2d(dcols, N, Cin, H, W, k, stride, pad, dx, strides)
        return dx, dconv_w, dconv_b, dgn_w, dgn_b, None, None


class MultiScaleXrayEncoder(nn.Module):
    """reference: direct_regression/progressive_cascade/model_progressive.py:16-83"""

    def __init__(self, img_size=512, in_channels=1, base_dim=512, num_views=2):
        super().__init__()
        self.xray_encoder = XrayConditioningModule(img_size=img_size, in_channels=in_channels, embed_dim=base_dim, num_views=num_views,
                                                   time_embed_dim=256, cond_dim=1024, share_view_weights=False)
        self.to_stage1 = nn.Sequential(
            nn.Conv2d(base_dim, base_dim, 3, stride=2, padding=1), nn.GroupNorm(32, base_dim), nn.GELU(),
            nn.Conv2d(base_dim, base_dim, 3, stride=2, padding=1), nn.GroupNorm(32, base_dim), nn.GELU())
        self.to_stage2 = nn.Sequential(
            nn.Conv2d(base_dim, base_dim, 3, stride=2, padding=1), nn.GroupNorm(32, base_dim), nn.GELU())

    @staticmethod
    def _down(feat, conv, gn):
        """feat: (B, C, H, W) with channels-last memory (what XrayConditioningModule returns) or any strides -> same, halved."""
        B, C, H, W = feat.shape
        f = feat.float()
        y = ConvGnGelu.apply(f, conv.weight, conv.bias, gn.weight, gn.bias, (B, C, H, W, 3, 2, 1, tuple(f.stride())), gn.num_groups)
        Ho, Wo = K.conv2d_out(H, 3, 2, 1), K.conv2d_out(W, 3, 2, 1)
        return y.view(B, Ho, Wo, conv.weight.shape[0]).permute(0, 3, 1, 2)

    def forward(self, xrays, stage=1):
        batch_size = xrays.shape[0]
        dummy_t = torch.zeros(batch_size, 256, device=xrays.device)
        xray_context, time_xray_cond, feats = self.xray_encoder(xrays, dummy_t)
        if stage == 1:
            feats = self._down(feats, self.to_stage1[0], self.to_stage1[1])
            feats = self._down(feats, self.to_stage1[3], self.to_stage1[4])
        elif stage == 2:
            feats = self._down(feats, self.to_stage2[0], self.to_stage2[1])
        return feats, time_xray_cond, xray_context


class Stage1Base64(nn.Module):
    """reference: direct_regression/progressive_cascade/model_progressive.py:86-150 (stage 1 of the cascade: 64^3 base volume)"""

    def __init__(self, volume_size=(64, 64, 64), xray_img_size=512, voxel_dim=256, vit_depth=4, num_heads=4, xray_feature_dim=512):
        super().__init__()
        self.volume_size = volume_size
        self.xray_encoder = MultiScaleXrayEncoder(img_size=xray_img_size, in_channels=1, base_dim=xray_feature_dim, num_views=2)
        self.vit_backbone = HybridViT3D(volume_size=volume_size, in_channels=1, voxel_dim=voxel_dim, depth=vit_depth,
                                        num_heads=num_heads, context_dim=xray_feature_dim, cond_dim=1024, use_prev_stage=False)
        D, H, W = volume_size
        self.initial_volume = nn.Parameter(torch.randn(1, 1, D, H, W) * 0.01)

    def forward(self, xrays):
        batch_size = xrays.shape[0]
        feats, time_xray_cond, _ = self.xray_encoder(xrays, stage=1)
        x = self.initial_volume.expand(batch_size, -1, -1, -1, -1)
        return self.vit_backbone(x=x, context=feats.flatten(2).transpose(1, 2), cond=time_xray_cond, prev_stage_embed=None)


# ------------------------------------------------------------------ cascade stage 2 (model_progressive.py:153-215)

class Interp3d(Function):
    """F.interpolate(x, size, mode="trilinear", align_corners=...) on (B, 1, D, H, W)."""

    @staticmethod
    def forward(ctx, x, size, align_corners):
        B, C, D, H, W = x.shape
        ctx.meta = (B * C, (D, H, W), tuple(size), bool(align_corners), x.dtype, tuple(x.shape))
        return K.interp3d_fwd(x.float().contiguous().view(B * C, D, H, W), B * C, (D, H, W), tuple(size), align_corners).view(B, C, *size)

    @staticmethod
    def backward(ctx, dy):
        n, grid, size, ac, dt, x_shape = ctx.meta
        dv = K.interp3d_bwd(dy.float().contiguous().view(n, *size), n, grid, size, ac)
        return dv.view(x_shape).to(dt), None, None


# patch-matrix budget of one Conv3d: above it the conv runs in depth slabs (one batch element, a range of output planes plus a
# one-plane halo each side), and the backward rebuilds each slab's patches instead of keeping them
CONV3D_COLS_BYTES = 8 << 30


# Conv3d with Cin % 64 == 0 runs as an implicit GEMM on the zero-padded channels-last volume (hvc_conv_taps): no patch matrix at all
CONV3D_IMPLICIT = True


def _pad_cl(src_cl, B, D, H, W, Cc, Cp):
    """(B, D, H, W, Cc) view (f32 or bf16) -> zero-padded bf16 (B, D+2, H+2, W+2, Cp) buffer, Cc <= Cp."""
    if src_cl.is_contiguous() and Cc % 8 == 0:
        return K.pad3d_cl(src_cl, B, D, H, W, Cc, Cp)
    out = torch.zeros(B, D + 2, H + 2, W + 2, Cp, device=src_cl.device, dtype=torch.bfloat16)       # other layouts: strided torch copy
    out[:, 1:-1, 1:-1, 1:-1, :Cc].copy_(src_cl)
    return out


def _conv3d_slabs(B, Cin, D, H, W):
    Kp = (Cin * 27 + 7) // 8 * 8
    plane = H * W * Kp * 2
    if B * D * plane <= CONV3D_COLS_BYTES:
        return None
    n = max(1, CONV3D_COLS_BYTES // plane - 2)
    return [(b, d0, min(D, d0 + n)) for b in range(B) for d0 in range(0, D, n)]


def _slab_cols(xf, b, d0, d1, Cin, D, H, W, tm):
    """Patch rows of output planes [d0, d1) of batch element b: im2col of the input planes [d0-1, d1+1) (clipped to the volume, where
    the kernel's own zero padding is the right one), minus the halo planes' rows."""
    lo, hi = max(d0 - 1, 0), min(d1 + 1, D)
    xs = xf[b:b + 1, :, lo:hi]
    cols = K.im2col3d(xs, 1, Cin, hi - lo, H, W, 1, tuple(xs.stride()), tap_major=tm)
    return cols[(d0 - lo) * H * W:(d1 - lo) * H * W], lo, hi


def _tap_major(x, Cin):
    """Channels-last inputs with whole 8-channel runs use the tap-major patch matrix (16-byte accesses in im2col and col2im)."""
    return Cin % 8 == 0 and x.stride(1) == 1 and all(s % 8 == 0 for i, s in enumerate(x.stride()) if i != 1)


class Conv3dGnGelu(Function):
    """gelu(group_norm(conv3d(x)))  -- Conv3d(k3, pad 1) + GroupNorm + GELU of the stage wrappers (model_progressive.py:170-172,
    :240-242, :260-265).  x: (B, Cin, D, H, W) any strides; returns the channels-last buffer viewed as (B, Cout, D, H, W), which is
    the layout the refiner ViT's voxel embedding (and the next Conv3dGnGelu) gathers from directly."""

    @staticmethod
    def forward(ctx, x, conv_w, conv_b, gn_w, gn_b, groups):
        B, Cin, D, H, W = x.shape
        Cout = conv_w.shape[0]
        V = D * H * W
        xf = x.float()
        tm = _tap_major(xf, Cin)
        Kp = (Cin * 27 + 7) // 8 * 8
        implicit = CONV3D_IMPLICIT and Cin % 64 == 0
        w_16 = ops.w16_taps(conv_w) if (tm or implicit) else ops.w16(conv_w, pad_to=Kp)
        slabs = None if implicit else _conv3d_slabs(B, Cin, D, H, W)
        if implicit:
            # rows = padded voxels; a tap (kd, kh, kw) is a shift of (kd-1)*(H+2)*(W+2) + (kh-1)*(W+2) + (kw-1) rows
            cols = _pad_cl(xf.permute(0, 2, 3, 4, 1), B, D, H, W, Cin, Cin)
            zpad = K.gemm(cols.view(-1, Cin), w_16, bias=conv_b, epilogue=K.EPI_F32, taps=(1, Cin, K.conv_tap_offsets(H, W)))
            z = K.unpad3d_cl(zpad, B, D, H, W, Cout).view(B * V, Cout)
            del zpad
        elif slabs is None:
            cols = K.im2col3d(xf, B, Cin, D, H, W, 1, tuple(xf.stride()), tap_major=tm)
            z = K.gemm(cols, w_16, bias=conv_b, epilogue=K.EPI_F32)                                             # [B*V, Cout]
        else:
            cols = None
            z = torch.empty(B * V, Cout, device=x.device, dtype=torch.float32)
            for (b, d0, d1) in slabs:
                rows, _, _ = _slab_cols(xf, b, d0, d1, Cin, D, H, W, tm)
                K.gemm(rows, w_16, bias=conv_b, epilogue=K.EPI_F32, out=z[(b * D + d0) * H * W:(b * D + d1) * H * W])
                del rows
        y, mean, rstd = K.norm_act_fwd(z, gn_w, gn_b, B, V, Cout, groups, K.ACT_GELU_ERF, torch.float32)
        ctx.save_for_backward(cols if slabs is None else xf, z, mean, rstd, conv_w, gn_w, gn_b)
        ctx.meta = (B, Cin, D, H, W, Cout, groups, tuple(x.shape), tuple(xf.stride()), slabs, tm, implicit)
        return y.view(B, D, H, W, Cout).permute(0, 4, 1, 2, 3)

    @staticmethod
    def backward(ctx, dy):
        cols, z, mean, rstd, conv_w, gn_w, gn_b = ctx.saved_tensors
        B, Cin, D, H, W, Cout, groups, x_shape, x_strides, slabs, tm, implicit = ctx.meta
        V, HW = D * H * W, H * W
        Kp = (Cin * 27 + 7) // 8 * 8
        dy_cl = dy.float().permute(0, 2, 3, 4, 1).contiguous().view(B * V, Cout)       # a no-op when dy already is channels-last
        dz, dgn_w, dgn_b = K.norm_act_bwd(dy_cl, z, gn_w, gn_b, mean, rstd, B, V, Cout, groups, K.ACT_GELU_ERF)
        dz16 = K.cast_bf16(dz)
        del dz
        dconv_b = K.colsum_bf16(dz16)
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty_strided(x_shape, x_strides, device=dy.device, dtype=torch.float32) if need_dx else None
        if implicit:
            # padded output gradient (zero rows at the padding positions, channels padded to a whole 64-wide k-block)
            Cp = (Cout + 63) // 64 * 64
            dzp = _pad_cl(dz16.view(B, D, H, W, Cout), B, D, H, W, Cout, Cp).view(-1, Cp)
            del dz16
            tiles = (27 * Cin + 127) // 128
            splits = max(1, min(dzp.shape[0] // 64, (16 * ops._sms(dy.device)) // tiles))     # short fp32 accumulation chains
            dw = K.gemm(dzp, cols.view(-1, Cin), a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=splits,
                        taps=(2, Cin, K.conv_tap_offsets(H, W)))[:Cout]
            if need_dx:
                dxp = K.gemm(dzp, ops.w16_taps_t(conv_w, Cp), epilogue=K.EPI_F32, taps=(1, Cp, K.conv_tap_offsets(H, W, -1)))
                dx_cl = dx.permute(0, 2, 3, 4, 1)
                if dx_cl.is_contiguous():
                    K.unpad3d_cl(dxp, B, D, H, W, Cin, out=dx_cl)
                else:
                    dx_cl.copy_(dxp.view(B, D + 2, H + 2, W + 2, Cin)[:, 1:-1, 1:-1, 1:-1])
                del dxp
            return dx, dw.view(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3).contiguous(), dconv_b, dgn_w, dgn_b, None
        w_16 = ops.w16_taps(conv_w) if tm else ops.w16(conv_w, pad_to=Kp)
        if slabs is None:
            dw = ops._wgrad(dz16, cols)
            if need_dx:
                K.col2im3d(ops._dgrad(dz16, w_16), B, Cin, D, H, W, 1, dx, x_strides, tap_major=tm)
        else:
            xf = cols                                                                   # the slab path saved the input instead
            dw = torch.zeros(Cout, Kp, device=dy.device, dtype=torch.float32)
            for (b, d0, d1) in slabs:
                rows, lo, hi = _slab_cols(xf, b, d0, d1, Cin, D, H, W, tm)
                T = (d1 - d0) * HW
                splits = max(1, min((T + 63) // 64, (4 * ops._sms(dy.device)) // max(((Cout + 127) // 128) * ((Kp + 127) // 128), 1)))
                K.gemm(dz16[(b * D + d0) * HW:(b * D + d1) * HW], rows, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC,
                       k_splits=splits, out=dw)
                del rows
                if need_dx:
                    # dx planes [d0, d1) gather from the patch gradients of output planes [d0-1, d1+1): col2im on that halo'd slab is
                    # exact for its interior planes, which are the ones copied out
                    dcols = ops._dgrad(dz16[(b * D + lo) * HW:(b * D + hi) * HW], w_16)
                    tmp = torch.empty(1, hi - lo, H, W, Cin, device=dy.device, dtype=torch.float32).permute(0, 4, 1, 2, 3)
                    K.col2im3d(dcols, 1, Cin, hi - lo, H, W, 1, tmp, tuple(tmp.stride()), tap_major=tm)
                    dx[b:b + 1, :, d0:d1].copy_(tmp[:, :, d0 - lo:d1 - lo])
                    del dcols, tmp
        if tm:
            dconv_w = dw.view(Cout, 3, 3, 3, Cin).permute(0, 4, 1, 2, 3).contiguous()
        else:
            dconv_w = dw[:, :Cin * 27].reshape(conv_w.shape)
        return dx, dconv_w, dconv_b, dgn_w, dgn_b, None


class ChanDot(Function):
    """Conv3d(C -> 1, kernel 1) on a channels-last activation (model_progressive.py:266).  y: (B, C, D, H, W) viewed from a
    channels-last buffer -> (B, 1, D, H, W)."""

    @staticmethod
    def forward(ctx, y, w, b):
        B, Cc, D, H, W = y.shape
        y2 = y.permute(0, 2, 3, 4, 1).contiguous().view(-1, Cc)                        # a view of the channels-last buffer
        w1 = w.reshape(-1).float().contiguous()
        out = K.chan_dot_fwd(y2, w1, b)
        ctx.save_for_backward(y2, w1)
        ctx.shape = (B, Cc, D, H, W, tuple(w.shape))
        return out.view(B, 1, D, H, W)

    @staticmethod
    def backward(ctx, dout):
        y2, w1 = ctx.saved_tensors
        B, Cc, D, H, W, w_shape = ctx.shape
        dy, dw, db = K.chan_dot_bwd(dout.float().contiguous().view(-1), y2, w1)
        return dy.view(B, D, H, W, Cc).permute(0, 4, 1, 2, 3), dw.view(w_shape), db


class Stage2Refiner128(nn.Module):
    """reference: direct_regression/progressive_cascade/model_progressive.py:153-215 (64^3 -> 128^3 refinement)"""

    def __init__(self, volume_size=(128, 128, 128), voxel_dim=256, vit_depth=6, num_heads=8, xray_feature_dim=512, token_grid="reference"):
        super().__init__()
        self.volume_size = volume_size
        self.upsample_from_64 = nn.Sequential(nn.Upsample(scale_factor=2, mode='trilinear', align_corners=False),
                                              nn.Conv3d(1, 32, 3, padding=1), nn.GroupNorm(8, 32), nn.GELU())
        self.vit_refiner = HybridViT3D(volume_size=volume_size, in_channels=32, voxel_dim=voxel_dim, depth=vit_depth,
                                       num_heads=num_heads, context_dim=xray_feature_dim, cond_dim=1024, use_prev_stage=False,
                                       token_grid=token_grid)
        self.residual_weight = nn.Parameter(torch.ones(1) * 0.5)

    def forward(self, volume_64, xray_features_2d, time_xray_cond):
        """volume_64: (B, 1, D/2, H/2, W/2); xray_features_2d: (B, C, h, w); time_xray_cond: (B, 1024) -> (B, 1, D, H, W)"""
        D2, H2, W2 = volume_64.shape[2:]
        up = Interp3d.apply(volume_64, (2 * D2, 2 * H2, 2 * W2), False)          # nn.Upsample(scale_factor=2, align_corners=False), :169
        conv, gn = self.upsample_from_64[1], self.upsample_from_64[2]
        x = Conv3dGnGelu.apply(up, conv.weight, conv.bias, gn.weight, gn.bias, gn.num_groups)
        refinement = self.vit_refiner(x=x, context=xray_features_2d.flatten(2).transpose(1, 2), cond=time_xray_cond,
                                      prev_stage_embed=None)
        # :211-213 -- F.interpolate(volume_64, size=volume_size, align_corners=False) is the same resize as `up` when volume_size is
        # twice the input (the only way the reference runs); the blend with the learned scalar is a two-op elementwise epilogue
        base = up if tuple(self.volume_size) == tuple(up.shape[2:]) else Interp3d.apply(volume_64, tuple(self.volume_size), False)
        return base + self.residual_weight * refinement


class Stage3Refiner256(nn.Module):
    """reference: direct_regression/progressive_cascade/model_progressive.py:218-315 (128^3 -> 256^3 refinement + high-frequency
    detail branch).  `use_gradient_checkpointing` is accepted for signature compatibility and changes nothing: the flash-style
    attention never stores the (B, h, N, M) probability matrices the reference checkpoints away (:286-293), and a B = 2 step fits
    the 180 GB of HBM without recomputation."""

    def __init__(self, volume_size=(256, 256, 256), voxel_dim=256, vit_depth=8, num_heads=8, xray_feature_dim=512,
                 use_gradient_checkpointing=True, token_grid="reference"):
        super().__init__()
        self.volume_size = volume_size
        self.use_gradient_checkpointing = use_gradient_checkpointing
        self.upsample_from_128 = nn.Sequential(nn.Upsample(scale_factor=2, mode='trilinear', align_corners=False),
                                               nn.Conv3d(1, 32, 3, padding=1), nn.GroupNorm(8, 32), nn.GELU())
        self.vit_refiner = HybridViT3D(volume_size=volume_size, in_channels=32, voxel_dim=voxel_dim, depth=vit_depth,
                                       num_heads=num_heads, context_dim=xray_feature_dim, cond_dim=1024, use_prev_stage=False,
                                       token_grid=token_grid)
        self.detail_enhancer = nn.Sequential(nn.Conv3d(1, 64, 3, padding=1), nn.GroupNorm(16, 64), nn.GELU(),
                                             nn.Conv3d(64, 32, 3, padding=1), nn.GroupNorm(8, 32), nn.GELU(), nn.Conv3d(32, 1, 1))
        self.residual_weight = nn.Parameter(torch.ones(1) * 0.5)
        self.detail_weight = nn.Parameter(torch.ones(1) * 0.3)

    def forward(self, volume_128, xray_features_2d, time_xray_cond):
        """volume_128: (B, 1, D/2, H/2, W/2); xray_features_2d: (B, C, h, w); time_xray_cond: (B, 1024) -> (B, 1, D, H, W)"""
        D2, H2, W2 = volume_128.shape[2:]
        up = Interp3d.apply(volume_128, (2 * D2, 2 * H2, 2 * W2), False)               # nn.Upsample(scale_factor=2), :239
        conv, gn = self.upsample_from_128[1], self.upsample_from_128[2]
        x = Conv3dGnGelu.apply(up, conv.weight, conv.bias, gn.weight, gn.bias, gn.num_groups)
        refinement = self._vit_forward(x, xray_features_2d, time_xray_cond)
        # :296-297 -- the same resize as `up` whenever volume_size is twice the input (the only way the reference runs)
        base = up if tuple(self.volume_size) == tuple(up.shape[2:]) else Interp3d.apply(volume_128, tuple(self.volume_size), False)
        de = self.detail_enhancer
        d = Conv3dGnGelu.apply(base, de[0].weight, de[0].bias, de[1].weight, de[1].bias, de[1].num_groups)     # :260-262
        d = Conv3dGnGelu.apply(d, de[3].weight, de[3].bias, de[4].weight, de[4].bias, de[4].num_groups)        # :263-265
        details = ChanDot.apply(d, de[6].weight, de[6].bias)                                                   # :266
        return base + self.residual_weight * refinement + self.detail_weight * details                        # :303-305

    def _vit_forward(self, x, xray_features_2d, time_xray_cond):
        return self.vit_refiner(x=x, context=xray_features_2d.flatten(2).transpose(1, 2), cond=time_xray_cond, prev_stage_embed=None)


class ProgressiveCascadeModel(nn.Module):
    """reference: direct_regression/progressive_cascade/model_progressive.py:318-432 (64^3 -> 128^3 -> 256^3; train stage by stage or
    end to end).  `stage2_token_grid`: the committed reference cannot run its stage 2 (HybridViT3D at 128^3 sizes pos_embed for 25^3
    tokens while the conv stack emits 32^3); "reference" keeps that behaviour (max_stage >= 2 raises the same
    shape error), "conv" or an int (16 = the author's recorded fix) makes the stage runnable."""

    def __init__(self, xray_img_size=512, xray_feature_dim=512, voxel_dim=256, use_gradient_checkpointing=True,
                 stage2_token_grid="reference"):
        super().__init__()
        self.xray_encoder = MultiScaleXrayEncoder(img_size=xray_img_size, in_channels=1, base_dim=xray_feature_dim, num_views=2)
        self.stage1 = Stage1Base64(volume_size=(64, 64, 64), xray_img_size=xray_img_size, voxel_dim=voxel_dim, vit_depth=4, num_heads=4,
                                   xray_feature_dim=xray_feature_dim)
        self.stage2 = Stage2Refiner128(volume_size=(128, 128, 128), voxel_dim=voxel_dim, vit_depth=6, num_heads=8,
                                       xray_feature_dim=xray_feature_dim, token_grid=stage2_token_grid)
        self.stage3 = Stage3Refiner256(volume_size=(256, 256, 256), voxel_dim=voxel_dim, vit_depth=8, num_heads=8,
                                       xray_feature_dim=xray_feature_dim, use_gradient_checkpointing=use_gradient_checkpointing)

    def forward(self, xrays, return_intermediate=False, max_stage=3):
        """xrays: (B, 2, 1, S, S) -> the volume of `max_stage`, or {'stage1': ..., 'stage2': ..., 'stage3': ...} up to it"""
        outputs = {}
        volume_64 = self.stage1(xrays)
        outputs['stage1'] = volume_64
        if max_stage == 1:
            return outputs if return_intermediate else volume_64
        feats2, cond, _ = self.xray_encoder(xrays, stage=2)
        volume_128 = self.stage2(volume_64, feats2, cond)
        outputs['stage2'] = volume_128
        if max_stage == 2:
            return outputs if return_intermediate else volume_128
        feats3, cond, _ = self.xray_encoder(xrays, stage=3)
        volume_256 = self.stage3(volume_128, feats3, cond)
        outputs['stage3'] = volume_256
        return outputs if return_intermediate else volume_256

    def _set_stage(self, stage, flag):
        for p in getattr(self, f"stage{stage}").parameters():
            p.requires_grad = flag

    def freeze_stage(self, stage):
        """reference :404-418"""
        if stage in (1, 2, 3):
            self._set_stage(stage, False)
            print(f"Stage {stage} ({(64, 128, 256)[stage - 1]}\u00b3) frozen")

    def unfreeze_stage(self, stage):
        """reference :420-434"""
        if stage in (1, 2, 3):
            self._set_stage(stage, True)
            print(f"Stage {stage} ({(64, 128, 256)[stage - 1]}\u00b3) unfrozen")
