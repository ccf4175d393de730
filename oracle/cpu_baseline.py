"""CPU baseline for bench.py: the reference algorithm (oracle port, plain torch on the host cores) timed on a
BOUNDED SAMPLE of the benchmark workload.

TEST/BENCH INFRASTRUCTURE ONLY -- see oracle/vit_oracle.py.  kind = "port": /root/reference does not exist on
the GPU box and the reference is pure Python with nothing to compile, so the oracle restatement (same ATen
calls, same order, materialised (h, rows, N) softmax exactly like vit_components.py:46-51) is what is timed.

Why a sample: at the 128^3 token grid (N = 32768) one sample needs 15.8 TFLOP forward+backward and the
reference materialises a 4 x 32768 x 32768 fp32 score tensor (17 GB, several copies with autograd) per
block -- minutes of CPU time and an OOM risk.  Every stage of the backbone except the K/V projection of
self-attention is separable over query rows, so the sample runs, per block,
    * LN1 + K/V projection of ALL N tokens                                       (timed in full)
    * LN/q-proj/self-attention (R query rows x N keys, materialised softmax)/proj,
      LN2/cross-attention (R rows x M context tokens)/proj, LN3/MLP on an R-row slab   (timed, scaled x N/R)
forward + backward through autograd, plus the conv embedding and the LN/proj/upsample head in full, and
reports   seconds per volume = depth * (t_kv + t_slab * N/R) + t_embed_head.
"""
import os
import time

import torch
import torch.nn.functional as F

from . import vit_oracle as O


def _block_sample(sd, pfx, cfg, x, ctx, cond, rows, drop=None):
    """Forward+backward of one block restricted to `rows` query rows (K/V of self-attention over all tokens).
    drop: None, or oracle.dropout_mask.TorchDropout = the six nn.Dropout(0.1) sites in train mode, as the reference trainers run."""
    C, H = cfg.voxel_dim, cfg.num_heads
    d = C // H
    N = x.shape[1]
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if k.startswith(pfx)}
    x = x.detach().clone().requires_grad_(True)
    t0 = time.perf_counter()
    shift_sa, scale_sa, gate_sa, shift_mlp, scale_mlp, gate_mlp = O.adaln(cond, params, pfx + "adaln.")

    def ln(t, name):
        return F.layer_norm(t, (C,), params[pfx + name + ".weight"], params[pfx + name + ".bias"], 1e-5)

    # --- K/V for every token (the only non-row-separable part)
    h_all = (1 + scale_sa) * ln(x, "norm1") + shift_sa
    w_qkv = params[pfx + "self_attn.qkv.weight"]
    kv = F.linear(h_all, w_qkv[C:]).reshape(1, N, 2, H, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    t_kv_f = time.perf_counter() - t0
    # --- the row slab
    t1 = time.perf_counter()
    xs = x[:, :rows]
    q = F.linear(h_all[:, :rows], w_qkv[:C]).reshape(1, rows, H, d).permute(0, 2, 1, 3)
    pm = drop.attn(0, 1, H, rows, N) if drop is not None else None
    o, _ = O.attention_core(q, k, v, d ** -0.5, pmask=pm)            # materialised (1,H,rows,N) softmax
    o = o.transpose(1, 2).reshape(1, rows, C)
    o = F.linear(o, params[pfx + "self_attn.proj.weight"], params[pfx + "self_attn.proj.bias"])
    if drop is not None:
        o = drop.apply_tokens(o)
    xs = xs + gate_sa * o
    xs = xs + O.cross_attention(ln(xs, "norm2"), ctx, params, pfx + "cross_attn.", H, drop=drop, site=2)
    hh = (1 + scale_mlp) * ln(xs, "norm3") + shift_mlp
    hh = F.gelu(F.linear(hh, params[pfx + "mlp.0.weight"], params[pfx + "mlp.0.bias"]))
    if drop is not None:
        hh = drop.apply_tokens(hh)
    hh = F.linear(hh, params[pfx + "mlp.3.weight"], params[pfx + "mlp.3.bias"])
    if drop is not None:
        hh = drop.apply_tokens(hh)
    xs = xs + gate_mlp * hh
    loss = xs.square().mean()
    t_slab_f = time.perf_counter() - t1
    t2 = time.perf_counter()
    loss.backward()
    t_b = time.perf_counter() - t2
    # backward time splits between the K/V projection (all tokens) and the slab in proportion to their flops
    f_kv = 2 * 2 * N * C * 2 * C          # dgrad + wgrad of the [N,C]x[C,2C] projection
    f_slab = 2 * (4 * rows * N * C + 4 * rows * ctx.shape[1] * C + 24 * rows * C * C + 4 * ctx.shape[1] * cfg.context_dim * C)
    t_kv = t_kv_f + t_b * f_kv / (f_kv + f_slab)
    t_slab = t_slab_f + t_b * f_slab / (f_kv + f_slab)
    return t_kv, t_slab


def _embed_head(sd, cfg, x_vol):
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()
              if k.startswith("voxel_embed") or k in ("pos_embed", "norm.weight", "norm.bias", "output_proj.weight",
                                                      "output_proj.bias")}
    C = cfg.voxel_dim
    D, H, W = cfg.volume_size
    Dd, Hd, Wd = cfg.downsampled_size
    t0 = time.perf_counter()
    t = O.voxel_embed(x_vol, params, "voxel_embed.", cfg).flatten(2).transpose(1, 2) + params["pos_embed"]
    t = F.layer_norm(t, (C,), params["norm.weight"], params["norm.bias"], 1e-5)
    t = F.linear(t, params["output_proj.weight"], params["output_proj.bias"])
    t = t.transpose(1, 2).reshape(1, 1, Dd, Hd, Wd)
    out = F.interpolate(t, size=(D, H, W), mode="trilinear", align_corners=True)
    out.abs().mean().backward()
    return time.perf_counter() - t0


class CpuBaseline:
    """Holds the sample's inputs so that repeated steps time only the compute."""

    def __init__(self, cfg: O.BackboneConfig, context_len: int, rows: int = 1024, threads: int = 0, seed: int = 1234,
                 train: bool = True):
        from .dropout_mask import TorchDropout
        self.cfg = cfg
        self.train = train
        self.drop = TorchDropout(0.1) if train else None
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.rows = min(rows, cfg.num_tokens)
        g = torch.Generator().manual_seed(seed)
        self.sd = O.init_state_dict(cfg, seed=0)
        N, C = cfg.num_tokens, cfg.voxel_dim
        self.x_tok = torch.randn(1, N, C, generator=g)
        self.ctx = torch.rand(1, context_len, cfg.context_dim, generator=g)
        self.cond = torch.randn(1, cfg.cond_dim, generator=g)
        self.x_vol = torch.randn(1, cfg.in_channels, *cfg.volume_size, generator=g) * 0.01
        self.t_embed_head = None

    def describe(self):
        c = self.cfg
        mode = "train mode (nn.Dropout(0.1) at the six sites per block, torch's bernoulli as in the reference)" if self.train else "dropout off"
        return (f"oracle port, fp32, B=1, {self.threads} threads, {mode}: one block fwd+bwd with K/V of all {c.num_tokens} tokens + "
                f"LN/q/self-attn({self.rows} query rows x {c.num_tokens} keys, materialised softmax)/cross-attn/MLP on a "
                f"{self.rows}-row slab, scaled x{c.num_tokens / self.rows:g} rows x{c.depth} blocks; conv embed + head in full")

    def step(self):
        """One bounded sample -> estimated seconds per volume (forward+backward of the whole backbone, B=1)."""
        c = self.cfg
        if self.t_embed_head is None:
            self.t_embed_head = _embed_head(self.sd, c, self.x_vol)
        t0 = time.perf_counter()
        t_kv, t_slab = _block_sample(self.sd, "blocks.0.", c, self.x_tok, self.ctx, self.cond, self.rows, self.drop)
        self.last_measured_s = time.perf_counter() - t0          # what this step really ran (the sample), not the estimate
        return c.depth * (t_kv + t_slab * c.num_tokens / self.rows) + self.t_embed_head

    def measured_fraction(self):
        """Share of one volume's algorithmic FLOPs that a sampled step executes (the estimate scales the rest)."""
        c = self.cfg
        return (self.rows / c.num_tokens) / c.depth


# --------------------------------------------------------------------------------------------------------------------------
# Config A0 (BASELINE.json configs[0]): the reference's own CPU-runnable case, run IN FULL -- nothing sampled,
# nothing extrapolated.  DirectCTRegression(**config_direct.json['model']) at 64^3 (direct_regression/config_direct.json:5-12:
# voxel_dim 256, vit_depth 4, num_heads 4, xray_feature_dim 512, 512^2 X-rays), batch 1, fp32, forward + DirectRegressionLoss
# (model_direct.py:110-131) + backward, in train() mode (dropout + BatchNorm batch statistics: what train_direct_4gpu.py:49-98
# runs) and with dropout off.
# --------------------------------------------------------------------------------------------------------------------------
class A0Full:
    def __init__(self, threads: int = 0, seed: int = 1234):
        from . import encoder_oracle as E
        self.E = E
        self.threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(self.threads)
        self.cfg = O.BackboneConfig(volume_size=(64, 64, 64), in_channels=1, voxel_dim=256, depth=4, num_heads=4, context_dim=512,
                                    cond_dim=1024)
        g = torch.Generator().manual_seed(seed)
        bsd = O.init_state_dict(self.cfg, seed=0)
        sd = {"vit_backbone." + k: v for k, v in bsd.items()}
        sd["initial_volume"] = torch.randn(1, 1, 64, 64, 64, generator=g) * 0.01

        def uni(shape, fan):
            return (torch.rand(shape, generator=g) * 2 - 1) / fan ** 0.5

        e = "xray_encoder."
        for ci, bi, cin, cout, k in ((0, 1, 1, 64, 7), (4, 5, 64, 128, 3), (8, 9, 128, 512, 3)):      # diagnostic_losses.py:81-93
            sd[f"{e}encoder.{ci}.weight"] = uni((cout, cin, k, k), cin * k * k)
            sd[f"{e}encoder.{ci}.bias"] = uni((cout,), cin * k * k)
            sd[f"{e}encoder.{bi}.weight"] = torch.ones(cout)
            sd[f"{e}encoder.{bi}.bias"] = torch.zeros(cout)
            sd[f"{e}encoder.{bi}.running_mean"] = torch.zeros(cout)
            sd[f"{e}encoder.{bi}.running_var"] = torch.ones(cout)
        for name, fo, fi in (("time_mlp.0", 512, 256), ("time_mlp.2", 1024, 512), ("to_cond", 1024, 512)):
            sd[f"{e}{name}.weight"] = uni((fo, fi), fi)
            sd[f"{e}{name}.bias"] = uni((fo,), fi)
        self.sd = sd
        self.xrays = torch.rand(1, 2, 1, 512, 512, generator=g) * 2 - 1
        self.target = torch.rand(1, 1, 64, 64, 64, generator=g) * 2 - 1

    def describe(self, train):
        return (f"config A0 in full: oracle port of DirectCTRegression(config_direct.json) 64^3, B=1, fp32, {self.threads} threads, "
                f"forward + DirectRegressionLoss + backward, {'train() (dropout 0.1 + BatchNorm batch statistics)' if train else 'dropout off'}; "
                f"nothing sampled or extrapolated")

    def step(self, train: bool):
        """One full training step (no optimizer: the reference's AdamW adds ~15 M-parameter elementwise work) -> seconds."""
        from .dropout_mask import TorchDropout
        params = {k: (v.detach().clone().requires_grad_(True) if "running" not in k else v) for k, v in self.sd.items()}
        t0 = time.perf_counter()
        B = 1
        dummy_t = torch.zeros(B, 256)
        _, cond, feats = self.E.xray_conditioning(self.xrays, dummy_t, params, "xray_encoder.", True)
        bsd = {k[len("vit_backbone."):]: v for k, v in params.items() if k.startswith("vit_backbone.")}
        y = O.backbone(params["initial_volume"].expand(B, -1, -1, -1, -1), feats.flatten(2).transpose(1, 2), cond, bsd, self.cfg,
                       drop=TorchDropout(0.1) if train else None)
        loss = self.E.direct_regression_loss(y, self.target)["total_loss"]
        loss.backward()
        return time.perf_counter() - t0
