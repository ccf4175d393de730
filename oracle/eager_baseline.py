"""GPU-eager baseline for bench.py: the reference algorithm (oracle port -- the same ATen calls in the same order as
models/vit_components.py / models/hybrid_vit_backbone.py) run ON THE B200 under ``torch.autocast('cuda', bfloat16)``, i.e. what
the reference trainers would do on this GPU with cuBLAS / ATen eager kernels (train_direct_4gpu.py:65 uses autocast; fp16 there,
bf16 here to match the kernels' operand type).  SURVEY.md section 0.1 / BASELINE.md section 3 name exactly this as the bar.

TEST/BENCH INFRASTRUCTURE ONLY (see oracle/vit_oracle.py): bench.py's ``--impl eager`` arm and the ``eager_b200`` key of the
default line time it BESIDE the hand-written kernels; nothing in the product imports it.

Two measurements, both forward + backward with the softmax materialised as the reference does (vit_components.py:46-51):
  * ``Direct64Step``   -- the full 64^3 direct-regression training step at batch 8 (BASELINE.json configs[1]): backbone,
                          DirectRegressionLoss, backward, clip_grad_norm_, AdamW(fused).  S is (8, 4, 4096, 4096).
  * ``Block32768``     -- ONE HybridViTBlock3D at the 32768-token grid of the headline 128^3 configuration, batch 1:
                          S is (1, h, 32768, 32768) -- 8.6 GB in bf16, 17 GB as the fp32 softmax autocast produces; more than one
                          block or sample of it at a time does not fit next to autograd's copies, which is why the reference
                          cannot train this configuration at all and the comparison is per block.
"""
import torch

from . import encoder_oracle as E
from . import vit_oracle as O
from .dropout_mask import TorchDropout


def _time(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


class Direct64Step:
    def __init__(self, dev, batch=8, train=True, seed=1234):
        self.cfg = O.BackboneConfig(volume_size=(64, 64, 64), in_channels=1, voxel_dim=256, depth=4, num_heads=4, context_dim=512,
                                    cond_dim=1024)
        self.sd = {k: v.to(dev).requires_grad_(True) for k, v in O.init_state_dict(self.cfg, seed=0).items()}
        self.initial_volume = (torch.randn(1, 1, 64, 64, 64, device=dev) * 0.01).requires_grad_(True)
        self.params = list(self.sd.values()) + [self.initial_volume]
        self.opt = torch.optim.AdamW(self.params, lr=1e-4, weight_decay=0.01, fused=True)
        g = torch.Generator(device=dev).manual_seed(seed)
        self.B = batch
        self.feat = torch.rand(batch, 512, 64, 64, device=dev, generator=g)
        self.cond = torch.randn(batch, 1024, device=dev, generator=g)
        self.target = torch.rand(batch, 1, 64, 64, 64, device=dev, generator=g) * 2 - 1
        self.drop = TorchDropout(0.1) if train else None

    def step(self):
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ctx = self.feat.flatten(2).transpose(1, 2)                                   # model_direct.py:80
            y = O.backbone(self.initial_volume.expand(self.B, -1, -1, -1, -1), ctx, self.cond, self.sd, self.cfg, drop=self.drop)
            loss = E.direct_regression_loss(y.float(), self.target)["total_loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0, foreach=True)                   # train_direct_4gpu.py:72-75
        self.opt.step()
        return loss

    def ms_per_step(self, steps=5, warmup=2):
        return _time(self.step, steps, warmup)


class Block32768:
    """One block, forward + backward of every parameter and of x, at N tokens (default 32768), batch 1."""

    def __init__(self, dev, heads=4, tokens=32768, ctx_tokens=4096, train=True, seed=1234):
        self.cfg = O.BackboneConfig(volume_size=(128, 128, 128), in_channels=1, voxel_dim=256, depth=1, num_heads=heads, context_dim=512,
                                    cond_dim=1024, token_grid="conv")
        sd = O.init_state_dict(self.cfg, seed=0)
        self.sd = {k: v.to(dev).requires_grad_(True) for k, v in sd.items() if k.startswith("blocks.0.")}
        g = torch.Generator(device=dev).manual_seed(seed)
        self.x = torch.randn(1, tokens, 256, device=dev, generator=g).requires_grad_(True)
        self.ctx = torch.rand(1, ctx_tokens, 512, device=dev, generator=g)
        self.cond = torch.randn(1, 1024, device=dev, generator=g)
        self.r = torch.randn(1, tokens, 256, device=dev, generator=g)
        self.heads = heads
        self.drop = TorchDropout(0.1) if train else None

    def step(self):
        for v in self.sd.values():
            v.grad = None
        self.x.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = O.block(self.x, self.ctx, self.cond, self.sd, "blocks.0.", self.heads, drop=self.drop)
        (y.float() * self.r).sum().backward()

    def ms_per_step(self, steps=3, warmup=1):
        return _time(self.step, steps, warmup)
