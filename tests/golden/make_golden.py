"""Generate golden fixtures from the REAL reference (run in the authoring container only).

    python tests/golden/make_golden.py

Imports ``models.vit_components`` / ``models.hybrid_vit_backbone`` from
``/root/reference`` (read-only), runs them on seeded inputs in eval mode
(dropout off -- bit parity is only defined without dropout, SURVEY.md finding 4)
with AdaLN re-randomised (finding 3), and stores weights, inputs, outputs and
gradients.  The fixtures travel to the GPU box; ``/root/reference`` does not.
"""
import json
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from models.vit_components import (  # noqa: E402
    AdaLNModulation, MultiHeadCrossAttention, MultiHeadSelfAttention, SinusoidalTimeEmbedding)
from models.hybrid_vit_backbone import HybridViT3D, HybridViTBlock3D  # noqa: E402


def randomise_adaln(mod, g):
    for name, p in mod.named_parameters():
        if "adaln.linear" in name or name.startswith("linear."):
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def grads_of(mod, inputs, out, r):
    params = [p for p in mod.parameters()]
    names = [n for n, _ in mod.named_parameters()]
    leaves = [t for t in inputs if t.requires_grad]
    gs = torch.autograd.grad((out * r).sum(), params + leaves, allow_unused=True)
    pg = {n: (g if g is not None else torch.zeros_like(p)) for n, g, p in zip(names, gs, params)}
    ig = [g for g in gs[len(params):]]
    return pg, ig


def components():
    g = torch.Generator().manual_seed(1234)
    out = {}

    torch.manual_seed(0)
    sa = MultiHeadSelfAttention(32, num_heads=4).eval()
    x = torch.randn(2, 48, 32, generator=g, requires_grad=True)
    y = sa(x)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(sa, [x], y, r)
    out["self_attn"] = dict(sd=sa.state_dict(), x=x.detach(), y=y.detach(), r=r, pgrad=pg, xgrad=ig[0],
                            num_heads=4)

    ca = MultiHeadCrossAttention(32, 40, num_heads=4, store_attention=True).eval()
    x = torch.randn(2, 48, 32, generator=g, requires_grad=True)
    ctx = torch.randn(2, 24, 40, generator=g, requires_grad=True)
    y = ca(x, ctx)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(ca, [x, ctx], y, r)
    out["cross_attn"] = dict(sd=ca.state_dict(), x=x.detach(), ctx=ctx.detach(), y=y.detach(), r=r,
                             pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], probs=ca.attention_weights.clone(),
                             num_heads=4)

    ad = AdaLNModulation(32, 48)
    zero_out = [t.detach().clone() for t in ad(None, torch.randn(3, 48, generator=g))]
    randomise_adaln(ad, g)
    cond = torch.randn(3, 48, generator=g)
    out["adaln"] = dict(sd=ad.state_dict(), cond=cond, chunks=[t.detach() for t in ad(None, cond)],
                        zero_init_max=max(float(t.abs().max()) for t in zero_out))

    te = SinusoidalTimeEmbedding(64)
    t = torch.tensor([0.0, 1.0, 17.0, 999.0])
    out["time_embed"] = dict(t=t, y=te(t))

    for name, prev in (("block", False), ("block_prev", True)):
        blk = HybridViTBlock3D(32, num_heads=4, context_dim=40, cond_dim=48, use_prev_stage=prev,
                               return_attention=prev).eval()
        randomise_adaln(blk, g)
        x = torch.randn(2, 48, 32, generator=g, requires_grad=True)
        ctx = torch.randn(2, 24, 40, generator=g, requires_grad=True)
        cond = torch.randn(2, 48, generator=g, requires_grad=True)
        res = blk(x, ctx, cond, None)
        attn_map = None
        if prev:
            res, attn_map = res
        r = torch.randn(res.shape, generator=g)
        pg, ig = grads_of(blk, [x, ctx, cond], res, r)
        out[name] = dict(sd=blk.state_dict(), x=x.detach(), ctx=ctx.detach(), cond=cond.detach(),
                         y=res.detach(), r=r, pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], condgrad=ig[2],
                         attn_map=attn_map, num_heads=4, use_prev_stage=prev)
    torch.save(out, os.path.join(HERE, "components.pt"))


def components_d64():
    """head_dim 64 cases (embed 128 / 2 heads) sized for the CUDA parity tests: ragged N, M."""
    g = torch.Generator().manual_seed(4321)
    out = {}
    torch.manual_seed(1)
    sa = MultiHeadSelfAttention(128, num_heads=2).eval()
    x = torch.randn(2, 200, 128, generator=g, requires_grad=True)
    y = sa(x)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(sa, [x], y, r)
    out["self_attn"] = dict(sd=sa.state_dict(), x=x.detach(), y=y.detach(), r=r, pgrad=pg, xgrad=ig[0], num_heads=2)

    ca = MultiHeadCrossAttention(128, 40, num_heads=2).eval()
    x = torch.randn(2, 200, 128, generator=g, requires_grad=True)
    ctx = torch.randn(2, 72, 40, generator=g, requires_grad=True)
    y = ca(x, ctx)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(ca, [x, ctx], y, r)
    out["cross_attn"] = dict(sd=ca.state_dict(), x=x.detach(), ctx=ctx.detach(), y=y.detach(), r=r,
                             pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], num_heads=2)

    for name, prev in (("block", False), ("block_prev", True)):
        blk = HybridViTBlock3D(64, num_heads=1, context_dim=40, cond_dim=48, use_prev_stage=prev).eval()
        randomise_adaln(blk, g)
        x = torch.randn(2, 200, 64, generator=g, requires_grad=True)
        ctx = torch.randn(2, 72, 40, generator=g, requires_grad=True)
        cond = torch.randn(2, 48, generator=g, requires_grad=True)
        prev_e = torch.randn(2, 256, generator=g) if prev else None
        res = blk(x, ctx, cond, prev_e)
        r = torch.randn(res.shape, generator=g)
        pg, ig = grads_of(blk, [x, ctx, cond], res, r)
        out[name] = dict(sd=blk.state_dict(), x=x.detach(), ctx=ctx.detach(), cond=cond.detach(), prev=prev_e,
                         y=res.detach(), r=r, pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], condgrad=ig[2],
                         num_heads=1, use_prev_stage=prev)
    for c in out.values():      # gradients of the weights in fp16: halves the fixture, ample for cos/3e-2 checks
        c["pgrad"] = {k: v.half() for k, v in c["pgrad"].items()}
    torch.save(out, os.path.join(HERE, "components_d64.pt"))


def components_d32():
    """head_dim 32 cases (cascade stages 2/3: C=256, 8 heads) incl. the materialised attention map
    (store_attention / return_attention, vit_components.py:106-108, hybrid_vit_backbone.py:130-143)."""
    g = torch.Generator().manual_seed(9876)
    out = {}
    torch.manual_seed(2)
    sa = MultiHeadSelfAttention(64, num_heads=2).eval()
    x = torch.randn(2, 200, 64, generator=g, requires_grad=True)
    y = sa(x)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(sa, [x], y, r)
    out["self_attn"] = dict(sd=sa.state_dict(), x=x.detach(), y=y.detach(), r=r, pgrad=pg, xgrad=ig[0], num_heads=2)

    ca = MultiHeadCrossAttention(96, 40, num_heads=3, store_attention=True).eval()
    x = torch.randn(2, 200, 96, generator=g, requires_grad=True)
    ctx = torch.randn(2, 72, 40, generator=g, requires_grad=True)
    y = ca(x, ctx)
    r = torch.randn(y.shape, generator=g)
    pg, ig = grads_of(ca, [x, ctx], y, r)
    out["cross_attn"] = dict(sd=ca.state_dict(), x=x.detach(), ctx=ctx.detach(), y=y.detach(), r=r,
                             pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], probs=ca.attention_weights.clone().half(),
                             num_heads=3)

    blk = HybridViTBlock3D(64, num_heads=2, context_dim=40, cond_dim=48, return_attention=True).eval()
    randomise_adaln(blk, g)
    x = torch.randn(2, 200, 64, generator=g, requires_grad=True)
    ctx = torch.randn(2, 72, 40, generator=g, requires_grad=True)
    cond = torch.randn(2, 48, generator=g, requires_grad=True)
    res, attn_map = blk(x, ctx, cond, None)
    r = torch.randn(res.shape, generator=g)
    pg, ig = grads_of(blk, [x, ctx, cond], res, r)
    out["block_attn"] = dict(sd=blk.state_dict(), x=x.detach(), ctx=ctx.detach(), cond=cond.detach(), prev=None,
                             y=res.detach(), r=r, pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], condgrad=ig[2],
                             attn_map=attn_map.half(), num_heads=2, use_prev_stage=False)
    for c in out.values():
        c["pgrad"] = {k: v.half() for k, v in c["pgrad"].items()}
    torch.save(out, os.path.join(HERE, "components_d32.pt"))


BACKBONES = {
    # name: ctor kwargs, context_len, batch
    "vit_s2": (dict(volume_size=(32, 16, 16), in_channels=1, voxel_dim=32, depth=2, num_heads=2,
                    context_dim=24, cond_dim=48), 16, 2),
    "vit_s4_quirk": (dict(volume_size=(64, 16, 16), in_channels=8, voxel_dim=32, depth=1, num_heads=4,
                          context_dim=24, cond_dim=48, use_prev_stage=True), 12, 1),
    "vit_s1": (dict(volume_size=(16, 16, 8), in_channels=2, voxel_dim=32, depth=1, num_heads=1,
                    context_dim=16, cond_dim=32), 8, 1),
    # head_dim 64 cases for the CUDA parity tests
    "vit_d64": (dict(volume_size=(32, 16, 16), in_channels=1, voxel_dim=64, depth=2, num_heads=1,
                     context_dim=24, cond_dim=48), 40, 2),
    "vit_d64_h2": (dict(volume_size=(16, 8, 4), in_channels=1, voxel_dim=128, depth=1, num_heads=2,
                        context_dim=24, cond_dim=48), 20, 1),
    "vit_d64_quirk": (dict(volume_size=(64, 8, 8), in_channels=16, voxel_dim=64, depth=1, num_heads=1,
                           context_dim=24, cond_dim=48, use_prev_stage=True), 12, 1),
    # head_dim 32 cases (cascade stage 2/3 style: multi-channel input, 8 heads at C=256)
    "vit_d32": (dict(volume_size=(32, 16, 16), in_channels=4, voxel_dim=64, depth=2, num_heads=2,
                     context_dim=24, cond_dim=48), 40, 2),
    "vit_d32_h4": (dict(volume_size=(16, 16, 16), in_channels=2, voxel_dim=128, depth=1, num_heads=4,
                        context_dim=32, cond_dim=48), 24, 1),
}


def backbones():
    out = {}
    for idx, (name, (kw, M, B)) in enumerate(BACKBONES.items()):
        g = torch.Generator().manual_seed(100 + idx)
        torch.manual_seed(0)
        m = HybridViT3D(**kw).eval()
        randomise_adaln(m, g)
        x = torch.randn(B, kw["in_channels"], *kw["volume_size"], generator=g, requires_grad=True)
        ctx = torch.randn(B, M, kw["context_dim"], generator=g, requires_grad=True)
        cond = torch.randn(B, kw["cond_dim"], generator=g, requires_grad=True)
        prev = torch.randn(B, 256, generator=g) if kw.get("use_prev_stage") else None
        y = m(x, ctx, cond, prev)
        r = torch.randn(y.shape, generator=g)
        pg, ig = grads_of(m, [x, ctx, cond], y, r)
        out[name] = dict(kwargs=kw, sd=m.state_dict(), x=x.detach(), ctx=ctx.detach(), cond=cond.detach(),
                         prev=prev, y=y.detach(), r=r, pgrad=pg, xgrad=ig[0], ctxgrad=ig[1], condgrad=ig[2],
                         downsampled_size=tuple(m.downsampled_size))
        if "d64" in name or "d32" in name:
            out[name]["pgrad"] = {k: v.half() for k, v in out[name]["pgrad"].items()}
    torch.save({k: v for k, v in out.items() if "d64" not in k and "d32" not in k}, os.path.join(HERE, "backbones.pt"))
    torch.save({k: v for k, v in out.items() if "d64" in k}, os.path.join(HERE, "backbones_d64.pt"))
    torch.save({k: v for k, v in out.items() if "d32" in k}, os.path.join(HERE, "backbones_d32.pt"))


CTOR_CASES = [
    dict(volume_size=(64, 64, 64), in_channels=1, voxel_dim=256, depth=1, num_heads=4),
    dict(volume_size=(128, 128, 128), in_channels=1, voxel_dim=256, depth=1, num_heads=4),
    dict(volume_size=(128, 128, 128), in_channels=32, voxel_dim=256, depth=1, num_heads=8),
    dict(volume_size=(256, 256, 256), in_channels=32, voxel_dim=256, depth=1, num_heads=8),
    dict(volume_size=(64, 64, 64), in_channels=17, voxel_dim=384, depth=1, num_heads=6, use_prev_stage=True),
    dict(volume_size=(32, 32, 32), in_channels=1, voxel_dim=64, depth=1, num_heads=2),
    dict(volume_size=(16, 16, 16), in_channels=1, voxel_dim=64, depth=1, num_heads=2),
    dict(volume_size=(64, 64, 64), in_channels=64, voxel_dim=256, depth=1, num_heads=4),
    dict(volume_size=(48, 64, 80), in_channels=3, voxel_dim=64, depth=1, num_heads=2),
]


def ctor_table():
    rows = []
    for kw in CTOR_CASES:
        m = HybridViT3D(**kw)
        convs = []
        for i, layer in enumerate(m.voxel_embed):
            if isinstance(layer, torch.nn.Conv3d):
                convs.append([i, layer.in_channels, layer.out_channels, layer.stride[0]])
        rows.append(dict(kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                         downsampled_size=list(m.downsampled_size),
                         convs=convs,
                         shapes={k: list(v.shape) for k, v in m.state_dict().items()}))
    with open(os.path.join(HERE, "ctor_table.json"), "w") as f:
        json.dump(rows, f, indent=0)


if __name__ == "__main__":
    torch.set_num_threads(8)
    components()
    components_d64()
    components_d32()
    backbones()
    ctor_table()
    for fn in ("components.pt", "components_d64.pt", "components_d32.pt", "backbones.pt", "backbones_d64.pt",
               "backbones_d32.pt", "ctor_table.json"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
