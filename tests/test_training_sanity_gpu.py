"""End-to-end sanity on the GPU: the whole stack (encoder -> backbone -> loss -> backward -> buckets -> optimizer -> LR scheduler) overfits a
fixed batch, with either optimizer backend, in train mode with the reference's dropout active; a cascade-style multi-term loss trains too."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _train(backend, steps=40, loss_kind="direct"):
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200.dp import GradientBuckets
    torch.manual_seed(0)
    m = hvc.DirectCTRegression(volume_size=(32, 32, 32), xray_img_size=64, voxel_dim=64, vit_depth=2, num_heads=1, xray_feature_dim=64).cuda().train()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaln.linear" in n:
                p.normal_(0, 0.02)
    params = list(m.parameters())
    gb = GradientBuckets(params)
    if backend == "flat":
        opt = hvc.FlatAdamW(gb, lr=2e-3, weight_decay=0.01, max_grad_norm=1.0, params=params)
    else:
        opt = torch.optim.AdamW(params, lr=2e-3, weight_decay=0.01)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=steps)
    g = torch.Generator(device="cuda").manual_seed(1)
    xr = torch.rand(4, 2, 1, 64, 64, device="cuda", generator=g) * 2 - 1
    zz = torch.linspace(-1, 1, 32, device="cuda")
    tgt = (zz[None, None, :, None, None] * zz[None, None, None, :, None] + 0.3 * zz[None, None, None, None, :]).expand(4, 1, 32, 32, 32).contiguous()
    crit = hvc.DirectRegressionLoss(1.0, 0.5) if loss_kind == "direct" else hvc.Stage3Loss()
    crit_xr = torch.rand(4, 2, 1, 512, 512, device="cuda", generator=g) * 2 - 1 if loss_kind != "direct" else None
    losses = []
    for _ in range(steps):
        gb.reset()
        out = m(xr)
        d = crit(out, tgt) if loss_kind == "direct" else crit(out, tgt, crit_xr)
        d["total_loss"].backward()
        gb.finish()
        if backend != "flat":
            torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        sched.step()
        losses.append(float(d["total_loss"].detach()))
    assert all(torch.isfinite(torch.tensor(losses)))
    return losses, opt


@pytest.mark.parametrize("backend", ["torch", "flat"])
def test_overfits_a_fixed_batch(backend):
    losses, opt = _train(backend)
    first, last = sum(losses[:3]) / 3, sum(losses[-3:]) / 3
    assert last < 0.75 * first, (first, last)
    assert opt.param_groups[0]["lr"] < 1e-4                      # the cosine schedule drove the learning rate of either backend down


def test_cascade_loss_trains():
    losses, _ = _train("flat", steps=25, loss_kind="stage3")
    assert sum(losses[-3:]) < sum(losses[:3])
