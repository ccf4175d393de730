"""CPU: the round-2 oracle additions against outputs of the real reference (tests/golden/make_golden_r02.py), and the host logic of the
checkpoint module.  No GPU, no /root/reference at run time."""
import os
import warnings

import pytest
import torch

from oracle import encoder_oracle as E
from oracle import loss_oracle as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gold(name):
    return torch.load(os.path.join(ROOT, "tests", "golden", name), weights_only=False)


def _val_grad(fn, pred):
    p = pred.clone().requires_grad_(True)
    v = fn(p)
    v = v["total_loss"] if isinstance(v, dict) else v
    g, = torch.autograd.grad(v, p)
    return float(v.detach()), g


@pytest.mark.parametrize("name", ["ssim", "tv_pred_only", "tv_vs_target", "tv_vs_target_smooth", "freq", "drr", "stage1"])
def test_loss_oracle_matches_reference(name):
    c = _gold("r02_losses.pt")
    t, x = c["target"], c["xrays"]
    fns = {"ssim": lambda p: L.ssim_loss(p, t), "tv_pred_only": lambda p: L.total_variation_loss(p),
           "tv_vs_target": lambda p: L.total_variation_loss(p, t), "tv_vs_target_smooth": lambda p: L.total_variation_loss(p, t),
           "freq": lambda p: L.frequency_loss(p, t, 2.0), "drr": lambda p: L.drr_reprojection_loss(p, x, c["img_size"]),
           "stage1": lambda p: L.stage1_loss(p, t)}
    pred = c["smooth"] if name.endswith("smooth") else c["pred"]
    v, g = _val_grad(fns[name], pred)
    v_ref, g_ref = c[name]
    assert abs(v - v_ref) <= 1e-6 * max(1.0, abs(v_ref))
    assert float((g - g_ref).abs().max()) <= 1e-6 * float(g_ref.abs().max())


def test_loss_oracle_metrics_and_stage_totals():
    c = _gold("r02_losses.pt")
    p, t, x = c["pred"], c["target"], c["xrays"]
    assert abs(L.psnr(p, t) - c["psnr"]) < 1e-4
    assert abs((1 - float(L.ssim_loss(p, t))) - c["ssim_metric"]) < 1e-6
    # stage totals are the weighted sums of the pinned terms (loss_multiscale.py:359-363, :404-430); vgg is an argument
    s2 = L.stage2_loss(p, t, vgg=0.25)
    want = c["stage1"][0] + 0.1 * 0.25 + 0.02 * c["tv_vs_target"][0] + 0.05 * c["freq"][0]
    assert abs(float(s2["total_loss"]) - want) < 1e-5
    s3 = L.stage3_loss(p, t, x, vgg=0.0, img_size=c["img_size"])
    want3 = c["stage1"][0] + 0.03 * c["tv_vs_target"][0] + 0.07 * c["freq"][0] + 0.3 * c["drr"][0]
    assert abs(float(s3["total_loss"]) - want3) < 1e-5 and "drr_loss" in s3


def test_triplanar_vgg_oracle_and_module_match_reference():
    """TriPlanarVGGLoss (loss_multiscale.py:54-137) with seeded stand-in VGG16 weights: the oracle restatement AND the shipped module (plain
    torch convolutions: it runs on the CPU too) against the reference class; Stage2Loss / Stage3Loss totals of the reference = the oracle's
    weighted sums with that VGG term."""
    c, v = _gold("r02_losses.pt"), _gold("r02_vgg.pt")
    p, t, x = c["pred"], c["target"], c["xrays"]
    sd = L.vgg16_features_state(v["vgg_seed"])
    val, g = _val_grad(lambda q: L.triplanar_vgg_loss(q, t, sd), p)
    assert abs(val - v["vgg"][0]) <= 1e-6 and float((g - v["vgg"][1]).abs().max()) <= 1e-5 * float(v["vgg"][1].abs().max())
    from hybrid_vit_cascade_b200.losses import TriPlanarVGGLoss
    mod = TriPlanarVGGLoss(weights=sd)
    assert not list(mod.parameters()) and not mod.state_dict()           # frozen, and absent from checkpoints
    val, g = _val_grad(lambda q: mod(q, t), p)
    assert abs(val - v["vgg"][0]) <= 2e-6 and float((g - v["vgg"][1]).abs().max()) <= 2e-5 * float(v["vgg"][1].abs().max())
    cube = torch.rand(1, 1, 16, 16, 16, generator=torch.Generator().manual_seed(3)) * 2 - 1          # the three slices as ONE batch
    tc = torch.rand(1, 1, 16, 16, 16, generator=torch.Generator().manual_seed(4)) * 2 - 1
    assert abs(float(mod(cube, tc)) - float(L.triplanar_vgg_loss(cube, tc, sd))) <= 2e-6
    s2, g2 = _val_grad(lambda q: L.stage2_loss(q, t, vgg=L.triplanar_vgg_loss(q, t, sd)), p)
    assert abs(s2 - v["stage2"][0]) <= 2e-6 * v["stage2"][0] and float((g2 - v["stage2"][1]).abs().max()) <= 1e-5 * float(v["stage2"][1].abs().max())
    s3, g3 = _val_grad(lambda q: L.stage3_loss(q, t, x, vgg=L.triplanar_vgg_loss(q, t, sd), img_size=c["img_size"]), p)
    assert abs(s3 - v["stage3"][0]) <= 2e-6 * v["stage3"][0] and float((g3 - v["stage3"][1]).abs().max()) <= 1e-5 * float(v["stage3"][1].abs().max())


def test_encoder_oracle_two_views_into_default_constructor():
    """XrayConditioningModule() defaults to num_views=1 but decides on the INPUT's view count (diagnostic_losses.py:118-125)."""
    c = _gold("r02_views.pt")
    ctx, cond, feats = E.xray_conditioning(c["xrays"], c["t"], c["sd"], "", training=True)
    for a, b in ((ctx, c["ctx"]), (cond, c["cond"]), (feats, c["feats"])):
        assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())


def test_checkpoint_layout_and_torch_resume_on_cpu():
    """The oracle-built checkpoint has the reference trainer's layout (recorded from the real reference), and the package's
    load_checkpoint resumes a torch.optim.AdamW from it (constructing the drop-in modules needs no GPU; running them does)."""
    warnings.filterwarnings("ignore")
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200 import checkpoint as CK
    from tools.ckpt_roundtrip import CONFIG, OracleTrainer, check_layout
    tr = OracleTrainer()
    tr.step()
    ck = tr.checkpoint(epoch=3)
    check_layout(ck)
    mc = CONFIG["model"]
    m = hvc.DirectCTRegression(volume_size=tuple(mc["volume_size"]), **{k: v for k, v in mc.items() if k != "volume_size"})
    assert [n for n, _ in m.named_parameters()] == tr.param_names          # same parameter order -> same optimizer indices
    opt = torch.optim.AdamW(m.parameters(), lr=1.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=10)
    wrapped = torch.nn.Module()
    wrapped.module = m                                                     # a DDP-style wrapper is unwrapped
    ck_prefixed = dict(ck, model_state_dict={"module." + k: v for k, v in ck["model_state_dict"].items()})   # ... and so are its keys
    for c in (ck, ck_prefixed):
        start, best, _ = CK.load_checkpoint(c, wrapped, opt, sched, strict=True)
        assert start == 4 and best == 21.5
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), ck["model_state_dict"][n])
    st = opt.state_dict()["state"]
    assert len(st) == len(tr.param_names) and torch.equal(st[0]["exp_avg"], ck["optimizer_state_dict"]["state"][0]["exp_avg"])
    assert opt.param_groups[0]["lr"] == ck["optimizer_state_dict"]["param_groups"][0]["lr"]
    # inference_direct.py's loader: nested config, flat config, no config
    for cfg in (ck["config"], ck["config"]["model"]):
        model, mcfg = CK.load_model(dict(ck, config=cfg), "cpu")
        assert not model.training and mcfg == mc
    with pytest.raises(RuntimeError):                                      # default config (64^3 / 256 / 4 / 4) does not fit this file: strict
        CK.load_model({k: v for k, v in ck.items() if k != "config"}, "cpu")


def test_stage_handoff_prefix_filter_and_freeze():
    """train_progressive_1gpu.py:213-225 / train_progressive_4gpu.py:223-232 on a stand-in with the cascade's attribute names."""
    from hybrid_vit_cascade_b200 import checkpoint as CK

    class Cascade(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.xray_encoder = torch.nn.Linear(4, 4)
            self.stage1, self.stage2, self.stage3 = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
            self.frozen = []

        def freeze_stage(self, k):
            self.frozen.append(k)
            for p in getattr(self, f"stage{k}").parameters():
                p.requires_grad_(False)

    torch.manual_seed(0)
    src, dst = Cascade(), Cascade()
    ck = {"model_state_dict": src.state_dict()}
    res = CK.load_previous_stage(dst, ck, stage=2, prefix_filtered=True)
    assert torch.equal(dst.stage1.weight, src.stage1.weight) and not torch.equal(dst.stage2.weight, src.stage2.weight)
    assert sorted(res.missing_keys) == ["stage2.bias", "stage2.weight", "stage3.bias", "stage3.weight"] and dst.frozen == [1]
    assert not dst.stage1.weight.requires_grad and dst.stage2.weight.requires_grad
    dst3 = Cascade()
    CK.load_previous_stage(dst3, ck, stage=3, prefix_filtered=False)       # the 4-GPU trainer loads everything non-strictly
    assert torch.equal(dst3.stage3.weight, src.stage3.weight) and dst3.frozen == [1, 2]
