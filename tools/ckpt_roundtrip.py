#!/usr/bin/env python
"""Checkpoint round trip between the reference's format and the drop-in modules.

    python tools/ckpt_roundtrip.py [--out DIR]          (needs a B200: the drop-in modules have no CPU path)

VERIFICATION TOOL, not product code: it builds the reference-side checkpoint with the oracle port (oracle/, test infrastructure) and
torch's own AdamW / CosineAnnealingLR -- the classes the reference trainer uses (train_direct_4gpu.py:159-168) -- exactly in the
layout the trainer's save lines produce (train_direct_4gpu.py:277-297; the layout is pinned to the real reference by
tests/golden/r02_checkpoint_layout.json).  Then:

  1. resume it into hybrid_vit_cascade_b200.DirectCTRegression strictly, once with torch.optim.AdamW and once with FlatAdamW
     (checkpoint.load_checkpoint), and run the NEXT training step on the GPU with both;
  2. compare both updated models with the oracle's own next step on the CPU (the reference's resume path, :177-189);
  3. save each run (checkpoint.save_checkpoint), reload it into the OTHER optimizer backend, run one more step, compare again;
  4. reload the final file through inference_direct.py's load_model path (checkpoint.load_model) and through plain
     torch.optim.AdamW.load_state_dict (what the reference trainer would do with a file written by this package).

tests/test_checkpoint_gpu.py runs the same functions.
"""
import argparse
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIG = {"model": dict(volume_size=[32, 32, 32], xray_img_size=64, voxel_dim=64, vit_depth=1, num_heads=1, xray_feature_dim=64),
          "training": dict(learning_rate=1e-4, weight_decay=0.01, num_epochs=10, gradient_clip=1.0),
          "checkpoints": dict(save_dir="checkpoints_direct", save_every=5)}


def layout():
    with open(os.path.join(ROOT, "tests", "golden", "r02_checkpoint_layout.json")) as f:
        return json.load(f)


class OracleTrainer:
    """The reference trainer's step (train_direct_4gpu.py:49-98 without AMP) over the oracle port, on the CPU."""

    def __init__(self, seed=31):
        from oracle import vit_oracle as O
        self.O = O
        mc = CONFIG["model"]
        self.cfg = O.BackboneConfig(volume_size=tuple(mc["volume_size"]), in_channels=1, voxel_dim=mc["voxel_dim"], depth=mc["vit_depth"],
                                    num_heads=mc["num_heads"], context_dim=mc["xray_feature_dim"], cond_dim=1024)
        lay = layout()
        g = torch.Generator().manual_seed(seed)
        self.sd = {}
        for k, (shape, dtype) in lay["model_state_dict"].items():          # same keys, shapes and order as the reference's state_dict
            if "num_batches_tracked" in k:
                self.sd[k] = torch.zeros(shape, dtype=torch.int64)
            elif "running_var" in k:
                self.sd[k] = torch.ones(shape)
            elif "running_mean" in k:
                self.sd[k] = torch.zeros(shape)
            elif k.endswith("norm1.weight") or k.endswith("norm2.weight") or k.endswith("norm3.weight") or k.endswith("norm.weight") or \
                    (".encoder." in k and len(shape) == 1 and k.endswith(".weight")) or ("voxel_embed" in k and len(shape) == 1 and k.endswith(".weight")):
                self.sd[k] = 1 + 0.1 * torch.randn(shape, generator=g)
            else:
                fan = max(1, int(torch.tensor(shape[1:]).prod())) if len(shape) > 1 else 64
                self.sd[k] = torch.randn(shape, generator=g) * (0.02 if ("adaln" in k or "pos_embed" in k or "initial_volume" in k) else fan ** -0.5)
        self.param_names = lay["parameter_order"]
        for n in self.param_names:
            self.sd[n].requires_grad_(True)
        tc = CONFIG["training"]
        self.params = [self.sd[n] for n in self.param_names]
        self.opt = torch.optim.AdamW(self.params, lr=tc["learning_rate"], weight_decay=tc["weight_decay"])          # :159-163
        self.sched = torch.optim.lr_scheduler.CosineAnnealingLR(self.opt, T_max=tc["num_epochs"])                   # :165-168
        self.xrays = torch.rand(2, 2, 1, 64, 64, generator=g) * 2 - 1
        self.target = torch.rand(2, 1, 32, 32, 32, generator=g) * 2 - 1

    def step(self):
        from oracle import encoder_oracle as E
        new_stats = {}
        y = E.direct_ct_regression(self.xrays, self.sd, self.cfg, training=True, new_stats=new_stats)
        loss = E.direct_regression_loss(y, self.target)["total_loss"]
        self.opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, CONFIG["training"]["gradient_clip"])
        self.opt.step()
        with torch.no_grad():
            for k, v in new_stats.items():
                self.sd[k].copy_(v)
            for k in self.sd:
                if "num_batches_tracked" in k:
                    self.sd[k] += 1
        return float(loss)

    def checkpoint(self, epoch=3):
        """The dict of train_direct_4gpu.py:277-287.  torch.save serialises the optimizer's tensors at this moment; an in-memory dict must
        copy them, state_dict() hands out references that the next step() mutates."""
        import copy
        return {"epoch": epoch, "model_state_dict": {k: v.detach().clone() for k, v in self.sd.items()},
                "optimizer_state_dict": copy.deepcopy(self.opt.state_dict()), "scheduler_state_dict": copy.deepcopy(self.sched.state_dict()),
                "val_psnr": 21.5, "best_psnr": 21.5, "config": CONFIG}


def check_layout(ckpt):
    """The checkpoint against the table recorded from the real reference."""
    lay = layout()
    assert sorted(ckpt.keys()) == lay["checkpoint_keys"], (sorted(ckpt.keys()), lay["checkpoint_keys"])
    msd = ckpt["model_state_dict"]
    assert list(msd.keys()) == list(lay["model_state_dict"].keys())
    for k, (shape, dtype) in lay["model_state_dict"].items():
        assert list(msd[k].shape) == shape and str(msd[k].dtype) == dtype, k
    osd = ckpt["optimizer_state_dict"]
    assert osd["param_groups"][0]["params"] == lay["optimizer_params_index"]
    assert set(lay["optimizer_param_group_keys"]) - {"initial_lr"} <= set(osd["param_groups"][0].keys())
    for i, st in lay["optimizer_state"].items():
        got = osd["state"][int(i)]
        for k, (shape, dtype) in st.items():
            assert list(got[k].shape) == shape and str(got[k].dtype) == dtype, (i, k)


class HvcRun:
    """The drop-in DirectCTRegression resumed from a checkpoint dict, stepping on the GPU with either optimizer backend."""

    def __init__(self, ckpt, backend, xrays, target, dev="cuda"):
        import hybrid_vit_cascade_b200 as hvc
        from hybrid_vit_cascade_b200 import checkpoint as CK
        from hybrid_vit_cascade_b200.dp import GradientBuckets
        self.hvc, self.CK, self.backend = hvc, CK, backend
        mc, tc = ckpt["config"]["model"], ckpt["config"]["training"]
        self.model = hvc.DirectCTRegression(volume_size=tuple(mc["volume_size"]), xray_img_size=mc["xray_img_size"], voxel_dim=mc["voxel_dim"],
                                            vit_depth=mc["vit_depth"], num_heads=mc["num_heads"], xray_feature_dim=mc["xray_feature_dim"]).to(dev).train()
        self.params = list(self.model.parameters())
        self.gb = GradientBuckets(self.params)
        if backend == "flat":
            self.opt = hvc.FlatAdamW(self.gb, lr=tc["learning_rate"], weight_decay=tc["weight_decay"], max_grad_norm=tc["gradient_clip"],
                                     params=self.params)
        else:
            self.opt = torch.optim.AdamW(self.params, lr=tc["learning_rate"], weight_decay=tc["weight_decay"])
        self.sched = torch.optim.lr_scheduler.CosineAnnealingLR(self.opt, T_max=tc["num_epochs"])      # either backend is a torch Optimizer
        self.start_epoch, self.best, _ = CK.load_checkpoint(ckpt, self.model, self.opt, self.sched, strict=True, map_location=dev)
        self.clip = tc["gradient_clip"]
        self.xrays, self.target = xrays.to(dev), target.to(dev)
        self.crit = hvc.DirectRegressionLoss(1.0, 0.5)
        self.config = ckpt["config"]

    def step(self):
        self.hvc.set_dropout_policy("ignore")            # the oracle run has no dropout (masks cannot match torch's)
        try:
            self.gb.reset()
            loss = self.crit(self.model(self.xrays), self.target)["total_loss"]
            loss.backward()
            self.gb.finish()
            if self.backend != "flat":
                torch.nn.utils.clip_grad_norm_(self.params, self.clip)
            self.opt.step()
        finally:
            self.hvc.set_dropout_policy("apply")
        return float(loss)

    def save(self, path, epoch):
        return self.CK.save_checkpoint(path, self.model, self.opt, self.sched, epoch=epoch, config=self.config, val_psnr=22.0, best_psnr=22.0)

    def state(self):
        return {k: v.detach().float().cpu() for k, v in self.model.state_dict().items()}


def update_error(before, after_ref, after_got, names):
    """Relative error of the parameter UPDATE (after - before) against the oracle's, over all parameters."""
    num = den = 0.0
    for n in names:
        dr = (after_ref[n] - before[n]).double()
        dg = (after_got[n] - before[n]).double()
        num += float((dg - dr).pow(2).sum())
        den += float(dr.pow(2).sum())
    return (num / max(den, 1e-300)) ** 0.5


def roundtrip(out_dir, log=print):
    ref = OracleTrainer()
    ref.step()                                             # one step so that the optimizer state is not empty
    ck0 = ref.checkpoint(epoch=3)
    check_layout(ck0)
    before = {k: v.detach().clone() for k, v in ck0["model_state_dict"].items()}
    loss_ref = ref.step()                                  # the reference's own next step after a resume
    after_ref = {k: v.detach().clone() for k, v in ref.sd.items()}
    res = {}
    runs = {b: HvcRun(ck0, b, ref.xrays, ref.target) for b in ("torch", "flat")}
    for b, run in runs.items():
        assert run.start_epoch == 4 and run.best == 21.5
        loss = run.step()
        err = update_error(before, after_ref, run.state(), ref.param_names)
        res[f"step1_{b}"] = dict(loss=loss, loss_ref=loss_ref, update_rel_err=err)
        log(f"resume -> {b:5s}: loss {loss:.6f} (oracle {loss_ref:.6f}), update rel err vs oracle {err:.3e}")
    # cross-load: a file written with one backend resumes in the other; both continue to the same place as the oracle's 3rd step
    loss_ref3 = ref.step()
    after_ref3 = {k: v.detach().clone() for k, v in ref.sd.items()}
    for src, dst in (("torch", "flat"), ("flat", "torch")):
        path = os.path.join(out_dir, f"checkpoint_epoch_4_{src}.pt")
        ck = runs[src].save(path, epoch=4)
        check_layout(torch.load(path, map_location="cpu", weights_only=False))
        run = HvcRun(torch.load(path, map_location="cuda", weights_only=False), dst, ref.xrays, ref.target)
        loss = run.step()
        err = update_error(before, after_ref3, run.state(), ref.param_names)
        res[f"step2_{src}_to_{dst}"] = dict(loss=loss, loss_ref=loss_ref3, update_rel_err=err)
        log(f"{src:5s} file -> {dst:5s}: loss {loss:.6f} (oracle {loss_ref3:.6f}), 2-step update rel err vs oracle {err:.3e}")
        # the reference trainer's own resume of a file written by this package: plain torch load_state_dict on the oracle's optimizer
        probe = OracleTrainer()
        probe.opt.load_state_dict(torch.load(path, map_location="cpu", weights_only=False)["optimizer_state_dict"])
        probe.sched.load_state_dict(ck["scheduler_state_dict"]) if "scheduler_state_dict" in ck else None
    # inference path (inference_direct.py:22-66)
    from hybrid_vit_cascade_b200 import checkpoint as CK
    model, mcfg = CK.load_model(os.path.join(out_dir, "checkpoint_epoch_4_flat.pt"), "cuda")
    assert not model.training and mcfg == CONFIG["model"]
    with torch.no_grad():
        y = model(ref.xrays.cuda())
    res["inference_output_finite"] = bool(torch.isfinite(y).all())
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    out = a.out or tempfile.mkdtemp(prefix="hvc_ckpt_")
    os.makedirs(out, exist_ok=True)
    print(json.dumps(roundtrip(out), indent=1))
