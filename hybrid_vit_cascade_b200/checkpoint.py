"""Checkpoint compatibility with the reference trainers.

The reference writes plain ``torch.save`` dicts

    direct:       {epoch, model_state_dict, optimizer_state_dict, scheduler_state_dict, val_psnr[, best_psnr], config}
                  (direct_regression/train_direct_4gpu.py:277-297; best_model.pt, checkpoint_epoch_N.pt; resume :177-189)
    progressive:  {epoch, model_state_dict, optimizer_state_dict, scheduler_state_dict, val_loss[, val_psnr, val_ssim], config}
                  (progressive_cascade/train_progressive_4gpu.py:338-364; stage{k}_best.pth, stage{k}_epoch{N}.pth)

with ``model_state_dict`` taken from ``model.module`` (DDP-unwrapped) and ``optimizer_state_dict`` from ``torch.optim.AdamW``.
The drop-in modules keep every ``state_dict`` key and shape, so these files load strictly; this module adds the pieces a
trainer needs around that:

  * ``save_checkpoint`` / ``load_checkpoint`` -- the same dict, either optimizer backend.  ``optim.FlatAdamW`` keeps its moments in
    flat buckets; ``flat_to_torch_state`` / ``torch_to_flat_state`` convert to and from the ``torch.optim.AdamW`` layout (per-parameter
    ``step`` / ``exp_avg`` / ``exp_avg_sq`` indexed by position in the optimizer's parameter list), so a run can be resumed with
    the other backend -- or by the reference trainer itself.
  * ``load_model`` -- inference_direct.py:22-66: rebuild ``DirectCTRegression`` from ``checkpoint['config']['model']`` (nested or flat
    config, default config when absent), strict load, eval mode.
  * ``load_previous_stage`` -- the cascade's stage hand-off: non-strict load of ``stage{k-1}_best.pth`` (train_progressive_4gpu.py:223-232)
    or the prefix-filtered variant (train_progressive_1gpu.py:213-225), then ``freeze_stage`` of the earlier stages.
  * keys saved from a still-wrapped model (``module.`` prefix) are accepted.
"""
from collections import OrderedDict
from pathlib import Path
from typing import Dict, Iterable, Optional

import torch


def unwrap(model: torch.nn.Module) -> torch.nn.Module:
    """DDP-style wrappers expose the real model as ``.module`` (train_direct_4gpu.py:280 saves model.module.state_dict())."""
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def strip_module_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    if state_dict and all(k.startswith("module.") for k in state_dict):
        return OrderedDict((k[len("module."):], v) for k, v in state_dict.items())
    return state_dict


# ------------------------------------------------------------------------------------------ optimizer state interchange
def flat_to_torch_state(opt, params: Iterable[torch.nn.Parameter]) -> dict:
    """``FlatAdamW`` -> the ``state_dict()`` a ``torch.optim.AdamW(params, ...)`` would hold after the same steps.
    `params`: the parameter list the torch optimizer is (or would be) built from, in order (the reference passes model.parameters())."""
    params = list(params)
    where = {}
    for bi, (members, offsets) in enumerate(zip(opt.gb._members, opt.gb._offsets)):
        for p, off in zip(members, offsets):
            where[id(p)] = (bi, off)
    step = opt.t.detach().clone().reshape(())
    state = {}
    for i, p in enumerate(params):
        if id(p) not in where:
            continue                      # frozen (not bucketed): torch keeps no state for a parameter that never got a gradient
        if float(step) == 0.0:
            continue
        bi, off = where[id(p)]
        n = p.numel()
        state[i] = {"step": step.clone().cpu(),       # torch keeps `step` as a CPU float tensor unless capturable/fused
                    "exp_avg": opt.exp_avg[bi][off:off + n].view_as(p).clone(),
                    "exp_avg_sq": opt.exp_avg_sq[bi][off:off + n].view_as(p).clone()}
    group = {"lr": opt.lr, "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": opt.weight_decay, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": True}
    if "initial_lr" in opt.param_groups[0]:              # written by an attached LR scheduler, as in the reference's files
        group["initial_lr"] = opt.param_groups[0]["initial_lr"]
    group["params"] = list(range(len(params)))
    return {"state": state, "param_groups": [group]}


def torch_to_flat_state(opt, torch_state: dict, params: Iterable[torch.nn.Parameter]) -> None:
    """Load a ``torch.optim.AdamW.state_dict()`` (e.g. a reference checkpoint's ``optimizer_state_dict``) into ``FlatAdamW``."""
    params = list(params)
    index_of = {}
    flat_index = [i for g in torch_state["param_groups"] for i in g["params"]]
    assert len(flat_index) == len(params), f"optimizer state describes {len(flat_index)} parameters, got {len(params)}"
    for pos, p in enumerate(params):
        index_of[id(p)] = flat_index[pos]
    steps = set()
    with torch.no_grad():
        for bi, (members, offsets) in enumerate(zip(opt.gb._members, opt.gb._offsets)):
            for p, off in zip(members, offsets):
                st = torch_state["state"].get(index_of[id(p)])
                n = p.numel()
                if st is None:
                    opt.exp_avg[bi][off:off + n].zero_()
                    opt.exp_avg_sq[bi][off:off + n].zero_()
                    continue
                opt.exp_avg[bi][off:off + n].copy_(st["exp_avg"].reshape(-1))
                opt.exp_avg_sq[bi][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(float(st["step"]))
        if len(steps) > 1:
            raise ValueError(f"FlatAdamW keeps one step count; the checkpoint holds several ({sorted(steps)})")
        opt.t.fill_(steps.pop() if steps else 0.0)
    g = torch_state["param_groups"][0]
    opt.lr, opt.betas, opt.eps, opt.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
    if "initial_lr" in g:
        opt.param_groups[0]["initial_lr"] = g["initial_lr"]


def _is_flat(opt) -> bool:
    return hasattr(opt, "gb") and hasattr(opt, "exp_avg_sq")


# ------------------------------------------------------------------------------------------ save / load
def save_checkpoint(path, model, optimizer=None, scheduler=None, epoch=0, config=None, optimizer_params=None, **metrics):
    """Write the reference's checkpoint dict.  `metrics`: val_psnr / best_psnr (direct) or val_loss / val_psnr / val_ssim (progressive).
    `optimizer_params`: the parameter order of the torch optimizer this file should be loadable into (default model.parameters(),
    what train_direct_4gpu.py:159 uses); only needed to convert FlatAdamW state."""
    m = unwrap(model)
    ckpt = {"epoch": epoch, "model_state_dict": m.state_dict()}
    if optimizer is not None:
        ckpt["optimizer_state_dict"] = (flat_to_torch_state(optimizer, optimizer_params) if _is_flat(optimizer) and optimizer_params is not None
                                        else optimizer.state_dict())
    if scheduler is not None:
        ckpt["scheduler_state_dict"] = scheduler.state_dict()
    ckpt.update(metrics)
    ckpt["config"] = config
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(ckpt, path)
    return ckpt


def load_checkpoint(path_or_dict, model, optimizer=None, scheduler=None, strict=True, map_location=None, optimizer_params=None):
    """Resume (train_direct_4gpu.py:177-189): weights, optimizer, scheduler -> (start_epoch, best_psnr, checkpoint dict)."""
    ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location=map_location, weights_only=False)
    m = unwrap(model)
    result = m.load_state_dict(strip_module_prefix(ckpt["model_state_dict"]), strict=strict)
    _after_weight_load()
    if optimizer is not None and "optimizer_state_dict" in ckpt:
        if _is_flat(optimizer) and optimizer_params is not None:
            torch_to_flat_state(optimizer, ckpt["optimizer_state_dict"], optimizer_params)
        else:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])      # FlatAdamW speaks the torch.optim.AdamW layout itself
    if scheduler is not None and "scheduler_state_dict" in ckpt:
        scheduler.load_state_dict(ckpt["scheduler_state_dict"])
    start_epoch = ckpt.get("epoch", 0) + 1
    best = ckpt.get("best_psnr", ckpt.get("val_psnr", 0))
    ckpt["_load_result"] = result
    return start_epoch, best, ckpt


DEFAULT_DIRECT_MODEL_CONFIG = {"volume_size": (64, 64, 64), "xray_img_size": 512, "voxel_dim": 256, "vit_depth": 4, "num_heads": 4,
                               "xray_feature_dim": 512}      # inference_direct.py:40-47


def load_model(checkpoint_path, device, token_grid="reference"):
    """inference_direct.py:22-66 -> (DirectCTRegression in eval mode on `device`, model config)."""
    from .xray_encoder import DirectCTRegression
    ckpt = checkpoint_path if isinstance(checkpoint_path, dict) else torch.load(checkpoint_path, map_location=device, weights_only=False)
    full = ckpt.get("config", None)
    cfg = (full["model"] if "model" in full else full) if full is not None else DEFAULT_DIRECT_MODEL_CONFIG
    model = DirectCTRegression(volume_size=tuple(cfg["volume_size"]), xray_img_size=cfg["xray_img_size"], voxel_dim=cfg["voxel_dim"],
                               vit_depth=cfg["vit_depth"], num_heads=cfg["num_heads"], xray_feature_dim=cfg["xray_feature_dim"],
                               token_grid=token_grid).to(device)
    model.load_state_dict(strip_module_prefix(ckpt["model_state_dict"]))       # strict, as the reference
    _after_weight_load()
    model.eval()
    return model, cfg


def load_previous_stage(model, checkpoint_path, stage: int, prefix_filtered: bool = False, map_location=None, freeze: bool = True):
    """Stage hand-off of the progressive cascade before training `stage` (2 or 3): weights of ``stage{stage-1}_best.pth`` into the
    cascade model, earlier stages frozen.  prefix_filtered=False: ``load_state_dict(strict=False)`` of everything in the file
    (train_progressive_4gpu.py:223-232); True: keys of the stages not yet trained are dropped first
    (train_progressive_1gpu.py:213-225 -- 'handles architecture changes in later stages')."""
    ckpt = checkpoint_path if isinstance(checkpoint_path, dict) else torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    sd = strip_module_prefix(ckpt["model_state_dict"])
    if prefix_filtered:
        drop = ("stage2.", "stage3.") if stage == 2 else ("stage3.",) if stage == 3 else ()
        sd = OrderedDict((k, v) for k, v in sd.items() if not k.startswith(drop))
    m = unwrap(model)
    result = m.load_state_dict(sd, strict=False)
    _after_weight_load()
    if freeze:
        for prev in range(1, stage):
            m.freeze_stage(prev)
    return result


def _after_weight_load():
    """load_state_dict copies into the parameters in place (bumping their version counters), so cached bf16 operands refresh on their
    own; the explicit clear also covers parameters whose storage FlatAdamW re-pointed."""
    from . import ops
    ops.clear_weight_cache()
