#!/usr/bin/env python
"""bench.py -- training throughput of the 3D ViT backbone hot path (volumes/sec), B200-native arm and
reference (CPU) arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload direct128|direct64|stage2|stage3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: HybridViT3D forward (voxel embed, L x (AdaLN
self-attention, cross-attention to the X-ray tokens, AdaLN MLP), LN/proj/upsample head) + loss +
backward of every parameter and of the learned input volume + (N>1) bucketed gradient all-reduce
overlapped with backward + gradient-norm clip + AdamW (loss = DirectRegressionLoss L1 + 0.5 (1 - SSIM3D), clip 1.0, AdamW lr 1e-4 wd 0.01:
config_direct.json).  Workload = BASELINE.json configs[2]: direct_regression at 128^3,
C=256, 4 heads (d=64), depth 4, 32^3 = 32768 volume tokens (what the conv stack emits), 4096 X-ray
context tokens x 512, batch 8 per GPU, synthetic inputs, random-init weights (AdaLN re-randomised so the
self-attention and MLP branches are live).  The modules are in train() mode with their default nn.Dropout(p=0.1) at all six
sites per block ACTIVE (what the reference trainers run: no caller overrides it); the same step with
dropout switched off is timed right after and reported as "dropout_off".  The CPU arm runs dropout-off, which favours it
(bernoulli_ is 49 % of the reference's train-mode CPU step).

Prints ONE JSON line on stdout (rank 0); diagnostics go to stderr.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "direct128": dict(volume=(128, 128, 128), token_grid="conv", voxel_dim=256, depth=4, heads=4, ctx_hw=64, ctx_dim=512,
                      cond_dim=1024, batch=8, desc="direct_regression 128^3 (32^3=32768 tokens, 4096 ctx tokens)"),
    "direct64": dict(volume=(64, 64, 64), token_grid="reference", voxel_dim=256, depth=4, heads=4, ctx_hw=64, ctx_dim=512,
                     cond_dim=1024, batch=8, desc="direct_regression 64^3 (16^3=4096 tokens, 4096 ctx tokens)"),
    # progressive cascade refiners (model_progressive.py:177-186, :247-256): 32-channel input volume from the
    # upsample conv, 8 heads (d=32); batch 2 per GPU (config_progressive.json:27,35); stage 3 runs the ViT under
    # torch.utils.checkpoint (model_progressive.py:286-291)
    "stage2": dict(volume=(128, 128, 128), token_grid="conv", in_channels=32, voxel_dim=256, depth=6, heads=8, ctx_hw=32,
                   ctx_dim=512, cond_dim=1024, batch=2,
                   desc="progressive_cascade stage 2 ViT 128^3 (32 ch in, 32768 tokens, 1024 ctx tokens, d=32)"),
    # the whole direct-regression model (model_direct.py: X-ray encoder -> context / cond -> backbone), X-rays in, volume out:
    # the backbone workloads above plus the encoder of SURVEY 8(f) row 1
    "direct128_model": dict(volume=(128, 128, 128), token_grid="conv", voxel_dim=256, depth=4, heads=4, ctx_hw=64, ctx_dim=512,
                            cond_dim=1024, batch=8, full_model=True, xray=512,
                            desc="DirectCTRegression 128^3: 2 x 512^2 X-rays -> encoder -> 3D ViT (32768 tokens, 4096 ctx tokens)"),
    "direct64_model": dict(volume=(64, 64, 64), token_grid="reference", voxel_dim=256, depth=4, heads=4, ctx_hw=64, ctx_dim=512,
                           cond_dim=1024, batch=8, full_model=True, xray=512,
                           desc="DirectCTRegression 64^3: 2 x 512^2 X-rays -> encoder -> 3D ViT (4096 tokens, 4096 ctx tokens)"),
    # the whole Stage2Refiner128 (model_progressive.py:153-215): 64^3 volume -> trilinear x2 -> Conv3d(1->32)+GroupNorm+GELU -> refiner
    # ViT -> residual blend; the stage-1 volume and the encoder features are inputs (stage 1 is frozen in the reference trainer)
    "stage2_model": dict(volume=(128, 128, 128), token_grid="conv", in_channels=32, voxel_dim=256, depth=6, heads=8, ctx_hw=32,
                         ctx_dim=512, cond_dim=1024, batch=2, stage_wrapper=2,
                         desc="Stage2Refiner128: 64^3 volume -> upsample + Conv3d/GN/GELU -> ViT 128^3 (32768 tokens, 1024 ctx tokens, d=32) -> blend"),
    # the whole Stage3Refiner256 (model_progressive.py:218-315): 128^3 volume -> trilinear x2 -> Conv3d(1->32)+GN+GELU -> refiner ViT, plus
    # the detail_enhancer CNN (Conv3d 1->64->32->1 at 256^3) and the three-way blend; no recomputation (fits the 180 GB)
    "stage3_model": dict(volume=(256, 256, 256), token_grid="reference", in_channels=32, voxel_dim=256, depth=8, heads=8, ctx_hw=64,
                         ctx_dim=512, cond_dim=1024, batch=2, stage_wrapper=3,
                         desc="Stage3Refiner256: 128^3 volume -> upsample + Conv3d/GN/GELU -> ViT 256^3 (32768 tokens, 4096 ctx tokens, d=32) + detail_enhancer CNN -> blend"),
    "stage3": dict(volume=(256, 256, 256), token_grid="reference", in_channels=32, voxel_dim=256, depth=8, heads=8, ctx_hw=64,
                   ctx_dim=512, cond_dim=1024, batch=2, checkpoint=True,
                   desc="progressive_cascade stage 3 ViT 256^3 (32 ch in, 32768 tokens, 4096 ctx tokens, d=32, checkpointed)"),
}


# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE self-attention backward launch per sample at 32768
# tokens, from the `ncu --set full` captures profiles/r02_ncu_full_attn_d64_b1.csv (d = 64: 104.1 MB read + 24.1 MB written) and
# profiles/r01_ncu_full_hot_kernels_b1.csv (d = 32) (one sample, C = 256): head_dim -> bytes.
# A launch over B samples moves B times that (each (batch, head) slice is touched by its own CTAs only).  The algorithmic
# bytes are Q, K, V, dO, O reads + dK, dV writes + the fp32 dQ reduction = 8 * N * C * 2 + N * C * 4 = 151 MB per sample.
NCU_ATTN_BWD_DRAM_BYTES_PER_SAMPLE_32K = {64: 128.2e6, 32: 141.4e6}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def oracle_cfg(w):
    from oracle import vit_oracle as O
    return O.BackboneConfig(volume_size=w["volume"], in_channels=w.get("in_channels", 1), voxel_dim=w["voxel_dim"], depth=w["depth"],
                            num_heads=w["heads"], context_dim=w["ctx_dim"], cond_dim=w["cond_dim"],
                            token_grid=w["token_grid"])


def step_flops(w):
    """Algorithmic FLOPs of one sample, forward+backward (3 x forward)."""
    from oracle import vit_oracle as O
    f = O.forward_flops(oracle_cfg(w), w["ctx_hw"] ** 2)
    if w.get("full_model"):     # X-ray encoder, per sample = 2 views: conv 7x7 1->64 s2, 3x3 64->128, 3x3 128->ctx_dim (+ small linears)
        x = w["xray"]
        enc = 2 * (2.0 * 49 * 64 * (x // 2) ** 2 + 2.0 * 576 * 128 * (x // 4) ** 2 + 2.0 * 1152 * w["ctx_dim"] * (x // 8) ** 2)
        f = dict(f, encoder=enc, total=f["total"] + enc)
    if w.get("stage_wrapper"):  # stage wrapper convs at full resolution: Conv3d 1->32 (k3); stage 3 adds the detail CNN 1->64->32 (k3) ->1 (k1)
        V = w["volume"][0] * w["volume"][1] * w["volume"][2]
        wr = 2.0 * 27 * 32 * V + (2.0 * 27 * 64 * V + 2.0 * 1728 * 32 * V + 2.0 * 32 * V if w["stage_wrapper"] == 3 else 0.0)
        f = dict(f, wrapper=wr, total=f["total"] + wr)
    return {k: 3.0 * v for k, v in f.items()}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.strip() == "Active":
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_baseline(w, seconds_budget=20.0, rows=2048, steps=None, warmup=1, train=True):
    """Oracle port on the host cores, bounded sample (oracle/cpu_baseline.py).  The sample executes `measured_fraction` of one
    volume's work and the per-volume time is EXTRAPOLATED from it (flagged as such in every line that carries it)."""
    from oracle.cpu_baseline import CpuBaseline
    cb = CpuBaseline(oracle_cfg(w), w["ctx_hw"] ** 2, rows=rows, train=train)
    for _ in range(warmup):
        cb.step()
    est, meas, t0 = [], [], time.perf_counter()
    while True:
        est.append(cb.step())
        meas.append(cb.last_measured_s)
        if steps is not None:
            if len(est) >= steps:
                break
        elif time.perf_counter() - t0 > seconds_budget or len(est) >= 8:
            break
    sec_per_vol = min(est) if steps is None else sum(est) / len(est)
    return {"value": 1.0 / sec_per_vol, "unit": "volumes/s", "cores": cb.threads, "kind": "port", "sample": cb.describe(),
            "extrapolated": True, "measured_fraction": cb.measured_fraction(), "measured_ms_per_sample_step": 1e3 * sum(meas) / len(meas),
            "dropout": 0.1 if train else 0.0, "sec_per_volume": sec_per_vol, "samples": len(est)}


def a0_full(best_of=2):
    """BASELINE.json configs[0] / SURVEY 8(d) 'config A0', measured IN FULL on the host cores: DirectCTRegression 64^3, B=1,
    forward + loss + backward, train() and dropout-off.  One warm-up step, then best of `best_of` per mode."""
    from oracle.cpu_baseline import A0Full
    a = A0Full()
    a.step(False)
    out = {"cores": a.threads, "kind": "port", "extrapolated": False, "steps_per_mode": best_of}
    for name, train in (("train", True), ("dropout_off", False)):
        t = min(a.step(train) for _ in range(best_of))
        out[name] = {"value": 1.0 / t, "unit": "volumes/s", "ms_per_step": t * 1e3, "sample": a.describe(train)}
    return out


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_wall = time.perf_counter()
    cb = cpu_baseline(w, rows=args.cpu_rows, steps=args.steps, warmup=max(1, min(args.warmup, 2)), train=True)
    cb_off = cpu_baseline(w, rows=args.cpu_rows, steps=2, warmup=1, train=False)      # the same sample with the nn.Dropout sites off
    full = a0_full() if not args.skip_a0 else None
    line = {
        "impl": "reference", "metric": "train volumes/sec (3D ViT backbone fwd+bwd)", "value": cb["value"], "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # what one timed step of THIS run took (the bounded sample), not the per-volume estimate: value = 1 / (ms_per_step / measured_fraction + embed/head)
        "ms_per_step": cb["measured_ms_per_sample_step"], "extrapolated": True, "measured_fraction": cb["measured_fraction"],
        "estimated_ms_per_volume": cb["sec_per_volume"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_gpu": 1, "dropout": 0.1,
                   "note": "CPU arm, train mode like the GPU arm: one step = one bounded sample of one volume (one block, a slab of query rows); "
                           "value is EXTRAPOLATED from it (the full 128^3 step needs 15.8 TFLOP and a 17 GB score tensor per block on the host); "
                           "a0_full is the reference's own CPU-runnable configuration measured in full"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "measured_fraction")},
        "e2e": {"value": cb["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["dropout_off"] = {k: cb_off[k] for k in ("value", "unit", "extrapolated", "measured_fraction", "measured_ms_per_sample_step", "sample")}
    if full:
        line["a0_full"] = full
    line["wall_s"] = time.perf_counter() - t_wall
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU-eager arm
def _run_self(extra, timeout=900):
    """Run this script again (fresh process, fresh CUDA context) and parse its one JSON line."""
    r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True, timeout=timeout)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError(f"bench.py {' '.join(extra)} printed no JSON line (rc {r.returncode}): {r.stderr[-400:]}")


def run_eager(args):
    """SURVEY 0.1 / BASELINE.md 3: 'the bar is PyTorch eager on the same B200'.  The oracle port (the reference's ATen calls) under
    torch.autocast(bf16) on cuda:0, timed beside the hand-written kernels:
      * the full 64^3 direct-regression training step at batch 8 (configs[1]) -- eager vs this package (a bench.py --workload direct64 run);
      * one block at the headline configuration's 32768 tokens, batch 1, forward+backward -- all of it that the reference's
        materialised softmax lets a 180 GB GPU hold."""
    import hybrid_vit_cascade_b200 as hvc
    from oracle import eager_baseline as EB
    from oracle import vit_oracle as O
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    out = {"impl": "eager", "what": "oracle port of the reference modules on the same B200, torch.autocast(bf16), cuBLAS/ATen eager kernels",
           "torch": torch.__version__}
    # ---- one block at 32768 tokens, B = 1, train mode and dropout off
    blk = {}
    for mode, train in (("train", True), ("dropout_off", False)):
        try:
            eb = EB.Block32768(dev, heads=4, train=train)
            ms_eager = eb.ms_per_step(steps=3, warmup=1)
            sd = {k[len("blocks.0."):]: v.detach() for k, v in eb.sd.items()}
            x, ctx, cond, r = eb.x.detach(), eb.ctx, eb.cond, eb.r
            del eb
            torch.cuda.empty_cache()
        except torch.OutOfMemoryError as e:     # pragma: no cover
            blk[mode] = {"eager": "out of memory", "detail": str(e)[:120]}
            continue
        m = hvc.HybridViTBlock3D(256, num_heads=4, context_dim=512, cond_dim=1024).to(dev).train()
        m.load_state_dict(sd, strict=True)
        hvc.set_dropout_policy("apply" if train else "ignore")
        xg = x.clone().requires_grad_(True)

        def hstep():
            m.zero_grad(set_to_none=True)
            xg.grad = None
            (m(xg, ctx, cond) * r).sum().backward()

        from oracle.eager_baseline import _time
        ms_hvc = _time(hstep, 5, 2)
        hvc.set_dropout_policy("apply")
        blk[mode] = {"eager_ms": ms_eager, "hvc_ms": ms_hvc, "speedup": ms_eager / ms_hvc}
        del m
        torch.cuda.empty_cache()
    out["block_32768_tokens_b1_fwd_bwd"] = dict(blk, shape="HybridViTBlock3D C=256, 4 heads (d=64), 32768 tokens, 4096 context tokens, batch 1")
    # ---- full 64^3 training step at batch 8
    d64 = {}
    for mode, train in (("train", True), ("dropout_off", False)):
        st = EB.Direct64Step(dev, batch=8, train=train)
        d64[mode] = {"eager_ms": st.ms_per_step(steps=args.steps, warmup=max(2, args.warmup))}
        del st
        torch.cuda.empty_cache()
    try:
        h = _run_self(["--workload", "direct64", "--skip-cpu-baseline", "--skip-eager", "--steps", str(max(args.steps, 10)), "--warmup", "5"])
        d64["train"]["hvc_ms"] = h["ms_per_step"]
        d64["dropout_off"]["hvc_ms"] = h["dropout_off"]["ms_per_step"]
        for k in ("train", "dropout_off"):
            d64[k]["speedup"] = d64[k]["eager_ms"] / d64[k]["hvc_ms"]
            d64[k]["eager_volumes_per_s"] = 8e3 / d64[k]["eager_ms"]
            d64[k]["hvc_volumes_per_s"] = 8e3 / d64[k]["hvc_ms"]
    except Exception as e:      # pragma: no cover
        d64["hvc_error"] = str(e)[:300]
    out["direct64_b8_training_step"] = dict(d64, shape="direct_regression 64^3 backbone, B=8, 4096 tokens: forward + DirectRegressionLoss + backward + clip + AdamW")
    out.update({"metric": "train volumes/sec (3D ViT backbone fwd+bwd)", "value": 8e3 / d64["train"]["eager_ms"], "unit": "volumes/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": d64["train"]["eager_ms"], "higher_is_better": True, "dtype": "bf16",
                "data": "synthetic", "config": {"workload": WORKLOADS["direct64"]["desc"], "batch_per_gpu": 8, "dropout": 0.1}})
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, w):
    import torch.distributed as dist
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200 import _lib, kernels as K
    from hybrid_vit_cascade_b200.dp import GradientBuckets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch or w["batch"]
    D, H, W = w["volume"]
    hvc.set_dropout_policy("apply" if args.dropout == "on" else "ignore")

    torch.manual_seed(0)
    cin = w.get("in_channels", 1)
    full = bool(w.get("full_model"))
    wrap2 = bool(w.get("stage_wrapper"))
    model = (hvc.Stage2Refiner128 if w.get("stage_wrapper") == 2 else hvc.Stage3Refiner256)(volume_size=w["volume"], voxel_dim=w["voxel_dim"], vit_depth=w["depth"], num_heads=w["heads"],
                                 xray_feature_dim=w["ctx_dim"], token_grid=w["token_grid"]).to(dev) if wrap2 else \
        hvc.DirectCTRegression(volume_size=w["volume"], xray_img_size=w["xray"], voxel_dim=w["voxel_dim"], vit_depth=w["depth"],
                                   num_heads=w["heads"], xray_feature_dim=w["ctx_dim"], token_grid=w["token_grid"]).to(dev) if full else \
        hvc.HybridViT3D(volume_size=w["volume"], in_channels=cin, voxel_dim=w["voxel_dim"], depth=w["depth"],
                            num_heads=w["heads"], context_dim=w["ctx_dim"], cond_dim=w["cond_dim"],
                            token_grid=w["token_grid"]).to(dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaln.linear" in n:
                p.normal_(0.0, 0.02)
    direct = cin == 1 and not full
    if full:
        params = list(model.parameters())
    elif direct:
        initial_volume = torch.nn.Parameter(torch.randn(1, 1, D, H, W, device=dev) * 0.01)   # model_direct.py:57
        params = list(model.parameters()) + [initial_volume]
    else:
        params = list(model.parameters())
    gb = GradientBuckets(params)
    gb.broadcast_parameters(params)
    flat_opt = args.optimizer == "flat"
    if flat_opt:    # one sum-of-squares + one fused clip+AdamW kernel per gradient bucket (hvc_optim.cu)
        opt = hvc.FlatAdamW(gb, lr=1e-4, weight_decay=0.01, max_grad_norm=args.clip)
    else:
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True, capturable=bool(args.graph))

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    hw = w["ctx_hw"]
    # backbone workloads: the encoder feature map (B, C, H', W'); full-model workloads: the AP + lateral X-rays in [-1, 1]
    feat = (torch.rand(B, 2, 1, w["xray"], w["xray"], device=dev, generator=g) * 2 - 1) if full else \
        torch.rand(B, w["ctx_dim"], hw, hw, device=dev, generator=g)
    cond = torch.randn(B, w["cond_dim"], device=dev, generator=g)
    target = torch.rand(B, 1, D, H, W, device=dev, generator=g) * 2 - 1
    # cascade refiners: the ViT input is the (B, 32, D, H, W) output of the stage's upsample conv; its gradient is needed
    vol_in = None if (direct or full) else \
        (torch.rand(B, 1, D // 2, H // 2, W // 2, device=dev, generator=g) * 2 - 1) if wrap2 else \
        (torch.randn(B, cin, D, H, W, device=dev, generator=g) * 0.5).requires_grad_(True)
    use_ckpt = bool(w.get("checkpoint"))
    if args.loss == "direct":       # DirectRegressionLoss: L1 + 0.5 (1 - SSIM3D), model_direct.py:110-131 / config_direct.json
        crit = hvc.DirectRegressionLoss(1.0, 0.5)

        def loss_fn(out, tgt):
            return crit(out, tgt)["total_loss"]
    else:
        def loss_fn(out, tgt):
            return (out - tgt).abs().mean()
    if use_ckpt:
        from torch.utils.checkpoint import checkpoint

    def step(feat_, cond_, target_):
        gb.reset()
        if full:
            if world > 1:
                gb.broadcast_buffers(model)          # DDP(broadcast_buffers=True), train_direct_4gpu.py:146: BatchNorm running stats follow rank 0
            out = model(feat_)
            loss = loss_fn(out, target_)
            loss.backward()
            gb.finish()
            if args.clip > 0 and not flat_opt:
                torch.nn.utils.clip_grad_norm_(params, args.clip, foreach=True)
            opt.step()
            return loss
        if wrap2:
            out = model(vol_in, feat_, cond_)
            loss = loss_fn(out, target_)
            loss.backward()
            gb.finish()
            if args.clip > 0 and not flat_opt:
                torch.nn.utils.clip_grad_norm_(params, args.clip, foreach=True)
            opt.step()
            return loss
        ctx = feat_.flatten(2).transpose(1, 2)                                    # model_direct.py:80 (a view, no copy)
        if direct:
            out = model(initial_volume.expand(B, -1, -1, -1, -1), ctx, cond_)
        elif use_ckpt:
            vol_in.grad = None
            out = checkpoint(model, vol_in, ctx, cond_, use_reentrant=False)
        else:
            vol_in.grad = None
            out = model(vol_in, ctx, cond_)
        loss = loss_fn(out, target_)
        loss.backward()
        gb.finish()
        if args.clip > 0 and not flat_opt:
            torch.nn.utils.clip_grad_norm_(params, args.clip, foreach=True)     # after the all-reduce, on the averaged gradients
        opt.step()
        return loss

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step(feat, cond, target)
    # ---- device-resident timing
    prof = {}
    K.set_profiler(prof)
    sampler = ClockSampler(local)
    sampler.start()                       # every rank samples its own GPU (per_rank evidence); rank 0's goes into "clocks"
    l0 = _lib.launch_count()
    ms_total = timed(lambda: step(feat, cond, target), args.steps)
    launches = _lib.launch_count() - l0
    local_clocks = sampler.stop()
    clocks = local_clocks if rank == 0 else {}
    K.set_profiler(None)
    torch.cuda.synchronize()
    kern = {}
    for name, evs in prof.items():
        t = [a.elapsed_time(b) for a, b, _ in evs]
        fl = [f for _, _, f in evs]
        kern[name] = {"calls": len(t), "ms_total": sum(t), "tflops": sum(fl) / (sum(t) * 1e9) if sum(t) > 0 else 0.0,
                      "ms_max_call": max(t), "tflops_max_call": max(f / (x * 1e9) for f, x in zip(fl, t) if x > 0)}
    # dominant kernel: self-attention backward launches (the largest calls of attn_bwd)
    bwd = sorted(((a.elapsed_time(b), f) for a, b, f in prof.get("attn_bwd", [])), key=lambda z: -z[1])
    big = [z for z in bwd if z[1] == bwd[0][1]] if bwd else []

    # ---- optional: the whole step (forward, loss, backward, clip, AdamW) captured once in a CUDA graph and replayed -- removes the
    # host-side launch gaps that dominate the small 64^3 configuration (225 kernel launches in ~11 ms)
    graph_info = None
    if args.graph:
        # N > 1: the bucket all-reduces are issued from the gradient hooks while the backward is being captured; NCCL collectives are
        # capturable, so they become nodes of the same graph (on the communicator's stream, joined back by finish()'s stream wait)
        s_feat, s_cond, s_target = feat.clone(), cond.clone(), target.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step(s_feat, s_cond, s_target)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, capture_error_mode="thread_local" if world > 1 else "global"):
            s_loss = step(s_feat, s_cond, s_target)

        def step(f, c, t, _eager=step):                      # noqa: F811  (replaces the eager step from here on)
            if f is not s_feat:
                s_feat.copy_(f, non_blocking=True)
                s_cond.copy_(c, non_blocking=True)
                s_target.copy_(t, non_blocking=True)
            cg.replay()
            return s_loss

        for _ in range(3):
            step(s_feat, s_cond, s_target)
        ms_graph = timed(lambda: step(s_feat, s_cond, s_target), args.steps)
        graph_info = {"ms_per_step_eager": ms_total / args.steps, "ms_per_step": ms_graph / args.steps}
        ms_total = ms_graph

    # ---- end-to-end: host buffers in, loss out, every step
    feat_h, cond_h, target_h = (t.cpu().pin_memory() for t in (feat, cond, target))
    loss_val = [0.0]

    def e2e_step():
        f = feat_h.to(dev, non_blocking=True)
        c = cond_h.to(dev, non_blocking=True)
        t = target_h.to(dev, non_blocking=True)
        loss_val[0] = float(step(f, c, t).item())

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    h2d = sum(t.numel() * t.element_size() for t in ((feat_h, target_h) if full else (feat_h, cond_h, target_h)))

    # ---- the same step with train-mode dropout switched off (BASELINE.md section 3 asks for both): untimed warm-up, then K steps
    off = None
    if args.dropout == "on" and not args.graph:
        hvc.set_dropout_policy("ignore")
        for _ in range(2):
            step(feat, cond, target)
        prof_off = {}
        K.set_profiler(prof_off)
        ms_off = timed(lambda: step(feat, cond, target), args.steps)
        K.set_profiler(None)
        torch.cuda.synchronize()
        hvc.set_dropout_policy("apply")
        bo = sorted(((a.elapsed_time(b), f) for a, b, f in prof_off.get("attn_bwd", [])), key=lambda z: -z[1])
        bo = [z for z in bo if z[1] == bo[0][1]] if bo else []
        off = {"ms_per_step": ms_off / args.steps, "attn_bwd_tflops": (sum(f for _, f in bo) / (sum(t for t, _ in bo) * 1e9)) if bo else None}

    # per-rank evidence for the scaling numbers: every rank's own device time in the hot kernels and its SM clock under load.  The step
    # time is the MAX over ranks and the ranks meet in every all-reduce, so the slowest GPU of the box (power-capped clocks differ by a
    # few per cent between the GPUs of one box) sets the pace of all of them.
    per_rank = None
    if world > 1:
        mine = torch.tensor([sum(v["ms_total"] for v in kern.values()) / args.steps, kern.get("attn_bwd", {}).get("ms_total", 0.0) / args.steps,
                             float(local_clocks.get("sm_mhz") or 0.0)], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"hot_kernel_ms_per_step": [round(float(t[0]), 3) for t in allr], "attn_bwd_ms_per_step": [round(float(t[1]), 3) for t in allr],
                    "sm_mhz_median": [float(t[2]) for t in allr],
                    "note": "device time of this rank's own gemm + attention launches (CUDA events on its stream) and its SM clock; "
                            "ms_per_step is the max over ranks"}
    def teardown():
        if world == 1:
            return
        if args.graph:
            # tearing down a communicator whose collectives live in an instantiated CUDA graph blocks in this torch/NCCL build (measured:
            # the run printed its line and then hung in destroy_process_group); leave together and let the process exit release it
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        teardown()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    fl = step_flops(w)
    if use_ckpt:    # the forward runs twice (x4 instead of x3)
        fl = {k: v * 4.0 / 3.0 for k, v in fl.items()}
    vols = world * B * args.steps
    value = vols / (ms_total / 1e3)
    roof = None
    if big:
        ach = sum(f for _, f in big) / (sum(t for t, _ in big) * 1e9)
        roof = {"kernel": f"attn_bwd_kernel<{w['voxel_dim'] // w['heads']}> (self-attention; call also includes the delta and dq-convert passes)",
                "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": (NCU_ATTN_BWD_DRAM_BYTES_PER_SAMPLE_32K.get(w["voxel_dim"] // w["heads"], 0.0) * B
                            if oracle_cfg(w).num_tokens == 32768 and w["voxel_dim"] == 256 else None),
                "traffic_note": "bytes per launch = ncu DRAM read+write of a one-sample launch x batch (profiles/r02_ncu_full_attn_d64_b1.csv; d=32: r01_ncu_full_hot_kernels_b1.csv)",
                "peak_source": peak_src, "launches": len(big), "ms_per_launch": sum(t for t, _ in big) / len(big),
                "frac_of_nominal_2250": ach / 2250.0}
    line = {
        "metric": "train volumes/sec (DirectCTRegression fwd+bwd)" if full else "train volumes/sec (3D ViT backbone fwd+bwd)",
        "value": value, "unit": "volumes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_gpu": B, "global_batch": B * world, "voxel_dim": w["voxel_dim"],
                   "heads": w["heads"], "depth": w["depth"], "tokens": oracle_cfg(w).num_tokens, "context_tokens": hw * hw,
                   "parallelism": f"dp{world}", "dropout": 0.1 if args.dropout == "on" else 0.0, "optimizer": "FlatAdamW (hvc_optim.cu, clip fused)" if flat_opt else "AdamW(fused)", "grad_clip": args.clip,
                   "loss": "L1 + 0.5*(1-SSIM3D) (DirectRegressionLoss, hvc_loss.cu)" if args.loss == "direct" else "L1",
                   "l2_note": "inputs+activations per step (>10 GB) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": vols / (ms_e2e / 1e3), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "loss": loss_val[0]},
        "gpu_launches": launches,
        "roofline": roof,
        "model_tflops": {"algorithmic_tflop_per_volume": fl["total"] / 1e12, "achieved": value * fl["total"] / 1e12 / world,
                         "frac_of_measured_peak": value * fl["total"] / 1e12 / world / peak_tf,
                         "attention_share_of_flops": (fl["self_attn"] + fl["cross_attn"]) / fl["total"]},
        "kernels": kern,
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")},
    }
    if off:
        line["dropout_off"] = {"value": vols / (off["ms_per_step"] * args.steps / 1e3), "unit": "volumes/s", "ms_per_step": off["ms_per_step"],
                               "roofline_achieved": off["attn_bwd_tflops"],
                               "roofline_frac": off["attn_bwd_tflops"] / peak_tf if off["attn_bwd_tflops"] else None,
                               "note": "same step, nn.Dropout sites skipped (the mode the parity tests and the CPU arm run in)"}
    if per_rank:
        line["per_rank"] = per_rank
    if graph_info:
        line["cuda_graph"] = graph_info
        line["config"]["step_launch"] = "one CUDA graph per step (value, e2e); roofline / kernels / clocks from the eager run before it"
    if world == 1 and not args.skip_cpu_baseline:
        cb = cpu_baseline(w, rows=args.cpu_rows, train=args.dropout == "on")
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "measured_fraction", "dropout")}
    if world == 1 and not args.skip_eager:
        # the reference's own algorithm as PyTorch eager on this same GPU (a fresh process; this one's tensors are released first)
        try:
            del model, opt, gb
            torch.cuda.empty_cache()
            eg = _run_self(["--impl", "eager", "--steps", "5", "--warmup", "2"])
            line["eager_b200"] = {k: eg[k] for k in ("what", "block_32768_tokens_b1_fwd_bwd", "direct64_b8_training_step", "torch") if k in eg}
        except Exception as e:      # pragma: no cover
            line["eager_b200"] = {"error": str(e)[:300]}
    print(json.dumps(line), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager"],
                    help="b200: this package; reference: the reference's algorithm on the host cores (oracle port); eager: the same algorithm "
                         "as PyTorch eager on the GPU under autocast(bf16)")
    ap.add_argument("--workload", default="direct128", choices=sorted(WORKLOADS),
                    help="direct128 = BASELINE.json's metric configuration (default); the others are the remaining configs")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--loss", default="direct", choices=["l1", "direct"],
                    help="direct (default) = DirectRegressionLoss L1 + 0.5 (1 - SSIM3D) on the package's loss kernels (config_direct.json); l1 = plain L1")
    ap.add_argument("--graph", action="store_true", help="capture the training step (incl. the bucket all-reduces at N > 1) in a CUDA graph and time its replay")
    ap.add_argument("--dropout", default="on", choices=["on", "off"],
                    help="train-mode nn.Dropout(p=0.1) of the six sites per block, as the reference trainers run (masks regenerated inside "
                         "the kernels); 'on' also reports the dropout-off step as line['dropout_off']")
    ap.add_argument("--optimizer", default="torch", choices=["torch", "flat"],
                    help="torch: clip_grad_norm_(foreach) + AdamW(fused); flat: hvc.FlatAdamW on the gradient buckets (hvc_optim.cu)")
    ap.add_argument("--clip", type=float, default=1.0, help="gradient-norm clip (config_direct.json: 1.0; 0 = off)")
    ap.add_argument("--cpu-rows", type=int, default=1024, help="query rows in the CPU baseline sample")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-eager", action="store_true", help="do not time the GPU-eager arm beside the kernels (eager_b200 key)")
    ap.add_argument("--skip-a0", action="store_true", help="--impl reference: skip the fully measured config-A0 steps")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    elif args.impl == "eager":
        if int(os.environ.get("RANK", "0")) == 0:
            run_eager(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (sm_100); there is no CPU fallback")
        run_b200(args, w)


if __name__ == "__main__":
    main()
