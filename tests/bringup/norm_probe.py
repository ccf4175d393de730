"""Bring-up probe for the HBM-bound kernels (LN fwd/bwd, residual bwd, casts, AdaLN)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from hybrid_vit_cascade_b200 import kernels as K

bad = 0


def report(name, a, b, tol):
    global bad
    e = float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))
    ok = e < tol
    bad += 0 if ok else 1
    print(f"{name}: relerr {e:.3e} {'OK' if ok else 'FAIL'}")


g = torch.Generator(device="cuda").manual_seed(5)
for (B, N, C) in ((2, 100, 32), (3, 257, 256), (2, 64, 512), (1, 40, 1024), (2, 33, 384)):
    T = B * N
    x = torch.randn(T, C, device="cuda", generator=g) * 2 + 0.5
    w = torch.randn(C, device="cuda", generator=g)
    b = torch.randn(C, device="cuda", generator=g)
    mod = torch.randn(B, 6 * C, device="cuda", generator=g) * 0.3
    shift, scale = mod[:, :C], mod[:, C:2 * C]
    # ---- modulated LN fwd
    y, mean, rstd = K.ln_fwd(x, w, b, shift, scale, 6 * C, N)
    xr = x.clone().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    modr = mod.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), wr, br, 1e-5).view(B, N, C) * (1 + modr[:, None, C:2 * C]) + modr[:, None, :C]
    report(f"ln_fwd mod C={C}", y, yr.reshape(T, C), 1e-2)
    yf, _, _ = K.ln_fwd(x, w, b, out_dtype=torch.float32)
    report(f"ln_fwd plain f32 C={C}", yf, F.layer_norm(x, (C,), w, b, 1e-5), 1e-5)
    # ---- modulated LN bwd
    dz = torch.randn(T, C, device="cuda", generator=g).bfloat16()
    resid_g = torch.randn(T, C, device="cuda", generator=g)
    out = K.ln_bwd(dz, x, mean, rstd, w, b, B, N, scale=scale, mod_ld=6 * C, dx_in=resid_g, want_mod=True)
    yr.backward(dz.float().view(B, N, C))
    report(f"ln_bwd dx C={C}", out["dx"], xr.grad + resid_g, 1e-4)
    report(f"ln_bwd dw C={C}", out["dw"], wr.grad, 1e-4)
    report(f"ln_bwd db C={C}", out["db"], br.grad, 1e-4)
    report(f"ln_bwd dshift C={C}", out["dmod"][:, 0], modr.grad[:, :C], 1e-4)
    report(f"ln_bwd dscale C={C}", out["dmod"][:, 1], modr.grad[:, C:2 * C], 1e-4)
    # ---- head: LN -> dot(wo)+bo, backward from a per-row scalar
    wo = torch.randn(C, device="cuda", generator=g)
    yh, mean2, rstd2 = K.ln_fwd(x, w, b, out_dtype=torch.float32)
    xr2 = x.clone().requires_grad_(True)
    wr2, br2, wor = w.clone().requires_grad_(True), b.clone().requires_grad_(True), wo.clone().requires_grad_(True)
    bo = torch.zeros(1, device="cuda", requires_grad=True)
    v = F.layer_norm(xr2, (C,), wr2, br2, 1e-5) @ wor + bo
    dv = torch.randn(T, device="cuda", generator=g)
    v.backward(dv)
    out = K.ln_bwd(dv, x, mean2, rstd2, w, b, B, N, mult_vec=wo, head=True)
    report(f"head dx C={C}", out["dx"], xr2.grad, 1e-4)
    report(f"head dlnw C={C}", out["dw"], wr2.grad, 1e-4)
    report(f"head dlnb C={C}", out["db"], br2.grad, 1e-4)
    report(f"head dwo C={C}", out["dvec"], wor.grad, 1e-4)
    report(f"head dbo C={C}", out["dscalar"], bo.grad, 1e-4)
    # ---- gated residual backward
    dout = torch.randn(T, C, device="cuda", generator=g)
    branch = torch.randn(T, C, device="cuda", generator=g).bfloat16()
    gate = mod[:, 2 * C:3 * C]
    dbr, dgate, dbias = K.resid_bwd(dout, B, N, branch=branch, gate=gate, gate_ld=6 * C)
    ref_dbr = dout.view(B, N, C) * gate[:, None, :]
    report(f"resid dbranch C={C}", dbr, ref_dbr.reshape(T, C), 1e-2)
    report(f"resid dgate C={C}", dgate, (dout.view(B, N, C) * branch.float().view(B, N, C)).sum(1), 1e-4)
    report(f"resid dbias C={C}", dbias, ref_dbr.sum((0, 1)), 1e-4)
    dbr2, dg2, dbias2 = K.resid_bwd(dout, B, N)
    report(f"resid nogate dbranch C={C}", dbr2, dout, 1e-2)
    report(f"resid nogate dbias C={C}", dbias2, dout.sum(0), 1e-4)
    # ---- colsum, casts
    report(f"colsum C={C}", K.colsum_bf16(branch), branch.float().sum(0), 1e-4)
    report(f"cast C={C}", K.cast_bf16(x), x.bfloat16(), 1e-7)
    ctx_nchw = torch.randn(B, C, N, device="cuda", generator=g)
    report(f"cast_tokens transposed C={C}", K.cast_tokens(ctx_nchw.transpose(1, 2)),
           ctx_nchw.transpose(1, 2).reshape(T, C).bfloat16(), 1e-7)
    report(f"cast_tokens contiguous C={C}", K.cast_tokens(x.view(B, N, C)), x.bfloat16(), 1e-7)

# ---- AdaLN linear
for (B, Kd, J) in ((8, 1024, 1536), (3, 48, 192), (2, 1280, 2304)):
    cond = torch.randn(B, Kd, device="cuda", generator=g)
    W = torch.randn(J, Kd, device="cuda", generator=g) * 0.02
    bias = torch.randn(J, device="cuda", generator=g)
    out = K.adaln_fwd(cond, W, bias)
    report(f"adaln fwd {B}x{Kd}x{J}", out, cond @ W.t() + bias, 1e-5)
    dp = torch.randn(B, J, device="cuda", generator=g)
    dW, db, dcond = K.adaln_bwd(dp, cond, W)
    report(f"adaln dW", dW, dp.t() @ cond, 1e-5)
    report(f"adaln db", db, dp.sum(0), 1e-5)
    report(f"adaln dcond", dcond, dp @ W, 1e-5)
torch.cuda.synchronize()
print("norm_probe failures:", bad)
sys.exit(1 if bad else 0)
