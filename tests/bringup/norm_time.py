"""HBM-bound kernels at the benchmark shape (T = 8 x 32768 tokens, C = 256): achieved GB/s of ln_fwd / ln_bwd / resid_bwd against the
measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).   python tests/bringup/norm_time.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

B, N, C = 8, 32768, 256
T = B * N
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(T, C, device="cuda", generator=g)
w, b = torch.randn(C, device="cuda", generator=g), torch.randn(C, device="cuda", generator=g)
mod = torch.randn(B, 6 * C, device="cuda", generator=g) * 0.3
shift, scale, gate = mod[:, :C], mod[:, C:2 * C], mod[:, 2 * C:3 * C]
dz = torch.randn(T, C, device="cuda", generator=g).bfloat16()
dres = torch.randn(T, C, device="cuda", generator=g)
branch = torch.randn(T, C, device="cuda", generator=g).bfloat16()
y, mean, rstd = K.ln_fwd(x, w, b, shift, scale, 6 * C, N)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, fn, bytes_per_elem in (
        ("ln_fwd (f32 in, bf16 out)", lambda: K.ln_fwd(x, w, b, shift, scale, 6 * C, N), 6),
        ("ln_bwd (x f32, dz bf16, dx_in f32 -> dx f32)", lambda: K.ln_bwd(dz, x, mean, rstd, w, b, B, N, scale=scale, mod_ld=6 * C, dx_in=dres, want_mod=True), 14),
        ("resid_bwd (dout f32, branch bf16 -> dbranch bf16)", lambda: K.resid_bwd(dres, B, N, branch=branch, gate=gate, gate_ld=6 * C), 8)):
    ms = t(fn)
    gbs = bytes_per_elem * T * C / ms / 1e6
    print(f"{name:55s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  = {gbs / peak:.2f} of {peak:.0f} GB/s measured copy bandwidth")
