"""Kernel-level timing of hvc_attn_fwd / hvc_attn_bwd at the benchmark shapes, dropout on and off (tuning probe; HVC_LIB selects the build).

    python tests/bringup/attn_time.py [d64|d32|both] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "both"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
SHAPES = {"d64": (8, 4, 32768, 64), "d32": (2, 8, 32768, 32)}
tag = os.path.basename(os.environ.get("HVC_LIB", "default"))


def t(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name in (("d64", "d32") if which == "both" else (which,)):
    B, H, N, d = SHAPES[name]
    C = H * d
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    do = torch.randn(B * N, C, device="cuda", generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    seed = torch.tensor([123, -456], dtype=torch.int32, device="cuda")
    out = []
    for drop in (None, K.Drop(seed, 5, 0.1)):
        o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop)
        f = t(lambda: K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop))
        b = t(lambda: K.attn_bwd(q, k, v, o, lse, do, B, H, N, N, d, d ** -0.5, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], drop=drop))
        fl = B * H * N * N * d / 1e9
        out.append(f"{'drop' if drop else 'off '}: fwd {f:7.3f} ms {4 * fl / f:6.1f} TF | bwd {b:7.3f} ms {10 * fl / b:6.1f} TF")
    print(f"[{tag}] {name} B={B} H={H}: " + "  ||  ".join(out), flush=True)
