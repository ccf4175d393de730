"""Small attention forward+backward (ragged sizes, dropout on and off, d = 64 and 32): the command run under `compute-sanitizer --tool memcheck`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
seed = torch.tensor([123, -456], dtype=torch.int32, device="cuda")
for d, H, nq, nk in ((64, 2, 300, 520), (32, 3, 129, 257)):
    C, B = H * d, 2
    q = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    k = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    v = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    do = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    for drop in (None, K.Drop(seed, 5, 0.1)):
        o, lse = K.attn_fwd(q, k, v, B, H, nq, nk, d, d ** -0.5, drop=drop)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        K.attn_bwd(q, k, v, o, lse, do, B, H, nq, nk, d, d ** -0.5, dq, dk, dv, drop=drop)
        torch.cuda.synchronize()
        assert torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()
print("ok")
