"""One forward + one backward attention launch at a given shape (ncu target)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

B, H, nq, nk, d = (int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (1, 4, 32768, 32768, 64)))
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 2
C = H * d
g = torch.Generator(device="cuda").manual_seed(3)
q = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
k = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
v = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
d_o = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(reps):
    o, lse2 = K.attn_fwd(q, k, v, B, H, nq, nk, d, d ** -0.5)
    K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
print("ok")
