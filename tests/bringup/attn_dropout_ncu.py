"""One forward + one backward attention launch with dropout 0.1 at the benchmark shape of one sample (H=4, N=M=32768, d=64), for
    ncu --set full -k regex:attn_(fwd|bwd)_kernel -c 2 python tests/bringup/attn_dropout_ncu.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

B, H, N, d = 1, 4, 32768, 64
C = H * d
g = torch.Generator(device="cuda").manual_seed(1)
q, k, v, do = (torch.randn(B * N, C, device="cuda", generator=g).bfloat16() for _ in range(4))
drop = K.Drop(torch.tensor([0x1234567, -0x3456789], dtype=torch.int32, device="cuda"), 5, 0.1)
o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop)
dq, dk, dv = (torch.empty_like(t) for t in (q, k, v))
K.attn_bwd(q, k, v, o, lse, do, B, H, N, N, d, d ** -0.5, dq, dk, dv, drop=drop)
torch.cuda.synchronize()
print("ok")
