"""One launch each of attn_fwd / attn_bwd, dropout off then on, at one sample of the benchmark shape: the command profiled with
`ncu --set full --import-source on -k regex:attn_(fwd|bwd)_kernel -c 4` (profiles/r02_ncu_*).   python tests/bringup/attn_ncu_one.py [d] [H]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256 // d
B, N = 1, 32768
C = H * d
g = torch.Generator(device="cuda").manual_seed(3)
qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
do = torch.randn(B * N, C, device="cuda", generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
seed = torch.tensor([123, -456], dtype=torch.int32, device="cuda")
for drop in (None, K.Drop(seed, 5, 0.1)):
    o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop)
    K.attn_bwd(q, k, v, o, lse, do, B, H, N, N, d, d ** -0.5, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], drop=drop)
torch.cuda.synchronize()
print("ok")
