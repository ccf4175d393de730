"""Attention backward timing + accuracy at full sequence length (bring-up / tuning probe).
    python tests/bringup/attn_bwd_time.py [B H N d]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

B, H, N, d = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (2, 4, 32768, 64)))
C = H * d
g = torch.Generator(device="cuda").manual_seed(3)
qkv = torch.randn(B * N, 3 * C, device="cuda", generator=g).bfloat16()
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
d_o = torch.randn(B * N, C, device="cuda", generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
dq, dk, dv = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:]
o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5)
for _ in range(2):
    K.attn_bwd(q, k, v, o, lse, d_o, B, H, N, N, d, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    K.attn_bwd(q, k, v, o, lse, d_o, B, H, N, N, d, d ** -0.5, dq, dk, dv)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# accuracy on a 4096-token problem against fp32 autograd
n = 4096
qs, ks, vs, gs = (t[:n, :d].float().clone().requires_grad_(True) for t in (q, k, v, d_o))
ref = ((qs @ ks.t()) * d ** -0.5).softmax(-1) @ vs
ref.backward(gs.detach())
o2, lse2 = K.attn_fwd(q[:n, :d].contiguous(), k[:n, :d].contiguous(), v[:n, :d].contiguous(), 1, 1, n, n, d, d ** -0.5)
a, b_, c_ = (torch.empty(n, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
K.attn_bwd(q[:n, :d].contiguous(), k[:n, :d].contiguous(), v[:n, :d].contiguous(), o2, lse2, d_o[:n, :d].contiguous(), 1, 1, n, n, d,
           d ** -0.5, a, b_, c_)
cos = [float(torch.nn.functional.cosine_similarity(x.float().flatten(), y.grad.flatten(), dim=0)) for x, y in ((a, qs), (b_, ks), (c_, vs))]
print(f"B={B} H={H} N={N} d={d}: bwd {ms:.3f} ms  {10.0 * B * H * N * N * d / ms / 1e9:.1f} TFLOP/s  cos(dq,dk,dv)={cos}")
