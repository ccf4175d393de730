"""Attention forward timing + accuracy at the full sequence length (bring-up / tuning probe).

    python tests/bringup/attn_fwd_time.py [B H N d]      # prints ms, TFLOP/s and the error against chunked fp32
With a library built with -DHVC_TUNE_FWD_EMU the polynomial-exp2 share is taken from $HVC_FWD_EMU.
"""
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
for _ in range(3):
    o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# accuracy: first head, 1024 query rows against fp32
qf, kf, vf = q[:1024, :d].float(), k[:N, :d].float(), v[:N, :d].float()
ref = ((qf @ kf.t()) * d ** -0.5).softmax(-1) @ vf
err = float((o[:1024, :d].float() - ref).abs().max() / ref.abs().max())
lse_ref = torch.logsumexp((qf @ kf.t()) * d ** -0.5, -1) * 1.4426950408889634
lerr = float((lse[0, 0, :1024] - lse_ref).abs().max())
print(f"emu={os.environ.get('HVC_FWD_EMU', 'default')} B={B} H={H} N={N} d={d}: {ms:.3f} ms  "
      f"{4.0 * B * H * N * N * d / ms / 1e9:.1f} TFLOP/s  max_rel_err={err:.2e} lse_err={lerr:.2e}")
