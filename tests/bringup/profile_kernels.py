"""One launch of each hot kernel at the benchmark's shapes with one sample (T = 32768 tokens, C = 256): the ncu target for
the per-kernel captures under profiles/ (tensor-pipe % for the GEMM / attention kernels, DRAM GB/s for the norm kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402
from hybrid_vit_cascade_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N, C, H = 32768, 256, 4
T = B * N
g = torch.Generator(device="cuda").manual_seed(1)


def rnd(*shape, dtype=torch.bfloat16):
    return (torch.randn(*shape, device="cuda", generator=g) * 0.5).to(dtype)


x32 = rnd(T, C, dtype=torch.float32)
mod = rnd(B, 6 * C, dtype=torch.float32) * 0.1
lnw, lnb = rnd(C, dtype=torch.float32), rnd(C, dtype=torch.float32)
w_qkv, w_p, w1, w2 = rnd(3 * C, C), rnd(C, C), rnd(4 * C, C), rnd(C, 4 * C)
b_c, b_4c = rnd(C, dtype=torch.float32), rnd(4 * C, dtype=torch.float32)
for rep in range(2):
    y, mean, rstd = K.ln_fwd(x32, lnw, lnb, mod[:, :C], mod[:, C:2 * C], 6 * C, N)                      # ln_fwd (modulated)
    qkv = K.gemm(y, w_qkv)                                                                              # gemm bf16
    o, lse = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, H, N, N, C // H, (C // H) ** -0.5)   # attn_fwd<64>
    branch = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
    out = K.gemm(o, w_p, epilogue=K.EPI_RESIDUAL, bias=b_c, resid=x32, gate=mod[:, 2 * C:3 * C], gate_ld=6 * C,
                 rows_per_batch=N, out2=branch)                                                         # gemm residual epilogue
    h = torch.empty(T, 4 * C, device="cuda", dtype=torch.bfloat16)
    gact = K.gemm(y, w1, bias=b_4c, activation=K.ACT_GELU, out2=h)                                      # gemm gelu epilogue
    dout = rnd(T, C, dtype=torch.float32)
    dbranch, dgate, dbias = K.resid_bwd(dout, B, N, branch=branch, gate=mod[:, 2 * C:3 * C], gate_ld=6 * C)   # resid_bwd
    d_o = K.gemm(dbranch, w_p, b_major=1)
    dqkv = torch.empty(T, 3 * C, device="cuda", dtype=torch.bfloat16)
    K.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse, d_o, B, H, N, N, C // H, (C // H) ** -0.5,
               dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:])                                          # attn_bwd<64>
    dw = ops._wgrad(dqkv, y)                                                                            # gemm wgrad (atomic)
    dy = K.gemm(dqkv, w_qkv, b_major=1)
    r = K.ln_bwd(dy, x32, mean, rstd, lnw, lnb, B, N, scale=mod[:, C:2 * C], mod_ld=6 * C, dx_in=dout, want_mod=True)   # ln_bwd
    # cascade heads (d = 32)
    o32, lse32 = K.attn_fwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, 8, N, N, 32, 32 ** -0.5)
    K.attn_bwd(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o32, lse32, d_o, B, 8, N, N, 32, 32 ** -0.5,
               dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:])
torch.cuda.synchronize()
print("ok")
