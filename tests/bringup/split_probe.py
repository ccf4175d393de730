"""Numerical probe: accuracy of tcgen05 fp32 accumulation and of the three-term bf16 split (bring-up)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402


def err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def parts(x):
    p0 = x.bfloat16()
    r1 = x - p0.float()
    p1 = r1.bfloat16()
    p2 = (r1 - p1.float()).bfloat16()
    return p0, p1, p2


g = torch.Generator(device="cuda").manual_seed(1)
for (M, N, Kd) in [(256, 128, 64), (256, 128, 1024), (256, 64, 32768), (130, 72, 264)]:
    a = torch.randn(M, Kd, device="cuda", generator=g)
    b = torch.randn(N, Kd, device="cuda", generator=g)
    ref = a.double() @ b.double().t()
    a0, a1, a2 = parts(a)
    b0, b1, b2 = parts(b)
    # (a) plain bf16 GEMM against the exact product of the bf16-rounded operands: accumulation error alone
    o = K.gemm(a0.contiguous(), b0.contiguous(), epilogue=K.EPI_F32)
    print(f"[{M}x{N}x{Kd}] a0*b0 vs exact(a0,b0): max_rel %.2e fro %.2e" % err(o, a0.double() @ b0.double().t()))
    # (b) six-term K concat
    o6 = K.gemm(K.split3(a, 0), K.split3(b, 1), epilogue=K.EPI_F32)
    print(f"[{M}x{N}x{Kd}] split3 K-concat: max_rel %.2e fro %.2e" % err(o6, ref))
    chk = torch.cat([a0, a1, a2, a0, a1, a0], 1)
    print("   split3 kernel == torch parts:", bool((K.split3(a, 0) == chk).all()),
          bool((K.split3(b, 1) == torch.cat([b0, b0, b0, b1, b1, b2], 1)).all()))
    # (c) separate GEMMs by magnitude class, summed in fp32 on the CUDA cores (small terms first)
    t2 = K.gemm(torch.cat([a2, a1, a0], 1).contiguous(), torch.cat([b0, b1, b2], 1).contiguous(), epilogue=K.EPI_F32)
    t1 = K.gemm(torch.cat([a1, a0], 1).contiguous(), torch.cat([b0, b1], 1).contiguous(), epilogue=K.EPI_F32)
    t0 = K.gemm(a0.contiguous(), b0.contiguous(), epilogue=K.EPI_F32)
    print(f"[{M}x{N}x{Kd}] 3 GEMMs summed: max_rel %.2e fro %.2e" % err((t2 + t1) + t0, ref))
    print(f"[{M}x{N}x{Kd}] torch fp32 matmul (TF32 off): max_rel %.2e fro %.2e" % err(a @ b.t(), ref))
    # (d) split-K chunks of 256 for the leading term
    if Kd >= 1024:
        acc = torch.zeros(M, N, device="cuda")
        for k0 in range(0, Kd, 256):
            acc += K.gemm(a0[:, k0:k0 + 256].contiguous(), b0[:, k0:k0 + 256].contiguous(), epilogue=K.EPI_F32)
        print(f"[{M}x{N}x{Kd}] a0*b0 in K=256 chunks: max_rel %.2e fro %.2e" % err(acc, a0.double() @ b0.double().t()))
