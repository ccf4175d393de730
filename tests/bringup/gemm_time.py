"""Timing of hvc_gemm on the shapes / epilogues of one direct_regression 128^3 training step (T = 8 x 32768 tokens, C = 256).
Prints per call: time, TFLOP/s, algorithmic GB/s (operands + outputs once)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402
from hybrid_vit_cascade_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
C = 256
g = torch.Generator(device="cuda").manual_seed(1)


def rnd(*shape, dtype=torch.bfloat16):
    return (torch.randn(*shape, device="cuda", generator=g) * 0.5).to(dtype)


def timeit(fn, reps=5):
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


x = rnd(T, C)
x4 = rnd(T, 4 * C)
x3 = rnd(T, 3 * C)
resid = rnd(T, C, dtype=torch.float32)
gate = rnd(8, C, dtype=torch.float32)
w_qkv, w_p, w1, w2 = rnd(3 * C, C), rnd(C, C), rnd(4 * C, C), rnd(C, 4 * C)
bias_c, bias_4c = rnd(C, dtype=torch.float32), rnd(4 * C, dtype=torch.float32)
ctx, w_kv = rnd(8 * 4096, 512), rnd(2 * C, 512)
out2_c = torch.empty(T, C, device="cuda", dtype=torch.bfloat16)
out2_4c = torch.empty(T, 4 * C, device="cuda", dtype=torch.bfloat16)
cases = [
    ("fwd qkv            [T,768,256]  bf16", lambda: K.gemm(x, w_qkv), (T, 768, 256), T * 256 * 2 + T * 768 * 2),
    ("fwd q / dgrad-like [T,256,256]  bf16", lambda: K.gemm(x, w_p), (T, 256, 256), T * 256 * 4),
    ("fwd proj residual  [T,256,256]  f32+bf16", lambda: K.gemm(x, w_p, epilogue=K.EPI_RESIDUAL, bias=bias_c, resid=resid, gate=gate,
                                                               gate_ld=C, rows_per_batch=T // 8, out2=out2_c), (T, 256, 256),
     T * 256 * 2 + T * 256 * (4 + 4 + 2)),
    ("fwd kv (context)   [32768,512,512] bf16", lambda: K.gemm(ctx, w_kv), (8 * 4096, 512, 512), 8 * 4096 * 512 * 4),
    ("fwd mlp1 gelu+pre  [T,1024,256] 2xbf16", lambda: K.gemm(x, w1, bias=bias_4c, activation=K.ACT_GELU, out2=out2_4c), (T, 1024, 256),
     T * 256 * 2 + 2 * T * 1024 * 2),
    ("fwd mlp2 residual  [T,256,1024] f32+bf16", lambda: K.gemm(x4, w2, epilogue=K.EPI_RESIDUAL, bias=bias_c, resid=resid, gate=gate,
                                                                gate_ld=C, rows_per_batch=T // 8, out2=out2_c), (T, 256, 1024),
     T * 1024 * 2 + T * 256 * 10),
    ("bwd dgrad qkv      [T,256,768]  bf16", lambda: K.gemm(x3, w_qkv, b_major=1), (T, 256, 768), T * 768 * 2 + T * 256 * 2),
    ("bwd dgrad mlp2 gelu'[T,1024,256] bf16", lambda: K.gemm(x, w2, b_major=1, activation=K.ACT_GELU_GRAD, aux=out2_4c), (T, 1024, 256),
     T * 256 * 2 + 2 * T * 1024 * 2),
    ("bwd dgrad mlp1     [T,256,1024] bf16", lambda: K.gemm(x4, w1, b_major=1), (T, 256, 1024), T * 1024 * 2 + T * 256 * 2),
    ("bwd wgrad qkv      [768,256,T]  f32 atomic", lambda: ops._wgrad(x3, x), (768, 256, T), T * 1024 * 2),
    ("bwd wgrad proj     [256,256,T]  f32 atomic", lambda: ops._wgrad(x, x), (256, 256, T), T * 512 * 2),
    ("bwd wgrad mlp1     [1024,256,T] f32 atomic", lambda: ops._wgrad(x4, x), (1024, 256, T), T * 1280 * 2),
    ("bwd wgrad mlp2     [256,1024,T] f32 atomic", lambda: ops._wgrad(x, x4), (256, 1024, T), T * 1280 * 2),
]
tot = 0.0
for name, fn, (M, N, Kd), nbytes in cases:
    ms = timeit(fn)
    tot += ms
    print(f"{name:44s} {ms * 1e3:8.1f} us  {2.0 * M * N * Kd / ms / 1e9:7.1f} TFLOP/s  {nbytes / ms / 1e6:7.0f} GB/s")
print(f"sum {tot:.3f} ms")
