"""One launch each of the MLP GEMMs with a GELU / GELU' epilogue and of the plain qkv projection (T = 262144, C = 256): the command
profiled with `ncu --set full -k regex:gemm_bf16 -c 3`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

T, C = 262144, 256
g = torch.Generator(device="cuda").manual_seed(1)
rnd = lambda *s, dtype=torch.bfloat16: (torch.randn(*s, device="cuda", generator=g) * 0.5).to(dtype)
x, w1, w2, w_qkv = rnd(T, C), rnd(4 * C, C), rnd(C, 4 * C), rnd(3 * C, C)
bias = rnd(4 * C, dtype=torch.float32)
pre = torch.empty(T, 4 * C, device="cuda", dtype=torch.bfloat16)
K.gemm(x, w1, bias=bias, activation=K.ACT_GELU, out2=pre)
K.gemm(x, w2, b_major=1, activation=K.ACT_GELU_GRAD, aux=pre)
K.gemm(x, w_qkv)
torch.cuda.synchronize()
print("ok")
