"""In-kernel timeline of attn_fwd_kernel (library built with HVC_EXTRA_NVCC_FLAGS=-DHVC_TRACE_FWD; switched on through
hvc_debug_fwd_trace_enable): SM-clock offsets of the protocol
points of softmax warpgroups A / B and the MMA warp for 8 steady-state key tiles of CTA (0,0)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import _lib, kernels as K  # noqa: E402

B, H, N, d = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 4, 8192, 64)))
Cc = H * d
g = torch.Generator(device="cuda").manual_seed(3)
q, k, v = (torch.randn(B * N, Cc, device="cuda", generator=g).bfloat16() for _ in range(3))
assert _lib.lib().hvc_debug_fwd_trace_enable(1) == 0
DROP = len(sys.argv) > 5 and sys.argv[5] == "drop"      # python fwd_trace.py 1 4 8192 64 drop
drop = K.Drop(torch.tensor([123, -456], dtype=torch.int32, device="cuda"), 5, 0.1) if DROP else None
for _ in range(2):
    K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop)
torch.cuda.synchronize()
IT, PTS = 8, 8
buf = (C.c_ulonglong * (3 * IT * PTS))()
assert _lib.lib().hvc_debug_fwd_trace(buf) == 0
t = [[[buf[(r * IT + i) * PTS + p] for p in range(PTS)] for i in range(IT)] for r in range(3)]
t0 = min(x for r in t for it in r for x in it if x)
sm = ["iter_top", "S_ready", "rowmax_done", "turn_start", "exp_done", "P_arrived"]
mm = ["iter_start", "V_ready", "P0_ready", "PV0+S0_issued", "P1_ready", "PV1+S1_issued"]
for r, name in enumerate(["wgA", "wgB", "mma"]):
    print(f"--- {name}  ({', '.join(mm if r == 2 else sm)})")
    for i in range(IT):
        print(f" it{i}: " + " ".join(f"{max(t[r][i][p] - t0, 0):7d}" for p in range(6)))
print("period (wgA turn_start):", [t[0][i + 1][3] - t[0][i][3] for i in range(IT - 1)])
