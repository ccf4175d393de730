"""In-kernel timeline of attn_bwd_kernel (switched on through hvc_debug_bwd_trace_enable): prints, for 8 steady-state iterations of
CTA (0,0), the SM-clock offsets of the protocol points of warpgroup 0, warpgroup 1 and the MMA warp."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import _lib, kernels as K  # noqa: E402

B, H, N, d = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 4, 8192, 64)))
DROP = len(sys.argv) > 5 and sys.argv[5] == "drop"      # python bwd_trace.py 1 4 8192 64 drop
Cc = H * d
g = torch.Generator(device="cuda").manual_seed(3)
q, k, v, d_o = (torch.randn(B * N, Cc, device="cuda", generator=g).bfloat16() for _ in range(4))
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
drop = K.Drop(torch.tensor([123, -456], dtype=torch.int32, device="cuda"), 5, 0.1) if DROP else None
o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5, drop=drop)
assert _lib.lib().hvc_debug_bwd_trace_enable(1) == 0
for _ in range(2):
    K.attn_bwd(q, k, v, o, lse, d_o, B, H, N, N, d, d ** -0.5, dq, dk, dv, drop=drop)
torch.cuda.synchronize()
IT, PTS, ROLES = 8, 12, 3
buf = (C.c_ulonglong * (ROLES * IT * PTS))()
rc = _lib.lib().hvc_debug_bwd_trace(buf)
assert rc == 0, rc
t = [[[buf[(r * IT + i) * PTS + p] for p in range(PTS)] for i in range(IT)] for r in range(ROLES)]
t0 = min(x for r in t for it in r for x in it if x)
names = {0: "wg0", 1: "wg1", 2: "mma"}
el = ["iter_top", "S_loaded", "turn_start", "A_done", "PT_arrived", "dP_loaded", "B_done", "DS_arrived", "(drain)DQF_ready", "drained"]
mm = ["iter_start", "ST0_issued", "(tail i-1)DS1_ready", "(tail i-1)dQ_issued", "PT0_ready", "ST1_issued", "DS0_ready", "PT1_ready"]
for r in range(ROLES):
    print(f"--- {names[r]}  ({', '.join(mm if r == 2 else el)})")
    for i in range(IT):
        n = 8 if r == 2 else 10
        print(f" it{i}: " + " ".join(f"{max(t[r][i][p] - t0, 0):7d}" for p in range(n)))
print("per-iteration period (wg0 iter_top):", [t[0][i + 1][0] - t[0][i][0] for i in range(IT - 1)])
