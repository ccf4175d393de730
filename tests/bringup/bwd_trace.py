"""In-kernel timeline of attn_bwd_kernel (library built with -DHVC_TRACE_BWD): prints, for 8 steady-state iterations of
CTA (0,0), the SM-clock offsets of the protocol points of warpgroup 0, warpgroup 1 and the MMA warp."""
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
q, k, v, d_o = (torch.randn(B * N, Cc, device="cuda", generator=g).bfloat16() for _ in range(4))
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
o, lse = K.attn_fwd(q, k, v, B, H, N, N, d, d ** -0.5)
for _ in range(2):
    K.attn_bwd(q, k, v, o, lse, d_o, B, H, N, N, d, d ** -0.5, dq, dk, dv)
torch.cuda.synchronize()
IT, PTS = 8, 12
buf = (C.c_ulonglong * (3 * IT * PTS))()
rc = _lib.lib().hvc_debug_bwd_trace(buf)
assert rc == 0, rc
t = [[[buf[(r * IT + i) * PTS + p] for p in range(PTS)] for i in range(IT)] for r in range(3)]
t0 = min(x for r in t for it in r for x in it if x)
names = {0: "wg0", 1: "wg1", 2: "mma"}
el = ["wait_ST", "S_ready", "S_loaded", "A_done", "PT_arrived", "drained", "DPT_ready", "dP_loaded", "B_done", "DS_arrived",
      "(drain)DQF_ready", "(drain)dQ_loaded"]
mm = ["iter_start", "STFREE+QF_ok", "ST_issued", "PT_ready", "dV_issued", "DS_ready", "dPT_dK_issued", "DQFREE_ok", "dQ_issued"]
for r in range(3):
    print(f"--- {names[r]}  ({', '.join(mm if r == 2 else el)})")
    for i in range(IT):
        n = 9 if r == 2 else 12
        print(f" it{i}: " + " ".join(f"{t[r][i][p] - t0:7d}" for p in range(n)))
print("per-iteration period (wg0 S_ready):", [t[0][i + 1][1] - t[0][i][1] for i in range(IT - 1)])
