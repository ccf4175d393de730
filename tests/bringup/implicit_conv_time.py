"""Timing of the implicit-GEMM Conv3d (hvc_conv_taps) products at the detail_enhancer shape: Cin = 64 -> Cout = 32 on a D^3 volume.
    python tests/bringup/implicit_conv_time.py [D] [B]
Prints per product: time, executed TFLOP/s (2*M*N*K with the real N), patch-matrix-equivalent GB/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hybrid_vit_cascade_b200 import kernels as K  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Cin, Cout, Cp = 64, 32, 64
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn(B, D, D, D, Cin, device="cuda", generator=g)
dz = torch.randn(B, D, D, D, Cout, device="cuda", generator=g).bfloat16()
w_taps = (torch.randn(Cout, 27 * Cin, device="cuda", generator=g) * 0.02).bfloat16()
w_t = (torch.randn(Cin, 27 * Cp, device="cuda", generator=g) * 0.02).bfloat16()
xp = K.pad3d_cl(x, B, D, D, D, Cin, Cin).view(-1, Cin)
dzp = K.pad3d_cl(dz, B, D, D, D, Cout, Cp).view(-1, Cp)
rows = xp.shape[0]
reps = int(os.environ.get("REPS", "3"))


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tiles = (27 * Cin + 127) // 128
splits = max(1, min(rows // 64, (16 * 148) // tiles))
cases = [
    ("forward   [rows, 32, 1728]", lambda: K.gemm(xp, w_taps, epilogue=K.EPI_F32, taps=(1, Cin, K.conv_tap_offsets(D, D))), 2.0 * rows * Cout * 27 * Cin),
    ("data grad [rows, 64, 1728]", lambda: K.gemm(dzp, w_t, epilogue=K.EPI_F32, taps=(1, Cp, K.conv_tap_offsets(D, D, -1))), 2.0 * rows * Cin * 27 * Cp),
    ("wgt grad  [64, 1728, rows]", lambda: K.gemm(dzp, xp, a_major=1, b_major=1, epilogue=K.EPI_F32_ATOMIC, k_splits=splits,
                                                  taps=(2, Cin, K.conv_tap_offsets(D, D))), 2.0 * rows * Cp * 27 * Cin),
    ("pad f32 -> bf16           ", lambda: K.pad3d_cl(x, B, D, D, D, Cin, Cin), 0.0),
]
print(f"D={D} B={B} rows={rows} ({rows * Cin * 2 / 1e9:.2f} GB padded volume)")
for name, fn, fl in cases:
    ms = timeit(fn)
    print(f"{name}  {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s   volume read once = {rows * Cin * 2 / ms / 1e6:7.0f} GB/s")
