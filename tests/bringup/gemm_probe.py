"""Bring-up probe for hvc_gemm: each case runs in its own process (a faulting kernel poisons the
CUDA context).  Usage: python tests/bringup/gemm_probe.py [case ...]   (no args = all, each in a subprocess)
Compares against torch.matmul in fp32 on the same bf16-rounded operands."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {
    # name: (M, N, K, a_major, b_major, mode)
    "kk_small": (128, 128, 64, 0, 0, "plain"),
    "kk_256": (256, 256, 256, 0, 0, "plain"),
    "kk_qkv": (4096, 768, 256, 0, 0, "bias"),
    "kk_odd": (200, 72, 40, 0, 0, "bias"),
    "km_dgrad": (4096, 256, 768, 0, 1, "plain"),
    "mk": (256, 384, 512, 1, 0, "plain"),
    "mm_wgrad": (768, 256, 8192, 1, 1, "atomic"),
    "mm_wgrad_split": (768, 256, 32768, 1, 1, "atomic_split"),
    "mm_odd": (72, 40, 200, 1, 1, "atomic"),
    "gelu": (512, 1024, 256, 0, 0, "gelu"),
    "gelu_grad": (512, 1024, 256, 0, 1, "gelu_grad"),
    "residual": (1024, 256, 1024, 0, 0, "residual"),
    "f32": (256, 128, 128, 0, 0, "f32"),
    "big": (32768, 1024, 256, 0, 0, "bias"),
}


def run_case(name):
    import torch
    from hybrid_vit_cascade_b200 import kernels as K
    M, N, Kd, am, bm, mode = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, Kd, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, Kd, device="cuda", generator=g).bfloat16()
    a_st = A if am == 0 else A.t().contiguous()
    b_st = B if bm == 0 else B.t().contiguous()
    ref = A.float() @ B.float().t()
    bias = torch.randn(N, device="cuda", generator=g)
    kw = {}
    if mode == "plain":
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm)
    elif mode == "bias":
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, bias=bias)
        ref = ref + bias
    elif mode in ("atomic", "atomic_split"):
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, epilogue=K.EPI_F32_ATOMIC,
                     k_splits=1 if mode == "atomic" else 37, alpha=0.5)
        ref = 0.5 * ref
    elif mode == "gelu":
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, bias=bias, activation=K.ACT_GELU, out2=out2)
        pre = ref + bias
        e2 = float((out2.float() - pre).abs().max() / pre.abs().max())
        print(f"  pre-activation relerr {e2:.3e}")
        ref = torch.nn.functional.gelu(pre)
    elif mode == "gelu_grad":
        aux = torch.randn(M, N, device="cuda", generator=g).bfloat16()
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, activation=K.ACT_GELU_GRAD, aux=aux)
        x = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        ref = ref * x.grad
    elif mode == "residual":
        resid = torch.randn(M, N, device="cuda", generator=g)
        gate = torch.randn(4, 3 * N, device="cuda", generator=g)
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, epilogue=K.EPI_RESIDUAL, bias=bias, resid=resid,
                     gate=gate[:, N:2 * N], gate_ld=3 * N, rows_per_batch=M // 4, out2=out2)
        gfull = gate[:, N:2 * N].repeat_interleave(M // 4, dim=0)
        ref = resid + gfull * (ref + bias)
    elif mode == "f32":
        out = K.gemm(a_st, b_st, a_major=am, b_major=bm, epilogue=K.EPI_F32, bias=bias)
        ref = ref + bias
    torch.cuda.synchronize()
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    ok = err < (1e-2 if out.dtype == torch.bfloat16 else 2e-3)
    print(f"CASE {name}: relerr {err:.3e} {'OK' if ok else 'FAIL'}")
    if name == "big":
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            K.gemm(a_st, b_st, bias=bias)
        ev[0].record()
        for _ in range(10):
            K.gemm(a_st, b_st, bias=bias)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        print(f"  big: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s  {(M*Kd*2+M*N*2)/ms/1e6:.0f} GB/s")
    return 0 if ok else 1


if __name__ == "__main__":
    names = sys.argv[1:]
    if names:
        sys.exit(max(run_case(n) for n in names))
    bad = 0
    for n in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, n], capture_output=True, text=True, timeout=120)
            print(r.stdout.strip() or f"CASE {n}: no output")
            if r.returncode != 0:
                bad += 1
                print("  rc", r.returncode, r.stderr.strip()[-600:])
        except subprocess.TimeoutExpired:
            bad += 1
            print(f"CASE {n}: TIMEOUT")
    print("gemm_probe failures:", bad)
    sys.exit(1 if bad else 0)
