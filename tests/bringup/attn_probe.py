"""Bring-up probe for hvc_attn_fwd / hvc_attn_bwd; each case in its own process."""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, H, nq, nk, d, packed_self)
    "fwd_1tile": (1, 1, 256, 128, 64, False),
    "fwd_multi": (1, 2, 512, 640, 64, False),
    "fwd_self_packed": (2, 4, 1024, 1024, 64, True),
    "fwd_ragged": (2, 2, 200, 72, 64, False),
    "fwd_bigrange": (1, 1, 256, 1024, 64, False),
    "fwd_4096": (2, 4, 4096, 4096, 64, True),
    "bwd_1tile": (1, 1, 128, 128, 64, False),
    "bwd_multi": (1, 2, 384, 256, 64, False),
    "bwd_ragged": (2, 2, 200, 72, 64, False),
    "bwd_self_packed": (2, 4, 1024, 1024, 64, True),
    "bwd_4096": (2, 4, 4096, 4096, 64, True),
    "fwd32_1tile": (1, 1, 256, 128, 32, False),
    "fwd32_multi": (1, 2, 512, 640, 32, False),
    "fwd32_ragged": (2, 2, 200, 72, 32, False),
    "bwd32_1tile": (1, 1, 128, 128, 32, False),
    "bwd32_multi": (1, 2, 384, 256, 32, False),
    "bwd32_ragged": (2, 2, 200, 72, 32, False),
    "bwd32_self_packed": (2, 8, 1024, 1024, 32, True),
    "perf_32k": (1, 4, 32768, 32768, 64, True),
    "perf32_32k": (1, 8, 32768, 32768, 32, True),
    "perf32_cross": (2, 8, 32768, 4096, 32, False),
    "perf_cross": (2, 4, 32768, 4096, 64, False),
}


def ref_attn(q, k, v, scale):
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return p @ v.float(), torch.logsumexp(s, -1)


def run_perf(name):
    import torch
    from hybrid_vit_cascade_b200 import kernels as K
    B, H, nq, nk, d, packed = CASES[name]
    C = H * d
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    k = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    v = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    d_o = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    scale = d ** -0.5
    o, lse2 = K.attn_fwd(q, k, v, B, H, nq, nk, d, scale)
    def t(fn, n=5):
        for _ in range(2):
            fn()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(n):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / n
    mf = t(lambda: K.attn_fwd(q, k, v, B, H, nq, nk, d, scale))
    mb = t(lambda: K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, scale, dq, dk, dv))
    fl = 4.0 * B * H * nq * nk * d
    print(f"PERF {name}: fwd {mf:.3f} ms {fl/mf/1e9:.1f} TFLOP/s | bwd {mb:.3f} ms {2.5*fl/mb/1e9:.1f} TFLOP/s")
    return 0


def run_case(name, bwd=False):
    global torch
    import torch
    from hybrid_vit_cascade_b200 import kernels as K
    if name.startswith("perf"):
        return run_perf(name)
    B, H, nq, nk, d, packed = CASES[name]
    g = torch.Generator(device="cuda").manual_seed(3)
    C = H * d
    amp = 4.0 if name == "fwd_bigrange" else 1.0
    if packed:
        qkv = (torch.randn(B * nq, 3 * C, device="cuda", generator=g) * amp).bfloat16()
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q = (torch.randn(B * nq, C, device="cuda", generator=g) * amp).bfloat16()
        k = (torch.randn(B * nk, C, device="cuda", generator=g) * amp).bfloat16()
        v = torch.randn(B * nk, C, device="cuda", generator=g).bfloat16()
    scale = d ** -0.5
    o, lse2 = K.attn_fwd(q, k, v, B, H, nq, nk, d, scale)
    torch.cuda.synchronize()
    q4 = q.reshape(B, nq, H, d).permute(0, 2, 1, 3)
    k4 = k.reshape(B, nk, H, d).permute(0, 2, 1, 3)
    v4 = v.reshape(B, nk, H, d).permute(0, 2, 1, 3)
    ro, rlse = ref_attn(q4, k4, v4, scale)
    ro = ro.permute(0, 2, 1, 3).reshape(B * nq, C)
    err = float((o.float() - ro).abs().max() / ro.abs().max())
    lerr = float((lse2[:, :, :nq] * math.log(2.0) - rlse).abs().max())
    ok = err < 2e-2 and lerr < 2e-2
    print(f"CASE {name}: out relerr {err:.3e} lse abserr {lerr:.3e} {'OK' if ok else 'FAIL'}")
    if name.startswith("bwd"):  # bwd / bwd32
        d_o = torch.randn(B * nq, C, device="cuda", generator=g).bfloat16()
        if packed:
            dqkv = torch.full((B * nq, 3 * C), float("nan"), device="cuda", dtype=torch.bfloat16)
            dq, dk, dv = dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:]
        else:
            dq = torch.full((B * nq, C), float("nan"), device="cuda", dtype=torch.bfloat16)
            dk = torch.full((B * nk, C), float("nan"), device="cuda", dtype=torch.bfloat16)
            dv = torch.full((B * nk, C), float("nan"), device="cuda", dtype=torch.bfloat16)
        K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, scale, dq, dk, dv)
        torch.cuda.synchronize()
        qf, kf, vf = [t.float().detach().requires_grad_(True) for t in (q4, k4, v4)]
        ro2, _ = ref_attn(qf, kf, vf, scale)
        ro2.backward(d_o.float().reshape(B, nq, H, d).permute(0, 2, 1, 3))
        for nm, mine, ref, n in (("dq", dq, qf.grad, nq), ("dk", dk, kf.grad, nk), ("dv", dv, vf.grad, nk)):
            ref2 = ref.permute(0, 2, 1, 3).reshape(B * n, C)
            e = float((mine.float() - ref2).abs().max() / ref2.abs().max())
            cs = float(torch.nn.functional.cosine_similarity(mine.float().flatten(), ref2.flatten(), dim=0))
            good = e < 3e-2 and cs > 0.999
            ok = ok and good
            print(f"  {nm}: relerr {e:.3e} cos {cs:.6f} {'OK' if good else 'FAIL'}")
        if name == "bwd_4096":
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            for _ in range(3):
                K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, scale, dq, dk, dv)
            ev[0].record()
            for _ in range(10):
                K.attn_bwd(q, k, v, o, lse2, d_o, B, H, nq, nk, d, scale, dq, dk, dv)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / 10
            fl = 10.0 * B * H * nq * nk * d
            print(f"  bwd_4096: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (incl. delta/zero/convert)")
    if name == "fwd_4096":
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            K.attn_fwd(q, k, v, B, H, nq, nk, d, scale)
        ev[0].record()
        for _ in range(10):
            K.attn_fwd(q, k, v, B, H, nq, nk, d, scale)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        fl = 4.0 * B * H * nq * nk * d
        print(f"  fwd_4096: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")
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
                print("  rc", r.returncode, r.stderr.strip()[-800:])
        except subprocess.TimeoutExpired:
            bad += 1
            print(f"CASE {n}: TIMEOUT")
    print("attn_probe failures:", bad)
    sys.exit(1 if bad else 0)
