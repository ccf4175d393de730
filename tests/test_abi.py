"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports exactly what include/hvc.h declares."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "hvc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|uint64_t|const char\*)\s+(hvc_\w+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    fns = _header_functions()
    for must in ("hvc_gemm", "hvc_attn_fwd", "hvc_attn_bwd", "hvc_ln_fwd", "hvc_ln_bwd", "hvc_version", "hvc_last_error"):
        assert must in fns


def test_library_exports_every_declared_symbol():
    from hybrid_vit_cascade_b200 import _lib
    from hybrid_vit_cascade_b200.build import build
    build()
    L = _lib.lib()
    fns = _header_functions()
    for name in fns:
        assert hasattr(L, name), f"{name} declared in include/hvc.h but not exported"
    assert sorted(_lib.EXPORTS) == fns, "python binding list and header disagree"
    assert L.hvc_version() == 100


def test_struct_sizes_match_c_layout():
    """ctypes mirrors must have the C struct sizes (checked by the library via the leading `size` field)."""
    from hybrid_vit_cascade_b200 import _lib
    L = _lib.lib()
    # a call with a correctly sized but empty struct must be rejected for its CONTENT (invalid), not its size
    for struct, fn in ((_lib.GemmArgs, "hvc_gemm"), (_lib.AttnArgs, "hvc_attn_fwd"), (_lib.LnArgs, "hvc_ln_fwd"),
                       (_lib.LnBwdArgs, "hvc_ln_bwd"), (_lib.ResidBwdArgs, "hvc_resid_bwd")):
        a = struct()
        a.size = ctypes.sizeof(struct)
        rc = getattr(L, fn)(ctypes.byref(a), ctypes.c_void_p(0))
        msg = L.hvc_last_error().decode()
        assert rc == -1 and "bad args struct" not in msg, (fn, rc, msg)
        a.size = 4
        rc = getattr(L, fn)(ctypes.byref(a), ctypes.c_void_p(0))
        assert rc == -1 and "bad args struct" in L.hvc_last_error().decode()


def test_no_cpu_fallback():
    import torch
    import hybrid_vit_cascade_b200 as hvc
    from hybrid_vit_cascade_b200._lib import HvcError
    m = hvc.MultiHeadSelfAttention(64, num_heads=1).eval()
    with pytest.raises(HvcError):
        m(torch.zeros(1, 8, 64))


def test_state_dict_layout_matches_golden_table():
    """Same keys and shapes as the reference for every constructor case recorded from it."""
    import json
    import hybrid_vit_cascade_b200 as hvc
    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "ctor_table.json")))
    for row in rows:
        kw = dict(row["kwargs"])
        kw["volume_size"] = tuple(kw["volume_size"])
        m = hvc.HybridViT3D(**kw)
        assert list(m.downsampled_size) == row["downsampled_size"]
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == row["shapes"]
        assert list(m.state_dict().keys()) == list(row["shapes"].keys())


def test_same_seed_same_init_as_reference_fixture():
    """Modules are built in the reference's order: same manual_seed -> same initial weights (vit_s2 fixture)."""
    import torch
    import hybrid_vit_cascade_b200 as hvc
    c = torch.load(os.path.join(ROOT, "tests", "golden", "backbones.pt"), weights_only=False)["vit_s2"]
    torch.manual_seed(0)
    m = hvc.HybridViT3D(**c["kwargs"])
    for k, v in m.state_dict().items():
        if "adaln" in k:
            assert float(v.abs().max()) == 0.0       # zero-init like the reference (fixture re-randomised it)
        else:
            assert torch.equal(v, c["sd"][k]), k
    m.load_state_dict(c["sd"], strict=True)


def test_weight_cache_is_not_fooled_by_id_reuse():
    """The bf16 / six-term weight caches are keyed by id(parameter); CPython reuses ids (and the allocator reuses
    addresses) once a model is garbage collected, so an entry must also prove it belongs to the live object."""
    import weakref

    import torch

    from hybrid_vit_cascade_b200 import ops
    cache = {}
    p = torch.nn.Parameter(torch.zeros(4, 4))
    ops._cache_put(cache, p, None, "A")
    assert ops._cache_get(cache, p, None) == "A"
    q = torch.nn.Parameter(torch.zeros(4, 4))

    class Dead:
        pass

    # what a collected parameter leaves behind when its id, version, address and shape all coincide with q's
    cache[(id(q), None)] = (weakref.ref(Dead()), q._version, q.data_ptr(), "stale", tuple(q.shape))
    assert ops._cache_get(cache, q, None) is None
    with torch.no_grad():
        p.add_(1)
    assert ops._cache_get(cache, p, None) is None       # an in-place update (optimizer step) invalidates the entry


def test_implicit_conv_tap_offsets_describe_conv3d():
    """Host logic of the implicit-GEMM Conv3d (include/hvc.h, hvc_conv_taps): the per-tap row shifts that kernels.conv_tap_offsets /
    conv_tap_offsets_s2 hand to the GEMM's TMA producer, applied here as plain row gathers on the same padded layouts (built with
    torch on the CPU), must reproduce F.conv3d for stride 1 and stride 2."""
    import torch.nn.functional as F
    from hybrid_vit_cascade_b200 import kernels as K
    g = torch.Generator().manual_seed(0)
    B, C, Co, D, H, W = 2, 3, 4, 4, 6, 8
    x = torch.randn(B, C, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, C, 3, 3, 3, generator=g, dtype=torch.float64)
    w_taps = w.permute(0, 2, 3, 4, 1).reshape(Co, 27, C)                       # [Cout, tap, cin]

    def gather(mat, rows, off):                                                # rows outside the matrix read as zero (TMA fill)
        idx = rows + off
        ok = (idx >= 0) & (idx < mat.shape[0])
        out = torch.zeros(len(rows), mat.shape[1], dtype=mat.dtype)
        out[ok] = mat[idx[ok]]
        return out

    # stride 1: (B, D+2, H+2, W+2, C) padded volume, output on the same padded grid
    xp = F.pad(x.permute(0, 2, 3, 4, 1), (0, 0, 1, 1, 1, 1, 1, 1)).reshape(-1, C)
    rows = torch.arange(xp.shape[0])
    z = sum(gather(xp, rows, o) @ w_taps[:, t].T for t, o in enumerate(K.conv_tap_offsets(H, W)))
    z = z.view(B, D + 2, H + 2, W + 2, Co)[:, 1:-1, 1:-1, 1:-1].permute(0, 4, 1, 2, 3)
    assert torch.allclose(z, F.conv3d(x, w, padding=1), atol=1e-12)
    # stride 2: eight parity volumes, each padded by one voxel on the low side, stacked along the rows
    Dh, Hh, Wh = D // 2, H // 2, W // 2
    vols = []
    for par in range(8):
        v = x[:, :, (par >> 2)::2, ((par >> 1) & 1)::2, (par & 1)::2].permute(0, 2, 3, 4, 1)
        vols.append(F.pad(v, (0, 0, 1, 0, 1, 0, 1, 0)))
    xs = torch.stack(vols).reshape(-1, C)
    rows_p = B * (Dh + 1) * (Hh + 1) * (Wh + 1)
    rows = torch.arange(rows_p)
    tt = K.conv_tap_offsets_s2(rows_p, Hh + 1, Wh + 1)
    z = sum(gather(xs, rows, par * rows_p + sh) @ w_taps[:, t].T for t, (par, sh) in enumerate(tt))
    z = z.view(B, Dh + 1, Hh + 1, Wh + 1, Co)[:, 1:, 1:, 1:].permute(0, 4, 1, 2, 3)
    assert torch.allclose(z, F.conv3d(x, w, stride=2, padding=1), atol=1e-12)
    # the parity grouping the data gradient uses (ops._S2_TAPS / _S2_BOUNDS) is the same table
    from hybrid_vit_cascade_b200 import ops
    for par in range(8):
        assert ops._S2_TAPS[ops._S2_BOUNDS[par]:ops._S2_BOUNDS[par + 1]] == [t for t, (pp, _) in enumerate(tt) if pp == par]
