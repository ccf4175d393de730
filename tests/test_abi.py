"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports exactly what include/hvc.h declares."""
import ctypes
import os
import re

import pytest

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
