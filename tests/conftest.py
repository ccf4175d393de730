import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def rebuild_from_seed(cls, c, **extra):
    """Fixtures of the larger models do not store their weights: the drop-in modules are built in the reference's order, so
    torch.manual_seed(seed) (+ the recorded AdaLN re-randomisation) reproduces the reference's weights bit for bit; the stored
    per-tensor checksums prove it."""
    import torch
    torch.manual_seed(c["seed"])
    m = cls(**dict(c.get("kwargs", {}), **extra))
    if "adaln_seed" in c:
        ga = torch.Generator().manual_seed(c["adaln_seed"])
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "adaln.linear" in n:
                    p.copy_(torch.randn(p.shape, generator=ga) * 0.02)
    sd = m.state_dict()
    for k, v in c["wsum"].items():
        got = float(sd[k].double().sum())
        assert abs(got - v) <= 1e-9 * max(1.0, abs(v)), f"weights of {cls.__name__} not reproduced from the seed: {k} {got} vs {v}"
    return m
