"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only dev container; `-m gpu` on a B200."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def coracle():
    from oracle import oracle as O
    return O.COracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own callbacks (oracle/_ref).  Prebuilt in the dev container; travels to the GPU box."""
    from oracle import oracle as O
    if not O.Ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    return O.Ref()


@pytest.fixture(scope="session")
def pg():
    """The product, through its C ABI (ctypes).  Fails loudly if the CUDA library is missing."""
    import __graft_entry__ as G
    G.ensure_built()
    return G.load_package()


@pytest.fixture(scope="session")
def gpu(pg):
    import torch
    assert torch.cuda.is_available(), "gpu-marked test on a box without CUDA"
    h = pg.PerseusGpu(device=0)
    yield h
    h.close()
