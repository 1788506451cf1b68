import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """Built artefacts are git-ignored; if a checkout arrives without them, build once (nvcc, g++, make)."""
    needed = [os.path.join(ROOT, "blume_capel_b200", "libbc_engine.so"), os.path.join(ROOT, "oracle", "libbc_oracle.so"),
              os.path.join(ROOT, "blume_capel_b200", "host", "bc_run"), os.path.join(ROOT, "blume_capel_b200", "host", "bc_runner"),
              os.path.join(ROOT, "blume_capel_b200", "host", "bc_gaptool"),
              os.path.join(ROOT, "blume_capel_b200", "host", "bc_magnet"), os.path.join(ROOT, "blume_capel_b200", "host", "bc_gamma_nu"),
              os.path.join(ROOT, "blume_capel_b200", "host", "bc_gap_stats")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    _ensure_built()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def bc():
    import blume_capel_b200 as b

    b.load_library()
    return b
