import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding
    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def rt():
    """The product binding. Fails loudly (no fallback) when the CUDA library cannot be loaded."""
    from build_up_phase_b200 import rtcore
    rtcore.load()
    return rtcore


@pytest.fixture(scope="session")
def ctx(rt):
    c = rt.Context(0)
    yield c
    c.close()
