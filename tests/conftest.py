import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def reflib():
    """The unmodified reference compiled headless (oracle/_ref). Built here where /root/reference exists; the
    prebuilt .so travels to the GPU box."""
    from oracle import refapi

    if not os.path.exists(refapi.REF_SO):
        pytest.skip("oracle/_ref/libmpros_ref.so not built (needs /root/reference)")
    return refapi.RefLib()


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    import mpros_b200

    made = []

    def make(device=0):
        c = mpros_b200.Context(device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
