import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement of the reference (oracle/spn_oracle.c) -- the checker."""
    from oracle.spn_oracle import COracle
    return COracle()


@pytest.fixture(scope="session")
def ref_oracle():
    """The unmodified reference CPU extension (oracle/_ref); skipped where it was never built."""
    from oracle import build_ref
    if build_ref.build() is None:
        pytest.skip("oracle/_ref not available (needs /root/reference or a prebuilt _ext)")
    from oracle.spn_oracle import RefOracle
    return RefOracle()


@pytest.fixture(scope="session")
def spn():
    """The product package with its CUDA library built."""
    from smoothparticlenets_b200 import build
    build.build_library()
    import smoothparticlenets_b200
    return smoothparticlenets_b200
