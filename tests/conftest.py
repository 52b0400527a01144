import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device: gpu-marked tests skip instead of erroring (the product
    has no CPU path to fall back to)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device: gpu-marked tests run on the B200 box (-m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Make sure the product library, the oracle and (when the reference tree is present) the
    reference builds exist.  Building the checker is not using it."""
    import harness
    harness.ensure_built()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import harness
    return harness.Oracle()


@pytest.fixture(scope="session")
def ref_serial(built):
    import harness
    s = harness.ref_shim("serial")
    if s is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return s


@pytest.fixture(scope="session")
def ref_omp(built):
    import harness
    s = harness.ref_shim("omp")
    if s is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return s


@pytest.fixture(scope="session")
def b200(built):
    """The product shim on a GPU box."""
    import lis_b200
    return lis_b200.load_shim()
