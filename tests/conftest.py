import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def manifest():
    import json
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/): test infrastructure only."""
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def pt():
    import ascendpathtracing_b200 as pt
    pt.lib()
    return pt


@pytest.fixture(scope="session")
def cuda(pt):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device")
    assert pt.device_count() >= 1, "libptb200 sees no CUDA device: the product path has no fallback"
    torch.cuda.set_device(0)
    return torch
