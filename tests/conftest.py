import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """Build libb200reg.so and the oracle once per session (no GPU needed)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from platipy_b200.engine import Engine

    return Engine.get(0)
