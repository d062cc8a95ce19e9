import os
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `-m gpu`")


@pytest.fixture(scope="session")
def golden():
    """Fixtures produced by the REAL reference modules (tests/golden/make_golden.py)."""
    path = os.path.join(REPO, "tests", "golden", "reference_golden.pt")
    return torch.load(path, weights_only=False)


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


# manual fixtures of the reference's tests/data_generator.py:129-159 (users -> articles)
FIXTURE_GRAPHS = {
    "random": (3, 6, [[0, 0, 0, 1, 1, 2, 2], [0, 2, 4, 1, 5, 3, 0]]),
    "star": (5, 4, [[0, 0, 0, 0, 1, 2, 3, 4], [0, 1, 2, 3, 0, 1, 2, 3]]),
}
