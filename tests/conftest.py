import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kb():
    import knockoffs_b200  # noqa: F401  (loader at the repo root)
    return sys.modules["knockoffs_b200"]


@pytest.fixture(scope="session")
def ctx(kb):
    """GPU context; the product has no CPU fallback, so a missing device/library is a hard failure."""
    return kb.Context(0)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")) as f:
        return json.load(f)


def toeplitz(p, rho):
    idx = np.arange(p)
    return np.asfortranarray(rho ** np.abs(idx[:, None] - idx[None, :]).astype(float))
