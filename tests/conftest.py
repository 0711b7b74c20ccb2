import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def check_weight_sums(sd, golden):
    """The fixture stores (sum, abs-sum) per tensor of the weights the reference outputs were made with."""
    for k, v in sd.items():
        ref = golden["wsum/" + k]
        got = np.array([float(v.double().sum()), float(v.double().abs().sum())])
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12), f"weights drifted from the golden fixture: {k}"


@pytest.fixture(scope="session")
def lib():
    import vitocm_b200
    return vitocm_b200._lib.load_library()
