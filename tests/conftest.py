import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np

    here = os.path.join(ROOT, "tests", "golden")
    arrays = np.load(os.path.join(here, "reference_vectors.npz"))
    with open(os.path.join(here, "reference_vectors.json")) as fh:
        meta = json.load(fh)
    return arrays, meta


@pytest.fixture(scope="session")
def native():
    """The CUDA library; a GPU test FAILS (never skips) when it cannot run."""
    from fastselect_b200 import _native

    _native.load()
    assert _native.device_count() >= 1, "no usable sm_100 GPU: GPU parity tests cannot run"
    return _native
