import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        from prlib_b200 import capi
        return capi.load().prl_cuda_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must FAIL loudly, not skip: only skip when gpu tests were not asked for
    pass


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def real_crops():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "real_crops.npz")))


@pytest.fixture(scope="session")
def ctx():
    import prlib_b200
    c = prlib_b200.Context(0)     # raises PrlCudaError when there is no GPU: gpu tests fail loudly
    yield c
    c.close()


@pytest.fixture(scope="session")
def noise_page():
    return np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
