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


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def real_crops():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "real_crops.npz")))


# Every GPU test that takes `ctx` runs twice: with the library's defaults (windows <= 31 take the fused small-window
# kernel) and with that kernel switched off (kernel 1 + kernel 2 for every window) -- both must equal the oracle.
@pytest.fixture(scope="session", params=["default_paths", "two_kernel_path"])
def ctx(request):
    import prlib_b200
    c = prlib_b200.Context(0)     # raises PrlCudaError when there is no GPU: gpu tests fail loudly
    c.fused_default = 1 if request.param == "default_paths" else 0
    yield c
    c.close()


@pytest.fixture(autouse=True)
def _path_options(request):
    """(re)apply the path selection before every test that uses `ctx`: to it and to the package's own default context"""
    if "ctx" in request.fixturenames:
        import prlib_b200
        c = request.getfixturevalue("ctx")
        for k in (c, prlib_b200.default_context(0)):
            k.set_option("enable_fused", c.fused_default)
            k.set_option("fused_page_cap", 128)
            k.set_option("fused_no_tier2", 0)
    yield


@pytest.fixture(scope="session")
def noise_page():
    return np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
