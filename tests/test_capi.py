"""CPU suite, part 2: the C-ABI library loads, exports what include/prlib_cuda.h declares, validates
arguments on the host, and fails loudly (never falls back to the CPU) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import prlib_b200
from prlib_b200 import capi
from oracle import prl_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "prlib_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(prl_cuda_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    L = capi.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/prlib_cuda.h but not exported"
    assert sorted(capi.SIGNATURES) == names, "ctypes binding table out of sync with the header"


def test_library_has_no_oracle_or_opencv_dependency():
    import subprocess
    out = subprocess.run(["ldd", capi.lib_path()], capture_output=True, text=True).stdout
    assert "opencv" not in out.lower() and "oracle" not in out.lower() and "torch" not in out.lower()
    sym = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True).stdout
    assert "oracle_" not in sym


def test_output_shape_matches_reference_geometry():
    L = capi.load()
    rng = np.random.default_rng(3)
    for _ in range(200):
        rows, cols = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        window = int(rng.integers(1, 60)) * 2 + 1
        for method in range(5):
            a, b = C.c_int(), C.c_int()
            rc = L.prl_cuda_output_shape(method, rows, cols, window, C.byref(a), C.byref(b))
            want = O.output_shape(method, rows, cols, window)
            if want[0] <= 0 or want[1] <= 0:
                assert rc == capi.PRL_E_EMPTY_ROI
            else:
                assert rc == capi.PRL_OK and (a.value, b.value) == want
    a, b = C.c_int(), C.c_int()
    assert L.prl_cuda_output_shape(0, 100, 100, 14, C.byref(a), C.byref(b)) == capi.PRL_E_INVALID
    assert L.prl_cuda_output_shape(0, 100, 100, 1, C.byref(a), C.byref(b)) == capi.PRL_E_INVALID
    assert L.prl_cuda_output_shape(0, 0, 100, 15, C.byref(a), C.byref(b)) == capi.PRL_E_INVALID
    assert L.prl_cuda_output_shape(7, 100, 100, 15, C.byref(a), C.byref(b)) == capi.PRL_E_INVALID


def test_host_mirror_validates_like_the_reference():
    img = np.zeros((32, 32), np.uint8)
    for fn in (prlib_b200.binarizeSauvola, prlib_b200.binarizeNiblack, prlib_b200.binarizeWolfJolion,
               prlib_b200.binarizeNICK, prlib_b200.binarizeFeng):
        with pytest.raises(ValueError):
            fn(img, 10)                       # even window -> std::invalid_argument
        with pytest.raises(ValueError):
            fn(img, 1)
        with pytest.raises(ValueError):
            fn(np.zeros((0, 0), np.uint8))    # empty image -> std::invalid_argument
    with pytest.raises(ValueError):
        prlib_b200.binarizeLocalOtsuRects(img, [(0, 0, 4, 4)], maxValue=300)


def test_defaults_match_reference_headers():
    import inspect
    d = lambda f: {k: v.default for k, v in inspect.signature(f).parameters.items() if v.default is not inspect._empty}
    assert d(prlib_b200.binarizeSauvola) == {"windowSize": 101, "thresholdCoefficient": 0.01, "morphIterationCount": 2, "device": 0}
    assert d(prlib_b200.binarizeNiblack) == {"windowSize": 101, "thresholdCoefficient": 0.01, "morphIterationCount": 2, "device": 0}
    assert d(prlib_b200.binarizeWolfJolion) == {"windowSize": 101, "thresholdCoefficient": 0.01, "morphIterationCount": 2, "device": 0}
    assert d(prlib_b200.binarizeNICK) == {"windowSize": 21, "thresholdCoefficient": -0.01, "morphIterationCount": 0, "device": 0}
    assert d(prlib_b200.binarizeFeng) == {"windowSize": 21, "thresholdCoefficient_alpha1": 0.75, "thresholdCoefficient_k1": 0.2,
                                          "thresholdCoefficient_k2": 0.03, "thresholdCoefficient_gamma": 2.0,
                                          "morphIterationCount": 2, "device": 0}


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    L = capi.load()
    if L.prl_cuda_device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    h = C.c_void_p()
    assert L.prl_cuda_create(0, C.byref(h)) == capi.PRL_E_CUDA
    assert b"no CPU fallback" in L.prl_cuda_last_error(None)
    with pytest.raises(prlib_b200.PrlCudaError):
        prlib_b200.binarizeSauvola(np.full((64, 64), 128, np.uint8), 15, 0.2, 0)
    with pytest.raises(prlib_b200.PrlCudaError):
        prlib_b200.binarize_batch(np.zeros((2, 64, 64), np.uint8), capi.SAUVOLA, 15, (0.2,))


def test_product_sources_never_touch_the_oracle():
    pkg = os.path.join(ROOT, "prlib_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/prl_oracle.py:synth_page", ""), fn


def test_committed_dram_traffic_agrees_with_the_byte_model():
    """profiles/traffic.json (ncu dram__bytes of kernel 1, kernel 2 and the fused kernel, regenerated by
    scripts/make_profiles.py with every kernel change) must stay within 3 % of the algorithmic bytes bench.py's roofline
    uses: traffic above the model means wasted re-reads, traffic below it means the model counts bytes nobody moves."""
    import json
    import sys
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    k1, k2 = bench.algorithmic_bytes(bench.ROWS, bench.COLS, bench.WINDOW, "compact")
    g = bench.geometry(bench.ROWS, bench.COLS, bench.WINDOW)
    pmin = bench.ROWS * bench.COLS + g["out_rows"] * g["out_cols"]
    for fam, per_page in (("integral", k1), ("threshold", k2), ("fused", pmin)):
        model = per_page * t[fam]["pages_per_launch"]
        assert abs(t[fam]["dram_bytes_per_launch"] / model - 1.0) < 0.03, (fam, t[fam]["dram_bytes_per_launch"], model)
    assert t.get("commit")


def test_host_side_mask_expansion_equals_the_pix_layout():
    """prl_cuda_unpack_mask_host, the host half of the batch loader's 1-bit return path (no GPU involved): PIX words
    (bit 31 - (x & 31) of word x >> 5, 1 = black) to 0/255 bytes, AVX2 and portable loops, widths around the 32-pixel groups"""
    import numpy as np
    from prlib_b200 import capi, unpack_lept1
    L = capi.load()
    rng = np.random.default_rng(4)
    for rows, cols in ((1, 1), (3, 31), (5, 32), (4, 33), (7, 64), (9, 100), (37, 421), (11, 2479)):
        wpl = (cols + 31) // 32
        bits = rng.integers(0, 1 << 32, (rows, wpl), dtype=np.uint64).astype(np.uint32)
        want = unpack_lept1(bits, cols)
        for scalar in (0, 1):
            got = np.full((rows, cols), 7, np.uint8)
            assert L.prl_cuda_unpack_mask_host(bits.ctypes.data, rows, cols, got.ctypes.data, scalar) == capi.PRL_OK
            assert np.array_equal(got, want), (rows, cols, scalar)
