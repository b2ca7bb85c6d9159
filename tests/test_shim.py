"""The cv::Mat-facing C++ shim (prlib_b200/shim): compile check against the mock OpenCV header on the CPU
box, and a real run on the GPU box (the test binary links libprlib_cuda + the plain-C oracle)."""
import os
import subprocess

import pytest

from oracle import c_oracle
from prlib_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(BUILD, "shim_test")


def _build():
    os.makedirs(BUILD, exist_ok=True)
    capi.load()
    oracle_so = c_oracle.build()
    libdir = os.path.dirname(capi.lib_path())
    cmd = ["g++", "-std=c++11", "-O2", "-Wall", "-Werror",
           "-I" + os.path.join(ROOT, "tests", "mock_opencv"), "-I" + os.path.join(ROOT, "prlib_b200", "shim"),
           os.path.join(ROOT, "tests", "cpp", "shim_test.cpp"), os.path.join(ROOT, "prlib_b200", "shim", "prl_binarize_cuda.cpp"),
           oracle_so, "-L" + libdir, "-lprlib_cuda", "-Wl,-rpath," + libdir, "-Wl,-rpath," + os.path.dirname(oracle_so),
           "-lm", "-lpthread", "-ldl", "-lrt", "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_shim_compiles_against_the_reference_signatures():
    _build()
    assert os.path.exists(EXE)
    # the shim object itself must not reference the oracle
    r = subprocess.run(["g++", "-std=c++11", "-c", "-I" + os.path.join(ROOT, "tests", "mock_opencv"),
                        os.path.join(ROOT, "prlib_b200", "shim", "prl_binarize_cuda.cpp"), "-o", os.path.join(BUILD, "shim.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sym = subprocess.run(["nm", "-C", os.path.join(BUILD, "shim.o")], capture_output=True, text=True).stdout
    assert "oracle_" not in sym
    for name in ("prl::binarizeSauvola(cv::Mat&, cv::Mat&, int, double, int)",
                 "prl::binarizeNiblack(cv::Mat&, cv::Mat&, int, double, int)",
                 "prl::binarizeWolfJolion(cv::Mat&, cv::Mat&, int, double, int)",
                 "prl::binarizeNICK(cv::Mat&, cv::Mat&, int, double, int)",
                 "prl::binarizeFeng(cv::Mat&, cv::Mat&, int, double, double, double, double, int)"):
        assert name in sym, name


@pytest.mark.gpu
def test_shim_runs_on_gpu_and_matches_oracle():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all ok" in r.stdout, r.stdout + r.stderr
