"""CPU suite: parity PINNED by the reference itself.

oracle/_ref is PRLib's own C++ (binarize{Sauvola,Niblack,WolfJolion,NICK,Feng,LocalOtsu}.cpp, removeLines.cpp,
imageLibCommon.cpp) compiled UNMODIFIED against the cv:: facade (oracle/cvfacade/), its OpenCV primitives executed by
the cv2 wheel.  tests/golden/ref_golden.json holds its outputs (tests/golden/make_ref_golden.py).  Here:
  * the built _ref reproduces the committed digests (it is deterministic, and still the same code);
  * oracle/prl_oracle.py (cv2 op-for-op port) and oracle/prl_oracle.c (first principles) equal _ref's outputs --
    so every test that compares the CUDA path with either oracle is anchored to the reference's own code;
  * the three places where OpenCV fuses a multiply-add (what round 1's port had wrong) are pinned by measurement.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import prl_oracle as O
from oracle import ref as R
from util import sha

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/_prl_ref.so not built (needs /root/reference)")

with open(os.path.join(HERE, "golden", "ref_golden.json")) as f:
    REF = json.load(f)
PAGES = dict(np.load(os.path.join(HERE, "golden", "real_pages.npz")))
METHOD = {"binarizeSauvola": 0, "binarizeNiblack": 1, "binarizeWolfJolion": 2, "binarizeNICK": 3, "binarizeFeng": 4}


def image(key):
    if key in PAGES:
        return PAGES[key]
    if key == "noise_512x640":
        return np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
    if key == "tiny_12x40":
        return np.random.default_rng(5).integers(0, 256, (12, 40), dtype=np.uint8)
    if key.startswith("a4_p"):
        return CO.synth_page(int(key[4:]))
    if key == "a3_600_p0":
        return CO.synth_page(0, 9921, 7016)
    if key == "black_120x130":
        return np.zeros((120, 130), np.uint8)
    if key == "white_120x130":
        return np.full((120, 130), 255, np.uint8)
    if key == "const77_97x131":
        return np.full((97, 131), 77, np.uint8)
    if key == "halfblack_140x150":
        a = np.zeros((140, 150), np.uint8)
        a[:, 75:] = 200
        return a
    if key == "sparse_150x160":
        a = np.zeros((150, 160), np.uint8)
        a[::37, ::41] = 255
        return a
    raise KeyError(key)


def call_args(name):
    fn, args = REF["calls"][name]
    return fn, args[:-1], args[-1]        # function, (window, params...), morph


SMALL_KEYS = [k for k in REF["images"] if not k.startswith(("a4_", "a3_"))]


def test_fixture_inputs_are_the_committed_ones():
    for key, e in REF["images"].items():
        if key.startswith("a3_"):
            continue
        assert sha(image(key)) == e["sha1"], key


@needs_ref
@pytest.mark.parametrize("key", SMALL_KEYS + ["a4_p1"])
def test_ref_reproduces_committed_digests(key):
    img, e = image(key), REF["images"][key]
    for name, want in e["masks"].items():
        fn, args = REF["calls"][name]
        if isinstance(want, str):
            if want == "cv2.error":
                with pytest.raises(cv2.error):
                    getattr(R, fn)(img, *args)
            continue
        out, after = getattr(R, fn)(img, *args, return_input=True)
        assert list(out.shape) == want["shape"] and sha(out) == want["sha1"], (key, name)
        assert sha(after) == want["input_after"]["sha1"], (key, name, "input side effect")
    if "removeLines" in e:
        assert sha(R.removeLines(img)) == e["removeLines"]
    if e.get("localOtsu") == "ValueError":
        with pytest.raises(ValueError):
            R.binarizeLocalOtsu(img)
    elif "localOtsu" in e:
        assert sha(R.binarizeLocalOtsu(img)) == e["localOtsu"]
        assert sha(R.binarizeLocalOtsu(img, 255.0, 2.0)) == e["localOtsu_clahe2"]


@pytest.mark.parametrize("key", SMALL_KEYS)
def test_both_oracles_equal_the_reference_outputs(key):
    img, e = image(key), REF["images"][key]
    for name, want in e["masks"].items():
        fn, pars, morph = call_args(name)
        m, window, params = METHOD[fn], pars[0], tuple(pars[1:])
        if isinstance(want, str):
            if want == "cv2.error":
                with pytest.raises(cv2.error):
                    O.binarize_local(img, m, window, params, morph)
                with pytest.raises(ValueError):
                    CO.binarize_local(O.to_gray(img), m, window, params, morph)
            continue
        out = O.binarize_local(img, m, window, params, morph)
        assert sha(out) == want["sha1"], (key, name, "cv2 port")
        out_c = CO.binarize_local(O.to_gray(img), m, window, params, morph)
        assert sha(out_c) == want["sha1"], (key, name, "C restatement")
    if "removeLines" in e:
        assert sha(O.removeLines(img)) == e["removeLines"]
    if e.get("localOtsu") == "ValueError":
        with pytest.raises(ValueError):
            O.binarizeLocalOtsu(img)
    elif "localOtsu" in e:
        assert sha(O.binarizeLocalOtsu(img)) == e["localOtsu"]
        assert sha(O.binarizeLocalOtsu(img, 255.0, 2.0)) == e["localOtsu_clahe2"]


@pytest.mark.parametrize("name", ["sauvola_w15_k0.2", "wolfjolion_w15_k0.5", "nick_w101_k-0.1", "feng_defaults", "sauvola_defaults"])
def test_a4_c_oracle_equals_the_reference_outputs(name):
    fn, pars, morph = call_args(name)
    out = CO.binarize_local(CO.synth_page(0), METHOD[fn], pars[0], tuple(pars[1:]), morph)
    assert sha(out) == REF["images"]["a4_p0"]["masks"][name]["sha1"]


@needs_ref
def test_ref_live_against_both_oracles_random_sweep():
    rng = np.random.default_rng(77)
    for i in range(12):
        rows, cols = int(rng.integers(40, 260)), int(rng.integers(40, 300))
        img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        if i % 3 == 0:
            img = (img // 8 + rng.integers(0, 200)).astype(np.uint8)     # low-contrast page
        m = int(rng.integers(0, 5))
        window = int(rng.choice([3, 5, 9, 15, 21, 31]))
        params = (0.75, 0.2, 0.03, 2.0) if m == 4 else (float(rng.uniform(-0.5, 0.5)),)
        morph = int(rng.integers(-2, 3))
        if m >= 2 and min(rows, cols) <= window:
            continue
        got = R.binarize_local(img, m, window, params, morph)
        assert np.array_equal(got, O.binarize_local(img, m, window, params, morph)), (i, m, window)
        assert np.array_equal(got, CO.binarize_local(img, m, window, params, morph)), (i, m, window)


@needs_ref
def test_ref_error_behaviour_is_the_reference_s():
    img = image("noise_512x640")
    with pytest.raises(ValueError):
        R.binarizeSauvola(np.zeros((0, 0), np.uint8))          # binarizeSauvola.cpp:38-41
    with pytest.raises(ValueError):
        R.binarizeSauvola(img, 14)                             # :43-47 even window
    with pytest.raises(ValueError):
        R.binarizeNICK(img, 1)
    with pytest.raises(ValueError):
        R.binarizeLocalOtsu(img, 256.0)                        # binarizeLocalOtsu.cpp:52-55
    with pytest.raises(cv2.error):
        R.binarizeWolfJolion(img[:10, :10], 15)                # empty processingRect -> cv::Exception
    with pytest.raises(ValueError):
        R.binarizeLocalOtsu(np.full((80, 90), 128, np.uint8))  # no contours: RemoveChildrenContours throws


@needs_ref
def test_ref_bgr_input_and_side_effect():
    bgr = PAGES[[k for k, v in PAGES.items() if v.ndim == 3][0]]
    out, after = R.binarizeSauvola(bgr, 15, 0.2, 0, return_input=True)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(after, cv2.copyMakeBorder(gray, 7, 7, 7, 7, cv2.BORDER_REPLICATE))   # binarizeSauvola.cpp:51,65
    assert np.array_equal(out, O.binarizeSauvola(bgr, 15, 0.2, 0))


def test_wheel_multiply_add_is_fused():
    """Mat::convertTo(alpha, beta), cv::scaleAdd: ONE rounding on this OpenCV build; cv::addWeighted's vector body is
    fma(a, alpha, fma(b, beta, gamma)); filter2D's taps are NOT fused.  This is what the C oracle and the CUDA exact
    path restate (prl_oracle.c header, decide.cuh:thr_value_p)."""
    if "fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU without FMA3: OpenCV dispatches to its two-rounding baseline")
    libm = ctypes.CDLL("libm.so.6")
    libm.fma.restype = ctypes.c_double
    libm.fma.argtypes = [ctypes.c_double] * 3
    fma = np.vectorize(lambda a, b, c: libm.fma(a, b, c))
    rng = np.random.default_rng(1)
    s, m = rng.random((16, 1024)) * 100, rng.random((16, 1024)) * 255
    a, b = 0.2 / 128, 0.8
    conv = O._scale_shift(s, a, b)
    assert np.array_equal(conv, fma(s, a, b)) and not np.array_equal(conv, s * a + b)
    sa = cv2.scaleAdd(s, -0.2, m)
    assert np.array_equal(sa, fma(s, -0.2, m)) and not np.array_equal(sa, s * -0.2 + m)
    aw = cv2.addWeighted(s, 0.37, m, -1.234, 5.5)
    assert np.array_equal(aw, fma(s, 0.37, fma(m, -1.234, 5.5)))
    assert np.array_equal(cv2.addWeighted(m, 1.0, s, -0.1, 0.0), m + s * -0.1)       # NICK: alpha == 1 -> plain
