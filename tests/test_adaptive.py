"""The adaptive-mean family: prl::binarizeNativeAdaptive / AT / AGT / GAT / PureAdaptive /
PureAdaptiveGaussian.  Oracles: the reference's own object code (oracle/_ref), its cv2 call sequence
(oracle/prl_oracle.py) and the real cv2 primitives for the two building blocks (medianBlur, adaptiveThreshold)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import prl_oracle as O
from oracle import ref as R

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/_prl_ref.so not built (needs /root/reference)")
REPL = cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED


def _images():
    rng = np.random.default_rng(11)
    crops = dict(np.load(os.path.join(HERE, "golden", "real_crops.npz")))
    out = {"noise_120x150": rng.integers(0, 256, (120, 150, 3), dtype=np.uint8),
           "synth_300x421": np.repeat(CO.synth_page(3, 300, 421)[:, :, None], 3, axis=2),
           "dark_90x131": (rng.integers(0, 256, (90, 131, 3)) // 6).astype(np.uint8),
           "bgra_64x70": rng.integers(0, 256, (64, 70, 4), dtype=np.uint8)}
    for k, v in crops.items():
        if v.ndim == 3:
            out["crop_" + k] = v
            break
    return out


IMAGES = _images()

import json  # noqa: E402
with open(os.path.join(HERE, "golden", "ref_adaptive_golden.json")) as _f:
    GOLD = json.load(_f)       # outputs of the reference's own code (tests/golden/make_ref_adaptive_golden.py)


def gold_image(key):
    rng = np.random.default_rng(11)
    gen = {"noise_bgr_120x150": lambda: rng.integers(0, 256, (120, 150, 3), dtype=np.uint8)}
    if key == "noise_bgr_120x150":
        return gen[key]()
    a = rng.integers(0, 256, (120, 150, 3), dtype=np.uint8)
    if key == "dark_bgr_90x131":
        return (rng.integers(0, 256, (90, 131, 3)) // 6).astype(np.uint8)
    b = (rng.integers(0, 256, (90, 131, 3)) // 6).astype(np.uint8)
    if key == "noise_bgra_64x70":
        return rng.integers(0, 256, (64, 70, 4), dtype=np.uint8)
    if key == "synth_gray_300x421":
        return CO.synth_page(3, 300, 421)
    if key == "a4_p2":
        return CO.synth_page(2)
    return dict(np.load(os.path.join(HERE, "golden", "real_pages.npz")))[key]


def digest_or_class(fn):
    from util import sha
    o = outcome(fn)
    return sha(o[1]) if o[0] == "ok" else o[0]


def outcome(fn):
    """('ok', array) or the exception class the call ends in: the auto block size int(diagonal / 333 + 7) of
    binarizeNativeAdaptive can come out even, and then the REFERENCE itself ends in cv::adaptiveThreshold's assertion"""
    from prlib_b200 import PrlCudaError
    try:
        return "ok", fn()
    except ValueError:
        return "invalid_argument", None
    except (cv2.error, PrlCudaError):
        return "cv::Exception", None


def same_outcome(a, b):
    return a[0] == b[0] and (a[1] is None or np.array_equal(a[1], b[1]))


# ---------------------------------------------------------------------------------------------- CPU suite
def test_float_gaussian_coefficients_equal_getGaussianKernel():
    from prlib_b200 import capi
    L = capi.load()
    for n in range(1, 256, 2):
        k = (C.c_float * 256)()
        assert L.prl_cuda_gauss_kernel_float(n, k) == 0
        assert np.array_equal(np.array(k[:n], np.float32), cv2.getGaussianKernel(n, 0, cv2.CV_32F).ravel()), n
    k = (C.c_float * 256)()
    assert L.prl_cuda_gauss_kernel_float(4, k) != 0 and L.prl_cuda_gauss_kernel_float(257, k) != 0


def test_box_mean_model_is_what_boxFilter_computes():
    rng = np.random.default_rng(0)
    with O._single_thread():
        for bs in list(range(3, 65, 4)) + [63, 99, 201, 255]:
            for shape in ((67, 131), (40, 33), (9, 300)):
                img = rng.integers(0, 256, shape, dtype=np.uint8)
                ref = cv2.boxFilter(img, -1, (bs, bs), normalize=True, borderType=REPL)
                assert np.array_equal(O.box_mean_model(img, bs), ref), (bs, shape)


def test_gaussian_mean_model_is_what_GaussianBlur_32f_computes():
    """the measured summation order of the wheel's sepFilter2D (vector bodies fused, scalar tails as compiled): float for float"""
    rng = np.random.default_rng(7)
    with O._single_thread():
        for shape in ((97, 131), (50, 133), (40, 64), (33, 70), (20, 7), (64, 12), (5, 3), (1, 9), (9, 1), (120, 257)):
            g = rng.integers(0, 256, shape, dtype=np.uint8)
            for bs in (3, 5, 7, 9, 11, 13, 19, 21, 31, 35, 51):
                ref = cv2.GaussianBlur(g.astype(np.float32), (bs, bs), 0, 0, borderType=REPL)
                assert np.array_equal(O.gaussian_mean_model(g, bs, as_float=True), ref), (shape, bs)


@needs_ref
def test_reference_family_quirks_and_port_equals_reference():
    """AT / AGT / PureAdaptiveGaussian only work on colour input; GAT / PureAdaptive raise for every non-empty input; the
    cv2 restatement equals the reference's own object code where that returns."""
    for name, img in IMAGES.items():
        gray = np.ascontiguousarray(img[:, :, 0])
        for fn, args in (("binarizeAT", (5, 255, 19, 9)), ("binarizeAGT", (3, 200, 11, -3)), ("binarizePureAdaptiveGaussian", (255, 15, 4))):
            assert np.array_equal(getattr(R, fn)(img, *args), getattr(O, fn)(img, *args)), (name, fn)
            with pytest.raises(cv2.error):
                getattr(R, fn)(gray, *args)
            with pytest.raises(ValueError):
                getattr(R, fn)(np.zeros((0, 0), np.uint8), *args)
        for im in (img, gray):
            with pytest.raises(cv2.error):
                R.binarizeGAT(im, 7, 1.5, 1.5, 255, 19, 9)
            with pytest.raises(cv2.error):
                R.binarizePureAdaptive(im, 255, 19, 9)
            for kw in ({}, {"isGaussianBlurReqiured": True}, {"isAdaptiveThresholdCalculatedByGaussian": False},
                       {"adaptiveThresholdingBlockSize": 0}, {"medianBlurKernelSize": 3, "adaptiveThresholdingShift": -2.5},
                       {"adaptiveThresholdingMaxValue": 77.6}):
                assert same_outcome(outcome(lambda: R.binarizeNativeAdaptive(im, **kw)), outcome(lambda: O.binarizeNativeAdaptive(im, **kw))), (name, kw)
    with pytest.raises(ValueError):
        R.binarizeNativeAdaptive(IMAGES["noise_120x150"], adaptiveThresholdingMaxValue=300)
    with pytest.raises(cv2.error):
        R.binarizeNativeAdaptive(IMAGES["noise_120x150"], medianBlurKernelSize=2)
    with pytest.raises(cv2.error):
        R.binarizeNativeAdaptive(IMAGES["noise_120x150"], adaptiveThresholdingBlockSize=20)


def test_host_mirror_raises_like_the_reference_without_touching_a_gpu():
    import prlib_b200
    from prlib_b200 import PrlCudaError, capi
    gray = np.zeros((40, 50), np.uint8)
    for call in (lambda: prlib_b200.binarizeAT(gray, 5, 255, 19, 9), lambda: prlib_b200.binarizeAGT(gray, 5, 255, 19, 9),
                 lambda: prlib_b200.binarizePureAdaptiveGaussian(gray, 255, 19, 9),
                 lambda: prlib_b200.binarizeGAT(np.zeros((40, 50, 3), np.uint8), 7, 1.0, 1.0, 255, 19, 9),
                 lambda: prlib_b200.binarizePureAdaptive(np.zeros((40, 50, 3), np.uint8), 255, 19, 9)):
        with pytest.raises(PrlCudaError) as e:
            call()
        assert e.value.code == capi.PRL_E_EMPTY_ROI
    for call in (lambda: prlib_b200.binarizeAT(np.zeros((0, 0, 3), np.uint8), 5, 255, 19, 9),
                 lambda: prlib_b200.binarizeNativeAdaptive(np.zeros((0, 0), np.uint8)),
                 lambda: prlib_b200.binarizeNativeAdaptive(gray, adaptiveThresholdingMaxValue=256)):
        with pytest.raises(ValueError):
            call()


def _two_level_neighbourhoods(radius, maxv, rng):
    """one (2 radius + 1)^2 block per subset of the taps of the disc (and per centre value): every neighbourhood a two-level
    image can show the filter; the corners outside the disc are random"""
    n = 2 * radius + 1
    taps = [(i, j) for i in range(-radius, radius + 1) for j in range(-radius, radius + 1) if i * i + j * j <= radius * radius]
    npat = 1 << len(taps)
    side = int(np.ceil(np.sqrt(npat)))
    img = ((rng.random((side * n, side * n)) < 0.5) * maxv).astype(np.uint8)
    p = np.arange(npat)
    by, bx = p // side, p % side
    for k, (i, j) in enumerate(taps):
        img[by * n + radius + i, bx * n + radius + j] = np.where((p >> k) & 1, maxv, 0)
    return img


BILATERAL_CASES = ((3, 150.0, 150.0), (7, 30.0, 4.0), (9, 100.0, 2.5), (11, 60.0, 2.2), (0, 40.0, 2.0), (3, 10.0, 0.5), (15, 150.0, 150.0))


def test_bilateral_model_is_what_cv2_computes():
    """the restatement csrc/adaptive.cu follows (float32 tables from OpenCV's SIMD exponential, taps in raster order) against
    the wheel's own C++ path: full-range gray for every radius but 2, and for radius 2 (where the wheel runs a special case that
    rounds ~1e-5 of full-range pixels the other way) every neighbourhood of a two-level image -- what the reference feeds it"""
    rng = np.random.default_rng(77)
    gray = rng.integers(0, 256, (400, 523), dtype=np.uint8)
    smooth = cv2.GaussianBlur(gray, (0, 0), 2.0)
    with O._single_thread():
        for d, sc, ss in BILATERAL_CASES:
            for img in (gray, smooth, gray[:1], gray[:, :1], gray[:3, :2]):
                assert np.array_equal(O.bilateral_model(img, d, sc, ss), cv2.bilateralFilter(img, d, sc, ss)), (d, sc, ss, img.shape)
        for maxv in (255, 180, 1):
            for d, sc, ss in ((5, 150.0, 150.0), (5, 20.0, 1.5), (4, 300.0, 0.8), (0, 75.0, 1.4)):
                img = _two_level_neighbourhoods(2, maxv, rng)
                assert np.array_equal(O.bilateral_model(img, d, sc, ss), cv2.bilateralFilter(img, d, sc, ss)), (maxv, d, sc, ss)


@pytest.mark.parametrize("key", [k for k in GOLD["images"] if k != "a4_p2"])
def test_cv2_restatement_reproduces_the_reference_digests(key):
    from util import sha
    img, e = gold_image(key), GOLD["images"][key]
    assert sha(img) == e["sha1"], key
    for name, want in e["out"].items():
        fn, args, kw = GOLD["calls"][name]
        assert digest_or_class(lambda: getattr(O, fn)(img, *args, **kw)) == want, (key, name)


# ---------------------------------------------------------------------------------------------- GPU suite
@pytest.mark.gpu
@pytest.mark.parametrize("key", list(GOLD["images"]))
def test_cuda_reproduces_the_reference_digests(ctx, key):
    import prlib_b200
    from util import sha
    img, e = gold_image(key), GOLD["images"][key]
    assert sha(img) == e["sha1"], key
    for name, want in e["out"].items():
        fn, args, kw = GOLD["calls"][name]
        assert digest_or_class(lambda: getattr(prlib_b200, fn)(img, *args, **kw)) == want, (key, name)


def test_median_selection_networks_by_the_zero_one_principle():
    """the compare-exchange lists csrc/adaptive.cu executes for 3 x 3 and 5 x 5 medians leave the median in wire 4 / 12 for
    every 0/1 input (all 2^9 / 2^25 of them), hence for every input"""
    import re
    text = open(os.path.join(HERE, "..", "prlib_b200", "csrc", "adaptive.cu")).read()
    for name, n, out in (("PRL_MED9_NETWORK", 9, 4), ("PRL_MED25_NETWORK", 25, 12)):
        lines = text[text.index("#define " + name):].split("\n")
        k = next(i for i, ln in enumerate(lines) if not ln.rstrip().endswith("\\"))
        body = "\n".join(lines[:k + 1])
        net = [(int(a), int(b)) for a, b in re.findall(r"CE\((\d+), (\d+)\)", body)]
        assert len(net) == (19 if n == 9 else 99)
        idx = np.arange(1 << n, dtype=np.uint32)
        wires = [((idx >> k) & 1).astype(bool) for k in range(n)]
        ones = sum(w.astype(np.uint8) for w in wires)
        for a, b in net:
            wires[a], wires[b] = wires[a] & wires[b], wires[a] | wires[b]
        assert np.array_equal(wires[out], ones >= n // 2 + 1), name


@pytest.mark.gpu
@pytest.mark.parametrize("ksize", [3, 5, 7, 9, 15, 31])
def test_median_blur_equals_cv2(ctx, ksize):
    rng = np.random.default_rng(ksize)
    if ksize <= 5:                                                   # both kernels: selection network (default) and radix select
        img = rng.integers(0, 256, (70, 131, 3), dtype=np.uint8)
        want = cv2.medianBlur(img, ksize)
        ctx.set_option("median_legacy", 1)
        assert np.array_equal(ctx.median_blur(img, ksize), want)
        ctx.set_option("median_legacy", 0)
        assert np.array_equal(ctx.median_blur(img, ksize), want)
    with O._single_thread():
        for shape in ((97, 131), (33, 70, 3), (64, 12, 4), (5, 3), (1, 40), (300, 257, 3)):
            if ksize > 5 and len(shape) == 3 and shape[0] * shape[1] > 20000:
                continue
            img = rng.integers(0, 256, shape, dtype=np.uint8)
            if len(shape) == 2:
                img[::3, ::2] //= 8                                    # repeated values: ties inside the window
            assert np.array_equal(ctx.median_blur(img, ksize), cv2.medianBlur(img, ksize)), (ksize, shape)


@pytest.mark.gpu
@pytest.mark.parametrize("method", [0, 1])
def test_adaptive_threshold_equals_cv2(ctx, method):
    rng = np.random.default_rng(20 + method)
    cvm = cv2.ADAPTIVE_THRESH_GAUSSIAN_C if method else cv2.ADAPTIVE_THRESH_MEAN_C
    with O._single_thread():
        for shape in ((97, 131), (50, 133), (40, 64), (33, 70), (20, 7), (64, 12), (5, 3), (1, 9), (9, 1), (300, 421)):
            g = rng.integers(0, 256, shape, dtype=np.uint8)
            if shape == (300, 421):
                g = CO.synth_page(1, 300, 421)
            for bs in (3, 5, 9, 11, 19, 21, 35, 63, 65, 101):
                for ttype, delta, maxval in ((0, 9, 255), (1, 9, 255), (0, -3.5, 200.4), (1, 2.5, 77.5), (0, 0, 255), (1, 300, 255)):
                    want = cv2.adaptiveThreshold(g, maxval, cvm, cv2.THRESH_BINARY_INV if ttype else cv2.THRESH_BINARY, bs, delta)
                    got = ctx.adaptive_threshold(g, maxval, method, ttype, bs, delta)
                    assert np.array_equal(got, want), (shape, bs, ttype, delta, maxval, int((got != want).sum()))
    assert not ctx.adaptive_threshold(np.full((20, 30), 9, np.uint8), -1.0, 0, 0, 5, 0).any()          # maxValue < 0 -> zeros


@pytest.mark.gpu
def test_gaussian_mean_single_kernel_equals_the_two_kernel_form(ctx):
    rng = np.random.default_rng(31)
    for shape in ((97, 131), (33, 64), (200, 257), (40, 9), (3, 300)):
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        for bs in (3, 7, 19, 35, 63):
            ctx.set_option("gauss_legacy", 1)
            want = ctx.adaptive_threshold(g, 255, 1, 0, bs, 4.0)
            ctx.set_option("gauss_legacy", 0)
            assert np.array_equal(ctx.adaptive_threshold(g, 255, 1, 0, bs, 4.0), want), (shape, bs)


@pytest.mark.gpu
def test_family_equals_the_reference(ctx):
    import prlib_b200
    from prlib_b200 import PrlCudaError
    use_ref = R.available()
    W = R if use_ref else O
    for name, img in IMAGES.items():
        gray = np.ascontiguousarray(img[:, :, 1])
        for fn, args in (("binarizeAT", (5, 255, 19, 9)), ("binarizeAT", (3, 180, 7, -2)), ("binarizeAGT", (5, 255, 19, 9)),
                         ("binarizeAGT", (7, 255, 35, 3)), ("binarizePureAdaptiveGaussian", (255, 15, 4))):
            assert np.array_equal(getattr(prlib_b200, fn)(img, *args), getattr(W, fn)(img, *args)), (name, fn, args)
            with pytest.raises(PrlCudaError):
                getattr(prlib_b200, fn)(gray, *args)
        for im in (img, gray):
            for kw in ({}, {"isGaussianBlurReqiured": True}, {"isAdaptiveThresholdCalculatedByGaussian": False},
                       {"adaptiveThresholdingBlockSize": 0}, {"medianBlurKernelSize": 3, "adaptiveThresholdingShift": -2.5},
                       {"isGaussianBlurReqiured": True, "GaussianBlurKernelSize": 11, "GaussianBlurSigma": 2.0,
                        "isAdaptiveThresholdCalculatedByGaussian": False, "adaptiveThresholdingMaxValue": 77.6}):
                assert same_outcome(outcome(lambda: prlib_b200.binarizeNativeAdaptive(im, **kw)), outcome(lambda: W.binarizeNativeAdaptive(im, **kw))), (name, kw)
    with pytest.raises(PrlCudaError):
        prlib_b200.binarizeNativeAdaptive(IMAGES["noise_120x150"], medianBlurKernelSize=2)
    with pytest.raises(PrlCudaError):
        prlib_b200.binarizeNativeAdaptive(IMAGES["noise_120x150"], adaptiveThresholdingBlockSize=20)


@pytest.mark.gpu
def test_native_adaptive_on_an_a4_page(ctx):
    import prlib_b200
    page = CO.synth_page(2)
    assert np.array_equal(prlib_b200.binarizeNativeAdaptive(page), O.binarizeNativeAdaptive(page))
    assert np.array_equal(prlib_b200.binarizeNativeAdaptive(page, isAdaptiveThresholdCalculatedByGaussian=False, adaptiveThresholdingBlockSize=0),
                          O.binarizeNativeAdaptive(page, isAdaptiveThresholdCalculatedByGaussian=False, adaptiveThresholdingBlockSize=0))


@pytest.mark.gpu
def test_bilateral_filter_equals_cv2(ctx):
    rng = np.random.default_rng(78)
    gray = rng.integers(0, 256, (300, 421), dtype=np.uint8)
    smooth = cv2.GaussianBlur(gray, (0, 0), 2.0)
    mask = ((rng.random((257, 300)) < 0.3) * 255).astype(np.uint8)
    with O._single_thread():
        for d, sc, ss in BILATERAL_CASES + ((31, 90.0, 6.0),):
            for img in (gray, smooth, mask, gray[:1], gray[:, :1], gray[:3, :2], gray[:40, :33]):
                assert np.array_equal(ctx.bilateral_filter(img, d, sc, ss), cv2.bilateralFilter(img, d, sc, ss)), (d, sc, ss, img.shape)
        for maxv in (255, 180, 1):
            for d, sc, ss in ((5, 150.0, 150.0), (5, 20.0, 1.5), (4, 300.0, 0.8), (0, 75.0, 1.4)):
                img = _two_level_neighbourhoods(2, maxv, rng)
                assert np.array_equal(ctx.bilateral_filter(img, d, sc, ss), cv2.bilateralFilter(img, d, sc, ss)), (maxv, d, sc, ss)
        page = CO.synth_page(1, 900, 700)
        m = cv2.adaptiveThreshold(page, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, 19, 9)
        for d, sc, ss in ((5, 150.0, 150.0), (9, 40.0, 3.0)):
            assert np.array_equal(ctx.bilateral_filter(m, d, sc, ss), cv2.bilateralFilter(m, d, sc, ss)), (d, sc, ss)


@pytest.mark.gpu
def test_native_adaptive_with_the_bilateral_step_equals_the_reference(ctx):
    import prlib_b200
    W = R if R.available() else O
    for name, img in IMAGES.items():
        for kw in ({"bilateralFilterBlockSize": 5}, {"bilateralFilterBlockSize": 3, "bilateralFilterColorSigma": 25.0, "bilateralFilterSpaceSigma": 0.9},
                   {"bilateralFilterBlockSize": 9, "bilateralFilterColorSigma": 40.0, "bilateralFilterSpaceSigma": 3.0, "adaptiveThresholdingMaxValue": 180},
                   {"bilateralFilterBlockSize": 2, "bilateralFilterColorSigma": 0.0},                    # below 3: the step is skipped, sigmas unchecked
                   {"bilateralFilterBlockSize": 7, "bilateralFilterColorSigma": 0.0},                    # std::invalid_argument
                   {"bilateralFilterBlockSize": 7, "bilateralFilterSpaceSigma": -2.0},
                   {"bilateralFilterBlockSize": 7, "bilateralFilterColorSigma": -1.0, "adaptiveThresholdingBlockSize": 20}):   # cv::Exception comes first
            assert same_outcome(outcome(lambda: prlib_b200.binarizeNativeAdaptive(img, **kw)), outcome(lambda: W.binarizeNativeAdaptive(img, **kw))), (name, kw)


@pytest.mark.gpu
def test_batched_family_equals_the_single_image_calls(ctx):
    """prl_cuda_binarize_adaptive_batch_dev: pages resident in HBM, per-page kernel sequences on the context's page lanes"""
    import torch
    rng = np.random.default_rng(5)
    n, rows, cols = 19, 211, 307                                     # more pages than lanes; odd sizes
    for channels, kws in ((1, ({"gray_first": 1, "blur": 1, "blur_ksize": 5, "assert_ksize": 1, "method": 1, "type": 1, "maxval": 255.0, "check_maxval": 1,
                                 "block_size": 19, "auto_block": 1, "delta": 9.0, "invert_if_dark": 1},
                                {"gray_first": 1, "blur": 2, "blur_ksize": 7, "blur_sigma": 150.0, "method": 0, "type": 1, "maxval": 200.0, "block_size": 21,
                                 "auto_block": 1, "delta": 3.0, "invert_if_dark": 1, "bilateral_d": 5, "bilateral_sigma_color": 150.0,
                                 "bilateral_sigma_space": 150.0})),
                          (3, ({"gray_first": 0, "blur": 1, "blur_ksize": 5, "method": 0, "type": 0, "maxval": 255.0, "block_size": 19, "delta": 9.0},
                               {"gray_first": 0, "blur": 0, "method": 1, "type": 0, "maxval": 255.0, "block_size": 15, "delta": 4.0}))):
        shape = (n, rows, cols) if channels == 1 else (n, rows, cols, channels)
        pages = rng.integers(0, 256, shape, dtype=np.uint8)
        pages[3] //= 5                                               # a dark page: the mean test of NativeAdaptive flips it
        step = (cols * channels + 15) // 16 * 16
        d_src = torch.zeros((n, rows, step), dtype=torch.uint8, device="cuda")
        d_src[:, :, :cols * channels] = torch.from_numpy(pages.reshape(n, rows, cols * channels)).cuda()
        ostep = (cols + 15) // 16 * 16
        for kw in kws:
            d_dst = torch.zeros((n, rows, ostep), dtype=torch.uint8, device="cuda")
            ctx.binarize_adaptive_batch_dev(d_src.data_ptr(), n, rows, cols, step, rows * step, channels, d_dst.data_ptr(), ostep, rows * ostep, **kw)
            got = d_dst[:, :, :cols].cpu().numpy()
            for i in range(n):
                assert np.array_equal(got[i], ctx.binarize_adaptive(pages[i], **kw)), (channels, kw, i)
    with pytest.raises(prlib_b200_error()):
        ctx.binarize_adaptive_batch_dev(d_src.data_ptr(), n, rows, cols, step, rows * step, 3, d_dst.data_ptr(), ostep, rows * ostep,
                                        method=0, type=0, maxval=255.0, block_size=20, delta=1.0)


def prlib_b200_error():
    from prlib_b200 import PrlCudaError
    return PrlCudaError
