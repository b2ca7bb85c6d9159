"""GPU parity against THE REFERENCE'S OWN OUTPUTS: tests/golden/ref_golden.json was produced by oracle/_ref (PRLib's own
C++ compiled unmodified, OpenCV primitives executed by the cv2 wheel; tests/golden/make_ref_golden.py).  Every call goes
through the host mirror of the reference interface (prlib_b200.binarize*) and so through the C-ABI.  Bit-exact: zero
differing pixels.  Where the built _ref module travelled to this box it is also run live on seeded random inputs."""
import numpy as np
import pytest

import prlib_b200
from prlib_b200 import capi
from oracle import ref as R
from test_ref import METHOD, PAGES, REF, call_args, image
from util import sha

pytestmark = pytest.mark.gpu
FN = {"binarizeSauvola": prlib_b200.binarizeSauvola, "binarizeNiblack": prlib_b200.binarizeNiblack,
      "binarizeWolfJolion": prlib_b200.binarizeWolfJolion, "binarizeNICK": prlib_b200.binarizeNICK,
      "binarizeFeng": prlib_b200.binarizeFeng}


@pytest.mark.parametrize("key", list(REF["images"]))
def test_cuda_equals_reference_outputs(key, ctx):
    img, e = image(key), REF["images"][key]
    for name, want in e["masks"].items():
        fn, args = REF["calls"][name]
        if isinstance(want, str):
            if want == "cv2.error":                      # empty processingRect: cv::Exception in the reference
                with pytest.raises(capi.PrlCudaError) as ei:
                    FN[fn](img, *args)
                assert ei.value.code == capi.PRL_E_EMPTY_ROI
            continue
        out = FN[fn](img, *args)
        assert list(out.shape) == want["shape"], (key, name)
        assert sha(out) == want["sha1"], (key, name, float((out == 255).mean()), want["white"])
        # the side effect the C++ shim reproduces: imageInput := replicate-padded gray (binarizeSauvola.cpp:51,65)
        assert sha(prlib_b200.padded_gray(img, args[0])) == want["input_after"]["sha1"], (key, name)
    if "removeLines" in e:
        assert sha(prlib_b200.removeLines(img)) == e["removeLines"], key
    if e.get("localOtsu") == "ValueError":
        with pytest.raises(ValueError):
            prlib_b200.binarizeLocalOtsu(img)
    elif "localOtsu" in e:
        assert sha(prlib_b200.binarizeLocalOtsu(img)) == e["localOtsu"], key
        assert sha(prlib_b200.binarizeLocalOtsu(img, 255.0, 2.0)) == e["localOtsu_clahe2"], key


@pytest.mark.parametrize("opt,val", [("exact_threshold", 1), ("enable_fused", 0), ("enable_fused", 1), ("disable_compact", 1)])
def test_cuda_validation_paths_equal_reference_outputs(ctx, opt, val):
    """the literal-FP64 kernel, the two-kernel path (compact and int64 planes) and the fused small-window kernel against the same reference digests"""
    d = prlib_b200.default_context(0)
    d.set_option(opt, val)
    try:
        for key in ("noise_512x640", "halfblack_140x150", "sparse_150x160", "page_0001", "page_0099"):
            img, e = image(key), REF["images"][key]
            for name, want in e["masks"].items():
                fn, args = REF["calls"][name]
                assert sha(FN[fn](img, *args)) == want["sha1"], (opt, key, name)
    finally:
        d.set_option(opt, ctx.fused_default if opt == "enable_fused" else 0)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/_prl_ref.so did not travel to this box")
def test_cuda_equals_live_reference_random_sweep(ctx):
    rng = np.random.default_rng(4242)
    n = 0
    for i in range(60):
        rows, cols = int(rng.integers(24, 700)), int(rng.integers(24, 900))
        img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        kind = i % 4
        if kind == 1:
            img = (img // 16 + rng.integers(0, 230)).astype(np.uint8)
        elif kind == 2:
            img[rng.random((rows, cols)) < 0.7] = 0                     # mostly black: the NaN / zero-window branches
        elif kind == 3:
            img = np.repeat(np.repeat(img[::8, ::8], 8, 0), 8, 1)[:rows, :cols].copy()   # blocky: flat windows
        m = int(rng.integers(0, 5))
        window = int(rng.choice([3, 7, 15, 21, 31, 51, 101]))
        params = (float(rng.uniform(0, 1)), 0.2, float(rng.uniform(0, 0.2)), 2.0) if m == 4 else (float(rng.uniform(-0.6, 0.6)),)
        morph = int(rng.integers(-3, 4)) if i % 2 else 0
        if m >= 2 and min(img.shape) <= window:
            continue
        want = R.binarize_local(img, m, window, params, morph)
        got = ctx.binarize_local(img, m, window, params, morph)
        assert np.array_equal(got, want), (i, m, window, params, morph, img.shape, int((got != want).sum()))
        n += 1
    assert n >= 40


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/_prl_ref.so did not travel to this box")
def test_cuda_equals_live_reference_bgr_and_widened(ctx):
    for key, img in PAGES.items():
        if img.ndim != 3:
            continue
        assert np.array_equal(prlib_b200.binarizeSauvola(img, 15, 0.2, 0), R.binarizeSauvola(img, 15, 0.2, 0)), key
        assert np.array_equal(prlib_b200.binarizeLocalOtsu(img), R.binarizeLocalOtsu(img)), key
        assert np.array_equal(prlib_b200.removeLines(img), R.removeLines(img)), key
