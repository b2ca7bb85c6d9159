"""CPU suite, part 1: pin the oracle.

The reference ships no tests or golden outputs, so the pins are:
  * oracle/prl_oracle.py executes the real OpenCV primitives (cv2) in the reference's order;
  * oracle/prl_oracle.c (plain C, first principles) must agree with it bit for bit;
  * both must reproduce the committed digests in tests/golden/golden.json.
"""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import prl_oracle as O
from util import CONFIGS, sha

cv2 = pytest.importorskip("cv2")


def test_synth_page_generators_agree(golden):
    e = golden["images"]["a4_p0"]
    assert sha(CO.synth_page(0)) == e["sha1"]
    small_np = O.synth_page(3, 300, 420, seed=7)
    small_c = CO.synth_page(3, 300, 420, seed=7)
    assert np.array_equal(small_np, small_c)
    assert sha(O.synth_page(1)) == golden["images"]["a4_p1"]["sha1"]


def test_integrals_golden_and_exact(golden, noise_page):
    e = golden["images"]["noise_512x640"]["integral_pad7"]
    for impl in (O.integrals_int64, CO.integrals_int64):
        S, Q = impl(noise_page, 7)
        assert sha(S) == e["S_sha1"] and sha(Q) == e["Q_sha1"]
        assert int(S[-1, -1]) == e["S_last"] and int(Q[-1, -1]) == e["Q_last"]
    a4 = CO.synth_page(0)
    S, Q = CO.integrals_int64(a4, 7)
    e = golden["images"]["a4_p0"]["integral_pad7"]
    assert sha(S) == e["S_sha1"] and sha(Q) == e["Q_sha1"]


@pytest.mark.parametrize("name", list(CONFIGS))
def test_noise_masks_match_golden_both_oracles(golden, noise_page, name):
    m, w, p = CONFIGS[name]
    e = golden["images"]["noise_512x640"]
    out_cv, aux = O.binarize_local(noise_page, m, w, p, 0, return_aux=True)
    out_c, t8_c, _ = CO.binarize_local(noise_page, m, w, p, 0, want_t8=True)
    assert list(out_cv.shape) == e["masks"][name]["shape"]
    assert sha(out_cv) == e["masks"][name]["sha1"]          # cv2 oracle == committed digest
    assert np.array_equal(out_c, out_cv)                     # C restatement == real OpenCV ops
    assert sha(t8_c) == e["t8"][name] == sha(aux["T8"])


@pytest.mark.parametrize("name", ["sauvola_w15_k0.2", "wolfjolion_w15_k0.5", "nick_w101_k-0.1", "feng_w21_default"])
def test_a4_masks_match_golden_c_oracle(golden, name):
    m, w, p = CONFIGS[name]
    a4 = CO.synth_page(0)
    out = CO.binarize_local(a4, m, w, p, 0)
    e = golden["images"]["a4_p0"]["masks"][name]
    assert list(out.shape) == e["shape"] and sha(out) == e["sha1"]


def test_a4_sauvola_cv2_oracle_matches_golden(golden):
    a4 = O.synth_page(0)
    out = O.binarizeSauvola(a4, 15, 0.2, 0)
    assert sha(out) == golden["images"]["a4_p0"]["masks"]["sauvola_w15_k0.2"]["sha1"]


def test_real_crops_both_oracles(golden, real_crops):
    for key in ("real_0037", "real_0018", "real_0064"):
        img = real_crops[key]
        e = golden["images"][key]
        assert sha(img) == e["sha1"]
        for name, ent in e["masks"].items():
            m, w, p = CONFIGS[name]
            assert sha(CO.binarize_local(img, m, w, p, 0)) == ent["sha1"]
        assert sha(O.binarize_local(img, *CONFIGS["sauvola_w15_k0.2"], 0)) == e["masks"]["sauvola_w15_k0.2"]["sha1"]


@pytest.mark.parametrize("iters", [2, -1, 1, -3])
def test_morph_tail(noise_page, iters):
    a = CO.binarize_local(noise_page, 0, 15, (0.2,), iters)
    b = O.binarize_local(noise_page, 0, 15, (0.2,), iters)
    assert np.array_equal(a, b)


def test_morph_golden(golden, noise_page):
    e = golden["images"]["noise_512x640"]["morph"]
    assert sha(CO.binarize_local(noise_page, 0, 15, (0.2,), 2)) == e["sauvola_w15_k0.2_morph2"]
    assert sha(CO.binarize_local(noise_page, 0, 15, (0.2,), -1)) == e["sauvola_w15_k0.2_morph-1"]


def test_first_principles_numpy_matches_cv2_ops(noise_page):
    for name in ("sauvola_w15_k0.2", "niblack_w15_k-0.2", "wolfjolion_w15_k0.5", "nick_w15_k-0.1", "feng_w21_default"):
        m, w, p = CONFIGS[name]
        assert np.array_equal(O.binarize_local_numpy(noise_page, m, w, p), O.binarize_local(noise_page, m, w, p, 0))


def test_degenerate_pages():
    # all-black page: q - m*m may round negative -> NaN -> T8 = 0 -> mask 0
    z = np.zeros((64, 80), np.uint8)
    for m, w, p in ((0, 15, (0.2,)), (1, 15, (-0.2,)), (2, 15, (0.5,)), (3, 15, (-0.1,)), (4, 21, (0.75, 0.2, 0.03, 2.0))):
        a = O.binarize_local(z, m, w, p, 0)
        b = CO.binarize_local(z, m, w, p, 0)
        assert np.array_equal(a, b) and not a.any()
    # constant page, page with a black block, saturated page
    c = np.full((70, 90), 200, np.uint8)
    blk = c.copy(); blk[10:50, 20:70] = 0
    sat = np.full((70, 90), 255, np.uint8)
    for img in (c, blk, sat):
        for m, w, p in ((0, 15, (0.2,)), (1, 15, (-0.2,)), (2, 15, (0.5,)), (3, 15, (-0.1,)), (4, 21, (0.75, 0.2, 0.03, 2.0))):
            assert np.array_equal(O.binarize_local(img, m, w, p, 0), CO.binarize_local(img, m, w, p, 0))


def test_geometry_and_validation():
    assert O.output_shape(O.SAUVOLA, 3508, 2480, 15) == (3507, 2479)
    assert O.output_shape(O.NICK, 3508, 2480, 15) == (3493, 2465)
    assert CO.output_shape(O.FENG, 9921, 7016, 101) == (9820, 6915)
    # window clamped to min(rows, cols): may become even; Sauvola/Niblack then return the full size
    assert O.output_shape(O.SAUVOLA, 40, 60, 101) == (40, 60) == CO.output_shape(O.SAUVOLA, 40, 60, 101)
    small = np.random.default_rng(1).integers(0, 256, (40, 60), dtype=np.uint8)
    assert np.array_equal(O.binarize_local(small, O.SAUVOLA, 101, (0.01,), 0), CO.binarize_local(small, O.SAUVOLA, 101, (0.01,), 0))
    with pytest.raises(ValueError):
        O.binarizeSauvola(small, 14)
    with pytest.raises(ValueError):
        O.binarizeSauvola(small, 1)
    with pytest.raises(ValueError):
        O.binarizeSauvola(np.zeros((0, 0), np.uint8))
    with pytest.raises(Exception):
        O.binarize_local(small, O.NICK, 101, (-0.1,), 0)     # empty processingRect -> cv::Exception
    with pytest.raises(ValueError):
        CO.binarize_local(small, O.NICK, 101, (-0.1,), 0)


def test_otsu_recurrence_matches_opencv():
    rng = np.random.default_rng(5)
    tiles = []
    for i in range(400):
        kind = i % 5
        if kind == 0:
            t = rng.integers(0, 256, (64, 64), dtype=np.uint8)
        elif kind == 1:   # bimodal
            t = np.where(rng.random((64, 64)) < 0.3, rng.integers(20, 60, (64, 64)), rng.integers(180, 230, (64, 64))).astype(np.uint8)
        elif kind == 2:   # two values only (ties across a range of thresholds)
            a, b = sorted(rng.integers(0, 256, 2).tolist())
            t = np.where(rng.random((48, 52)) < rng.random(), a, b).astype(np.uint8)
        elif kind == 3:   # near-constant
            t = (rng.integers(100, 103, (64, 64))).astype(np.uint8)
        else:             # constant
            t = np.full((13, 17), int(rng.integers(0, 256)), np.uint8)
        tiles.append(t)
    for t in tiles:
        hist = np.bincount(t.ravel(), minlength=256)
        want = O.otsu_threshold_cv(t)
        assert O.otsu_threshold_from_hist(hist) == want
        assert CO.otsu_from_hist(hist) == want


def test_otsu_exact_ties_follow_opencv_cpp_not_ipp():
    """Symmetric histograms make sigma(t) == sigma(254 - t) exactly: the winner is decided by the
    rounding of the literal FP64 recurrence.  OpenCV's C++ code (IPP off) is the target."""
    rng = np.random.default_rng(123)
    for k in range(200):
        n = int(rng.integers(3, 18))
        a = rng.integers(0, 128, n)
        vals = np.concatenate([a, 255 - a]).astype(np.uint8)
        if k % 2:
            vals = np.concatenate([vals, [127, 128]]).astype(np.uint8)
        t = vals.reshape(1, -1).copy()
        want = O.otsu_threshold_cv(t)
        hist = np.bincount(t.ravel(), minlength=256)
        assert O.otsu_threshold_from_hist(hist) == want
        assert CO.otsu_from_hist(hist) == want
    noise = np.random.default_rng(0).integers(0, 256, (512, 640), dtype=np.uint8)
    assert np.array_equal(O.otsu_tiles(noise, 7, 5), CO.otsu_rects(noise, O.tile_rects(512, 640, 7, 5)))


def test_otsu_global_and_tiles_golden(golden):
    a4 = CO.synth_page(0)
    e = golden["images"]["a4_p0"]
    thr, mask = CO.otsu_global(a4)
    assert thr == e["otsu_global"]["thr"] and sha(mask) == e["otsu_global"]["sha1"]
    thr2, mask2 = O.otsu_global(a4)
    assert thr2 == thr and np.array_equal(mask, mask2)
    assert sha(CO.otsu_rects(a4, O.tile_rects(*a4.shape))) == e["otsu_tiles64"]


def test_otsu_rects_overlap_and_maxval(noise_page):
    rects = [(0, 0, 100, 80), (50, 40, 200, 200), (300, 100, 340, 412), (10, 300, 77, 33), (639, 511, 1, 1)]
    for mv in (255.0, 200.0, 0.0):
        a = O.otsu_rects(noise_page, rects, mv)
        b = CO.otsu_rects(noise_page, rects, mv)
        assert np.array_equal(a, b)
    # order independence (only zeros are written)
    assert np.array_equal(O.otsu_rects(noise_page, rects[::-1]), O.otsu_rects(noise_page, rects))
