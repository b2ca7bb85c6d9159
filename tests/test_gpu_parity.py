"""GPU parity tests: every call goes through the C-ABI (libprlib_cuda) and is checked against the
oracle (oracle/c_oracle.py = plain-C restatement, oracle/prl_oracle.py = real OpenCV ops) and the
committed golden digests.  Bars: integrals bit-exact; masks bit-exact on every case here (the allowed
tolerance is <= 1e-5 mismatching pixels, each with T within 1e-6 of a rounding boundary n + 1/2 --
checked by `assert_mask_parity`); Otsu thresholds exactly equal."""
import numpy as np
import pytest

import prlib_b200
from prlib_b200 import capi
from oracle import c_oracle as CO
from oracle import prl_oracle as O
from util import CONFIGS, sha

pytestmark = pytest.mark.gpu

MASK_MISMATCH_MAX = 1e-5   # fraction of pixels (north_star)
T_EPS = 1e-6               # |T - (n + 1/2)| bound for any mismatching pixel


def assert_mask_parity(gpu_mask, img, method, window, params):
    want = CO.binarize_local(img, method, window, params, 0)
    assert gpu_mask.shape == want.shape
    bad = gpu_mask != want
    n_bad = int(bad.sum())
    if n_bad == 0:
        return 0
    assert n_bad / bad.size <= MASK_MISMATCH_MAX, f"{n_bad} mismatching pixels"
    T, _, _, _ = O.threshold_map_f64(img, method, window, params)
    frac = np.abs(T[bad] - (np.floor(T[bad]) + 0.5))
    assert np.all(frac < T_EPS), "a mismatching pixel is not at a rounding boundary of T"
    return n_bad


def test_integral_bit_exact_noise_and_golden(ctx, golden, noise_page):
    for pad in (0, 1, 7, 50):
        S, Q = ctx.integral(noise_page, pad)
        Sw, Qw = CO.integrals_int64(noise_page, pad)
        assert np.array_equal(S, Sw) and np.array_equal(Q, Qw)
    e = golden["images"]["noise_512x640"]["integral_pad7"]
    S, Q = ctx.integral(noise_page, 7)
    assert sha(S) == e["S_sha1"] and sha(Q) == e["Q_sha1"]


def test_integral_bit_exact_a4_golden(ctx, golden):
    a4 = CO.synth_page(0)
    S, Q = ctx.integral(a4, 7)
    e = golden["images"]["a4_p0"]["integral_pad7"]
    assert sha(S) == e["S_sha1"] and sha(Q) == e["Q_sha1"]
    assert int(S[-1, -1]) == e["S_last"] and int(Q[-1, -1]) == e["Q_last"]


@pytest.mark.parametrize("shape", [(1, 1), (1, 300), (300, 1), (2, 2), (33, 257), (64, 256), (65, 511), (97, 1031),
                                   (100, 2049), (37, 4096), (19, 4100), (9, 8150)])
def test_integral_ragged_shapes(ctx, shape):
    rng = np.random.default_rng(shape[0] * 10007 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    for pad in (0, 3, 10):
        if shape[1] + 2 * pad > 8192:
            continue
        S, Q = ctx.integral(img, pad)
        Sw, Qw = CO.integrals_int64(img, pad)
        assert np.array_equal(S, Sw) and np.array_equal(Q, Qw)


def test_integral_borders_wider_than_a_warp_strip(ctx):
    rng = np.random.default_rng(21)
    for shape, pad in (((40, 50), 150), ((33, 300), 130), ((20, 700), 260), ((25, 1), 140), ((9, 2600), 200)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        S, Q = ctx.integral(img, pad)
        Sw, Qw = CO.integrals_int64(img, pad)
        assert np.array_equal(S, Sw) and np.array_equal(Q, Qw), (shape, pad)


def test_integral_saturated_large_needs_64_bit(ctx):
    img = np.full((3000, 3000), 255, np.uint8)     # S exceeds 2^31, Q exceeds 2^32
    S, Q = ctx.integral(img, 0)
    assert int(S[-1, -1]) == 255 * 9_000_000 and int(Q[-1, -1]) == 65025 * 9_000_000
    Sw, Qw = CO.integrals_int64(img, 0)
    assert np.array_equal(S, Sw) and np.array_equal(Q, Qw)


def test_integral_strided_input(ctx):
    rng = np.random.default_rng(11)
    big = rng.integers(0, 256, (200, 700), dtype=np.uint8)
    view = big[10:150, 33:600]          # step != cols, unaligned base
    S, Q = ctx.integral(view, 7)
    Sw, Qw = CO.integrals_int64(np.ascontiguousarray(view), 7)
    assert np.array_equal(S, Sw) and np.array_equal(Q, Qw)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_noise_masks_and_threshold_maps(ctx, golden, noise_page, name):
    m, w, p = CONFIGS[name]
    e = golden["images"]["noise_512x640"]
    mask = ctx.binarize_local(noise_page, m, w, p, 0)
    t8, aux = ctx.threshold_map(noise_page, m, w, p)
    assert sha(t8) == e["t8"][name], "u8 threshold surface differs from the reference arithmetic"
    assert assert_mask_parity(mask, noise_page, m, w, p) == 0
    assert sha(mask) == e["masks"][name]["sha1"]
    if m == capi.WOLFJOLION:
        assert aux["imin"] == e["aux"][name]["imin"] and aux["smax"] == e["aux"][name]["smax"]


@pytest.mark.parametrize("name", ["sauvola_w15_k0.2", "niblack_w15_k-0.2", "wolfjolion_w15_k0.5", "nick_w15_k-0.1",
                                  "feng_w21_default", "sauvola_w101_k0.01"])
def test_a4_masks_golden(ctx, golden, name):
    m, w, p = CONFIGS[name]
    a4 = CO.synth_page(0)
    mask = ctx.binarize_local(a4, m, w, p, 0)
    e = golden["images"]["a4_p0"]["masks"][name]
    assert list(mask.shape) == e["shape"]
    assert sha(mask) == e["sha1"]


@pytest.mark.parametrize("name", ["nick_w101_k-0.1", "feng_w101_default"])
def test_a3_600dpi_large_window_golden(ctx, golden, name):
    m, w, p = CONFIGS[name]
    a3 = CO.synth_page(0, 9921, 7016)
    assert sha(a3) == golden["images"]["a3_600_p0"]["sha1"]
    mask = ctx.binarize_local(a3, m, w, p, 0)
    e = golden["images"]["a3_600_p0"]["masks"][name]
    assert list(mask.shape) == e["shape"] and sha(mask) == e["sha1"]


def test_a3_integral_golden(ctx, golden):
    a3 = CO.synth_page(0, 9921, 7016)
    S, Q = ctx.integral(a3, 50)
    e = golden["images"]["a3_600_p0"]["integral_pad50"]
    assert sha(S) == e["S_sha1"] and sha(Q) == e["Q_sha1"]


def test_real_page_crops(ctx, golden, real_crops):
    for key in ("real_0037", "real_0018", "real_0064"):
        img = real_crops[key]
        for name, ent in golden["images"][key]["masks"].items():
            m, w, p = CONFIGS[name]
            assert sha(ctx.binarize_local(img, m, w, p, 0)) == ent["sha1"], (key, name)


def test_bgr_front_step(ctx, golden, real_crops):
    bgr = real_crops["bgr_0037"]
    g = ctx.bgr2gray(bgr)
    assert sha(g) == golden["bgr_0037_gray_sha1"]
    assert np.array_equal(prlib_b200.binarizeSauvola(bgr, 15, 0.2, 0), O.binarizeSauvola(bgr, 15, 0.2, 0))
    bgra = np.dstack([bgr, np.full(bgr.shape[:2], 255, np.uint8)])
    assert np.array_equal(ctx.bgr2gray(bgra), g)


@pytest.mark.parametrize("iters", [1, 2, 3, -1, -2])
def test_morph_tail(ctx, noise_page, iters):
    for m, w, p in ((0, 15, (0.2,)), (3, 15, (-0.1,))):
        got = ctx.binarize_local(noise_page, m, w, p, iters)
        assert np.array_equal(got, CO.binarize_local(noise_page, m, w, p, iters))
    raw = CO.binarize_local(noise_page, 0, 15, (0.2,), 0)
    assert np.array_equal(ctx.morph(raw, iters), O.morph(raw, iters))


@pytest.mark.parametrize("shape", [(497, 625), (70, 2000), (300, 961), (33, 31), (5, 1000)])
def test_morph_bit_kernels_equal_byte_kernels_and_oracle(ctx, shape):
    """Bit-packed closing/opening (binary masks) vs the byte kernels vs cv2 dilate/erode, strip edges and n up to 15."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    mask = np.where(rng.random(shape) < 0.35, 0, 255).astype(np.uint8)
    mask[:, : shape[1] // 3][rng.random((shape[0], shape[1] // 3)) < 0.5] = 255
    for iters in (1, 2, 3, 4, 8, 15, -1, -2, -5, -15):
        want = O.morph(mask, iters)
        got = ctx.morph(mask, iters)
        assert np.array_equal(got, want), iters
        ctx.set_option("morph_bytes", 1)
        try:
            assert np.array_equal(ctx.morph(mask, iters), want), iters
        finally:
            ctx.set_option("morph_bytes", 0)
    gray = rng.integers(0, 256, shape, dtype=np.uint8)              # not a mask: byte kernels, still cv2's result
    assert np.array_equal(ctx.morph(gray, 2), O.morph(gray, 2))


def test_pages_wider_than_8192_columns(ctx):
    """Kernel 1 chains column passes of 2560 columns: integrals bit-exact and masks identical on a 9000-column strip."""
    rng = np.random.default_rng(3)
    page = rng.integers(0, 256, (120, 9000), dtype=np.uint8)
    page[:, 4000:4700] = 255
    S, Q = ctx.integral(page, 10)
    Sw, Qw = CO.integrals_int64(page, 10)
    assert np.array_equal(S, Sw) and np.array_equal(Q, Qw)
    for m, w, p in ((0, 21, (0.2,)), (2, 21, (0.5,)), (3, 51, (-0.1,))):
        assert np.array_equal(ctx.binarize_local(page, m, w, p, 0), CO.binarize_local(page, m, w, p, 0)), m


def test_header_defaults_end_to_end(ctx, golden):
    a4 = CO.synth_page(0)
    got = prlib_b200.binarizeSauvola(a4)            # w=101, k=0.01, morph=2 (binarizeSauvola.h:45-47)
    assert sha(got) == golden["images"]["a4_p0"]["morph"]["sauvola_w101_k0.01_morph2"]
    small = CO.synth_page(2, 400, 520)
    assert np.array_equal(prlib_b200.binarizeNICK(small), O.binarizeNICK(small))
    assert np.array_equal(prlib_b200.binarizeFeng(small), O.binarizeFeng(small))
    assert np.array_equal(prlib_b200.binarizeWolfJolion(small), O.binarizeWolfJolion(small))
    assert np.array_equal(prlib_b200.binarizeNiblack(small), O.binarizeNiblack(small))


def test_degenerate_pages(ctx):
    z = np.zeros((64, 80), np.uint8)
    c = np.full((70, 90), 200, np.uint8)
    blk = c.copy(); blk[10:50, 20:70] = 0
    sat = np.full((70, 90), 255, np.uint8)
    for img in (z, c, blk, sat):
        for m, w, p in ((0, 15, (0.2,)), (1, 15, (-0.2,)), (2, 15, (0.5,)), (3, 15, (-0.1,)), (4, 21, (0.75, 0.2, 0.03, 2.0))):
            got = ctx.binarize_local(img, m, w, p, 0)
            assert np.array_equal(got, O.binarize_local(img, m, w, p, 0)), (m, img[0, 0])


def test_extreme_coefficients_saturate_like_cvround(ctx, noise_page):
    for m, w, p in ((0, 15, (1e12,)), (0, 15, (-1e12,)), (1, 15, (1e11,)), (1, 15, (-50.0,)), (3, 15, (40.0,)), (2, 15, (1e13,))):
        got, _ = ctx.threshold_map(noise_page, m, w, p)
        want = O.threshold_map(noise_page, m, w, p)
        assert np.array_equal(got, want), (m, p)


def test_window_clamped_to_small_image(ctx):
    rng = np.random.default_rng(4)
    small = rng.integers(0, 256, (40, 60), dtype=np.uint8)
    for m in (capi.SAUVOLA, capi.NIBLACK):              # w clamps to 40 (even): full-size output
        got = ctx.binarize_local(small, m, 101, (0.05,), 0)
        assert got.shape == (40, 60)
        assert np.array_equal(got, O.binarize_local(small, m, 101, (0.05,), 0))
    for m in (capi.WOLFJOLION, capi.NICK, capi.FENG):   # empty processingRect: the reference throws cv::Exception
        with pytest.raises(prlib_b200.PrlCudaError) as ei:
            ctx.binarize_local(small, m, 101, (0.05, 0.2, 0.03, 2.0), 0)
        assert ei.value.code == capi.PRL_E_EMPTY_ROI
    with pytest.raises(ValueError):
        ctx.binarize_local(small, capi.SAUVOLA, 16, (0.2,), 0)


def test_random_shapes_all_methods(ctx):
    rng = np.random.default_rng(2024)
    for i in range(12):
        rows, cols = int(rng.integers(30, 500)), int(rng.integers(30, 900))
        kind = i % 3
        if kind == 0:
            img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        elif kind == 1:
            img = CO.synth_page(i, rows, cols, seed=99)
        else:
            img = (rng.integers(0, 2, (rows, cols)) * 255).astype(np.uint8)     # pure black/white
        w = int(rng.integers(1, 14)) * 2 + 1
        if min(rows, cols) <= w:
            continue
        for m, p in ((0, (0.2,)), (1, (-0.2,)), (2, (0.5,)), (3, (-0.1,)), (4, (0.75, 0.2, 0.03, 2.0))):
            got = ctx.binarize_local(img, m, w, p, 0)
            assert assert_mask_parity(got, img, m, w, p) == 0, (i, m, w, rows, cols)


def test_fast_decision_path_equals_literal_fp64_path(ctx, noise_page, real_crops):
    """Kernel 2 decides most pixels from exact integer window sums + an FP32 estimate and runs the
    reference's FP64 formula only within a proven margin of the rounding boundary; forcing the FP64
    formula for every pixel must give byte-identical masks (and both must equal the oracle)."""
    rng = np.random.default_rng(77)
    ramp = (np.add.outer(np.arange(300), np.arange(420)) % 256).astype(np.uint8)
    flat = np.repeat(np.arange(256, dtype=np.uint8), 3)[None, :].repeat(64, 0)            # every gray level, flat columns
    sparse = np.zeros((200, 260), np.uint8); sparse[rng.integers(0, 200, 40), rng.integers(0, 260, 40)] = rng.integers(1, 256, 40)
    border = real_crops["real_0018"].copy(); border[:, :60] = 0; border[:50, :] = 0          # scanner-style black border
    images = [noise_page, ramp, flat, sparse, border, real_crops["real_0037"]]
    cases = [(0, 15, (0.2,)), (0, 101, (0.01,)), (0, 15, (-3.0,)), (1, 15, (-0.2,)), (1, 31, (2.5,)), (2, 15, (0.5,)),
             (2, 41, (0.01,)), (3, 15, (-0.1,)), (3, 51, (0.7,)), (4, 21, (0.75, 0.2, 0.03, 2.0)), (4, 15, (0.2, 0.2, 0.9, 2.0))]
    for img in images:
        for m, w, p in cases:
            if min(img.shape) <= w:
                continue
            fast = ctx.binarize_local(img, m, w, p, 0)
            ctx.set_option("exact_threshold", 1)
            exact = ctx.binarize_local(img, m, w, p, 0)
            ctx.set_option("exact_threshold", 0)
            assert np.array_equal(fast, exact), (m, w, p, img.shape)
            assert assert_mask_parity(fast, img, m, w, p) == 0, (m, w, p, img.shape)


def test_integral_generic_kernel_equals_tma_kernel(ctx, noise_page):
    a4 = CO.synth_page(5, 900, 2480)
    for img, pad in ((noise_page, 7), (a4, 7), (a4, 50), (noise_page[:, :333], 10)):
        S1, Q1 = ctx.integral(img, pad)
        ctx.set_option("disable_tma", 1)
        S2, Q2 = ctx.integral(img, pad)
        ctx.set_option("disable_tma", 0)
        Sw, Qw = CO.integrals_int64(np.ascontiguousarray(img), pad)
        assert np.array_equal(S1, Sw) and np.array_equal(Q1, Qw)
        assert np.array_equal(S2, Sw) and np.array_equal(Q2, Qw)


# ---- Otsu ------------------------------------------------------------------------------------
def test_otsu_global_golden_and_cv(ctx, golden, noise_page, real_crops):
    a4 = CO.synth_page(0)
    thr, mask = ctx.otsu_global(a4)
    e = golden["images"]["a4_p0"]["otsu_global"]
    assert thr == e["thr"] and sha(mask) == e["sha1"]
    assert ctx.otsu_threshold(a4) == e["thr"]
    for img in (noise_page, real_crops["real_0037"], real_crops["real_0018"], real_crops["real_0064"]):
        for mv in (255.0, 128.0):
            t, m = ctx.otsu_global(img, mv)
            tw, mw = O.otsu_global(img, mv)
            assert t == tw and np.array_equal(m, mw)


def test_otsu_thresholds_exact_on_adversarial_tiles(ctx):
    rng = np.random.default_rng(5)
    for i in range(60):
        kind = i % 5
        if kind == 0:
            t = rng.integers(0, 256, (64, 64), dtype=np.uint8)
        elif kind == 1:
            t = np.where(rng.random((64, 64)) < 0.3, rng.integers(20, 60, (64, 64)), rng.integers(180, 230, (64, 64))).astype(np.uint8)
        elif kind == 2:
            a, b = sorted(rng.integers(0, 256, 2).tolist())
            t = np.where(rng.random((48, 52)) < rng.random(), a, b).astype(np.uint8)
        elif kind == 3:
            t = (rng.integers(100, 103, (64, 64))).astype(np.uint8)
        else:
            t = np.full((13, 17), int(rng.integers(0, 256)), np.uint8)
        assert ctx.otsu_threshold(t) == O.otsu_threshold_cv(t), i


def test_otsu_exact_ties(ctx):
    rng = np.random.default_rng(123)
    for k in range(60):
        n = int(rng.integers(3, 18))
        a = rng.integers(0, 128, n)
        vals = np.concatenate([a, 255 - a]).astype(np.uint8)
        if k % 2:
            vals = np.concatenate([vals, [127, 128]]).astype(np.uint8)
        t = vals.reshape(1, -1).copy()
        assert ctx.otsu_threshold(t) == O.otsu_threshold_cv(t), k


def test_otsu_tiles_golden(ctx, golden, noise_page):
    a4 = CO.synth_page(0)
    assert sha(ctx.otsu_tiles(a4, 64, 64)) == golden["images"]["a4_p0"]["otsu_tiles64"]
    for tw, th in ((64, 64), (48, 52), (100, 37), (640, 512), (7, 5)):
        assert np.array_equal(ctx.otsu_tiles(noise_page, tw, th), O.otsu_tiles(noise_page, tw, th)), (tw, th)
    odd = noise_page[:301, :333]
    assert np.array_equal(ctx.otsu_tiles(odd, 64, 64), O.otsu_tiles(np.ascontiguousarray(odd), 64, 64))


def test_otsu_rects_overlapping_and_maxval(ctx, noise_page, real_crops):
    rects = [(0, 0, 100, 80), (50, 40, 200, 200), (300, 100, 340, 412), (10, 300, 77, 33), (639, 511, 1, 1), (0, 0, 640, 512)]
    for mv in (255.0, 200.0, 0.0):
        got, thr = ctx.otsu_rects(noise_page, rects, mv, return_thresholds=True)
        assert np.array_equal(got, O.otsu_rects(noise_page, rects, mv))
        for (x, y, w, h), t in zip(rects, thr):
            assert int(t) == O.otsu_threshold_cv(noise_page[y:y + h, x:x + w])
    img = real_crops["real_0018"]
    rects = [(5, 7, 300, 200), (250, 150, 500, 400), (0, 0, 801, 640)]
    assert np.array_equal(ctx.otsu_rects(img, rects), O.otsu_rects(img, rects))
    assert np.array_equal(ctx.otsu_rects(img, []), np.full(img.shape, 255, np.uint8))
    with pytest.raises(ValueError):
        ctx.otsu_rects(img, [(700, 600, 200, 100)])
