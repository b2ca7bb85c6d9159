"""Seeded random sweep of the five local-statistics binarizers against the plain-C oracle: window sizes from 3 up to
beyond the image, coefficients of both signs, morphology of both signs, image shapes down to the smallest the
reference accepts, flat and bimodal content.  Every mask must be byte-identical (or both sides must refuse)."""
import numpy as np
import pytest

import prlib_b200
from prlib_b200 import capi
from oracle import c_oracle as CO


def _image(rng, rows, cols, kind):
    if kind == 0:
        return rng.integers(0, 256, (rows, cols), dtype=np.uint8)
    if kind == 1:                                             # bimodal text-like
        img = np.full((rows, cols), 200, np.int32) + rng.integers(-12, 13, (rows, cols))
        ink = rng.random((rows, cols)) < 0.12
        img[ink] = rng.integers(10, 70, int(ink.sum()))
        return np.clip(img, 0, 255).astype(np.uint8)
    if kind == 2:                                             # flat with a few outliers: zero variance almost everywhere
        img = np.full((rows, cols), int(rng.integers(0, 256)), np.uint8)
        for _ in range(5):
            img[rng.integers(0, rows), rng.integers(0, cols)] = rng.integers(0, 256)
        return img
    return CO.synth_page(int(rng.integers(0, 1000)), rows, cols)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("seed", range(6))
def test_random_parameters_and_shapes(ctx, seed, fused):
    """fused = 1: windows up to 31 take the opt-in fused small-window path (planes never reach HBM)."""
    ctx.set_option("enable_fused", fused)
    try:
        _sweep(ctx, seed)
    finally:
        ctx.set_option("enable_fused", 0)


def _sweep(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    for case in range(14):
        rows, cols = int(rng.integers(2, 400)), int(rng.integers(2, 700))
        if case % 5 == 0:
            rows, cols = int(rng.integers(2, 40)), int(rng.integers(2, 60))
        window = int(rng.choice([3, 5, 7, 15, 21, 31, 33, 51, 101, 151, 255]))
        method = int(rng.integers(0, 5))
        k = float(rng.choice([-1.5, -0.5, -0.2, -0.01, 0.01, 0.2, 0.5, 0.75, 1.3]))
        params = {capi.FENG: (float(rng.choice([0.1, 0.75, 1.9])), 0.2, float(rng.choice([0.03, 0.5, 2.0])), 2.0)}.get(method, (k,))
        morph = int(rng.choice([0, 0, 1, 2, -1, -3]))
        img = _image(rng, rows, cols, int(rng.integers(0, 4)))
        tag = (seed, case, rows, cols, window, method, params, morph)
        try:
            want = CO.binarize_local(img, method, window, params, morph)
        except ValueError:
            with pytest.raises((ValueError, prlib_b200.PrlCudaError)):
                ctx.binarize_local(img, method, window, params, morph)
            continue
        got = ctx.binarize_local(img, method, window, params, morph)
        assert got.shape == want.shape and np.array_equal(got, want), tag


@pytest.mark.gpu
def test_random_otsu_tiles_rects_and_global(ctx):
    """Random tile sizes (vector path, byte path, tiles larger than the image, area above 65535), random rectangles with
    overlaps, random maxValue -- against cv2's THRESH_OTSU (IPP off)."""
    from oracle import prl_oracle as O
    rng = np.random.default_rng(77)
    for case in range(16):
        rows, cols = int(rng.integers(1, 300)), int(rng.integers(1, 500))
        img = _image(rng, rows, cols, int(rng.integers(0, 4)))
        tw = int(rng.choice([1, 3, 16, 17, 32, 48, 64, 100, 128, 256, 300, 512, 600]))
        th = int(rng.choice([1, 2, 7, 16, 33, 64, 128, 130, 400]))
        assert np.array_equal(ctx.otsu_tiles(img, tw, th), O.otsu_tiles(img, tw, th)), (case, rows, cols, tw, th)
        mv = float(rng.choice([255.0, 255.0, 100.0, 0.0, 254.6]))
        thr, mask = ctx.otsu_global(img, mv)
        thr_w, mask_w = O.otsu_global(img, mv)
        assert thr == thr_w and np.array_equal(mask, mask_w), (case, rows, cols, mv)
        rects = []
        for _ in range(int(rng.integers(0, 9))):
            x, y = int(rng.integers(0, cols)), int(rng.integers(0, rows))
            rects.append((x, y, int(rng.integers(1, cols - x + 1)), int(rng.integers(1, rows - y + 1))))
        assert np.array_equal(ctx.otsu_rects(img, rects, mv), O.otsu_rects(img, rects, mv)), (case, rects, mv)
