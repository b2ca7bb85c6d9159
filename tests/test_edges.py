"""Edge front-end of prl::binarizeLocalOtsu: GaussianBlur, Canny, CannyEdgeDetection and the
whole function against the real OpenCV calls in the reference's order (oracle/prl_oracle.py)."""
import numpy as np
import pytest

import prlib_b200
from oracle import c_oracle as CO
from oracle import prl_oracle as O


def test_gauss_fixed_point_coefficients_reproduce_cv2_blur():
    """Host-side coefficient derivation (no GPU): a numpy 8.8 / 16.16 convolution with the library's coefficients equals
    cv2.GaussianBlur bit for bit, for many kernel sizes and sigmas."""
    import ctypes as C
    import cv2
    L = prlib_b200.capi.load()
    fn = L.prl_cuda_gauss_kernel_fixed
    fn.restype = C.c_int; fn.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_int)]
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (90, 130), dtype=np.uint8)
    for k in (3, 5, 7, 9, 13, 19, 25, 31, 41, 63):
        for sigma in (0.0, 0.8, 2.0, 3.2, 9.5):
            buf = (C.c_int * 63)()
            assert fn(k, sigma, buf) == 0
            kk = np.array(buf[:k], np.int64)
            assert kk.sum() == 256
            r = k // 2
            p = cv2.copyMakeBorder(src, r, r, r, r, cv2.BORDER_REFLECT_101).astype(np.int64)
            h = sum(p[:, i:i + src.shape[1]] * kk[i] for i in range(k))
            v = sum(h[j:j + src.shape[0], :] * kk[j] for j in range(k))
            mine = np.clip((v + 32768) >> 16, 0, 255).astype(np.uint8)
            assert np.array_equal(mine, O.gaussian_blur(src, k, sigma)), (k, sigma)
    buf = (C.c_int * 63)()
    assert fn(4, 0.0, buf) != 0 and fn(65, 0.0, buf) != 0


@pytest.mark.gpu
def test_gaussian_blur_bit_exact(ctx, noise_page):
    page = CO.synth_page(2, 700, 900)
    for img in (noise_page, page, noise_page[:7, :300], noise_page[:200, :5], noise_page[:1, :1]):
        img = np.ascontiguousarray(img)
        for k, sigma in ((3, 0), (5, 0), (7, 0), (19, 0), (31, 0), (9, 1.7), (63, 0)):
            assert np.array_equal(ctx.gaussian_blur(img, k, sigma), O.gaussian_blur(img, k, sigma)), (img.shape, k, sigma)
    with pytest.raises(ValueError):
        ctx.gaussian_blur(noise_page, 4)


@pytest.mark.gpu
def test_canny_bit_exact(ctx, noise_page, real_crops):
    blurred = O.gaussian_blur(noise_page, 9)
    page = O.gaussian_blur(CO.synth_page(1, 900, 700), 19)
    flat = np.full((50, 60), 77, np.uint8)
    steps = np.zeros((64, 64), np.uint8); steps[:, 32:] = 200; steps[40:, :] = 90
    imgs = [noise_page, blurred, page, flat, steps, np.ascontiguousarray(noise_page[:3, :40]), np.ascontiguousarray(noise_page[:40, :2])]
    imgs += [np.ascontiguousarray(v) for k, v in real_crops.items() if getattr(v, "ndim", 0) == 2][:2]
    for img in imgs:
        for lo, hi in ((0.28, 28.05), (10.5, 40.2), (50, 150), (150, 50), (0, 0), (300, 900)):
            assert np.array_equal(ctx.canny(img, lo, hi), O.canny(img, lo, hi)), (img.shape, lo, hi)


@pytest.mark.gpu
def test_canny_edge_detection_chain(ctx, noise_page):
    page = CO.synth_page(0, 1100, 900)
    for img in (page, noise_page):
        for k, up, lo, it, post in ((19, 0.15, 0.01, 1, 3), (19, 0.15, 0.01, 1, 0), (5, 0.5, 0.2, 0, 0), (9, 0.3, 0.1, -1, 2),
                                    (19, 1.0, 1.0, 2, 3)):
            want = O.local_otsu_edges(img, k, up, lo, it, post)
            assert np.array_equal(ctx.canny_edge_detection(img, k, up, lo, it, post), want), (img.shape, k, up, lo, it, post)
    for bad in ((2, 0.15, 0.01), (19, 1.5, 0.01), (19, 0.1, 0.2), (19, 0.15, -0.1)):     # the reference's std::invalid_argument checks
        with pytest.raises(ValueError):
            ctx.canny_edge_detection(page, *bad)
    with pytest.raises(prlib_b200.PrlCudaError) as ei:                                   # even size: cv::GaussianBlur's assertion
        ctx.canny_edge_detection(page, 4, 0.15, 0.01)
    assert ei.value.code == prlib_b200.capi.PRL_E_EMPTY_ROI


@pytest.mark.gpu
def test_contour_rectangles_equal_findcontours(ctx, noise_page, real_crops):
    """Device rectangles == cv2.findContours(RETR_EXTERNAL) + boundingRect as a set, including pages where components
    sit inside holes of other components (those must be left out)."""
    import cv2
    pages = [CO.synth_page(0, 1500, 1300), CO.synth_page(5, 700, 2000), noise_page] + \
            [np.ascontiguousarray(v) for k, v in real_crops.items() if getattr(v, "ndim", 0) == 2]
    nested_seen = False
    for page in pages:
        edges = O.local_otsu_edges(page)
        want = sorted(O.contour_rects(edges))
        n_cc, _ = cv2.connectedComponents((edges > 0).astype(np.uint8), connectivity=8)
        nested_seen |= (n_cc - 1) > len(want)
        mask, rects = ctx.binarize_local_otsu(page, return_rects=True)
        assert sorted(map(tuple, rects.tolist())) == want, page.shape
        assert np.array_equal(mask, O.binarizeLocalOtsu(page)), page.shape
    assert nested_seen                                  # the enclosure rule was exercised


@pytest.mark.gpu
def test_binarize_local_otsu_end_to_end(ctx, real_crops):
    """prl::binarizeLocalOtsu with its header defaults: device edge map + host findContours + device rect loop equals the
    reference's OpenCV sequence."""
    page = CO.synth_page(3, 1200, 1000)
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(page), O.binarizeLocalOtsu(page))
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(page, 255.0, 0.0, 9, 0.3, 0.1, 2), O.binarizeLocalOtsu(page, 255.0, 0.0, 9, 0.3, 0.1, 2))
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(page, 100.0), O.binarizeLocalOtsu(page, 100.0))
    bgr = real_crops["bgr_0037"]
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(bgr), O.binarizeLocalOtsu(bgr))
    with pytest.raises(ValueError):
        prlib_b200.binarizeLocalOtsu(page, 300.0)
    with pytest.raises(ValueError):
        prlib_b200.binarizeLocalOtsu(np.full((80, 80), 200, np.uint8))      # no edges -> no contours -> invalid_argument


@pytest.mark.gpu
def test_binarize_local_otsu_one_pixel_wide_images(ctx):
    """1 x N and N x 1 images: every contour is a straight run with fewer than 3 CHAIN_APPROX_SIMPLE points, which
    CheckHierarhyLevelRecursively ignores (imageLibCommon.cpp:729-736) -> the result stays 255; 2 x N is
    the first shape whose contours survive."""
    rng = np.random.default_rng(3)
    for shape in ((1, 200), (200, 1), (2, 200), (200, 2), (1, 37), (3, 64)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        img[..., :max(1, shape[1] // 2)] //= 4
        img[:max(1, shape[0] // 2)] //= 2
        want = O.binarizeLocalOtsu(img)
        got = prlib_b200.binarizeLocalOtsu(img)
        assert np.array_equal(got, want), shape
    for shape in ((1, 1), (3, 3)):
        with pytest.raises(ValueError):
            prlib_b200.binarizeLocalOtsu(rng.integers(0, 256, shape, dtype=np.uint8))


def _page_with_rules(rows, cols, seed):
    page = CO.synth_page(seed, rows, cols).copy()
    rng = np.random.default_rng(seed)
    for _ in range(6):
        y = int(rng.integers(10, rows - 10)); x0 = int(rng.integers(0, cols // 2)); x1 = int(rng.integers(cols // 2, cols))
        page[y:y + int(rng.integers(1, 4)), x0:x1] = rng.integers(0, 60)
    for _ in range(5):
        x = int(rng.integers(10, cols - 10)); y0 = int(rng.integers(0, rows // 2)); y1 = int(rng.integers(rows // 2, rows))
        page[y0:y1, x:x + int(rng.integers(1, 4))] = rng.integers(0, 60)
    return page


@pytest.mark.gpu
def test_remove_lines_equals_the_opencv_sequence(ctx, noise_page, real_crops):
    """prl::removeLines (a Global-Otsu caller): inverted Otsu threshold, long 1-D openings with even and odd
    element lengths (asymmetric anchors), subtraction, inversion."""
    imgs = [_page_with_rules(900, 1237, 1), _page_with_rules(1500, 1000, 2), _page_with_rules(700, 2000, 3),
            _page_with_rules(411, 333, 4), noise_page, np.ascontiguousarray(noise_page[:50, :50]),
            np.ascontiguousarray(noise_page[:99, :150]), real_crops["bgr_0037"]]
    imgs += [np.ascontiguousarray(v) for k, v in real_crops.items() if getattr(v, "ndim", 0) == 2]
    for img in imgs:
        assert np.array_equal(prlib_b200.removeLines(img), O.removeLines(img)), img.shape
    import cv2
    small = np.zeros((40, 200), np.uint8)
    with pytest.raises(cv2.error):                       # the reference: zero-sized structuring element -> cv::Exception
        O.removeLines(small)
    with pytest.raises(prlib_b200.PrlCudaError) as ei:   # here: PRL_E_EMPTY_ROI, which the shim rethrows as cv::Exception
        prlib_b200.removeLines(small)
    assert ei.value.code == prlib_b200.capi.PRL_E_EMPTY_ROI


@pytest.mark.gpu
def test_external_rects_on_adversarial_topologies(ctx):
    """The union-find labelling against cv2.findContours(RETR_EXTERNAL): nested rings (components inside holes are
    dropped, components inside components inside holes too), diagonal-only links (8-connected foreground), diagonal
    gaps that do NOT open a ring (4-connected background), components touching every border, random masks."""
    import cv2
    def want(mask):
        cs, _ = cv2.findContours(mask.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        return sorted(cv2.boundingRect(c) for c in cs)
    masks = []
    m = np.zeros((40, 50), np.uint8)
    m[2:30, 3:40] = 255; m[5:27, 6:37] = 0; m[8:24, 9:34] = 255; m[11:21, 12:31] = 0; m[14:18, 15:28] = 255      # ring in ring in ring
    m[33:38, 1:9] = 255; m[35, 4] = 0                                                                              # ring with a 1-pixel hole
    masks.append(m)
    m = np.zeros((30, 30), np.uint8)
    for i in range(25): m[i, i] = 255                                                                              # a pure diagonal: one component
    m[3, 20] = 255; m[4, 21] = 255; m[5, 20] = 255; m[4, 19] = 255                                                 # diamond: the centre hole is closed
    m[4, 20] = 0
    masks.append(m)
    m = np.zeros((20, 20), np.uint8)
    m[5:15, 5:15] = 255; m[7:13, 7:13] = 0; m[9:11, 9:11] = 255                                                    # square ring with an island ...
    m[5, 5] = 0                                                                                                    # ... and a missing corner pixel: still closed for 4-connected background
    masks.append(m)
    m = np.full((16, 23), 255, np.uint8); m[3:13, 4:19] = 0; m[6:9, 8:12] = 255                                    # frame touching all borders, island inside
    masks.append(m)
    masks.append(np.zeros((9, 9), np.uint8))
    masks.append(np.full((7, 5), 255, np.uint8))
    rng = np.random.default_rng(7)
    for shape, p in (((64, 64), 0.5), ((64, 64), 0.62), ((120, 333), 0.55), ((257, 300), 0.4), ((33, 1000), 0.6), ((1, 50), 0.5),
                     ((50, 1), 0.5), ((200, 200), 0.7)):
        masks.append(np.where(rng.random(shape) < p, 255, 0).astype(np.uint8))
    coarse = np.kron(np.where(rng.random((40, 40)) < 0.55, 255, 0).astype(np.uint8), np.ones((5, 5), np.uint8))   # blobs with holes
    masks.append(np.ascontiguousarray(coarse))
    for i, m in enumerate(masks):
        got = sorted(map(tuple, ctx.external_rects(m).tolist()))
        assert got == want(m), (i, m.shape)


@pytest.mark.gpu
def test_clahe_and_local_otsu_with_clahe(ctx, noise_page, real_crops):
    """EnhanceLocalContrastByCLAHE (cv::CLAHE 8x8 tiles + equalizeHist) bit for bit, sizes that are and are not multiples
    of the tile grid, a one-grey-level image; then prl::binarizeLocalOtsu with CLAHEClipLimit > 0."""
    imgs = [noise_page, CO.synth_page(2, 701, 903), CO.synth_page(4, 160, 240), np.ascontiguousarray(noise_page[:37, :53]),
            np.full((64, 80), 93, np.uint8)]
    imgs += [np.ascontiguousarray(v) for k, v in real_crops.items() if getattr(v, "ndim", 0) == 2][:2]
    for img in imgs:
        for clip in (0.5, 2.0, 4.0, 40.0):
            for eq in (False, True):
                assert np.array_equal(ctx.clahe(img, clip, eq), O.enhance_local_contrast_clahe(img, clip, eq)), (img.shape, clip, eq)
    page = CO.synth_page(3, 1000, 800)
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(page, 255.0, 2.0), O.binarizeLocalOtsu(page, 255.0, 2.0))
    bgr = real_crops["bgr_0037"]
    assert np.array_equal(prlib_b200.binarizeLocalOtsu(bgr, 255.0, 4.0), O.binarizeLocalOtsu(bgr, 255.0, 4.0))


@pytest.mark.gpu
def test_batched_local_otsu_and_remove_lines_equal_the_per_image_oracle(ctx):
    """prl_cuda_binarize_local_otsu_batch_dev / prl_cuda_remove_lines_batch_dev: pages resident in HBM, 16 page lanes side by
    side; every page must equal the OpenCV call sequence of the reference for that page, a page without contours reports
    the reference's std::invalid_argument through its status word and leaves the others alone."""
    import torch
    from prlib_b200 import capi
    n, rows, cols = 37, 420, 610                       # more pages than lanes, odd pitch-free sizes
    host = np.stack([CO.synth_page(60 + p, rows, cols) for p in range(n)])
    host[5] = 200                                      # flat page: no edges -> "Contours array is empty"
    host[9] = np.random.default_rng(9).integers(0, 256, (rows, cols), dtype=np.uint8)
    step = (cols + 15) // 16 * 16
    buf = torch.zeros((n, rows, step), dtype=torch.uint8, device="cuda:0")
    buf[:, :, :cols] = torch.from_numpy(host).to("cuda:0")
    out = torch.full((n, rows, step), 7, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    n_rects, status = ctx.binarize_local_otsu_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, out.data_ptr(), step, rows * step)
    got = out.cpu().numpy()[:, :, :cols]
    for p in range(n):
        if p == 5:
            assert status[p] == capi.PRL_E_INVALID and n_rects[p] == 0 and (got[p] == 7).all()
            with pytest.raises(ValueError):
                O.binarizeLocalOtsu(host[p])
            continue
        assert status[p] == 0 and n_rects[p] > 0, p
        assert np.array_equal(got[p], O.binarizeLocalOtsu(host[p])), p
    # with the CLAHE pre-step and a maxValue below 255 (the whole rectangle turns black, binarizeLocalOtsu.cpp:159)
    n2 = 5
    ctx.binarize_local_otsu_batch_dev(buf.data_ptr(), n2, rows, cols, step, rows * step, out.data_ptr(), step, rows * step,
                                      maxval=200.0, clahe_clip_limit=2.0, ksize=11, morph_iters=2)
    got = out.cpu().numpy()[:, :, :cols]
    for p in range(n2):
        assert np.array_equal(got[p], O.binarizeLocalOtsu(host[p], 200.0, 2.0, 11, 0.15, 0.01, 2)), p
    out.fill_(7)
    ctx.remove_lines_batch_dev(buf.data_ptr(), n, rows, cols, step, rows * step, out.data_ptr(), step, rows * step)
    got = out.cpu().numpy()[:, :, :cols]
    for p in range(n):
        assert np.array_equal(got[p], O.removeLines(host[p])), p
    with pytest.raises(prlib_b200.PrlCudaError) as e:
        ctx.remove_lines_batch_dev(buf.data_ptr(), 2, 40, cols, step, rows * step, out.data_ptr(), step, rows * step)
    assert e.value.code == capi.PRL_E_EMPTY_ROI
