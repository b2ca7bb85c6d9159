"""CPU oracle for PRLib's local-statistics + Otsu binarization path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(prlib_b200 -> libprlib_cuda.so) never routes through this file.

What it is
----------
An op-for-op restatement, in Python on top of the cv2 wheel (OpenCV 4.13.0 -- the same
OpenCV primitives the C++ reference links against), of

  src/binarizations/binarizeSauvola.cpp:32-135
  src/binarizations/binarizeNiblack.cpp:32-127
  src/binarizations/binarizeWolfJolion.cpp:33-148
  src/binarizations/binarizeNICK.cpp:33-144
  src/binarizations/binarizeFeng.cpp:31-164
  src/binarizations/binarizeLocalOtsu.cpp:138-162   (per-rectangle Otsu loop)
  cv::threshold(..., THRESH_BINARY|THRESH_OTSU) call sites (src/deskew/deskew.cpp:224,
  src/removeLines.cpp:45, src/imageLibCommon.cpp:295-296)          ("Global Otsu")
  src/binarizations/binarizeLocalOtsu.cpp:38-163, src/removeLines.cpp:30-77 (the "next" rows F3 / F4)
  src/binarizations/binarizeNativeAdaptive.cpp:34-135 (bilateral step included), binarizeAT / AGT / GAT / PureAdaptive /
  PureAdaptiveGaussian.cpp                                          (the adaptive-mean family, F4)
plus first-principles numpy models of the OpenCV primitives whose arithmetic csrc/ restates (box and Gaussian means of
cv::adaptiveThreshold, cv::bilateralFilter), each checked against the real cv2 call in tests/test_adaptive.py.

Every arithmetic primitive the reference calls (copyMakeBorder, integral, filter2D, mul,
subtract, sqrt, convertTo, compare, minMaxLoc, addWeighted, divide, pow, threshold, dilate,
erode) is executed by the real OpenCV through cv2 with the same dtypes and in the same order.

Parity pinning status
---------------------
PINNED since round 2: oracle/_ref is the reference's OWN binarize*.cpp / removeLines.cpp / imageLibCommon.cpp compiled
unmodified against a cv:: facade whose primitives run in the cv2 wheel (oracle/Makefile, oracle/ref.py).  This file must
equal it output for output (tests/test_ref.py, digests in tests/golden/ref_golden.json); it stays because it also
travels to places the compiled module cannot (it needs no /root/reference) and because it exposes intermediate
results (integrals, T maps).  A second, independent plain-C restatement (oracle/prl_oracle.c) is checked bit for bit
against this one in tests/test_oracle.py.  The reference itself ships no tests and no golden outputs.
"""
from __future__ import annotations

import numpy as np
import cv2

cv2.setNumThreads(1)  # oracle-wide: one OpenCV thread.  cv2 4.13's parallel GaussianBlur was caught returning run-to-run varying rows
                       # on the GPU box (see _single_thread below); the reference's arithmetic is the single-thread result, and the
                       # CPU baseline runs one single-thread worker per core anyway.

cv2.ocl.setUseOpenCL(False)  # a UMat must never reach an OpenCL filter2D (different arithmetic)

SAUVOLA, NIBLACK, WOLFJOLION, NICK, FENG = 0, 1, 2, 3, 4
METHOD_NAMES = {SAUVOLA: "sauvola", NIBLACK: "niblack", WOLFJOLION: "wolfjolion", NICK: "nick", FENG: "feng"}

# header defaults of the reference: (windowSize, params, morphIterationCount)
#   binarizeSauvola.h:43-47, binarizeNiblack.h:43-47, binarizeWolfJolion.h:43-47,
#   binarizeNICK.h:43-47, binarizeFeng.h:46-53
DEFAULTS = {
    SAUVOLA: (101, (0.01,), 2),
    NIBLACK: (101, (0.01,), 2),
    WOLFJOLION: (101, (0.01,), 2),
    NICK: (21, (-0.01,), 0),
    FENG: (21, (0.75, 0.2, 0.03, 2.0), 2),
}


# --------------------------------------------------------------------------------------
# synthpage-v2: the deterministic synthetic page generator.
# The device generator in prlib_b200/csrc/synth.cu must produce identical bytes.
# --------------------------------------------------------------------------------------
def _mix(x):
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _key(s, p, a, b):
    s = np.uint32(s & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        return (s * np.uint32(0x9E3779B1) + np.uint32(p) * np.uint32(0x85EBCA77)
                + a.astype(np.uint32) * np.uint32(0xC2B2AE3D) + b.astype(np.uint32) * np.uint32(0x27D4EB2F))


def synth_page(page: int, rows: int = 3508, cols: int = 2480, seed: int = 2024) -> np.ndarray:
    """synthpage-v2 page `page` as a rows x cols uint8 array (values in [24, 234])."""
    with np.errstate(over="ignore"):
        y = np.arange(rows, dtype=np.uint32)[:, None]
        x = np.arange(cols, dtype=np.uint32)[None, :]
        yy = np.broadcast_to(y, (rows, cols))
        xx = np.broadcast_to(x, (rows, cols))
        noise = _mix(_key(seed, page, yy, xx)) & np.uint32(15)
        illum = (np.uint32(40) * xx) // np.uint32(cols) + (np.uint32(24) * yy) // np.uint32(rows)
        stain = (_mix(_key(seed ^ 0x5BD1E995, page, yy // np.uint32(64), xx // np.uint32(64))) >> np.uint32(4)) & np.uint32(31)
        bg = np.uint32(235) - illum - stain - noise
        band = ((yy % np.uint32(48)) < 26) & (xx >= 150) & (xx < cols - 150) & (yy >= 200) & (yy < rows - 200)
        b = _mix(_key(seed ^ 0xA5A5A5A5, page, yy // np.uint32(3), xx // np.uint32(3)))
        ink = band & ((b & np.uint32(0xFF)) < 56)
        inkv = np.uint32(24) + ((b >> np.uint32(8)) & np.uint32(127)) + (noise >> np.uint32(1))
        return np.where(ink, inkv, bg).astype(np.uint8)


# --------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------
def clamp_window(window: int, rows: int, cols: int) -> int:
    """w = min(windowSize, min(cols, rows))   binarizeSauvola.cpp:57"""
    return min(window, min(cols, rows))


def processing_rect(method: int, rows: int, cols: int, w: int):
    """(x, y, width, height) of processingRect in padded coordinates.

    Sauvola/Niblack compute it AFTER copyMakeBorder (binarizeSauvola.cpp:65-66,
    binarizeNiblack.cpp:65-66); WJ/NICK/Feng BEFORE (binarizeWolfJolion.cpp:69,
    binarizeNICK.cpp:69, binarizeFeng.cpp:66).
    """
    h = w // 2
    if method in (SAUVOLA, NIBLACK):
        return h, h, (cols + 2 * h) - w, (rows + 2 * h) - w
    return h, h, cols - w, rows - w


def output_shape(method: int, rows: int, cols: int, window: int):
    w = clamp_window(window, rows, cols)
    _, _, rw, rh = processing_rect(method, rows, cols, w)
    return rh, rw


def validate(window: int):
    """binarizeSauvola.cpp:43-47"""
    if not (window > 1 and window % 2 == 1):
        raise ValueError("Window size must satisfy the following condition: ( (windowSize > 1) && ((windowSize % 2) == 1) )")


def to_gray(image: np.ndarray) -> np.ndarray:
    """binarizeSauvola.cpp:49-52 (cvtColor BGR2GRAY when channels != 1)."""
    if image.ndim == 3 and image.shape[2] != 1:
        code = cv2.COLOR_BGR2GRAY if image.shape[2] == 3 else cv2.COLOR_BGRA2GRAY
        return cv2.cvtColor(image, code)
    return np.ascontiguousarray(image.reshape(image.shape[0], image.shape[1]))


# --------------------------------------------------------------------------------------
# the local-statistics block, shared by the five methods
# --------------------------------------------------------------------------------------
def padded_integrals(gray: np.ndarray, w: int):
    """copyMakeBorder + integral + crop (binarizeSauvola.cpp:65-77).

    Returns (pad u8, S, Q) with S, Q the INCLUSIVE (Hp x Wp) float64 integrals.
    """
    h = w // 2
    pad = cv2.copyMakeBorder(gray, h, h, h, h, cv2.BORDER_REPLICATE)
    I, I2 = cv2.integral2(pad, sdepth=cv2.CV_64F, sqdepth=cv2.CV_64F)
    return pad, I, I2  # I[1:,1:] is the inclusive integral; keep parents for the ROI trick


def integrals_int64(gray: np.ndarray, pad_h: int):
    """Inclusive padded integrals as int64 (the bit-exactness target of kernel 1)."""
    pad = cv2.copyMakeBorder(gray, pad_h, pad_h, pad_h, pad_h, cv2.BORDER_REPLICATE) if pad_h > 0 else gray
    I, I2 = cv2.integral2(pad, sdepth=cv2.CV_64F, sqdepth=cv2.CV_64F)
    S = I[1:, 1:].astype(np.int64)
    Q = I2[1:, 1:].astype(np.int64)
    assert np.array_equal(S.astype(np.float64), I[1:, 1:]) and np.array_equal(Q.astype(np.float64), I2[1:, 1:])
    return S, Q


def _roi(parent: np.ndarray, x: int, y: int, rw: int, rh: int):
    """integralImage(Rect(1,1,..))(processingRect) as a UMat ROI.

    Keeping the ROI offset makes OpenCV 4.x take the direct (non-DFT) FilterEngine path and
    read the taps from the parent matrix -- what the C++ reference executes
    (binarizeSauvola.cpp:89-90, :106-107).
    """
    return cv2.UMat(cv2.UMat(parent), [1 + y, 1 + y + rh], [1 + x, 1 + x + rw])


def local_stats(gray: np.ndarray, method: int, window: int):
    """Returns (pad, rect, m, s): local mean / deviation maps over processingRect."""
    rows, cols = gray.shape
    w = clamp_window(window, rows, cols)
    wsqr_back = 1.0 / float(w * w)
    rect = processing_rect(method, rows, cols, w)
    x, y, rw, rh = rect
    if rw <= 0 or rh <= 0:
        raise cv2.error("empty processingRect (min(rows, cols) <= windowSize)")
    pad, I, I2 = padded_integrals(gray, w)
    K = np.zeros((w, w), np.float64)
    K[0, 0] = wsqr_back
    K[w - 1, 0] = -wsqr_back
    K[w - 1, w - 1] = wsqr_back
    K[0, w - 1] = -wsqr_back
    m = cv2.filter2D(_roi(I, x, y, rw, rh), cv2.CV_64F, K, anchor=(-1, -1), delta=0.0,
                     borderType=cv2.BORDER_REFLECT).get()
    msq = cv2.multiply(m, m)
    q = cv2.filter2D(_roi(I2, x, y, rw, rh), cv2.CV_64F, K).get()
    q = cv2.subtract(q, msq)
    s = cv2.sqrt(q)
    return pad, rect, m, s


def _scale_shift(a: np.ndarray, alpha: float, beta: float) -> np.ndarray:
    """Mat::convertTo(same type, alpha, beta).  cv2 has no direct binding; G-API's convertTo on the default OpenCV CPU
    backend is a one-line wrapper around Mat::convertTo, so this IS the real call.  On FMA-capable CPUs OpenCV fuses the
    multiply-add (one rounding): round 1 spelled it cv2.add(cv2.multiply(a, alpha), beta), two roundings, which
    oracle/_ref (the reference's own C++) showed to be the wrong arithmetic on 6.5 % of the elements (last bit)."""
    g = cv2.GMat()
    comp = cv2.GComputation(cv2.GIn(g), cv2.GOut(cv2.gapi.convertTo(g, cv2.CV_64F, float(alpha), float(beta))))
    return comp.apply(cv2.gin(np.ascontiguousarray(a)))


def threshold_map_f64(gray: np.ndarray, method: int, window: int, params):
    """FP64 threshold surface T (before the u8 cast) + (pad, rect, aux dict)."""
    pad, rect, m, s = local_stats(gray, method, window)
    aux = {}
    if method == SAUVOLA:  # binarizeSauvola.cpp:115-118
        k = float(params[0])
        T = cv2.multiply(m, _scale_shift(s, k * (1.0 / 128.0), 1.0 - k))
    elif method == NIBLACK:  # binarizeNiblack.cpp:108
        k = float(params[0])
        T = cv2.scaleAdd(s, k, m)
    elif method == WOLFJOLION:  # binarizeWolfJolion.cpp:115-130
        k = float(params[0])
        imin = float(cv2.minMaxLoc(pad)[0])
        smax = float(cv2.minMaxLoc(s)[1])
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = float(np.float64(k) / np.float64(smax))
        d = _scale_shift(s, coeff, -k)
        d = cv2.multiply(d, cv2.subtract(m, imin))
        T = cv2.add(m, d)
        aux.update(imin=imin, smax=smax)
    elif method == NICK:  # binarizeNICK.cpp:121-126
        k = float(params[0])
        C = cv2.sqrt(cv2.add(cv2.multiply(m, m), cv2.multiply(s, s)))
        T = cv2.addWeighted(m, 1.0, C, k, 0.0)
    elif method == FENG:  # binarizeFeng.cpp:110-142
        alpha1, k1, k2, gamma = (float(v) for v in params)
        imin = float(cv2.minMaxLoc(pad)[0])
        t1 = cv2.divide(s, s)
        t2 = cv2.pow(t1, gamma)
        alpha3 = _scale_shift(t2, k2, 0.0)  # MatExpr k2 * tmpAlpha2 -> convertTo(alpha = k2)
        c1 = 1.0 - alpha1
        c2 = cv2.multiply(t2, t1)
        c3 = cv2.addWeighted(alpha3, imin, c2, -imin, 0.0)
        T = cv2.add(c2, c1)
        T = cv2.multiply(T, m)
        T = cv2.add(T, c3)
        aux.update(imin=imin)
    else:
        raise ValueError("unknown method")
    return T, pad, rect, aux


def to_u8(T: np.ndarray) -> np.ndarray:
    """Mat::convertTo(CV_8UC1): rint-half-even, saturate, NaN/inf -> 0."""
    return cv2.add(T, 0.0, dtype=cv2.CV_8U)


def morph(mask: np.ndarray, iters: int) -> np.ndarray:
    """binarizeSauvola.cpp:125-134"""
    if iters > 0:
        mask = cv2.dilate(mask, None, anchor=(-1, -1), iterations=iters)
        mask = cv2.erode(mask, None, anchor=(-1, -1), iterations=iters)
    elif iters < 0:
        mask = cv2.erode(mask, None, anchor=(-1, -1), iterations=-iters)
        mask = cv2.dilate(mask, None, anchor=(-1, -1), iterations=-iters)
    return mask


def threshold_map(gray: np.ndarray, method: int, window: int, params) -> np.ndarray:
    T, _, _, _ = threshold_map_f64(gray, method, window, params)
    return to_u8(T)


def binarize_local(image: np.ndarray, method: int, window=None, params=None, morph_iters=None,
                   return_aux: bool = False):
    """prl::binarize{Sauvola,Niblack,WolfJolion,NICK,Feng} restated op for op."""
    dw, dp, dm = DEFAULTS[method]
    window = dw if window is None else window
    params = dp if params is None else tuple(np.atleast_1d(params).tolist())
    morph_iters = dm if morph_iters is None else morph_iters
    if image is None or image.size == 0:
        raise ValueError("Input image for binarization is empty")
    validate(window)
    gray = to_gray(image)
    T, pad, (x, y, rw, rh), aux = threshold_map_f64(gray, method, window, params)
    T8 = to_u8(T)
    out = cv2.compare(np.ascontiguousarray(pad[y:y + rh, x:x + rw]), T8, cv2.CMP_GT)
    out = morph(out, morph_iters)
    if return_aux:
        aux.update(T=T, T8=T8)
        return out, aux
    return out


def binarizeSauvola(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2):
    return binarize_local(image, SAUVOLA, windowSize, (thresholdCoefficient,), morphIterationCount)


def binarizeNiblack(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2):
    return binarize_local(image, NIBLACK, windowSize, (thresholdCoefficient,), morphIterationCount)


def binarizeWolfJolion(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2):
    return binarize_local(image, WOLFJOLION, windowSize, (thresholdCoefficient,), morphIterationCount)


def binarizeNICK(image, windowSize=21, thresholdCoefficient=-0.01, morphIterationCount=0):
    return binarize_local(image, NICK, windowSize, (thresholdCoefficient,), morphIterationCount)


def binarizeFeng(image, windowSize=21, thresholdCoefficient_alpha1=0.75, thresholdCoefficient_k1=0.2,
                 thresholdCoefficient_k2=0.03, thresholdCoefficient_gamma=2.0, morphIterationCount=2):
    return binarize_local(image, FENG, windowSize,
                          (thresholdCoefficient_alpha1, thresholdCoefficient_k1,
                           thresholdCoefficient_k2, thresholdCoefficient_gamma), morphIterationCount)


# --------------------------------------------------------------------------------------
# Otsu
# --------------------------------------------------------------------------------------
class _no_ipp:
    """Run OpenCV's own C++ getThreshVal_Otsu_8u, not Intel IPP's closed ippiComputeThreshold_Otsu.

    The cv2 wheel bundles IPP (ICV 2022.2) and routes THRESH_OTSU through it; IPP resolves EXACT
    ties of the between-class variance differently from OpenCV's published recurrence (measured:
    ~15 % of deliberately symmetric histograms, 4 of 9476 7x5 noise tiles; never without a tie).
    The reference links an un-pinned distro OpenCV (CMakeLists.txt:17, .travis.yml:15: Ubuntu
    libopencv-dev, built without IPP), so the published C++ recurrence is the parity target.
    With IPP off cv2 agrees with the literal transcription below on every tie case tried.
    """

    def __enter__(self):
        self.prev = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(False)

    def __exit__(self, *a):
        cv2.ipp.setUseIPP(self.prev)


class _single_thread(_no_ipp):
    """... and one OpenCV thread.  Found on the GPU box (16 host threads, AVX-512 dispatch): cv2 4.13's fixed-point
    GaussianBlur returns a wrong, run-to-run varying row 1 for 5x5 kernels on images 7 rows high when its parallel_for
    splits the rows; with one thread it equals the exact 8.8 / 16.16 integer model on every input tried
    (scripts/debug_blur.py).  The reference's arithmetic is the single-thread result."""

    def __enter__(self):
        super().__enter__()
        self.threads = cv2.getNumThreads()
        cv2.setNumThreads(1)

    def __exit__(self, *a):
        cv2.setNumThreads(self.threads)
        super().__exit__(*a)


def otsu_threshold_cv(gray: np.ndarray) -> int:
    """The threshold cv::threshold(THRESH_OTSU) picks (real OpenCV C++ code path, IPP off)."""
    with _no_ipp():
        thr, _ = cv2.threshold(np.ascontiguousarray(gray), 128, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    return int(thr)


def otsu_threshold_from_hist(hist) -> int:
    """Literal transcription of OpenCV's getThreshVal_Otsu_8u recurrence.

    Pure Python floats == IEEE FP64, same operation order.  Used to pin the recurrence the
    CUDA kernel implements; checked against otsu_threshold_cv in tests/test_oracle.py.
    """
    FLT_EPSILON = 1.1920928955078125e-07
    n = int(sum(int(v) for v in hist))
    if n == 0:
        return 0
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * float(int(hist[i]))
    mu *= scale
    mu1 = 0.0
    q1 = 0.0
    max_sigma = 0.0
    max_val = 0
    for i in range(256):
        p_i = int(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < FLT_EPSILON or max(q1, q2) > 1.0 - FLT_EPSILON:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma = sigma
            max_val = i
    return max_val


def otsu_global(gray: np.ndarray, maxval: float = 255.0):
    """cv::threshold(src, dst, 128, maxval, THRESH_BINARY|THRESH_OTSU)  (deskew.cpp:224)."""
    with _no_ipp():
        thr, dst = cv2.threshold(np.ascontiguousarray(gray), 128, maxval, cv2.THRESH_BINARY | cv2.THRESH_OTSU)
    return int(thr), dst


def otsu_rects(gray: np.ndarray, rects, maxval: float = 255.0) -> np.ndarray:
    """Per-rectangle Otsu loop, binarizeLocalOtsu.cpp:138-162.  rects: iterable of (x, y, w, h)."""
    out = np.full(gray.shape, 255, np.uint8)
    with _no_ipp():
        for (x, y, w, h) in rects:
            tmp = gray[y:y + h, x:x + w].copy()
            _, tmp = cv2.threshold(tmp, 128, maxval, cv2.THRESH_OTSU)
            sel = (tmp ^ 255) != 0
            out[y:y + h, x:x + w][sel] = 0
    return out


def tile_rects(rows: int, cols: int, tile_w: int = 64, tile_h: int = 64):
    """Regular grid of rects (BASELINE config 4's "64x64 tiles"); edge tiles are clipped."""
    return [(x, y, min(tile_w, cols - x), min(tile_h, rows - y))
            for y in range(0, rows, tile_h) for x in range(0, cols, tile_w)]


def otsu_tiles(gray: np.ndarray, tile_w: int = 64, tile_h: int = 64, maxval: float = 255.0) -> np.ndarray:
    return otsu_rects(gray, tile_rects(gray.shape[0], gray.shape[1], tile_w, tile_h), maxval)


# --------------------------------------------------------------------------------------
# Edge front-end of prl::binarizeLocalOtsu: the real OpenCV calls, in the reference's order.
# --------------------------------------------------------------------------------------
def gaussian_blur(gray: np.ndarray, ksize: int, sigma: float = 0.0) -> np.ndarray:
    """cv::GaussianBlur(src, dst, Size(k, k), sigma)  (imageLibCommon.cpp:289)."""
    with _single_thread():
        return cv2.GaussianBlur(np.ascontiguousarray(gray), (ksize, ksize), sigma)


def canny(gray: np.ndarray, low: float, high: float) -> np.ndarray:
    """cv::Canny(src, dst, low, high): aperture 3, L1 gradient  (imageLibCommon.cpp:306)."""
    with _single_thread():
        return cv2.Canny(np.ascontiguousarray(gray), low, high)


def canny_edge_detection(gray: np.ndarray, ksize: int = 19, upper_coeff: float = 0.15, lower_coeff: float = 0.01,
                         morph_iters: int = 1) -> np.ndarray:
    """CannyEdgeDetection, imageLibCommon.cpp:244-324, for a single-channel image."""
    if gray.size == 0:
        raise ValueError("Image for histograms extraction is empty")
    if ksize < 3:
        raise ValueError("Gaussian blur kernel size is lesser than 3")
    if not (0 <= upper_coeff <= 1) or not (0 <= lower_coeff <= 1) or lower_coeff > upper_coeff:
        raise ValueError("Canny threshold coefficients")
    with _single_thread():
        channel = cv2.GaussianBlur(np.ascontiguousarray(gray), (ksize, ksize), 0)          # :289
        otsu, _ = cv2.threshold(channel, 0, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)      # :294-296
        upper = upper_coeff * otsu                                                          # :302
        lower = lower_coeff * upper                                                         # :303
        channel = cv2.Canny(channel, lower, upper)                                          # :306
        if morph_iters > 0:                                                                 # :309-313
            channel = cv2.dilate(channel, None, iterations=morph_iters)
            channel = cv2.erode(channel, None, iterations=morph_iters)
        if morph_iters < 0:                                                                 # :315-319
            channel = cv2.erode(channel, None, iterations=-morph_iters)
            channel = cv2.dilate(channel, None, iterations=-morph_iters)
    return np.zeros(gray.shape, np.uint8) | channel                                         # :322


def local_otsu_edges(gray: np.ndarray, ksize: int = 19, upper_coeff: float = 0.15, lower_coeff: float = 0.01,
                     morph_iters: int = 1, post_dilate: int = 3) -> np.ndarray:
    """... followed by cv::dilate(resultCanny, ..., 3)  (binarizeLocalOtsu.cpp:88-92)."""
    e = canny_edge_detection(gray, ksize, upper_coeff, lower_coeff, morph_iters)
    with _single_thread():
        return cv2.dilate(e, None, iterations=post_dilate) if post_dilate > 0 else e


def enhance_local_contrast_clahe(gray: np.ndarray, clip_limit: float, equalize: bool = True) -> np.ndarray:
    """EnhanceLocalContrastByCLAHE_1, imageLibCommon.cpp:326-346."""
    with _single_thread():
        clahe = cv2.createCLAHE()
        clahe.setClipLimit(clip_limit)
        dst = clahe.apply(np.ascontiguousarray(gray))
        return cv2.equalizeHist(dst) if equalize else dst


def contour_rects(edges: np.ndarray):
    """cv::findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + RemoveChildrenContours + cv::boundingRect
    (binarizeLocalOtsu.cpp:104-110,150; imageLibCommon.cpp:643-678,717-737: with a flat hierarchy every contour
    of at least 3 points is approved, shorter ones are dropped)."""
    contours, hierarchy = cv2.findContours(edges.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) == 0:
        raise ValueError("Contours array is empty")
    return [cv2.boundingRect(c) for c in contours if len(c) >= 3]


def binarizeLocalOtsu(image: np.ndarray, maxValue: float = 255.0, CLAHEClipLimit: float = 0.0, GaussianBlurKernelSize: int = 19,
                      CannyUpperThresholdCoeff: float = 0.15, CannyLowerThresholdCoeff: float = 0.01, CannyMorphIters: int = 1):
    """prl::binarizeLocalOtsu, binarizeLocalOtsu.cpp:38-163 (CLAHE branch not restated: the default is off)."""
    if image.size == 0:
        raise ValueError("Input image for binarization is empty")
    if not (0 <= maxValue <= 255):
        raise ValueError("Max value must be in range [0; 255]")
    gray = cv2.cvtColor(image, cv2.COLOR_RGB2GRAY) if image.ndim == 3 else image.copy()      # :60-68 (RGB2GRAY, as written)
    if CLAHEClipLimit > 0:
        gray = enhance_local_contrast_clahe(gray, CLAHEClipLimit, True)                       # :79-82
    edges = local_otsu_edges(gray, GaussianBlurKernelSize, CannyUpperThresholdCoeff, CannyLowerThresholdCoeff, CannyMorphIters, 3)
    return otsu_rects(gray, contour_rects(edges), maxValue)


def removeLines(image: np.ndarray) -> np.ndarray:
    """prl::removeLines, src/removeLines.cpp:30-77, call for call."""
    with _single_thread():
        gray = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 and image.shape[2] == 3 else image   # :32-40
        _, bw = cv2.threshold(cv2.bitwise_not(gray), 255, 255, cv2.THRESH_BINARY | cv2.THRESH_OTSU)             # :45
        horizontal, vertical = bw.copy(), bw.copy()
        hs = horizontal.shape[1] // 50                                                                           # :51
        he = cv2.getStructuringElement(cv2.MORPH_RECT, (hs, 1))                                                  # :54
        horizontal = cv2.dilate(cv2.erode(horizontal, he, anchor=(-1, -1)), he, anchor=(-1, -1))                 # :57-58
        vs = vertical.shape[0] // 50                                                                             # :61
        ve = cv2.getStructuringElement(cv2.MORPH_RECT, (1, vs))                                                  # :64
        vertical = cv2.dilate(cv2.erode(vertical, ve, anchor=(-1, -1)), ve, anchor=(-1, -1))                     # :67-68
        out = cv2.subtract(cv2.subtract(bw, horizontal), vertical)                                               # :70-72
        return cv2.bitwise_not(out)                                                                              # :74


# --------------------------------------------------------------------------------------
# first-principles restatement (numpy, no cv2 in the arithmetic) -- SURVEY Appendix A.
# Independent cross-check of the op-for-op path above; also what oracle/prl_oracle.c states.
# --------------------------------------------------------------------------------------
def threshold_map_f64_numpy(gray: np.ndarray, method: int, window: int, params):
    rows, cols = gray.shape
    w = clamp_window(window, rows, cols)
    h = w // 2
    kw = 1.0 / float(w * w)
    x0, y0, rw, rh = processing_rect(method, rows, cols, w)
    pad = np.pad(gray, h, mode="edge")
    S = np.cumsum(np.cumsum(pad.astype(np.int64), 0), 1).astype(np.float64)
    Q = np.cumsum(np.cumsum(pad.astype(np.int64) ** 2, 0), 1).astype(np.float64)
    d = w - 1

    def taps(A):
        a = A[0:rh, 0:rw]; b = A[0:rh, d:d + rw]; c = A[d:d + rh, 0:rw]; e = A[d:d + rh, d:d + rw]
        return kw * a + (-kw) * b + (-kw) * c + kw * e

    m = taps(S)
    q = taps(Q)
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.sqrt(q - m * m)
        if method == SAUVOLA:
            k = float(params[0]); T = m * (s * (k * (1.0 / 128.0)) + (1.0 - k))
        elif method == NIBLACK:
            k = float(params[0]); T = m + k * s
        elif method == WOLFJOLION:
            k = float(params[0]); imin = float(pad.min()); smax = float(np.nanmax(s)) if np.isfinite(s).any() else float("nan")
            coeff = np.float64(k) / np.float64(smax)
            T = m + (s * coeff + (-k)) * (m - imin)
        elif method == NICK:
            k = float(params[0]); T = m * 1.0 + np.sqrt(m * m + s * s) * k + 0.0
        else:
            alpha1, k1, k2, gamma = (float(v) for v in params)
            imin = float(pad.min())
            t1 = s / s
            t2 = np.power(t1, gamma)
            alpha3 = t2 * k2
            c2 = t2 * t1
            c3 = alpha3 * imin + c2 * (-imin) + 0.0
            T = (c2 + (1.0 - alpha1)) * m + c3
    return T, pad, (x0, y0, rw, rh)


def to_u8_numpy(T: np.ndarray) -> np.ndarray:
    with np.errstate(invalid="ignore"):
        r = np.rint(T)
        r = np.where(np.isfinite(T), r, 0.0)
        return np.clip(r, 0, 255).astype(np.uint8)


def binarize_local_numpy(gray: np.ndarray, method: int, window: int, params) -> np.ndarray:
    T, pad, (x, y, rw, rh) = threshold_map_f64_numpy(gray, method, window, params)
    T8 = to_u8_numpy(T)
    return np.where(pad[y:y + rh, x:x + rw] > T8, 255, 0).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# The adaptive-mean family: the reference's call sequences on cv2, op for op.
# cv2.error plays the role of cv::Exception.  (oracle/_ref runs the reference's own object code for the same functions.)
# ------------------------------------------------------------------------------------------------
def _empty_mat_adaptive():
    """what cv::adaptiveThreshold does with the empty Mat these functions hand it: CV_Assert(src.type() == CV_8UC1) ..."""
    raise cv2.error("adaptiveThreshold on an empty Mat (the reference never assigned it)")


def binarizeAT(image, medianKernelSize, maxValue, blockSize, shift):          # binarizeAT.cpp:33-67
    with _single_thread():
        if image is None or image.size == 0:
            raise ValueError("Input image for binarization is empty")
        tmp = cv2.medianBlur(image, medianKernelSize)
        if image.ndim == 2 or image.shape[2] == 1:
            _empty_mat_adaptive()
        g = cv2.cvtColor(tmp, cv2.COLOR_BGR2GRAY)
        return cv2.adaptiveThreshold(g, maxValue, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, blockSize, int(shift))


def binarizeAGT(image, medianKernelSize, maxValue, blockSize, shift):         # binarizeAGT.cpp:33-58
    with _single_thread():
        if image is None or image.size == 0:
            raise ValueError("Input image for binarization is empty")
        tmp = cv2.medianBlur(image, medianKernelSize)
        if image.ndim == 2 or image.shape[2] == 1:
            _empty_mat_adaptive()
        g = cv2.cvtColor(tmp, cv2.COLOR_BGR2GRAY)
        return cv2.adaptiveThreshold(g, maxValue, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, blockSize, int(shift))


def binarizePureAdaptiveGaussian(image, maxValue, blockSize, shift):          # binarizePureAdaptiveGaussian.cpp:31-71
    with _single_thread():
        if image is None or image.size == 0:
            raise ValueError("Input image for binarization is empty")
        if image.ndim == 2 or image.shape[2] == 1:
            _empty_mat_adaptive()
        g = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
        return cv2.adaptiveThreshold(g, maxValue, cv2.ADAPTIVE_THRESH_GAUSSIAN_C, cv2.THRESH_BINARY, blockSize, int(shift))


def binarizeNativeAdaptive(image, isGaussianBlurReqiured=False, medianBlurKernelSize=5, GaussianBlurKernelSize=7, GaussianBlurSigma=150.0,
                           isAdaptiveThresholdCalculatedByGaussian=True, adaptiveThresholdingMaxValue=255.0,
                           adaptiveThresholdingBlockSize=19, adaptiveThresholdingShift=9, bilateralFilterBlockSize=0,
                           bilateralFilterColorSigma=150.0, bilateralFilterSpaceSigma=150.0):     # binarizeNativeAdaptive.cpp:34-135
    with _single_thread():
        if image is None or image.size == 0:
            raise ValueError("Input image for binarization is empty")
        if not (0 <= adaptiveThresholdingMaxValue <= 255):
            raise ValueError("Max value must be in range [0; 255]")
        g = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY) if image.ndim == 3 and image.shape[2] > 1 else image.reshape(image.shape[:2])
        if not isGaussianBlurReqiured:
            if medianBlurKernelSize < 3:
                raise cv2.error("medianBlurKernelSize >= 3")
            out = cv2.medianBlur(g, medianBlurKernelSize)
        else:
            if GaussianBlurKernelSize < 3 or not GaussianBlurSigma > 0:
                raise cv2.error("GaussianBlurKernelSize >= 3, GaussianBlurSigma > 0")
            out = cv2.GaussianBlur(g, (GaussianBlurKernelSize, GaussianBlurKernelSize), GaussianBlurSigma)
        method = cv2.ADAPTIVE_THRESH_GAUSSIAN_C if isAdaptiveThresholdCalculatedByGaussian else cv2.ADAPTIVE_THRESH_MEAN_C
        bs = adaptiveThresholdingBlockSize
        if bs < 3:
            bs = int(np.sqrt(float(out.shape[0] * out.shape[0] + out.shape[1] * out.shape[1])) / 333 + 7)
        out = cv2.adaptiveThreshold(out, adaptiveThresholdingMaxValue, method, cv2.THRESH_BINARY_INV, bs, adaptiveThresholdingShift)
        if cv2.mean(out)[0] < 128:
            out = 255 - out
        if bilateralFilterBlockSize >= 3:                                                          # :116-134
            if bilateralFilterColorSigma <= 0:
                raise ValueError("Color sigma for bilateral filtration must be greater than 0")
            if bilateralFilterSpaceSigma <= 0:
                raise ValueError("Space sigma for bilateral filtration must be greater than 0")
            out = cv2.bilateralFilter(out, bilateralFilterBlockSize, bilateralFilterColorSigma, bilateralFilterSpaceSigma)   # IPP off, see _no_ipp
        return out


# ---- first-principles models of the two means cv::adaptiveThreshold uses (what csrc/adaptive.cu restates); checked
# against the real cv2 calls in tests/test_adaptive.py
def box_mean_model(gray: np.ndarray, bs: int) -> np.ndarray:
    """boxFilter(normalize, BORDER_REPLICATE) of a u8 image: round(sum / bs^2), never a tie for odd bs"""
    h = bs // 2
    p = np.pad(gray.astype(np.int64), h, mode="edge")
    I = np.zeros((p.shape[0] + 1, p.shape[1] + 1), np.int64)
    I[1:, 1:] = p.cumsum(0).cumsum(1)
    H, W = gray.shape
    s = I[bs:bs + H, bs:bs + W] - I[0:H, bs:bs + W] - I[bs:bs + H, 0:W] + I[0:H, 0:W]
    return ((2 * s + bs * bs) // (2 * bs * bs)).astype(np.uint8)


def _vexp32(x):
    """OpenCV's v_exp for float32 (the Cephes expf: two-piece ln 2 reduction, degree-5 polynomial) with every multiply-add
    unfused, as the baseline (SSE) build that fills bilateralFilter's colour table evaluates it.  x: float32 array."""
    f = np.float32
    madd = lambda a, b, c: (a * b).astype(f) + c
    x = np.minimum(np.maximum(x.astype(f), f(-88.3762626647949)), f(89.0))
    fx = madd(x, f(1.44269504088896341), f(0.5)).astype(f)
    mm = np.floor(fx)
    fx = mm.astype(f)
    x = madd(fx, f(-6.93359375E-1), x).astype(f)
    x = madd(fx, f(2.12194440E-4), x).astype(f)
    xx = (x * x).astype(f)
    y = madd(x, f(1.9875691500E-4), f(1.3981999507E-3)).astype(f)
    for c in (8.3334519073E-3, 4.1665795894E-2, 1.6666665459E-1, 5.0000001201E-1):
        y = madd(y, x, f(c)).astype(f)
    y = madd(y, xx, x).astype(f)
    y = (y + f(1)).astype(f)
    return (y * np.exp2(mm).astype(f)).astype(f)


def bilateral_tables(d: int, sigma_color: float, sigma_space: float):
    """(radius, colour weights[256], [(dy, dx, space weight)]) of cv::bilateralFilter for CV_8UC1 as OpenCV 4.13 builds them
    (measured against the cv2 wheel, IPP off; what csrc/adaptive.cu restates)."""
    if sigma_color <= 0:
        sigma_color = 1
    if sigma_space <= 0:
        sigma_space = 1
    gc = np.float32(-0.5 / (sigma_color * sigma_color))
    gs = np.float32(-0.5 / (sigma_space * sigma_space))
    radius = d // 2 if d > 0 else int(np.rint(sigma_space * 1.5))
    radius = max(radius, 1)
    i = np.arange(256)
    cw = _vexp32((i * i).astype(np.float32) * gc)
    taps = [(dy, dx, np.float32(np.exp(float(dy * dy + dx * dx) * float(gs))))
            for dy in range(-radius, radius + 1) for dx in range(-radius, radius + 1) if dy * dy + dx * dx <= radius * radius]
    return radius, cw, taps


def bilateral_model(gray: np.ndarray, d: int, sigma_color: float, sigma_space: float) -> np.ndarray:
    """cv::bilateralFilter(gray, d, sigmaColor, sigmaSpace), CV_8UC1, BORDER_REFLECT_101: taps in raster order,
    w = sw * cw[|v - centre|], wsum += w, sum = fma(v, w, sum) in float32, dst = cvRound(sum / wsum).  Identical to the wheel
    for radius != 2 and, for radius 2, on two-level images (all the reference feeds it); see csrc/adaptive.cu."""
    radius, cw, taps = bilateral_tables(d, sigma_color, sigma_space)
    t = cv2.copyMakeBorder(gray, radius, radius, radius, radius, cv2.BORDER_REFLECT_101)
    H, W = gray.shape
    c = gray.astype(np.int32)
    s = np.zeros((H, W), np.float32)
    ws = np.zeros((H, W), np.float32)
    for dy, dx, sw in taps:
        v = t[radius + dy:radius + dy + H, radius + dx:radius + dx + W].astype(np.int32)
        w = (sw * cw[np.abs(v - c)]).astype(np.float32)
        ws = (ws + w).astype(np.float32)
        s = _fma32(v.astype(np.float32), w, s)
    return np.rint((s / ws).astype(np.float32)).astype(np.uint8)


def _fma32(a, b, c):
    # exact fused multiply-add in float32: the 48-bit product and the sum are exact in 64-bit-mantissa long doubles
    return (a.astype(np.longdouble) * b.astype(np.longdouble) + c.astype(np.longdouble)).astype(np.float32)


def _mad32(a, b, c):
    return (c + (a * b).astype(np.float32)).astype(np.float32)


def gaussian_mean_model(gray: np.ndarray, bs: int, as_float: bool = False) -> np.ndarray:
    """GaussianBlur(float32(gray), (bs, bs), 0, BORDER_REPLICATE) -> u8 as the cv2 wheel's AVX2 build evaluates it:
    fused multiply-adds in the vector bodies (row pass: columns below cols & ~3, column pass: below cols & ~7), the
    scalar tails as gcc compiled them (csrc/adaptive.cu header)."""
    H, W = gray.shape
    f = gray.astype(np.float32)
    # cv::GaussianBlur: a one-pixel-wide / -high image is not filtered along that axis (ksize.width / .height := 1)
    k = cv2.getGaussianKernel(bs if W > 1 else 1, 0, cv2.CV_32F).ravel().astype(np.float32)
    bs_x = len(k)
    h = bs_x // 2
    p = np.pad(f, ((0, 0), (h, h)), mode="edge")
    nu = 4 * ((bs_x - 1) // 4)
    body = (p[:, 0:W] * k[0]).astype(np.float32)
    tail = body.copy()
    for i in range(1, bs_x):
        x = p[:, i:i + W]
        body = _fma32(x, np.float32(k[i]), body)
        tail = (_mad32 if i <= nu else _fma32)(x, np.float32(k[i]), tail)
    rowp = np.where(np.arange(W)[None, :] < (W & ~3), body, tail)
    k = cv2.getGaussianKernel(bs if H > 1 else 1, 0, cv2.CV_32F).ravel().astype(np.float32)
    h = len(k) // 2
    r = np.pad(rowp, ((h, h), (0, 0)), mode="edge")
    body = (r[h:h + H] * k[h]).astype(np.float32)
    tail = body.copy()
    for j in range(1, h + 1):
        sm = (r[h + j:h + j + H] + r[h - j:h - j + H]).astype(np.float32)
        body = _fma32(sm, np.float32(k[h + j]), body)
        tail = _mad32(sm, np.float32(k[h + j]), tail)
    out = np.where(np.arange(W)[None, :] < (W & ~7), body, tail)
    if as_float:
        return out
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)
