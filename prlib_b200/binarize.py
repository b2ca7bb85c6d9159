"""prl::binarize* mirrored on the host side: same names, parameters, defaults and error behaviour
as the reference headers, bodies replaced by calls into libprlib_cuda (no CPU fallback).

  prl::binarizeSauvola     src/binarizations/binarizeSauvola.h:43-47
  prl::binarizeNiblack     src/binarizations/binarizeNiblack.h:43-47
  prl::binarizeWolfJolion  src/binarizations/binarizeWolfJolion.h:43-47
  prl::binarizeNICK        src/binarizations/binarizeNICK.h:43-47
  prl::binarizeFeng        src/binarizations/binarizeFeng.h:46-53
  prl::binarizeLocalOtsu   src/binarizations/binarizeLocalOtsu.h:50-57  (per-rectangle core only: the
                           Canny/contour front-end stays with the caller)

cv::Mat in / cv::Mat out becomes numpy in / numpy out.  std::invalid_argument -> ValueError;
cv::Exception (empty processingRect) -> PrlCudaError(PRL_E_EMPTY_ROI).  The C++ reference also
overwrites its input Mat with the padded gray image; `padded_gray()` returns that side effect for
callers that relied on it (the C++ shim in prlib_b200/shim reproduces it in place).
"""
from __future__ import annotations

import numpy as np

from . import capi
from .context import Context, default_context


def _gray(image, ctx: Context) -> np.ndarray:
    im = np.asarray(image)
    if im.size == 0:
        raise ValueError("Input image for binarization is empty")       # binarizeSauvola.cpp:38-41
    if im.ndim == 3 and im.shape[2] != 1:
        return ctx.bgr2gray(im)                                           # cvtColor(BGR2GRAY) :49-52
    return im


def _check_window(window: int):
    if not (window > 1 and window % 2 == 1):                             # :43-47
        raise ValueError("Window size must satisfy the following condition: "
                         "( (windowSize > 1) && ((windowSize % 2) == 1) ) ")


def _local(image, method, window, params, morph, device):
    im = np.asarray(image)
    if im.size == 0:
        raise ValueError("Input image for binarization is empty")
    _check_window(window)                      # argument errors come first, as in the reference
    ctx = default_context(device)
    return ctx.binarize_image(im, method, window, params, morph)     # cvtColor (if needed) happens on the device


def binarizeSauvola(imageInput, windowSize: int = 101, thresholdCoefficient: float = 0.01,
                    morphIterationCount: int = 2, device: int = 0) -> np.ndarray:
    return _local(imageInput, capi.SAUVOLA, windowSize, (thresholdCoefficient,), morphIterationCount, device)


def binarizeNiblack(imageInput, windowSize: int = 101, thresholdCoefficient: float = 0.01,
                    morphIterationCount: int = 2, device: int = 0) -> np.ndarray:
    return _local(imageInput, capi.NIBLACK, windowSize, (thresholdCoefficient,), morphIterationCount, device)


def binarizeWolfJolion(imageInput, windowSize: int = 101, thresholdCoefficient: float = 0.01,
                       morphIterationCount: int = 2, device: int = 0) -> np.ndarray:
    return _local(imageInput, capi.WOLFJOLION, windowSize, (thresholdCoefficient,), morphIterationCount, device)


def binarizeNICK(imageInput, windowSize: int = 21, thresholdCoefficient: float = -0.01,
                 morphIterationCount: int = 0, device: int = 0) -> np.ndarray:
    return _local(imageInput, capi.NICK, windowSize, (thresholdCoefficient,), morphIterationCount, device)


def binarizeFeng(imageInput, windowSize: int = 21, thresholdCoefficient_alpha1: float = 0.75,
                 thresholdCoefficient_k1: float = 0.2, thresholdCoefficient_k2: float = 0.03,
                 thresholdCoefficient_gamma: float = 2.0, morphIterationCount: int = 2, device: int = 0) -> np.ndarray:
    return _local(imageInput, capi.FENG, windowSize,
                  (thresholdCoefficient_alpha1, thresholdCoefficient_k1, thresholdCoefficient_k2,
                   thresholdCoefficient_gamma), morphIterationCount, device)


def padded_gray(imageInput, windowSize: int, device: int = 0) -> np.ndarray:
    """What the reference leaves in `imageInput` after a local-statistics call: the gray image
    replicate-padded by w/2 (binarizeSauvola.cpp:49-52, :65)."""
    ctx = default_context(device)
    g = _gray(imageInput, ctx)
    w = min(windowSize, min(g.shape[0], g.shape[1]))
    return np.pad(g, w // 2, mode="edge")


def otsuThreshold(image, maxValue: float = 255.0, device: int = 0):
    """cv::threshold(src, dst, 128, maxValue, THRESH_BINARY|THRESH_OTSU) -> (thr, dst)
    (src/deskew/deskew.cpp:224, src/removeLines.cpp:45, src/imageLibCommon.cpp:295-296)."""
    ctx = default_context(device)
    return ctx.otsu_global(_gray(image, ctx), maxValue)


def binarizeLocalOtsuRects(image, rects, maxValue: float = 255.0, device: int = 0) -> np.ndarray:
    """The per-contour-rectangle loop of prl::binarizeLocalOtsu (binarizeLocalOtsu.cpp:138-162);
    `rects` are the cv::boundingRect(x, y, w, h) of the caller's contours."""
    if not (0 <= maxValue <= 255):                                        # binarizeLocalOtsu.cpp:52-55
        raise ValueError("Max value must be in range [0; 255]")
    ctx = default_context(device)
    return ctx.otsu_rects(_gray(image, ctx), rects, maxValue)


def binarizeLocalOtsu(image, maxValue: float = 255.0, CLAHEClipLimit: float = 0.0, GaussianBlurKernelSize: int = 19,
                      CannyUpperThresholdCoeff: float = 0.15, CannyLowerThresholdCoeff: float = 0.01, CannyMorphIters: int = 1,
                      device: int = 0) -> np.ndarray:
    """prl::binarizeLocalOtsu (binarizeLocalOtsu.h:50-57, same defaults), all on the device: blur, Otsu value, Canny,
    closing, dilation, bounding rectangles of the top-level contours and the per-rectangle Otsu loop
    (prl_cuda_binarize_local_otsu)."""
    image = np.asarray(image)
    if image.size == 0:
        raise ValueError("Input image for binarization is empty")                   # binarizeLocalOtsu.cpp:47-50
    if not (0 <= maxValue <= 255):
        raise ValueError("Max value must be in range [0; 255]")                     # :52-55
    if GaussianBlurKernelSize < 3:
        raise ValueError("Gaussian blur kernel size is lesser than 3")              # imageLibCommon.cpp:253-256
    ctx = default_context(device)
    # (3/4-channel input is converted with COLOR_RGB2GRAY on the device, as binarizeLocalOtsu.cpp:63 does)
    return ctx.binarize_local_otsu(image, maxValue, GaussianBlurKernelSize, CannyUpperThresholdCoeff, CannyLowerThresholdCoeff,
                                   CannyMorphIters, clahe_clip_limit=max(CLAHEClipLimit, 0.0))


def removeLines(image, device: int = 0) -> np.ndarray:
    """prl::removeLines (src/removeLines.h, removeLines.cpp:30-77)."""
    return default_context(device).remove_lines(np.asarray(image))


def binarizeLocalOtsuTiles(image, tileWidth: int = 64, tileHeight: int = 64, maxValue: float = 255.0,
                           device: int = 0) -> np.ndarray:
    """Same loop with rects := the regular tile grid (BASELINE.json config 4)."""
    if not (0 <= maxValue <= 255):
        raise ValueError("Max value must be in range [0; 255]")
    ctx = default_context(device)
    return ctx.otsu_tiles(_gray(image, ctx), tileWidth, tileHeight, maxValue)


# ---- the adaptive-mean family -----------------------------------------------------------
# The reference's own quirks are part of the interface (tests/test_adaptive.py runs the reference's own object code and shows the same):
#   * binarizeAT / binarizeAGT / binarizePureAdaptiveGaussian only assign the image adaptiveThreshold reads inside
#     `if (channels != 1)`: a 1-channel input ends in the cv::Exception of cv::adaptiveThreshold on an empty Mat
#     (binarizeAT.cpp:53-65, binarizeAGT.cpp:46-58, binarizePureAdaptiveGaussian.cpp:47-69);
#   * binarizeGAT and binarizePureAdaptive convert to gray first, so that branch is never taken: they raise for EVERY
#     non-empty input (binarizeGAT.cpp:37-64, binarizePureAdaptive.cpp:38-60).
_CV_EMPTY = "adaptiveThreshold: src.type() == CV_8UC1 fails on the empty Mat the reference passes (cv::Exception)"


def _nonempty(image):
    im = np.asarray(image)
    if im.size == 0:
        raise ValueError("Input image for binarization is empty")
    return im


def _colour_only(im):
    if im.ndim == 2 or im.shape[2] == 1:
        raise capi.PrlCudaError(capi.PRL_E_EMPTY_ROI, _CV_EMPTY)
    return im


def binarizeAT(inputImage, medianKernelSize: int, maxValue: float, blockSize: int, shift: int, device: int = 0) -> np.ndarray:
    """prl::binarizeAT (binarizeAT.h:33-34): medianBlur on the colour image -> BGR2GRAY -> adaptiveThreshold(MEAN_C, BINARY)."""
    im = _colour_only(_nonempty(inputImage))
    return default_context(device).binarize_adaptive(im, gray_first=0, blur=1, blur_ksize=int(medianKernelSize), method=0, type=0,
                                                     maxval=float(maxValue), block_size=int(blockSize), delta=float(int(shift)))


def binarizeAGT(inputImage, medianKernelSize: int, maxValue: float, blockSize: int, shift: int, device: int = 0) -> np.ndarray:
    """prl::binarizeAGT (binarizeAGT.h:32-33): as binarizeAT with ADAPTIVE_THRESH_GAUSSIAN_C."""
    im = _colour_only(_nonempty(inputImage))
    return default_context(device).binarize_adaptive(im, gray_first=0, blur=1, blur_ksize=int(medianKernelSize), method=1, type=0,
                                                     maxval=float(maxValue), block_size=int(blockSize), delta=float(int(shift)))


def binarizePureAdaptiveGaussian(inputImage, maxValue: float, blockSize: int, shift: int, device: int = 0) -> np.ndarray:
    """prl::binarizePureAdaptiveGaussian (binarizePureAdaptiveGaussian.h:33-34): BGR2GRAY -> adaptiveThreshold(GAUSSIAN_C, BINARY)."""
    im = _colour_only(_nonempty(inputImage))
    return default_context(device).binarize_adaptive(im, gray_first=1, blur=0, method=1, type=0, maxval=float(maxValue),
                                                     block_size=int(blockSize), delta=float(int(shift)))


def binarizeGAT(inputImage, gaussianKernelSize: int, sigmaX: float, sigmaY: float, maxValue: float, blockSize: int, shift: int,
                device: int = 0):
    """prl::binarizeGAT (binarizeGAT.h:33-35) raises for every non-empty input (see above)."""
    _nonempty(inputImage)
    raise capi.PrlCudaError(capi.PRL_E_EMPTY_ROI, _CV_EMPTY)


def binarizePureAdaptive(inputImage, maxValue: float, blockSize: int, shift: int, device: int = 0):
    """prl::binarizePureAdaptive (binarizePureAdaptive.h:33-34) raises for every non-empty input (see above)."""
    _nonempty(inputImage)
    raise capi.PrlCudaError(capi.PRL_E_EMPTY_ROI, _CV_EMPTY)


def binarizeNativeAdaptive(inputImage, isGaussianBlurReqiured: bool = False, medianBlurKernelSize: int = 5,
                           GaussianBlurKernelSize: int = 7, GaussianBlurSigma: float = 150.0,
                           isAdaptiveThresholdCalculatedByGaussian: bool = True, adaptiveThresholdingMaxValue: float = 255.0,
                           adaptiveThresholdingBlockSize: int = 19, adaptiveThresholdingShift: float = 9,
                           bilateralFilterBlockSize: int = 0, bilateralFilterColorSigma: float = 150.0,
                           bilateralFilterSpaceSigma: float = 150.0, device: int = 0) -> np.ndarray:
    """prl::binarizeNativeAdaptive (binarizeNativeAdaptive.h:63-75, same defaults): gray -> median or Gaussian blur ->
    adaptiveThreshold(BINARY_INV) -> 255 - image when its mean is below 128 -> cv::bilateralFilter of the result when
    bilateralFilterBlockSize >= 3 (off by default; its sigma checks come last, :116-127, so a cv::Exception of the steps before
    wins, as in the reference)."""
    im = _nonempty(inputImage)
    if not (0 <= adaptiveThresholdingMaxValue <= 255):
        raise ValueError("Max value must be in range [0; 255]")                                    # :53-56
    return default_context(device).binarize_adaptive(
        im, gray_first=1, blur=2 if isGaussianBlurReqiured else 1,
        blur_ksize=int(GaussianBlurKernelSize if isGaussianBlurReqiured else medianBlurKernelSize), blur_sigma=float(GaussianBlurSigma),
        assert_ksize=1, method=1 if isAdaptiveThresholdCalculatedByGaussian else 0, type=1, maxval=float(adaptiveThresholdingMaxValue),
        check_maxval=1, block_size=int(adaptiveThresholdingBlockSize), auto_block=1, delta=float(adaptiveThresholdingShift), invert_if_dark=1,
        bilateral_d=int(bilateralFilterBlockSize) if bilateralFilterBlockSize >= 3 else 0,
        bilateral_sigma_color=float(bilateralFilterColorSigma), bilateral_sigma_space=float(bilateralFilterSpaceSigma))
