"""oracle/_ref: the reference's OWN C++ (binarize{Sauvola,Niblack,WolfJolion,NICK,Feng,LocalOtsu}.cpp, removeLines.cpp,
imageLibCommon.cpp), compiled UNMODIFIED from /root/reference/src against the cv:: facade of oracle/cvfacade/ whose
primitives are executed by the cv2 wheel's OpenCV.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU arm).

Build recipe: oracle/Makefile target `_ref/_prl_ref.so` (needs /root/reference, i.e. the build container; the built
file travels to the GPU box with the snapshot).  `available()` says whether the built module is there.
"""
from __future__ import annotations

import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "_prl_ref.so")
REFERENCE_SRC = "/root/reference/src"
_mod = None


def build(force: bool = False) -> str | None:
    """Compile the reference sources where they lie (only possible where /root/reference exists)."""
    if not os.path.isdir(REFERENCE_SRC):
        return _SO if os.path.exists(_SO) else None
    subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["_ref/_prl_ref.so"])
    return _SO


def available() -> bool:
    return os.path.exists(_SO)


def module():
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("oracle/_ref/_prl_ref.so is not built (make -C oracle _ref/_prl_ref.so, needs /root/reference)")
        import cv2
        from oracle.cvfacade import cvcalls
        spec = importlib.util.spec_from_file_location("_prl_ref", _SO)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        m.register(cvcalls, cv2.error)
        _mod = m
    return _mod


def _img(image):
    if image is None:
        return None
    a = np.ascontiguousarray(image)
    if a.size == 0:
        return None  # cv::Mat::empty()
    return a.copy()  # the reference overwrites its input Mat


def _k(name, image, window, k, morph, return_input=False):
    out, after = getattr(module(), name)(_img(image), int(window), float(k), int(morph))
    return (out, after) if return_input else out


# defaults = the reference headers (binarizeSauvola.h:43-47 etc.)
def binarizeSauvola(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2, return_input=False):
    return _k("binarizeSauvola", image, windowSize, thresholdCoefficient, morphIterationCount, return_input)


def binarizeNiblack(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2, return_input=False):
    return _k("binarizeNiblack", image, windowSize, thresholdCoefficient, morphIterationCount, return_input)


def binarizeWolfJolion(image, windowSize=101, thresholdCoefficient=0.01, morphIterationCount=2, return_input=False):
    return _k("binarizeWolfJolion", image, windowSize, thresholdCoefficient, morphIterationCount, return_input)


def binarizeNICK(image, windowSize=21, thresholdCoefficient=-0.01, morphIterationCount=0, return_input=False):
    return _k("binarizeNICK", image, windowSize, thresholdCoefficient, morphIterationCount, return_input)


def binarizeFeng(image, windowSize=21, thresholdCoefficient_alpha1=0.75, thresholdCoefficient_k1=0.2,
                 thresholdCoefficient_k2=0.03, thresholdCoefficient_gamma=2.0, morphIterationCount=2, return_input=False):
    out, after = module().binarizeFeng(_img(image), int(windowSize), float(thresholdCoefficient_alpha1),
                                       float(thresholdCoefficient_k1), float(thresholdCoefficient_k2),
                                       float(thresholdCoefficient_gamma), int(morphIterationCount))
    return (out, after) if return_input else out


def binarizeLocalOtsu(image, maxValue=255.0, CLAHEClipLimit=0.0, GaussianBlurKernelSize=19, CannyUpperThresholdCoeff=0.15,
                      CannyLowerThresholdCoeff=0.01, CannyMorphIters=1):
    out, _ = module().binarizeLocalOtsu(_img(image), float(maxValue), float(CLAHEClipLimit), int(GaussianBlurKernelSize),
                                        float(CannyUpperThresholdCoeff), float(CannyLowerThresholdCoeff), int(CannyMorphIters))
    return out


def removeLines(image):
    out, _ = module().removeLines(_img(image))
    return out


# ---- the adaptive-mean family; no defaults in the reference headers except NativeAdaptive's
def binarizeAT(image, medianKernelSize, maxValue, blockSize, shift):
    return module().binarizeAT(_img(image), int(medianKernelSize), float(maxValue), int(blockSize), int(shift))[0]


def binarizeAGT(image, medianKernelSize, maxValue, blockSize, shift):
    return module().binarizeAGT(_img(image), int(medianKernelSize), float(maxValue), int(blockSize), int(shift))[0]


def binarizeGAT(image, gaussianKernelSize, sigmaX, sigmaY, maxValue, blockSize, shift):
    return module().binarizeGAT(_img(image), int(gaussianKernelSize), float(sigmaX), float(sigmaY), float(maxValue), int(blockSize), int(shift))[0]


def binarizePureAdaptive(image, maxValue, blockSize, shift):
    return module().binarizePureAdaptive(_img(image), float(maxValue), int(blockSize), int(shift))[0]


def binarizePureAdaptiveGaussian(image, maxValue, blockSize, shift):
    return module().binarizePureAdaptiveGaussian(_img(image), float(maxValue), int(blockSize), int(shift))[0]


def binarizeNativeAdaptive(image, isGaussianBlurReqiured=False, medianBlurKernelSize=5, GaussianBlurKernelSize=7, GaussianBlurSigma=150.0,
                           isAdaptiveThresholdCalculatedByGaussian=True, adaptiveThresholdingMaxValue=255.0,
                           adaptiveThresholdingBlockSize=19, adaptiveThresholdingShift=9, bilateralFilterBlockSize=0,
                           bilateralFilterColorSigma=150.0, bilateralFilterSpaceSigma=150.0, return_input=False):
    out, after = module().binarizeNativeAdaptive(_img(image), int(bool(isGaussianBlurReqiured)), int(medianBlurKernelSize),
                                                 int(GaussianBlurKernelSize), float(GaussianBlurSigma),
                                                 int(bool(isAdaptiveThresholdCalculatedByGaussian)), float(adaptiveThresholdingMaxValue),
                                                 int(adaptiveThresholdingBlockSize), float(adaptiveThresholdingShift),
                                                 int(bilateralFilterBlockSize), float(bilateralFilterColorSigma), float(bilateralFilterSpaceSigma))
    return (out, after) if return_input else out


_BY_METHOD = {0: "binarizeSauvola", 1: "binarizeNiblack", 2: "binarizeWolfJolion", 3: "binarizeNICK", 4: "binarizeFeng"}


def binarize_local(image, method, window, params, morph_iters=0):
    """Same calling convention as prl_oracle.binarize_local / c_oracle.binarize_local."""
    params = tuple(float(v) for v in np.atleast_1d(params))
    if method == 4:
        return binarizeFeng(image, window, *params, morphIterationCount=morph_iters)
    return _k(_BY_METHOD[method], image, window, params[0], morph_iters)
