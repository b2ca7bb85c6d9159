"""prlib_b200 -- B200 (sm_100a) implementation of PRLib's local-statistics + Otsu binarization path.

Product path: these host functions -> C-ABI libprlib_cuda.so (include/prlib_cuda.h) -> hand-written
CUDA kernels.  There is no CPU fallback: without the built library or without a Blackwell GPU every
call raises.
"""
from . import capi
from .capi import PrlCudaError, SAUVOLA, NIBLACK, WOLFJOLION, NICK, FENG
from .context import Context, default_context, binarize_batch, unpack_lept1, PinnedArray, set_global_option
from .binarize import (binarizeSauvola, binarizeNiblack, binarizeWolfJolion, binarizeNICK, binarizeFeng,
                       padded_gray, otsuThreshold, binarizeLocalOtsuRects, binarizeLocalOtsuTiles, binarizeLocalOtsu, removeLines,
                       binarizeAT, binarizeAGT, binarizeGAT, binarizePureAdaptive, binarizePureAdaptiveGaussian, binarizeNativeAdaptive)

__all__ = ["capi", "PrlCudaError", "Context", "default_context", "binarize_batch", "unpack_lept1", "PinnedArray", "set_global_option", "SAUVOLA", "NIBLACK",
           "WOLFJOLION", "NICK", "FENG", "binarizeSauvola", "binarizeNiblack", "binarizeWolfJolion",
           "binarizeNICK", "binarizeFeng", "padded_gray", "otsuThreshold", "binarizeLocalOtsuRects",
           "binarizeLocalOtsuTiles", "binarizeLocalOtsu", "removeLines", "binarizeAT", "binarizeAGT", "binarizeGAT",
           "binarizePureAdaptive", "binarizePureAdaptiveGaussian", "binarizeNativeAdaptive"]
