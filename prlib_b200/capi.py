"""ctypes binding of the C-ABI declared in include/prlib_cuda.h (no torch types cross it)."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

PRL_OK, PRL_E_INVALID, PRL_E_EMPTY_ROI, PRL_E_CUDA, PRL_E_NOMEM, PRL_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
SAUVOLA, NIBLACK, WOLFJOLION, NICK, FENG = 0, 1, 2, 3, 4

_u8p = C.POINTER(C.c_uint8)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_intp = C.POINTER(C.c_int)
_ctx = C.c_void_p

# name -> (restype, argtypes); every symbol include/prlib_cuda.h declares
SIGNATURES = {
    "prl_cuda_device_count": (C.c_int, []),
    "prl_cuda_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "prl_cuda_destroy": (None, [_ctx]),
    "prl_cuda_last_error": (C.c_char_p, [_ctx]),
    "prl_cuda_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "prl_cuda_synchronize": (C.c_int, [_ctx]),
    "prl_cuda_set_workspace_limit": (C.c_int, [_ctx, C.c_size_t]),
    "prl_cuda_set_option": (C.c_int, [_ctx, C.c_char_p, C.c_longlong]),
    "prl_cuda_output_shape": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _intp, _intp]),
    "prl_cuda_integral_u8": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]),
    "prl_cuda_binarize_local": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, _f64p, C.c_int,
                                          C.c_void_p, C.c_size_t, _intp, _intp]),
    "prl_cuda_binarize_local_image": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, _f64p, C.c_int,
                                                C.c_void_p, C.c_size_t, _intp, _intp, C.c_void_p, C.c_size_t]),
    "prl_cuda_threshold_map": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, _f64p,
                                         C.c_void_p, C.c_size_t, _intp, _intp, _f64p]),
    "prl_cuda_bgr2gray": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_morph": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_otsu_threshold": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, _intp]),
    "prl_cuda_otsu_global": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_void_p, C.c_size_t, _intp]),
    "prl_cuda_otsu_rects": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_double,
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "prl_cuda_otsu_tiles": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_double,
                                      C.c_void_p, C.c_size_t]),
    "prl_cuda_binarize_local_batch_dev": (C.c_int, [_ctx, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                                    C.c_int, _f64p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "prl_cuda_integral_u8_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]),
    "prl_cuda_otsu_global_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_double,
                                                 C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "prl_cuda_otsu_tiles_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                                C.c_double, C.c_void_p, C.c_size_t, C.c_size_t]),
    "prl_cuda_synth_pages_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32]),
    "prl_cuda_binarize_batch": (C.c_int, [_intp, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int,
                                          C.c_void_p]),
    "prl_cuda_binarize_batch_packed": (C.c_int, [_intp, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, C.c_int,
                                                 C.c_void_p]),
    "prl_cuda_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "prl_cuda_host_free": (C.c_int, [C.c_void_p]),
    "prl_cuda_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "prl_cuda_host_unregister": (C.c_int, [C.c_void_p]),
    "prl_cuda_set_global_option": (C.c_int, [C.c_char_p, C.c_longlong]),
    "prl_cuda_pack_mask_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p]),
    "prl_cuda_clahe": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_binarize_local_otsu": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                               C.c_int, C.c_void_p, C.c_size_t, _intp, C.c_void_p, C.c_int]),
    "prl_cuda_external_rects": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, _intp]),
    "prl_cuda_remove_lines": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_gauss_kernel_fixed": (C.c_int, [C.c_int, C.c_double, _intp]),
    "prl_cuda_gaussian_blur": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_void_p, C.c_size_t]),
    "prl_cuda_canny": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_size_t]),
    "prl_cuda_canny_edge_detection": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                                C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_canny_edge_detection_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                                    C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_timing_enable": (C.c_int, [_ctx, C.c_int]),
    "prl_cuda_timing_reset": (C.c_int, [_ctx]),
    "prl_cuda_timing_get": (C.c_int, [_ctx, C.c_char_p, _f64p, C.POINTER(C.c_longlong)]),
    "prl_cuda_launch_count": (C.c_longlong, [_ctx]),
    "prl_cuda_fused_redo_count": (C.c_longlong, [_ctx]),
}



class AdaptiveParams(C.Structure):
    """struct prl_adaptive_params (include/prlib_cuda.h)"""
    _fields_ = [("gray_first", C.c_int), ("blur", C.c_int), ("blur_ksize", C.c_int), ("blur_sigma", C.c_double),
                ("assert_ksize", C.c_int), ("method", C.c_int), ("type", C.c_int), ("maxval", C.c_double),
                ("check_maxval", C.c_int), ("block_size", C.c_int), ("auto_block", C.c_int), ("delta", C.c_double),
                ("invert_if_dark", C.c_int), ("bilateral_d", C.c_int), ("bilateral_sigma_color", C.c_double),
                ("bilateral_sigma_space", C.c_double)]


SIGNATURES.update({
    "prl_cuda_binarize_local_otsu_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_double, C.c_double,
                                                         C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t,
                                                         C.c_void_p, C.c_void_p]),
    "prl_cuda_remove_lines_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                                  C.c_size_t]),
    "prl_cuda_median_blur": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "prl_cuda_adaptive_threshold": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int,
                                              C.c_double, C.c_void_p, C.c_size_t]),
    "prl_cuda_gauss_kernel_float": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "prl_cuda_binarize_adaptive_batch_dev": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                                       C.POINTER(AdaptiveParams), C.c_void_p, C.c_size_t, C.c_size_t]),
    "prl_cuda_batch_unpack_threads": (C.c_int, []),
    "prl_cuda_unpack_mask_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "prl_cuda_bilateral_filter": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                            C.c_size_t]),
    "prl_cuda_binarize_adaptive": (C.c_int, [_ctx, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(AdaptiveParams),
                                             C.c_void_p, C.c_size_t]),
})

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """dlopen libprlib_cuda.so (building it first if the sources are newer and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m prlib_b200._build` (there is no CPU fallback)")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)   # AttributeError if the library does not export it
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


class PrlCudaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libprlib_cuda error {code}: {msg}")
        self.code = code
