"""Host-side handle on a libprlib_cuda context (one device, one stream, scratch planes)."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import capi
from .capi import PrlCudaError

_FAMILIES = ("integral", "threshold", "smax", "morph", "otsu_hist", "otsu_search", "otsu_apply", "otsu_tiles",
             "synth", "bgr2gray", "band_carry", "fused", "fused_pre", "fused_fix", "pack", "edges", "lines")


def _params4(params) -> "C.Array":
    arr = (C.c_double * 4)(0.0, 0.0, 0.0, 0.0)
    for i, v in enumerate(np.atleast_1d(np.asarray(params, dtype=np.float64)).tolist()[:4]):
        arr[i] = v
    return arr


def _as_u8_2d(a, name="image") -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name} must be uint8")
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    if a.ndim != 2:
        raise ValueError(f"{name} must be a single-channel 2-D array")
    if a.strides[1] != 1 or a.strides[0] < a.shape[1]:
        a = np.ascontiguousarray(a)
    return a


class Context:
    """Owns one prl_cuda_ctx.  Not thread-safe: use one Context per host thread."""

    def __init__(self, device: int = 0):
        self._L = capi.load()
        h = C.c_void_p()
        rc = self._L.prl_cuda_create(int(device), C.byref(h))
        if rc != capi.PRL_OK:
            raise PrlCudaError(rc, (self._L.prl_cuda_last_error(None) or b"").decode())
        self._h = h
        self.device = int(device)

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.prl_cuda_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != capi.PRL_OK:
            msg = (self._L.prl_cuda_last_error(self._h) or b"").decode()
            if rc == capi.PRL_E_INVALID:
                raise ValueError(msg)          # the reference throws std::invalid_argument here
            raise PrlCudaError(rc, msg)

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        """Borrow a cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream); None gives the
        context its own stream back.  Handle 0 (torch's default stream) is passed as cudaStreamLegacy."""
        if cuda_stream is None:
            h = 0
        else:
            h = int(cuda_stream) or 1          # cudaStreamLegacy == (cudaStream_t)0x1
        self._check(self._L.prl_cuda_set_stream(self._h, C.c_void_p(h)))

    def synchronize(self):
        self._check(self._L.prl_cuda_synchronize(self._h))

    def set_workspace_limit(self, nbytes: int):
        self._check(self._L.prl_cuda_set_workspace_limit(self._h, int(nbytes)))

    def set_option(self, name: str, value: int):
        """Validation switches: "exact_threshold", "disable_tma" (results are identical, speed differs)."""
        self._check(self._L.prl_cuda_set_option(self._h, name.encode(), int(value)))

    def timing_enable(self, on: bool = True):
        self._check(self._L.prl_cuda_timing_enable(self._h, int(on)))

    def timing_reset(self):
        self._check(self._L.prl_cuda_timing_reset(self._h))

    def timing(self) -> dict:
        out = {}
        for fam in _FAMILIES:
            ms, n = C.c_double(), C.c_longlong()
            self._check(self._L.prl_cuda_timing_get(self._h, fam.encode(), C.byref(ms), C.byref(n)))
            if n.value:
                out[fam] = {"ms": ms.value, "launches": n.value}
        return out

    def launch_count(self) -> int:
        return int(self._L.prl_cuda_launch_count(self._h))

    def fused_redo_count(self) -> int:
        """Pages the fused small-window path handed back to the two-kernel path."""
        return int(self._L.prl_cuda_fused_redo_count(self._h))

    # -- host-pointer entry points -------------------------------------------------------------
    def integral(self, gray, pad: int):
        g = _as_u8_2d(gray)
        r, c = g.shape
        S = np.empty((r + 2 * pad, c + 2 * pad), np.int64)
        Q = np.empty_like(S)
        self._check(self._L.prl_cuda_integral_u8(self._h, g.ctypes.data, r, c, g.strides[0], int(pad),
                                                 S.ctypes.data, Q.ctypes.data))
        return S, Q

    def output_shape(self, method: int, rows: int, cols: int, window: int):
        orow, ocol = C.c_int(), C.c_int()
        rc = self._L.prl_cuda_output_shape(method, rows, cols, window, C.byref(orow), C.byref(ocol))
        return rc, orow.value, ocol.value

    def binarize_local(self, gray, method: int, window: int, params, morph_iters: int = 0):
        g = _as_u8_2d(gray)
        r, c = g.shape
        rc, orow, ocol = self.output_shape(method, r, c, window)
        out = np.empty((max(orow, 0), max(ocol, 0)), np.uint8)
        a, b = C.c_int(), C.c_int()
        self._check(self._L.prl_cuda_binarize_local(self._h, method, g.ctypes.data, r, c, g.strides[0], int(window),
                                                    _params4(params), int(morph_iters), out.ctypes.data,
                                                    max(ocol, 1), C.byref(a), C.byref(b)))
        return out

    def binarize_image(self, image, method: int, window: int, params, morph_iters: int = 0, want_gray: bool = False):
        """1-, 3- (BGR) or 4-channel (BGRA) image in, mask out; the cvtColor front step runs on the device."""
        im = np.asarray(image)
        if im.dtype != np.uint8:
            raise TypeError("image must be uint8")
        if im.ndim == 2:
            im = im[:, :, None]
        if im.ndim != 3 or im.shape[2] not in (1, 3, 4):
            raise ValueError("expected an HxW, HxWx1, HxWx3 or HxWx4 uint8 image")
        im = np.ascontiguousarray(im)
        r, c, ch = im.shape
        rc, orow, ocol = self.output_shape(method, r, c, window)
        out = np.empty((max(orow, 0), max(ocol, 0)), np.uint8)
        gray = np.empty((r, c), np.uint8) if want_gray else None
        a, b = C.c_int(), C.c_int()
        self._check(self._L.prl_cuda_binarize_local_image(self._h, method, im.ctypes.data, r, c, im.strides[0], ch, int(window),
                                                          _params4(params), int(morph_iters), out.ctypes.data, max(ocol, 1),
                                                          C.byref(a), C.byref(b), gray.ctypes.data if want_gray else None, c))
        return (out, gray) if want_gray else out

    def threshold_map(self, gray, method: int, window: int, params):
        g = _as_u8_2d(gray)
        r, c = g.shape
        rc, orow, ocol = self.output_shape(method, r, c, window)
        out = np.empty((max(orow, 0), max(ocol, 0)), np.uint8)
        a, b = C.c_int(), C.c_int()
        aux = (C.c_double * 2)()
        self._check(self._L.prl_cuda_threshold_map(self._h, method, g.ctypes.data, r, c, g.strides[0], int(window),
                                                   _params4(params), out.ctypes.data, max(ocol, 1), C.byref(a),
                                                   C.byref(b), aux))
        return out, {"imin": aux[0], "smax": aux[1]}

    def bgr2gray(self, image):
        im = np.asarray(image)
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] not in (3, 4):
            raise ValueError("expected an HxWx3 or HxWx4 uint8 image")
        im = np.ascontiguousarray(im)
        r, c, ch = im.shape
        out = np.empty((r, c), np.uint8)
        self._check(self._L.prl_cuda_bgr2gray(self._h, im.ctypes.data, r, c, im.strides[0], ch, out.ctypes.data, c))
        return out

    def morph(self, mask, iters: int):
        m = _as_u8_2d(mask, "mask")
        r, c = m.shape
        out = np.empty((r, c), np.uint8)
        self._check(self._L.prl_cuda_morph(self._h, m.ctypes.data, r, c, m.strides[0], int(iters), out.ctypes.data, c))
        return out

    def otsu_threshold(self, gray) -> int:
        g = _as_u8_2d(gray)
        t = C.c_int()
        self._check(self._L.prl_cuda_otsu_threshold(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0], C.byref(t)))
        return t.value

    def otsu_global(self, gray, maxval: float = 255.0):
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        t = C.c_int()
        self._check(self._L.prl_cuda_otsu_global(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0],
                                                 float(maxval), out.ctypes.data, g.shape[1], C.byref(t)))
        return t.value, out

    def otsu_rects(self, gray, rects, maxval: float = 255.0, return_thresholds: bool = False):
        g = _as_u8_2d(gray)
        xywh = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4))
        out = np.empty(g.shape, np.uint8)
        thr = np.zeros(max(len(xywh), 1), np.int32)
        self._check(self._L.prl_cuda_otsu_rects(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0],
                                                xywh.ctypes.data if len(xywh) else None, len(xywh), float(maxval),
                                                out.ctypes.data, g.shape[1], thr.ctypes.data))
        return (out, thr[:len(xywh)]) if return_thresholds else out

    def otsu_tiles(self, gray, tile_w: int = 64, tile_h: int = 64, maxval: float = 255.0):
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        self._check(self._L.prl_cuda_otsu_tiles(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0],
                                                int(tile_w), int(tile_h), float(maxval), out.ctypes.data, g.shape[1]))
        return out

    # -- edge front-end of prl::binarizeLocalOtsu ---------------------------
    def gaussian_blur(self, gray, ksize: int, sigma: float = 0.0):
        """cv::GaussianBlur on CV_8U (fixed point, BORDER_REFLECT_101)."""
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        self._check(self._L.prl_cuda_gaussian_blur(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0], int(ksize),
                                                   float(sigma), out.ctypes.data, g.shape[1]))
        return out

    def canny(self, gray, low: float, high: float):
        """cv::Canny, aperture 3, L1 gradient."""
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        self._check(self._L.prl_cuda_canny(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0], float(low), float(high),
                                           out.ctypes.data, g.shape[1]))
        return out

    def canny_edge_detection(self, gray, ksize: int = 19, upper_coeff: float = 0.15, lower_coeff: float = 0.01,
                             morph_iters: int = 1, post_dilate: int = 0):
        """CannyEdgeDetection (imageLibCommon.cpp:244-324) + `post_dilate` dilations (binarizeLocalOtsu.cpp:92)."""
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        self._check(self._L.prl_cuda_canny_edge_detection(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0], int(ksize),
                                                          float(upper_coeff), float(lower_coeff), int(morph_iters),
                                                          int(post_dilate), out.ctypes.data, g.shape[1]))
        return out

    def clahe(self, gray, clip_limit: float, equalize: bool = True):
        """EnhanceLocalContrastByCLAHE for one channel: cv::createCLAHE()->apply (+ cv::equalizeHist)."""
        g = _as_u8_2d(gray)
        out = np.empty(g.shape, np.uint8)
        self._check(self._L.prl_cuda_clahe(self._h, g.ctypes.data, g.shape[0], g.shape[1], g.strides[0], float(clip_limit), int(equalize),
                                           out.ctypes.data, g.shape[1]))
        return out

    def binarize_local_otsu(self, image, maxval: float = 255.0, ksize: int = 19, upper_coeff: float = 0.15, lower_coeff: float = 0.01,
                            morph_iters: int = 1, return_rects: bool = False, clahe_clip_limit: float = 0.0):
        """prl::binarizeLocalOtsu, all on the device (prl_cuda_binarize_local_otsu).  image: (H, W) gray or (H, W, 3|4)."""
        im = np.ascontiguousarray(image)
        if im.dtype != np.uint8 or im.ndim not in (2, 3):
            raise ValueError("image must be uint8, (H, W) or (H, W, C)")
        ch = 1 if im.ndim == 2 else im.shape[2]
        out = np.empty(im.shape[:2], np.uint8)
        n = C.c_int()
        cap = 65535 if return_rects else 0
        rects = np.zeros((max(cap, 1), 4), np.int32)
        self._check(self._L.prl_cuda_binarize_local_otsu(self._h, im.ctypes.data, im.shape[0], im.shape[1], im.strides[0], ch,
                                                         float(maxval), float(clahe_clip_limit), int(ksize), float(upper_coeff), float(lower_coeff),
                                                         int(morph_iters), out.ctypes.data, im.shape[1], C.byref(n),
                                                         rects.ctypes.data if cap else None, cap))
        return (out, rects[:n.value]) if return_rects else out

    def external_rects(self, mask, cap: int = 65535):
        """(x, y, w, h) of the top-level contours of a binary image: cv::findContours(RETR_EXTERNAL) + cv::boundingRect as a set."""
        m = _as_u8_2d(mask, "mask")
        rects = np.zeros((max(cap, 1), 4), np.int32)
        n = C.c_int()
        self._check(self._L.prl_cuda_external_rects(self._h, m.ctypes.data, m.shape[0], m.shape[1], m.strides[0], rects.ctypes.data,
                                                    int(cap), C.byref(n)))
        return rects[:min(n.value, cap)]

    def remove_lines(self, image):
        """prl::removeLines (removeLines.cpp:30-77).  image: (H, W) gray or (H, W, 3) BGR."""
        im = np.ascontiguousarray(image)
        if im.dtype != np.uint8 or im.ndim not in (2, 3):
            raise ValueError("image must be uint8, (H, W) or (H, W, 3)")
        ch = 1 if im.ndim == 2 else im.shape[2]
        out = np.empty(im.shape[:2], np.uint8)
        self._check(self._L.prl_cuda_remove_lines(self._h, im.ctypes.data, im.shape[0], im.shape[1], im.strides[0], ch,
                                                  out.ctypes.data, im.shape[1]))
        return out

    # ---- the adaptive-mean family
    def median_blur(self, image, ksize: int):
        im = np.ascontiguousarray(image)
        if im.dtype != np.uint8 or im.ndim not in (2, 3):
            raise TypeError("image must be a uint8 HxW or HxWxC array")
        r, c = im.shape[:2]
        ch = 1 if im.ndim == 2 else im.shape[2]
        out = np.empty_like(im)
        self._check(self._L.prl_cuda_median_blur(self._h, im.ctypes.data, r, c, im.strides[0], ch, int(ksize), out.ctypes.data, out.strides[0]))
        return out

    def adaptive_threshold(self, gray, maxval: float, method: int, thresh_type: int, block_size: int, delta: float):
        g = _as_u8_2d(gray)
        r, c = g.shape
        out = np.empty((r, c), np.uint8)
        self._check(self._L.prl_cuda_adaptive_threshold(self._h, g.ctypes.data, r, c, g.strides[0], float(maxval), int(method),
                                                        int(thresh_type), int(block_size), float(delta), out.ctypes.data, c))
        return out

    def bilateral_filter(self, gray, d: int, sigma_color: float, sigma_space: float):
        """cv::bilateralFilter(gray, d, sigmaColor, sigmaSpace) for CV_8UC1 (prl_cuda_bilateral_filter)."""
        g = _as_u8_2d(gray)
        r, c = g.shape
        out = np.empty((r, c), np.uint8)
        self._check(self._L.prl_cuda_bilateral_filter(self._h, g.ctypes.data, r, c, g.strides[0], int(d), float(sigma_color), float(sigma_space),
                                                      out.ctypes.data, c))
        return out

    def binarize_adaptive(self, image, **kw):
        """prl_cuda_binarize_adaptive; keyword arguments = the fields of struct prl_adaptive_params"""
        im = np.ascontiguousarray(image)
        if im.dtype != np.uint8 or im.ndim not in (2, 3):
            raise TypeError("image must be a uint8 HxW or HxWxC array")
        r, c = im.shape[:2]
        ch = 1 if im.ndim == 2 else im.shape[2]
        p = capi.AdaptiveParams(**kw)
        out = np.empty((r, c), np.uint8)
        self._check(self._L.prl_cuda_binarize_adaptive(self._h, im.ctypes.data, r, c, im.strides[0], ch, C.byref(p), out.ctypes.data, c))
        return out

    # -- device-pointer entry points (raw addresses: torch .data_ptr() or cudaMalloc) ----------
    def binarize_local_batch_dev(self, method, d_src, n_pages, rows, cols, src_step, src_page_stride, window, params,
                                 morph_iters, d_dst, dst_step, dst_page_stride):
        self._check(self._L.prl_cuda_binarize_local_batch_dev(self._h, method, d_src, n_pages, rows, cols, src_step,
                                                              src_page_stride, window, _params4(params), morph_iters,
                                                              d_dst, dst_step, dst_page_stride))

    def pack_mask_dev(self, d_mask, n_pages, rows, cols, step, page_stride, d_bits):
        """u8 masks in HBM -> 1 bit per pixel, Leptonica PIX layout (prl_cuda_pack_mask_dev)."""
        self._check(self._L.prl_cuda_pack_mask_dev(self._h, d_mask, n_pages, rows, cols, step, page_stride, d_bits))

    def integral_batch_dev(self, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, d_sum, d_sqsum,
                           plane_pitch, plane_page_stride):
        self._check(self._L.prl_cuda_integral_u8_batch_dev(self._h, d_src, n_pages, rows, cols, src_step, src_page_stride,
                                                           pad, d_sum, d_sqsum, plane_pitch, plane_page_stride))

    def otsu_global_batch_dev(self, d_src, n_pages, rows, cols, src_step, src_page_stride, maxval, d_dst, dst_step,
                              dst_page_stride, d_thr):
        self._check(self._L.prl_cuda_otsu_global_batch_dev(self._h, d_src, n_pages, rows, cols, src_step, src_page_stride,
                                                           float(maxval), d_dst, dst_step, dst_page_stride, d_thr))

    def otsu_tiles_batch_dev(self, d_src, n_pages, rows, cols, src_step, src_page_stride, tile_w, tile_h, maxval,
                             d_dst, dst_step, dst_page_stride):
        self._check(self._L.prl_cuda_otsu_tiles_batch_dev(self._h, d_src, n_pages, rows, cols, src_step, src_page_stride,
                                                          tile_w, tile_h, float(maxval), d_dst, dst_step, dst_page_stride))

    def binarize_local_otsu_batch_dev(self, d_gray, n_pages, rows, cols, step, page_stride, d_dst, dst_step, dst_page_stride,
                                      maxval=255.0, clahe_clip_limit=0.0, ksize=19, upper_coeff=0.15, lower_coeff=0.01, morph_iters=1):
        """-> (n_rects[n_pages], status[n_pages]) as numpy int32 arrays; synchronous"""
        n_rects = np.zeros(n_pages, np.int32)
        status = np.zeros(n_pages, np.int32)
        self._check(self._L.prl_cuda_binarize_local_otsu_batch_dev(self._h, d_gray, n_pages, rows, cols, step, page_stride, float(maxval),
                                                                   float(clahe_clip_limit), int(ksize), float(upper_coeff), float(lower_coeff),
                                                                   int(morph_iters), d_dst, dst_step, dst_page_stride, n_rects.ctypes.data,
                                                                   status.ctypes.data))
        return n_rects, status

    def binarize_adaptive_batch_dev(self, d_src, n_pages, rows, cols, src_step, src_page_stride, channels, d_dst, dst_step, dst_page_stride, **kw):
        """prl_cuda_binarize_adaptive_batch_dev: the adaptive family over pages resident in HBM; keyword arguments = the fields of
        struct prl_adaptive_params"""
        p = capi.AdaptiveParams(**kw)
        self._check(self._L.prl_cuda_binarize_adaptive_batch_dev(self._h, d_src, n_pages, rows, cols, src_step, src_page_stride, channels,
                                                                 C.byref(p), d_dst, dst_step, dst_page_stride))

    def remove_lines_batch_dev(self, d_gray, n_pages, rows, cols, step, page_stride, d_dst, dst_step, dst_page_stride):
        self._check(self._L.prl_cuda_remove_lines_batch_dev(self._h, d_gray, n_pages, rows, cols, step, page_stride, d_dst, dst_step, dst_page_stride))

    def synth_pages_dev(self, d_dst, n_pages, rows, cols, step, page_stride, seed=2024, first_page=0):
        self._check(self._L.prl_cuda_synth_pages_dev(self._h, d_dst, n_pages, rows, cols, step, page_stride, seed, first_page))


_tls = threading.local()


def default_context(device: int = 0) -> Context:
    """Per-thread, per-device lazily created context used by the prl-style free functions."""
    d = getattr(_tls, "ctxs", None)
    if d is None:
        d = _tls.ctxs = {}
    if device not in d:
        d[device] = Context(device)
    return d[device]


class PinnedArray:
    """A page-locked host buffer of the batch loader (prl_cuda_host_alloc) seen as a numpy array (`.array`).
    Keep the object alive while the array is in use; close() (or garbage collection) frees the range."""

    def __init__(self, shape, dtype=np.uint8):
        L = capi.load()
        self._n = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
        p = C.c_void_p()
        rc = L.prl_cuda_host_alloc(self._n, C.byref(p))
        if rc != capi.PRL_OK:
            raise PrlCudaError(rc, (L.prl_cuda_last_error(None) or b"").decode())
        self._ptr = p.value
        self.array = np.frombuffer((C.c_uint8 * self._n).from_address(self._ptr), dtype=dtype,
                                   count=int(np.prod(shape))).reshape(shape)

    def close(self):
        if getattr(self, "_ptr", None):
            self.array = None
            capi.load().prl_cuda_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_global_option(name: str, value: int):
    """prl_cuda_set_global_option: "batch_chunk_pages", "batch_stage_pageable", "batch_unpack_threads" (-1 = automatic), "batch_unpack_nt", "batch_unpack_lag"."""
    rc = capi.load().prl_cuda_set_global_option(name.encode(), int(value))
    if rc != capi.PRL_OK:
        raise ValueError(f"unknown global option {name!r}")


def unpack_lept1(bits: np.ndarray, cols: int) -> np.ndarray:
    """(..., rows, wpl) uint32 PIX words (pixel x at bit 31 - (x & 31) of word x >> 5, 1 = black) -> 0/255 u8 masks."""
    b = np.unpackbits(np.ascontiguousarray(bits).astype(">u4").view(np.uint8), axis=-1)[..., :cols]
    return ((1 - b) * 255).astype(np.uint8)


def binarize_batch(pages: np.ndarray, method: int, window: int, params, morph_iters: int = 0, devices=None,
                   out: np.ndarray | None = None, packed: bool = False) -> np.ndarray:
    """Host batch + page dispatcher (prl_cuda_binarize_batch): (N, rows, cols) u8 -> (N, out_rows, out_cols).
    packed=True (prl_cuda_binarize_batch_packed): (N, out_rows, wpl) uint32 words, 1 bit per pixel, PIX layout."""
    L = capi.load()
    pages = np.asarray(pages)
    if pages.dtype != np.uint8 or pages.ndim != 3 or not pages.flags.c_contiguous:
        raise ValueError("pages must be a C-contiguous (N, rows, cols) uint8 array")
    n, r, c = pages.shape
    orow, ocol = C.c_int(), C.c_int()
    rc = L.prl_cuda_output_shape(method, r, c, window, C.byref(orow), C.byref(ocol))
    if rc == capi.PRL_E_INVALID:
        raise ValueError("empty image or window not (>1 and odd)")
    if rc != capi.PRL_OK:
        raise PrlCudaError(rc, "empty processingRect: min(rows, cols) <= windowSize")
    shape, dt = ((n, orow.value, (ocol.value + 31) // 32), np.uint32) if packed else ((n, orow.value, ocol.value), np.uint8)
    if out is None:
        out = np.empty(shape, dt)
    elif out.shape != shape or out.dtype != dt or not out.flags.c_contiguous:
        raise ValueError("out has the wrong shape/dtype")
    devs = None
    nd = 0
    if devices is not None:
        nd = len(devices)
        devs = (C.c_int * nd)(*devices)
    fn = L.prl_cuda_binarize_batch_packed if packed else L.prl_cuda_binarize_batch
    rc = fn(devs, nd, method, pages.ctypes.data, n, r, c, int(window), _params4(params), int(morph_iters), out.ctypes.data)
    if rc != capi.PRL_OK:
        msg = (L.prl_cuda_last_error(None) or b"").decode()
        if rc == capi.PRL_E_INVALID:
            raise ValueError(msg)
        raise PrlCudaError(rc, msg)
    return out
