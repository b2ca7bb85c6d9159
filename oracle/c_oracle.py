"""ctypes loader for the plain-C oracle (oracle/prl_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libprl_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "prl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libprl_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i64p, i32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_double))
        L.oracle_synth_page.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
        L.oracle_synth_page.restype = None
        L.oracle_integral_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, i64p, i64p]
        L.oracle_integral_u8.restype = None
        L.oracle_output_shape.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_binarize_local.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, f64p, u8p, u8p, f64p]
        L.oracle_morph.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.oracle_morph.restype = None
        L.oracle_otsu_from_hist.argtypes = [i32p]
        L.oracle_otsu_threshold.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t]
        L.oracle_otsu_global.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_double, u8p, C.c_size_t]
        L.oracle_otsu_rects.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, i32p, C.c_int, C.c_double, u8p, C.c_size_t]
        L.oracle_otsu_rects.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def synth_page(page, rows=3508, cols=2480, seed=2024):
    out = np.empty((rows, cols), np.uint8)
    lib().oracle_synth_page(_p(out, C.c_uint8), cols, rows, cols, seed, page)
    return out


def integrals_int64(gray, pad):
    gray = np.ascontiguousarray(gray)
    r, c = gray.shape
    S = np.empty((r + 2 * pad, c + 2 * pad), np.int64)
    Q = np.empty_like(S)
    lib().oracle_integral_u8(_p(gray, C.c_uint8), r, c, c, pad, _p(S, C.c_int64), _p(Q, C.c_int64))
    return S, Q


def output_shape(method, rows, cols, window):
    orow, ocol = C.c_int(), C.c_int()
    lib().oracle_output_shape(method, rows, cols, window, C.byref(orow), C.byref(ocol))
    return orow.value, ocol.value


def binarize_local(gray, method, window, params, morph_iters=0, want_t8=False):
    gray = np.ascontiguousarray(gray)
    r, c = gray.shape
    orow, ocol = output_shape(method, r, c, window)
    if orow <= 0 or ocol <= 0:
        raise ValueError("empty processingRect")
    par = np.zeros(4, np.float64)
    par[:len(params)] = params
    t8 = np.empty((orow, ocol), np.uint8)
    mask = np.empty((orow, ocol), np.uint8)
    aux = np.zeros(2, np.float64)
    rc = lib().oracle_binarize_local(_p(gray, C.c_uint8), r, c, c, method, window, _p(par, C.c_double),
                                     _p(t8, C.c_uint8), _p(mask, C.c_uint8), _p(aux, C.c_double))
    if rc != 0:
        raise RuntimeError(f"oracle_binarize_local rc={rc}")
    if morph_iters:
        lib().oracle_morph(_p(mask, C.c_uint8), orow, ocol, morph_iters)
    return (mask, t8, aux) if want_t8 else mask


def otsu_from_hist(hist):
    h = np.ascontiguousarray(hist, dtype=np.int32)
    return lib().oracle_otsu_from_hist(_p(h, C.c_int32))


def otsu_global(gray, maxval=255.0):
    gray = np.ascontiguousarray(gray)
    r, c = gray.shape
    dst = np.empty_like(gray)
    thr = lib().oracle_otsu_global(_p(gray, C.c_uint8), r, c, c, maxval, _p(dst, C.c_uint8), c)
    return thr, dst


def otsu_rects(gray, rects, maxval=255.0):
    gray = np.ascontiguousarray(gray)
    r, c = gray.shape
    xywh = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4))
    dst = np.empty_like(gray)
    lib().oracle_otsu_rects(_p(gray, C.c_uint8), r, c, c, _p(xywh, C.c_int32), len(xywh), maxval, _p(dst, C.c_uint8), c)
    return dst
