"""Forwarding layer of the cv:: facade (oracle/cvfacade/): one function per OpenCV primitive the reference calls.

TEST INFRASTRUCTURE.  The C++ facade (`opencv2/core/core.hpp` + `facade.cpp`) gives the reference's UNMODIFIED
sources a `cv::Mat` that wraps a numpy array; every `cv::` free function and every `Mat` method that does arithmetic
lands here and is executed by the real OpenCV inside the cv2 wheel (4.13.0) -- nothing in this file computes pixels
itself.  Three spellings need a word:

* filter2D on a sub-matrix: the reference filters `integralImage(processingRect)` (binarizeSauvola.cpp:89).  OpenCV 4.x
  looks at the ROI offset of the source (it reads the taps from the parent and skips the DFT path); a numpy slice loses
  that offset, a `cv2.UMat(cv2.UMat(parent), rowRange, colRange)` keeps it.
* Mat::convertTo(rtype, alpha, beta) has no direct binding; `cv2.gapi.convertTo` on the default (OpenCV CPU) backend is
  a one-line wrapper around it (GCPUConvertTo::run -> in.convertTo(out, rtype, alpha, beta)).  Measured here: on this
  AVX2/AVX-512 build it evaluates fma(x, alpha, beta), NOT round(round(x*alpha) + beta) -- tests/test_ref.py pins that.
* THRESH_OTSU: IPP is switched off so that OpenCV's own published getThreshVal_Otsu_8u runs (the wheel's closed IPP
  routine resolves exact ties differently; the reference links a distro OpenCV without IPP, .travis.yml:15).
"""
import numpy as np
import cv2

cv2.setNumThreads(1)
cv2.ocl.setUseOpenCL(False)
cv2.ipp.setUseIPP(False)

_DEPTH = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}


def _dtype(cvtype):
    return _DEPTH[cvtype & 7], (cvtype >> 3) + 1


def _shape(rows, cols, cn):
    return (rows, cols) if cn == 1 else (rows, cols, cn)


def empty(rows, cols, cvtype):
    dt, cn = _dtype(cvtype)
    return np.empty(_shape(rows, cols, cn), dt)


def zeros(rows, cols, cvtype):
    dt, cn = _dtype(cvtype)
    return np.zeros(_shape(rows, cols, cn), dt)


def ones(rows, cols, cvtype):
    dt, cn = _dtype(cvtype)
    a = np.zeros(_shape(rows, cols, cn), dt)
    a[..., 0] = 1  # Mat::ones sets the first channel only (Scalar(1))
    return a


def roi(a, y0, y1, x0, x1):
    return a[y0:y1, x0:x1]


def clone(a):
    return np.array(a, copy=True, order="C")


def copy_into(dst, src):
    np.copyto(dst, src)


def same_layout(dst, src):
    return dst.shape == src.shape and dst.dtype == src.dtype


def set_to(a, scalar, mask):
    # Mat::setTo(Scalar, mask): every channel c of the selected pixels := scalar[c]
    v = scalar[0] if a.ndim == 2 else np.asarray(scalar[:a.shape[2]])
    if mask is None:
        a[...] = np.asarray(v).astype(a.dtype) if a.ndim == 3 else a.dtype.type(_saturate(v, a.dtype))
    else:
        a[mask != 0] = np.asarray(v).astype(a.dtype) if a.ndim == 3 else a.dtype.type(_saturate(v, a.dtype))


def _saturate(v, dt):
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        return int(min(max(round(v), info.min), info.max))
    return v


def _src(root, oy, ox, rows, cols):
    """The Mat as OpenCV sees it: a whole array, or a sub-matrix that remembers its parent (UMat ROI)."""
    if oy == 0 and ox == 0 and root.shape[0] == rows and root.shape[1] == cols:
        return root
    return cv2.UMat(cv2.UMat(np.ascontiguousarray(root)), [oy, oy + rows], [ox, ox + cols])


def _get(r):
    return r.get() if isinstance(r, cv2.UMat) else r


def filter2D(root, oy, ox, rows, cols, ddepth, kernel, ax, ay, delta, border):
    return _get(cv2.filter2D(_src(root, oy, ox, rows, cols), ddepth, kernel, anchor=(ax, ay), delta=delta,
                             borderType=border))


def convert_to(a, rtype, alpha, beta):
    depth = rtype & 7 if rtype >= 0 else {np.dtype(v): k for k, v in _DEPTH.items()}[a.dtype]
    g = cv2.GMat()
    comp = cv2.GComputation(cv2.GIn(g), cv2.GOut(cv2.gapi.convertTo(g, depth, alpha, beta)))
    return comp.apply(cv2.gin(np.ascontiguousarray(a)))


def cvtColor(a, code):
    return cv2.cvtColor(a, code)


def copyMakeBorder(a, t, b, l, r, border):
    return cv2.copyMakeBorder(a, t, b, l, r, border)


def integral2(a, sdepth, sqdepth):
    return cv2.integral2(a, sdepth=sdepth, sqdepth=sqdepth)


def sqrt(a):
    return cv2.sqrt(a)


def pow_(a, p):
    return cv2.pow(a, p)


def minMaxLoc(a):
    return cv2.minMaxLoc(a)


def add(a, b):
    return cv2.add(a, b)


def subtract(a, b):
    return cv2.subtract(a, b)


def add_scalar(a, s):
    return cv2.add(a, s)


def rsub_scalar(s, a):
    return cv2.subtract(s, a)


def multiply(a, b, scale):
    return cv2.multiply(a, b, scale=scale)


def divide(a, b, scale):
    return cv2.divide(a, b, scale=scale)


def scaleAdd(a, alpha, b):
    return cv2.scaleAdd(a, alpha, b)


def addWeighted(a, alpha, b, beta, gamma):
    return cv2.addWeighted(a, alpha, b, beta, gamma)


def compare(a, b, op):
    return cv2.compare(np.ascontiguousarray(a), np.ascontiguousarray(b), op)


def bitwise_xor_scalar(a, s):
    return cv2.bitwise_xor(a, s)


def bitwise_or(a, b):
    return cv2.bitwise_or(a, b)


def bitwise_not(a):
    return cv2.bitwise_not(a)


def dilate(a, kernel, ax, ay, iterations):
    return cv2.dilate(a, kernel, anchor=(ax, ay), iterations=iterations)


def erode(a, kernel, ax, ay, iterations):
    return cv2.erode(a, kernel, anchor=(ax, ay), iterations=iterations)


def threshold(a, thresh, maxval, ttype):
    return cv2.threshold(np.ascontiguousarray(a), thresh, maxval, ttype)


def getStructuringElement(shape, w, h):
    return cv2.getStructuringElement(shape, (w, h))


def split(a):
    return list(cv2.split(a))


def merge(chs):
    return cv2.merge(chs)


def GaussianBlur(a, kw, kh, sx, sy):
    return cv2.GaussianBlur(a, (kw, kh), sx, sigmaY=sy)


def Canny(a, lo, hi):
    return cv2.Canny(a, lo, hi)


def equalizeHist(a):
    return cv2.equalizeHist(a)


def clahe_apply(a, clip, tx, ty):
    return cv2.createCLAHE(clipLimit=clip, tileGridSize=(tx, ty)).apply(a)


def findContours(a, mode, method, ox, oy):
    contours, hierarchy = cv2.findContours(np.ascontiguousarray(a), mode, method, offset=(ox, oy))
    pts = [np.ascontiguousarray(c.reshape(-1, 2), dtype=np.int32) for c in contours]
    hier = np.zeros((0, 4), np.int32) if hierarchy is None else np.ascontiguousarray(hierarchy.reshape(-1, 4), dtype=np.int32)
    return pts, hier


def boundingRect(pts):
    return cv2.boundingRect(pts)


def contourArea(pts, oriented):
    return float(cv2.contourArea(pts, oriented))


def adaptiveThreshold(a, maxval, method, ttype, block, c):
    return cv2.adaptiveThreshold(a, maxval, method, ttype, block, c)


def medianBlur(a, k):
    return cv2.medianBlur(a, k)


def bilateralFilter(a, d, sigma_color, sigma_space):
    return cv2.bilateralFilter(a, d, sigma_color, sigma_space)


def mean(a):
    return tuple(float(v) for v in cv2.mean(a))


def inRange(a, lo, hi):
    return cv2.inRange(a, lo, hi)


def points_array(pts):
    return np.asarray(pts, dtype=np.int32).reshape(-1, 1, 2)


def copy_masked(src, dst, mask):
    out = np.zeros_like(src) if dst is None or dst.shape != src.shape or dst.dtype != src.dtype else dst.copy()
    out[mask != 0] = src[mask != 0]
    return out
