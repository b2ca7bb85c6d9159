/*
 * cv:: facade, implementation -- TEST INFRASTRUCTURE (oracle/_ref).  See opencv2/core/core.hpp.
 * Every cv:: primitive below is ONE call into oracle/cvfacade/cvcalls.py, i.e. into the cv2 wheel's OpenCV.
 * Runs with the GIL held (the module is only entered from Python).
 */
#include <Python.h>

#include <cstdarg>
#include <cstring>
#include <stdexcept>

#include "facade.h"
#include "opencv2/imgproc/imgproc.hpp"

namespace cvfacade {

static PyObject* g_calls = 0;  // the cvcalls module (owned)

void set_calls(PyObject* module) {
    Py_XINCREF(module);
    Py_XDECREF(g_calls);
    g_calls = module;
}

[[noreturn]] static void throw_python_error() {
    PyObject *type = 0, *value = 0, *tb = 0;
    PyErr_Fetch(&type, &value, &tb);
    PyErr_NormalizeException(&type, &value, &tb);
    std::string text = "python error";
    if (value) {
        PyObject* s = PyObject_Str(value);
        if (s) {
            const char* c = PyUnicode_AsUTF8(s);
            if (c) text = c;
            Py_DECREF(s);
        }
    }
    Py_XDECREF(type);
    Py_XDECREF(value);
    Py_XDECREF(tb);
    PyErr_Clear();
    // whatever OpenCV (or the binding's argument check) rejected surfaces as cv::Exception, like in the C++ API
    throw cv::Exception(-2, text, "cvcalls", __FILE__, __LINE__);
}

PyObject* call(const char* name, const char* fmt, ...) {
    if (!g_calls) throw std::runtime_error("cvfacade: cvcalls module not registered");
    PyObject* fn = PyObject_GetAttrString(g_calls, name);
    if (!fn) throw_python_error();
    va_list va;
    va_start(va, fmt);
    PyObject* args = Py_VaBuildValue(fmt, va);
    va_end(va);
    if (!args) {
        Py_DECREF(fn);
        throw_python_error();
    }
    if (!PyTuple_Check(args)) {
        PyObject* t = PyTuple_Pack(1, args);
        Py_DECREF(args);
        args = t;
    }
    PyObject* r = PyObject_CallObject(fn, args);
    Py_DECREF(fn);
    Py_DECREF(args);
    if (!r) throw_python_error();
    return r;
}

static inline PyObject* py(const cv::Mat& m) { return m.arr ? m.arr : Py_None; }

static PyObject* scalar_tuple(const cv::Scalar& s) { return Py_BuildValue("(dddd)", s[0], s[1], s[2], s[3]); }

static PyObject* points_array(const std::vector<cv::Point>& pts) {
    PyObject* lst = PyList_New(static_cast<Py_ssize_t>(pts.size()));
    for (size_t i = 0; i < pts.size(); ++i) PyList_SET_ITEM(lst, i, Py_BuildValue("(ii)", pts[i].x, pts[i].y));
    PyObject* r = call("points_array", "(O)", lst);
    Py_DECREF(lst);
    return r;
}

}  // namespace cvfacade

using cvfacade::call;
using cvfacade::py;

namespace cv {

void error(int code, const std::string& err, const char* func, const char* file, int line) {
    throw Exception(code, err, func ? func : "", file ? file : "", line);
}

/* ------------------------------------------------------------------ Mat: header bookkeeping only */
Mat::Mat() : flags(0), dims(0), rows(0), cols(0), data(0), step(0), arr(0), root(0), ox(0), oy(0) {}

Mat::Mat(const Mat& m) : flags(m.flags), dims(m.dims), rows(m.rows), cols(m.cols), data(m.data), step(m.step),
                         arr(m.arr), root(m.root), ox(m.ox), oy(m.oy) {
    Py_XINCREF(arr);
    Py_XINCREF(root);
}

Mat::~Mat() {
    Py_XDECREF(arr);
    Py_XDECREF(root);
}

Mat& Mat::operator=(const Mat& m) {
    if (this != &m) {
        Py_XINCREF(m.arr);
        Py_XINCREF(m.root);
        Py_XDECREF(arr);
        Py_XDECREF(root);
        flags = m.flags; dims = m.dims; rows = m.rows; cols = m.cols; data = m.data; step = m.step;
        arr = m.arr; root = m.root; ox = m.ox; oy = m.oy;
    }
    return *this;
}

void Mat::bind(PyObject* a, PyObject* r, int ox_, int oy_) {
    Py_XDECREF(arr);
    Py_XDECREF(root);
    arr = a;
    if (!a) {
        Py_XDECREF(r);
        root = 0; flags = 0; dims = 0; rows = cols = 0; data = 0; step = 0; ox = oy = 0;
        return;
    }
    if (r) {
        root = r;
    } else {
        root = a;
        Py_INCREF(root);
    }
    ox = ox_;
    oy = oy_;
    Py_buffer view;
    if (PyObject_GetBuffer(a, &view, PyBUF_RECORDS_RO) != 0) cvfacade::throw_python_error();
    int depth = -1;
    const char* f = view.format ? view.format : "B";
    while (*f == '<' || *f == '=' || *f == '@' || *f == '|') ++f;
    switch (*f) {
        case 'B': depth = CV_8U; break;
        case 'b': depth = CV_8S; break;
        case 'H': depth = CV_16U; break;
        case 'h': depth = CV_16S; break;
        case 'i': case 'l': depth = view.itemsize == 4 ? CV_32S : -1; break;
        case 'f': depth = CV_32F; break;
        case 'd': depth = CV_64F; break;
        case '?': depth = CV_8U; break;
        default: depth = -1;
    }
    int nd = view.ndim;
    if (depth < 0 || nd < 2 || nd > 3) {
        PyBuffer_Release(&view);
        throw Exception(-210, "cvfacade: unsupported array layout", "Mat::bind", __FILE__, __LINE__);
    }
    int cn = nd == 3 ? static_cast<int>(view.shape[2]) : 1;
    rows = static_cast<int>(view.shape[0]);
    cols = static_cast<int>(view.shape[1]);
    dims = 2;
    flags = CV_MAKETYPE(depth, cn);
    step = static_cast<size_t>(view.strides[0]);
    data = static_cast<uchar*>(view.buf);
    PyBuffer_Release(&view);  // the array object stays alive through `arr`, so the pointer stays valid
}

void Mat::store(PyObject* result) {
    if (arr && result != arr) {
        PyObject* same = call("same_layout", "(OO)", arr, result);
        bool keep = PyObject_IsTrue(same) == 1;
        Py_DECREF(same);
        if (keep) {  // Mat::create() is a no-op: the result lands in the existing buffer
            PyObject* r = call("copy_into", "(OO)", arr, result);
            Py_DECREF(r);
            Py_DECREF(result);
            return;
        }
    }
    bind(result);
}

Mat::Mat(int r, int c, int type) : Mat() { create(r, c, type); }
Mat::Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
Mat::Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
Mat::Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); setTo(s); }
Mat::Mat(const Mat& m, const Rect& roi) : Mat() { *this = m(roi); }

void Mat::create(int r, int c, int type) {
    if (arr && rows == r && cols == c && this->type() == (type & 0xFFF)) return;
    CV_Assert(r >= 0 && c >= 0);
    bind(call("empty", "(iii)", r, c, type));
}
void Mat::create(Size sz, int type) { create(sz.height, sz.width, type); }

void Mat::release() { bind(0); }

size_t Mat::elemSize() const {
    static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2};
    return static_cast<size_t>(sz[depth()]) * channels();
}
bool Mat::isContinuous() const { return step == elemSize() * cols || rows <= 1; }
bool Mat::isSubmatrix() const { return arr != root; }

Mat Mat::operator()(const Rect& r) const {
    // Mat::Mat(const Mat& m, const Rect& roi) asserts exactly this (modules/core/src/matrix.cpp)
    CV_Assert(0 <= r.x && 0 <= r.width && r.x + r.width <= cols && 0 <= r.y && 0 <= r.height && r.y + r.height <= rows);
    Mat out;
    PyObject* v = call("roi", "(Oiiii)", py(*this), r.y, r.y + r.height, r.x, r.x + r.width);
    Py_XINCREF(root);
    out.bind(v, root, ox + r.x, oy + r.y);
    return out;
}

Mat Mat::clone() const {
    Mat out;
    if (arr) out.bind(call("clone", "(O)", arr));
    return out;
}

void Mat::copyTo(Mat& dst) const {
    if (!arr) { dst.release(); return; }
    dst.store(call("clone", "(O)", arr));
}

void Mat::copyTo(Mat& dst, const Mat& mask) const {
    if (!mask.arr) { copyTo(dst); return; }
    dst.store(call("copy_masked", "(OOO)", py(*this), py(dst), py(mask)));
}

Mat& Mat::setTo(const Scalar& s, const Mat& mask) {
    if (!arr) return *this;
    PyObject* t = cvfacade::scalar_tuple(s);
    PyObject* r = call("set_to", "(OOO)", arr, t, py(mask));
    Py_DECREF(t);
    Py_DECREF(r);
    return *this;
}

Mat& Mat::operator=(const Scalar& s) { return setTo(s); }

Mat Mat::zeros(int r, int c, int type) { Mat m; m.bind(call("zeros", "(iii)", r, c, type)); return m; }
Mat Mat::zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }
Mat Mat::ones(int r, int c, int type) { Mat m; m.bind(call("ones", "(iii)", r, c, type)); return m; }
Mat Mat::ones(Size sz, int type) { return ones(sz.height, sz.width, type); }

void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    CV_Assert(arr != 0);
    if (rtype < 0) rtype = type();
    dst.store(call("convert_to", "(Oidd)", arr, CV_MAT_DEPTH(rtype), alpha, beta));
}

/* ------------------------------------------------------------------ MatExpr: OpenCV's lowering (matop.cpp) */
static MatExpr addex(const Mat& a, const Mat& b, double alpha, double beta, const Scalar& s = Scalar()) {
    MatExpr e;
    e.kind = MatExpr::ADDEX; e.a = a; e.b = b; e.alpha = alpha; e.beta = b.data ? beta : 0; e.s = s;
    return e;
}
static MatExpr binop(int op, const Mat& a, const Mat& b, double scale = 1, const Scalar& s = Scalar()) {
    MatExpr e;
    e.kind = MatExpr::BIN; e.op = op; e.a = a; e.b = b; e.alpha = scale; e.s = s;
    return e;
}

MatExpr::operator Mat() const {
    Mat m;
    assign(m);
    return m;
}

Mat& Mat::operator=(const MatExpr& e) {
    e.assign(*this);
    return *this;
}

void MatExpr::assign(Mat& m) const {
    switch (kind) {
    case IDENTITY:
        m = a;
        return;
    case ADDEX:  // MatOp_AddEx::assign with _type == -1
        if (b.data) {
            if (s == Scalar() || !s.isReal()) {
                if (alpha == 1) {
                    if (beta == 1) cv::add(a, b, m);
                    else if (beta == -1) cv::subtract(a, b, m);
                    else cv::scaleAdd(b, beta, a, m);
                } else if (beta == 1) {
                    if (alpha == -1) cv::subtract(b, a, m);
                    else cv::scaleAdd(a, alpha, b, m);
                } else {
                    cv::addWeighted(a, alpha, b, beta, 0, m);
                }
                if (!s.isReal()) CV_Error(-213, "cvfacade: non-real scalar in MatExpr");
            } else {
                cv::addWeighted(a, alpha, b, beta, s[0], m);
            }
        } else if (s.isReal() && std::fabs(alpha) != 1) {
            a.convertTo(m, -1, alpha, s[0]);
        } else if (alpha == 1) {
            PyObject* t = cvfacade::scalar_tuple(s);
            PyObject* r = call("add_scalar", "(OO)", py(a), t);
            Py_DECREF(t);
            m.store(r);
        } else if (alpha == -1) {
            PyObject* t = cvfacade::scalar_tuple(s);
            PyObject* r = call("rsub_scalar", "(OO)", t, py(a));
            Py_DECREF(t);
            m.store(r);
        } else {
            CV_Error(-213, "cvfacade: MatExpr form not produced by the reference");
        }
        return;
    case BIN:
        if (op == '*') cv::multiply(a, b, m, alpha);
        else if (op == '/') cv::divide(a, b, m, alpha);
        else if (op == '|') cv::bitwise_or(a, b, m);
        else if (op == '~') cv::bitwise_not(a, m);
        else if (op == '^' && !b.data) {
            PyObject* t = cvfacade::scalar_tuple(s);
            PyObject* r = call("bitwise_xor_scalar", "(OO)", py(a), t);
            Py_DECREF(t);
            m.store(r);
        } else CV_Error(-213, "cvfacade: binary MatExpr not produced by the reference");
        return;
    case CMP:
        cv::compare(a, b, m, op);
        return;
    }
}

// MatOp::add / MatOp::subtract: a term of the form alpha*A + s is folded, anything else is evaluated first
static MatExpr fold_add(const MatExpr& e1, const MatExpr& e2, double sign) {
    double alpha = 1, beta = sign;
    Scalar s;
    Mat m1, m2;
    if (e1.kind == MatExpr::ADDEX && (!e1.b.data || e1.beta == 0)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; }
    else e1.assign(m1);
    if (e2.kind == MatExpr::ADDEX && (!e2.b.data || e2.beta == 0)) {
        m2 = e2.a; beta = sign * e2.alpha;
        for (int i = 0; i < 4; ++i) s[i] += sign * e2.s[i];
    } else e2.assign(m2);
    return addex(m1, m2, alpha, beta, s);
}

MatExpr operator+(const Mat& a, const Mat& b) { return addex(a, b, 1, 1); }
MatExpr operator+(const Mat& a, const MatExpr& e) { return fold_add(e, MatExpr(a), 1); }   // e.op->add(e, MatExpr(a), en)
MatExpr operator+(const MatExpr& e, const Mat& b) { return fold_add(e, MatExpr(b), 1); }
MatExpr operator+(const MatExpr& e1, const MatExpr& e2) { return fold_add(e1, e2, 1); }
MatExpr operator+(const Mat& a, const Scalar& s) { return addex(a, Mat(), 1, 0, s); }
MatExpr operator+(const Scalar& s, const Mat& a) { return addex(a, Mat(), 1, 0, s); }
MatExpr operator-(const Mat& a, const Mat& b) { return addex(a, b, 1, -1); }
MatExpr operator-(const Mat& a, const Scalar& s) { return addex(a, Mat(), 1, 0, -s); }
MatExpr operator-(const Scalar& s, const Mat& a) { return addex(a, Mat(), -1, 0, s); }
MatExpr operator-(const MatExpr& e, const Mat& b) { return fold_add(e, MatExpr(b), -1); }
MatExpr operator*(const Mat& a, double s) { return addex(a, Mat(), s, 0); }
MatExpr operator*(double s, const Mat& a) { return addex(a, Mat(), s, 0); }
MatExpr operator>(const Mat& a, const Mat& b) { MatExpr e; e.kind = MatExpr::CMP; e.op = CMP_GT; e.a = a; e.b = b; return e; }
MatExpr operator<(const Mat& a, const Mat& b) { MatExpr e; e.kind = MatExpr::CMP; e.op = CMP_LT; e.a = a; e.b = b; return e; }
MatExpr operator^(const Mat& a, const Scalar& s) { return binop('^', a, Mat(), 1, s); }
MatExpr operator|(const Mat& a, const Mat& b) { return binop('|', a, b); }
MatExpr operator~(const Mat& a) { return binop('~', a, Mat()); }
Mat& operator-=(Mat& a, const Mat& b) { cv::subtract(a, b, a); return a; }
Mat& operator+=(Mat& a, const Mat& b) { cv::add(a, b, a); return a; }
Mat& operator+=(Mat& a, const MatExpr& e) { Mat t = e; cv::add(a, t, a); return a; }   // e.op->augAssignAdd -> assign + add
Mat& operator*=(Mat& a, double s) { a.convertTo(a, -1, s); return a; }
Mat& operator/=(Mat& a, double s) { a.convertTo(a, -1, 1. / s); return a; }
void Mat::not_forwarded(const char* what) { cv::error(-213, std::string("cvfacade: not forwarded (off the binarization path): ") + what, "", __FILE__, __LINE__); }

MatExpr Mat::mul(const Mat& m, double scale) const { return binop('*', *this, m, scale); }
MatExpr Mat::mul(const MatExpr& e, double scale) const {  // MatOp::multiply(MatExpr(*this), e, res, scale)
    if (e.scaled()) return binop('*', *this, e.a, scale * e.alpha);
    Mat m2 = e;
    return binop('*', *this, m2, scale);
}
MatExpr MatExpr::mul(const Mat& m, double scale) const { Mat m1 = *this; return binop('*', m1, m, scale); }
MatExpr MatExpr::mul(const MatExpr& e, double scale) const { Mat m1 = *this; return m1.mul(e, scale); }

/* ------------------------------------------------------------------ core functions: one forwarded call each */
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar&) {
    dst.store(call("copyMakeBorder", "(Oiiiii)", py(src), top, bottom, left, right, borderType));
}
void sqrt(const Mat& src, Mat& dst) { dst.store(call("sqrt", "(O)", py(src))); }
void pow(const Mat& src, double power, Mat& dst) { dst.store(call("pow_", "(Od)", py(src), power)); }
void minMaxLoc(const Mat& src, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc, const Mat& mask) {
    CV_Assert(mask.arr == 0);
    PyObject* r = call("minMaxLoc", "(O)", py(src));
    double mn, mx;
    int x0, y0, x1, y1;
    if (!PyArg_ParseTuple(r, "dd(ii)(ii)", &mn, &mx, &x0, &y0, &x1, &y1)) { Py_DECREF(r); cvfacade::throw_python_error(); }
    Py_DECREF(r);
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = Point(x0, y0);
    if (maxLoc) *maxLoc = Point(x1, y1);
}
Scalar mean(const Mat& src) {
    PyObject* r = call("mean", "(O)", py(src));
    Scalar s;
    if (!PyArg_ParseTuple(r, "dddd", &s.val[0], &s.val[1], &s.val[2], &s.val[3])) { Py_DECREF(r); cvfacade::throw_python_error(); }
    Py_DECREF(r);
    return s;
}
void add(const Mat& a, const Mat& b, Mat& dst) { dst.store(call("add", "(OO)", py(a), py(b))); }
void subtract(const Mat& a, const Mat& b, Mat& dst) { dst.store(call("subtract", "(OO)", py(a), py(b))); }
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale) { dst.store(call("multiply", "(OOd)", py(a), py(b), scale)); }
void divide(const Mat& a, const Mat& b, Mat& dst, double scale) { dst.store(call("divide", "(OOd)", py(a), py(b), scale)); }
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst) { dst.store(call("scaleAdd", "(OdO)", py(a), alpha, py(b))); }
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst) {
    dst.store(call("addWeighted", "(OdOdd)", py(a), alpha, py(b), beta, gamma));
}
void compare(const Mat& a, const Mat& b, Mat& dst, int cmpop) { dst.store(call("compare", "(OOi)", py(a), py(b), cmpop)); }
void bitwise_not(const Mat& src, Mat& dst) { dst.store(call("bitwise_not", "(O)", py(src))); }
void bitwise_or(const Mat& a, const Mat& b, Mat& dst) { dst.store(call("bitwise_or", "(OO)", py(a), py(b))); }
void inRange(const Mat& src, const Scalar& lo, const Scalar& hi, Mat& dst) {
    PyObject *l = cvfacade::scalar_tuple(lo), *h = cvfacade::scalar_tuple(hi);
    PyObject* r = call("inRange", "(OOO)", py(src), l, h);
    Py_DECREF(l);
    Py_DECREF(h);
    dst.store(r);
}
void split(const Mat& src, std::vector<Mat>& mv) {
    PyObject* r = call("split", "(O)", py(src));
    Py_ssize_t n = PyList_Size(r);
    mv.resize(static_cast<size_t>(n));
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* it = PyList_GetItem(r, i);
        Py_INCREF(it);
        mv[static_cast<size_t>(i)].store(it);
    }
    Py_DECREF(r);
}
void merge(const std::vector<Mat>& mv, Mat& dst) {
    PyObject* lst = PyList_New(static_cast<Py_ssize_t>(mv.size()));
    for (size_t i = 0; i < mv.size(); ++i) {
        PyObject* it = py(mv[i]);
        Py_INCREF(it);
        PyList_SET_ITEM(lst, i, it);
    }
    PyObject* r = call("merge", "(O)", lst);
    Py_DECREF(lst);
    dst.store(r);
}

/* ------------------------------------------------------------------ imgproc */
void cvtColor(const Mat& src, Mat& dst, int code, int) { dst.store(call("cvtColor", "(Oi)", py(src), code)); }

void integral(const Mat& src, Mat& sum, Mat& sqsum, int sdepth, int sqdepth) {
    PyObject* r = call("integral2", "(Oii)", py(src), sdepth, sqdepth);
    PyObject *s = PyTuple_GetItem(r, 0), *q = PyTuple_GetItem(r, 1);
    Py_INCREF(s);
    Py_INCREF(q);
    Py_DECREF(r);
    sum.store(s);
    sqsum.store(q);
}

void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor, double delta, int borderType) {
    // the source goes down with its parent and offset: OpenCV looks at both (see cvcalls.filter2D)
    dst.store(call("filter2D", "(OiiiiiOiidi)", src.root ? src.root : Py_None, src.oy, src.ox, src.rows, src.cols,
                   ddepth, py(kernel), anchor.x, anchor.y, delta, borderType));
}
void dilate(const Mat& src, Mat& dst, const Mat& kernel, Point anchor, int iterations) {
    dst.store(call("dilate", "(OOiii)", py(src), py(kernel), anchor.x, anchor.y, iterations));
}
void erode(const Mat& src, Mat& dst, const Mat& kernel, Point anchor, int iterations) {
    dst.store(call("erode", "(OOiii)", py(src), py(kernel), anchor.x, anchor.y, iterations));
}
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
    PyObject* r = call("threshold", "(Oddi)", py(src), thresh, maxval, type);
    double t = PyFloat_AsDouble(PyTuple_GetItem(r, 0));
    PyObject* m = PyTuple_GetItem(r, 1);
    Py_INCREF(m);
    Py_DECREF(r);
    dst.store(m);
    return t;
}
void adaptiveThreshold(const Mat& src, Mat& dst, double maxValue, int adaptiveMethod, int thresholdType, int blockSize, double C) {
    dst.store(call("adaptiveThreshold", "(Odiiid)", py(src), maxValue, adaptiveMethod, thresholdType, blockSize, C));
}
Mat getStructuringElement(int shape, Size ksize, Point) {
    Mat m;
    m.bind(call("getStructuringElement", "(iii)", shape, ksize.width, ksize.height));
    return m;
}
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int) {
    dst.store(call("GaussianBlur", "(Oiidd)", py(src), ksize.width, ksize.height, sigmaX, sigmaY));
}
void medianBlur(const Mat& src, Mat& dst, int ksize) { dst.store(call("medianBlur", "(Oi)", py(src), ksize)); }
void bilateralFilter(const Mat& src, Mat& dst, int d, double sigmaColor, double sigmaSpace, int) {
    dst.store(call("bilateralFilter", "(Oidd)", py(src), d, sigmaColor, sigmaSpace));
}
void Canny(const Mat& image, Mat& edges, double t1, double t2, int aperture, bool l2) {
    CV_Assert(aperture == 3 && !l2);
    edges.store(call("Canny", "(Odd)", py(image), t1, t2));
}
void equalizeHist(const Mat& src, Mat& dst) { dst.store(call("equalizeHist", "(O)", py(src))); }

void findContours(const Mat& image, std::vector<std::vector<Point> >& contours, std::vector<Vec4i>& hierarchy, int mode,
                  int method, Point offset) {
    PyObject* r = call("findContours", "(Oiiii)", py(image), mode, method, offset.x, offset.y);
    PyObject *pts = PyTuple_GetItem(r, 0), *hier = PyTuple_GetItem(r, 1);
    Py_ssize_t n = PyList_Size(pts);
    contours.assign(static_cast<size_t>(n), std::vector<Point>());
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_buffer v;
        if (PyObject_GetBuffer(PyList_GetItem(pts, i), &v, PyBUF_C_CONTIGUOUS) != 0) { Py_DECREF(r); cvfacade::throw_python_error(); }
        const int* p = static_cast<const int*>(v.buf);
        size_t m = static_cast<size_t>(v.len / (2 * sizeof(int)));
        contours[static_cast<size_t>(i)].resize(m);
        for (size_t j = 0; j < m; ++j) contours[static_cast<size_t>(i)][j] = Point(p[2 * j], p[2 * j + 1]);
        PyBuffer_Release(&v);
    }
    Py_buffer hv;
    if (PyObject_GetBuffer(hier, &hv, PyBUF_C_CONTIGUOUS) != 0) { Py_DECREF(r); cvfacade::throw_python_error(); }
    const int* h = static_cast<const int*>(hv.buf);
    size_t hn = static_cast<size_t>(hv.len / (4 * sizeof(int)));
    hierarchy.resize(hn);
    for (size_t j = 0; j < hn; ++j) hierarchy[j] = Vec4i(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
    PyBuffer_Release(&hv);
    Py_DECREF(r);
}

Rect boundingRect(const std::vector<Point>& points) {
    PyObject* a = cvfacade::points_array(points);
    PyObject* r = call("boundingRect", "(O)", a);
    Py_DECREF(a);
    int x, y, w, h;
    if (!PyArg_ParseTuple(r, "iiii", &x, &y, &w, &h)) { Py_DECREF(r); cvfacade::throw_python_error(); }
    Py_DECREF(r);
    return Rect(x, y, w, h);
}

double contourArea(const std::vector<Point>& contour, bool oriented) {
    PyObject* a = cvfacade::points_array(contour);
    PyObject* r = call("contourArea", "(Oi)", a, oriented ? 1 : 0);
    Py_DECREF(a);
    double v = PyFloat_AsDouble(r);
    Py_DECREF(r);
    return v;
}

void convexHull(const std::vector<Point>&, std::vector<Point>&, bool, bool) { CV_Error(-213, "cvfacade: convexHull is not forwarded (off the binarization path)"); }
void convexHull(const Mat&, std::vector<Point2f>&, bool, bool) { Mat::not_forwarded("convexHull"); }
void convexHull(const std::vector<Point>&, std::vector<int>&, bool, bool) { Mat::not_forwarded("convexHull"); }
RotatedRect minAreaRect(const std::vector<Point>&) { CV_Error(-213, "cvfacade: minAreaRect is not forwarded (off the binarization path)"); }
void RotatedRect::points(Point2f[]) const { CV_Error(-213, "cvfacade: RotatedRect::points is not forwarded"); }
void calcHist(const Mat*, int, const int*, const Mat&, Mat&, int, const int*, const float**, bool, bool) {
    CV_Error(-213, "cvfacade: calcHist is not forwarded (off the binarization path)");
}

void CLAHE::apply(const Mat& src, Mat& dst) { dst.store(call("clahe_apply", "(Odii)", py(src), clip, tiles.width, tiles.height)); }
Ptr<CLAHE> createCLAHE(double clipLimit, Size tileGridSize) {
    Ptr<CLAHE> p = std::make_shared<CLAHE>();
    p->setClipLimit(clipLimit);
    p->setTilesGridSize(tileGridSize);
    return p;
}

}  // namespace cv
