/*
 * _prl_ref: Python entry points to the reference's OWN functions, compiled UNMODIFIED from /root/reference/src
 * against the cv:: facade (oracle/cvfacade).  TEST INFRASTRUCTURE (oracle/_ref), never linked into the product.
 *
 * Each wrapper binds the numpy input to a cv::Mat, calls prl::<function> with exactly the arguments given, and hands
 * back (output, imageInput-after-the-call): the five local-statistics binarizers overwrite their input Mat with the
 * padded gray image (binarizeSauvola.cpp:49-52,65), which is observable and therefore part of the parity target.
 * std::invalid_argument -> ValueError, cv::Exception -> cv2.error (registered by oracle/ref.py).
 */
#include <Python.h>

#include <stdexcept>

#include "facade.h"
#include "opencv2/core/core.hpp"

#include "binarizeSauvola.h"
#include "binarizeNiblack.h"
#include "binarizeWolfJolion.h"
#include "binarizeNICK.h"
#include "binarizeFeng.h"
#include "binarizeLocalOtsu.h"
#include "removeLines.h"
#include "binarizeAT.h"
#include "binarizeAGT.h"
#include "binarizeGAT.h"
#include "binarizePureAdaptive.h"
#include "binarizePureAdaptiveGaussian.h"
#include "binarizeNativeAdaptive.h"

static PyObject* g_cv_error = 0;  // cv2.error, or NULL -> RuntimeError

static PyObject* translate_exception() {
    try {
        throw;
    } catch (const std::invalid_argument& e) {
        PyErr_SetString(PyExc_ValueError, e.what());
    } catch (const cv::Exception& e) {
        PyErr_SetString(g_cv_error ? g_cv_error : PyExc_RuntimeError, e.what());
    } catch (const std::exception& e) {
        PyErr_SetString(PyExc_RuntimeError, e.what());
    }
    return 0;
}

static PyObject* mat_or_none(const cv::Mat& m) {
    PyObject* o = m.arr ? m.arr : Py_None;
    Py_INCREF(o);
    return o;
}

static PyObject* result_pair(const cv::Mat& out, const cv::Mat& in_after) {
    PyObject* t = PyTuple_New(2);
    PyTuple_SET_ITEM(t, 0, mat_or_none(out));
    PyTuple_SET_ITEM(t, 1, mat_or_none(in_after));
    return t;
}

static void bind_input(cv::Mat& m, PyObject* image) {
    if (image == Py_None) return;  // empty cv::Mat
    Py_INCREF(image);
    m.bind(image);
}

#define PRL_KWRAP(NAME, CALL)                                                                   \
    static PyObject* py_##NAME(PyObject*, PyObject* args) {                                     \
        PyObject* image;                                                                        \
        int window, morph;                                                                      \
        double k;                                                                               \
        if (!PyArg_ParseTuple(args, "Oidi", &image, &window, &k, &morph)) return 0;             \
        try {                                                                                   \
            cv::Mat in, out;                                                                    \
            bind_input(in, image);                                                              \
            CALL(in, out, window, k, morph);                                                    \
            return result_pair(out, in);                                                        \
        } catch (...) {                                                                         \
            return translate_exception();                                                       \
        }                                                                                       \
    }

PRL_KWRAP(binarizeSauvola, prl::binarizeSauvola)
PRL_KWRAP(binarizeNiblack, prl::binarizeNiblack)
PRL_KWRAP(binarizeWolfJolion, prl::binarizeWolfJolion)
PRL_KWRAP(binarizeNICK, prl::binarizeNICK)

static PyObject* py_binarizeFeng(PyObject*, PyObject* args) {
    PyObject* image;
    int window, morph;
    double a1, k1, k2, gamma;
    if (!PyArg_ParseTuple(args, "Oiddddi", &image, &window, &a1, &k1, &k2, &gamma, &morph)) return 0;
    try {
        cv::Mat in, out;
        bind_input(in, image);
        prl::binarizeFeng(in, out, window, a1, k1, k2, gamma, morph);
        return result_pair(out, in);
    } catch (...) {
        return translate_exception();
    }
}

static PyObject* py_binarizeLocalOtsu(PyObject*, PyObject* args) {
    PyObject* image;
    double maxval, clahe, upper, lower;
    int gauss, morph;
    if (!PyArg_ParseTuple(args, "Oddiddi", &image, &maxval, &clahe, &gauss, &upper, &lower, &morph)) return 0;
    try {
        cv::Mat in, out;
        bind_input(in, image);
        prl::binarizeLocalOtsu(in, out, maxval, clahe, gauss, upper, lower, morph);
        return result_pair(out, in);
    } catch (...) {
        return translate_exception();
    }
}

static PyObject* py_removeLines(PyObject*, PyObject* args) {
    PyObject* image;
    if (!PyArg_ParseTuple(args, "O", &image)) return 0;
    try {
        cv::Mat in, out;
        bind_input(in, image);
        prl::removeLines(in, out);
        return result_pair(out, in);
    } catch (...) {
        return translate_exception();
    }
}

// the adaptive-mean family: (image, [kernel size,] maxValue, blockSize, shift)
#define PRL_ATWRAP(NAME, FMT, DECL, PARSE, CALLARGS)                                            \
    static PyObject* py_##NAME(PyObject*, PyObject* args) {                                     \
        PyObject* image;                                                                        \
        DECL;                                                                                   \
        if (!PyArg_ParseTuple(args, FMT, &image, PARSE)) return 0;                              \
        try {                                                                                   \
            cv::Mat in, out;                                                                    \
            bind_input(in, image);                                                              \
            prl::NAME CALLARGS;                                                                 \
            return result_pair(out, in);                                                        \
        } catch (...) {                                                                         \
            return translate_exception();                                                       \
        }                                                                                       \
    }
#define PRL_COMMA ,
PRL_ATWRAP(binarizeAT, "Oidii", int mk PRL_COMMA bs PRL_COMMA sh; double mv, &mk PRL_COMMA &mv PRL_COMMA &bs PRL_COMMA &sh, (in, out, mk, mv, bs, sh))
PRL_ATWRAP(binarizeAGT, "Oidii", int mk PRL_COMMA bs PRL_COMMA sh; double mv, &mk PRL_COMMA &mv PRL_COMMA &bs PRL_COMMA &sh, (in, out, mk, mv, bs, sh))
PRL_ATWRAP(binarizeGAT, "Oidddii", int gk PRL_COMMA bs PRL_COMMA sh; double sx PRL_COMMA sy PRL_COMMA mv,
           &gk PRL_COMMA &sx PRL_COMMA &sy PRL_COMMA &mv PRL_COMMA &bs PRL_COMMA &sh, (in, out, gk, sx, sy, mv, bs, sh))
PRL_ATWRAP(binarizePureAdaptive, "Odii", int bs PRL_COMMA sh; double mv, &mv PRL_COMMA &bs PRL_COMMA &sh, (in, out, mv, bs, sh))
PRL_ATWRAP(binarizePureAdaptiveGaussian, "Odii", int bs PRL_COMMA sh; double mv, &mv PRL_COMMA &bs PRL_COMMA &sh, (in, out, mv, bs, sh))

static PyObject* py_binarizeNativeAdaptive(PyObject*, PyObject* args) {
    PyObject* image;
    int gauss_blur, median_k, gauss_k, by_gaussian, block, bil_k;
    double gauss_sigma, maxval, shift, bil_color, bil_space;
    if (!PyArg_ParseTuple(args, "Oiiididididd", &image, &gauss_blur, &median_k, &gauss_k, &gauss_sigma, &by_gaussian, &maxval, &block, &shift,
                          &bil_k, &bil_color, &bil_space))
        return 0;
    try {
        cv::Mat in, out;
        bind_input(in, image);
        prl::binarizeNativeAdaptive(in, out, gauss_blur != 0, median_k, gauss_k, gauss_sigma, by_gaussian != 0, maxval, block, shift, bil_k,
                                    bil_color, bil_space);
        return result_pair(out, in);
    } catch (...) {
        return translate_exception();
    }
}

static PyObject* py_register(PyObject*, PyObject* args) {
    PyObject *calls, *cv_error;
    if (!PyArg_ParseTuple(args, "OO", &calls, &cv_error)) return 0;
    cvfacade::set_calls(calls);
    Py_XINCREF(cv_error);
    Py_XDECREF(g_cv_error);
    g_cv_error = cv_error == Py_None ? 0 : cv_error;
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"register", py_register, METH_VARARGS, "register(cvcalls_module, cv2.error)"},
    {"binarizeSauvola", py_binarizeSauvola, METH_VARARGS, "prl::binarizeSauvola(image, windowSize, k, morph) -> (out, image_after)"},
    {"binarizeNiblack", py_binarizeNiblack, METH_VARARGS, "prl::binarizeNiblack"},
    {"binarizeWolfJolion", py_binarizeWolfJolion, METH_VARARGS, "prl::binarizeWolfJolion"},
    {"binarizeNICK", py_binarizeNICK, METH_VARARGS, "prl::binarizeNICK"},
    {"binarizeFeng", py_binarizeFeng, METH_VARARGS, "prl::binarizeFeng(image, windowSize, alpha1, k1, k2, gamma, morph)"},
    {"binarizeLocalOtsu", py_binarizeLocalOtsu, METH_VARARGS, "prl::binarizeLocalOtsu(image, maxValue, clahe, gauss, upper, lower, morph)"},
    {"removeLines", py_removeLines, METH_VARARGS, "prl::removeLines(image)"},
    {"binarizeAT", py_binarizeAT, METH_VARARGS, "prl::binarizeAT(image, medianKernelSize, maxValue, blockSize, shift)"},
    {"binarizeAGT", py_binarizeAGT, METH_VARARGS, "prl::binarizeAGT(image, medianKernelSize, maxValue, blockSize, shift)"},
    {"binarizeGAT", py_binarizeGAT, METH_VARARGS, "prl::binarizeGAT(image, gaussianKernelSize, sigmaX, sigmaY, maxValue, blockSize, shift)"},
    {"binarizePureAdaptive", py_binarizePureAdaptive, METH_VARARGS, "prl::binarizePureAdaptive(image, maxValue, blockSize, shift)"},
    {"binarizePureAdaptiveGaussian", py_binarizePureAdaptiveGaussian, METH_VARARGS, "prl::binarizePureAdaptiveGaussian(image, maxValue, blockSize, shift)"},
    {"binarizeNativeAdaptive", py_binarizeNativeAdaptive, METH_VARARGS,
     "prl::binarizeNativeAdaptive(image, gaussBlur, medianK, gaussK, gaussSigma, byGaussian, maxValue, blockSize, shift, bilateralK, colorSigma, spaceSigma)"},
    {0, 0, 0, 0}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_prl_ref",
                                       "PRLib's own binarize*.cpp compiled unmodified against the cv:: facade", -1, methods,
                                       0, 0, 0, 0};

PyMODINIT_FUNC PyInit__prl_ref(void) { return PyModule_Create(&moduledef); }
