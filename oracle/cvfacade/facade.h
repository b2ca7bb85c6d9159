/* cv:: facade internals shared by facade.cpp and module.cpp -- TEST INFRASTRUCTURE (oracle/_ref). */
#ifndef PRL_CVFACADE_FACADE_H
#define PRL_CVFACADE_FACADE_H

struct _object;
typedef struct _object PyObject;

namespace cvfacade {
void set_calls(PyObject* cvcalls_module);            /* registers oracle/cvfacade/cvcalls.py */
PyObject* call(const char* name, const char* fmt, ...); /* cvcalls.<name>(*args); new reference; throws cv::Exception */
}

#endif
