/* Forced with -include before coviar_data_loader.c (a Python-2 extension module): the two
 * Python-2-only spellings its module-init / main() use, so that the file compiles against the
 * CPython 3 headers of this image.  Neither is reached by the tests, which call
 * create_and_load_mv_residual directly. */
#ifndef LSFA_SHIM_PY2_COMPAT_H
#define LSFA_SHIM_PY2_COMPAT_H
#define NPY_NO_DEPRECATED_API_WARNING_SILENCE 1
#include <Python.h>
#define Py_InitModule3(name, methods, doc) ((PyObject*)NULL)
#endif
