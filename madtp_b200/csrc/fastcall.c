/* Low-overhead Python -> C-ABI call path for libmadtp_b200.so (CPython extension, x86-64 System V only).
 *
 * ctypes spends ~5 us converting the ~20 arguments of one entry point; the text encoder launches ~250 kernels of
 * 5-15 us each per step, so the binding, not the GPU, set the pace there. Every entry point of include/madtp_b200.h
 * takes only integer-class arguments (pointers, int, int64_t) and `float`s and returns int. Under the System V
 * x86-64 calling convention integer-class arguments are assigned to rdi, rsi, rdx, rcx, r8, r9 and then to stack
 * slots in order AMONG THEMSELVES, and float arguments to xmm0..7 in order among themselves, independent of how the
 * two kinds interleave in the prototype -- so one trampoline type (28 integers followed by 4 floats) reaches every
 * entry point: surplus trailing arguments are ignored by the callee (caller cleans the stack).
 *
 *   _fastcall.call(fn_address, *args) -> int     args: int / None (-> 0) in integer slots, float in float slots
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#if !defined(__x86_64__)
#error "fastcall.c relies on the System V x86-64 calling convention"
#endif

#define NI 28
#define NF 4
typedef int (*tramp_t)(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                       int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                       int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float, float, float,
                       float);

static PyObject* fc_call(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  (void)self;
  if (nargs < 1) {
    PyErr_SetString(PyExc_TypeError, "call(fn_address, *args)");
    return NULL;
  }
  tramp_t fn = (tramp_t)PyLong_AsVoidPtr(args[0]);
  if (fn == NULL) {
    if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "null function address");
    return NULL;
  }
  int64_t iv[NI] = {0};
  float fv[NF] = {0.f, 0.f, 0.f, 0.f};
  int ni = 0, nf = 0;
  for (Py_ssize_t i = 1; i < nargs; ++i) {
    PyObject* o = args[i];
    if (o == Py_None) {
      if (ni >= NI) goto too_many;
      iv[ni++] = 0;
    } else if (PyFloat_CheckExact(o)) {
      if (nf >= NF) goto too_many;
      fv[nf++] = (float)PyFloat_AS_DOUBLE(o);
    } else {
      if (ni >= NI) goto too_many;
      long long v = PyLong_AsLongLong(o);
      if (v == -1 && PyErr_Occurred()) {
        PyErr_Clear();
        unsigned long long u = PyLong_AsUnsignedLongLongMask(o);
        if (u == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
        v = (long long)u;
      }
      iv[ni++] = (int64_t)v;
    }
  }
  {
    /* The GIL is released around the call exactly as ctypes does: bench.py samples clocks from a second Python thread,
     * and the CUDA driver may block inside a launch. */
    int st;
    Py_BEGIN_ALLOW_THREADS
    st = fn(iv[0], iv[1], iv[2], iv[3], iv[4], iv[5], iv[6], iv[7], iv[8], iv[9], iv[10], iv[11], iv[12], iv[13],
                iv[14], iv[15], iv[16], iv[17], iv[18], iv[19], iv[20], iv[21], iv[22], iv[23], iv[24], iv[25], iv[26],
                iv[27], fv[0], fv[1], fv[2], fv[3]);
    Py_END_ALLOW_THREADS
    return PyLong_FromLong(st);
  }
too_many:
  PyErr_SetString(PyExc_TypeError, "too many arguments for the fastcall trampoline");
  return NULL;
}

static PyMethodDef methods[] = {{"call", (PyCFunction)(void (*)(void))fc_call, METH_FASTCALL,
                                 "call(fn_address, *args) -> int status"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastcall", "fast C-ABI trampoline", -1, methods,
                                    NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__fastcall(void) { return PyModule_Create(&moddef); }
