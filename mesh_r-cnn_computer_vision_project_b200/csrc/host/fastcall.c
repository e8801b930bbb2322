/* Host-side call shim for the C ABI of libmeshrcnn_b200 (x86-64 System V only).
 *
 * ctypes spends 3-5 us per call converting ~15 arguments through libffi; a refinement step makes ~150 calls, and the step is
 * within 25 % of being bound by the launching thread.  Entry points whose arguments are all pointers / int / long long (all
 * but a handful) can be called through a plain cast: on x86-64 SysV every such argument travels in a 64-bit integer register
 * or an 8-byte stack slot, so `int f(T0, ..., Tn)` is call-compatible with `int f(long long x (n+1))`.  Entry points with
 * float / double arguments keep going through ctypes (meshrcnn_b200/_lib.py).
 *
 *   call_ints(fn_address, (a0, a1, ..., an-1)) -> int        None -> 0; ints (also tensor.data_ptr()) -> long long
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

#define MAXARGS 24
typedef long long ll;

static int dispatch(void* fn, int n, const ll* a) {
    switch (n) {
#define A(i) a[i]
        case 0: return ((int (*)(void))fn)();
        case 1: return ((int (*)(ll))fn)(A(0));
        case 2: return ((int (*)(ll, ll))fn)(A(0), A(1));
        case 3: return ((int (*)(ll, ll, ll))fn)(A(0), A(1), A(2));
        case 4: return ((int (*)(ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3));
        case 5: return ((int (*)(ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4));
        case 6: return ((int (*)(ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5));
        case 7: return ((int (*)(ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6));
        case 8: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7));
        case 9: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8));
        case 10: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9));
        case 11: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10));
        case 12: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11));
        case 13: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12));
        case 14: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13));
        case 15: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14));
        case 16: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15));
        case 17: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16));
        case 18: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17));
        case 19: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18));
        case 20: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19));
        case 21: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20));
        case 22: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21));
        case 23: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21), A(22));
        case 24: return ((int (*)(ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll, ll))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21), A(22), A(23));
#undef A
    }
    return -9999;
}

static PyObject* call_ints(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
    (void)self;
    if (nargs != 2 || !PyTuple_Check(args[1])) {
        PyErr_SetString(PyExc_TypeError, "call_ints(fn_address, args_tuple)");
        return NULL;
    }
    void* fn = PyLong_AsVoidPtr(args[0]);
    if (!fn && PyErr_Occurred()) return NULL;
    const Py_ssize_t n = PyTuple_GET_SIZE(args[1]);
    if (n > MAXARGS) {
        PyErr_SetString(PyExc_ValueError, "call_ints: too many arguments");
        return NULL;
    }
    ll a[MAXARGS];
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = PyTuple_GET_ITEM(args[1], i);
        if (o == Py_None) {
            a[i] = 0;
        } else {
            /* pointers come as unsigned 64-bit Python ints (tensor.data_ptr()); everything else fits a signed long long */
            unsigned long long u = PyLong_AsUnsignedLongLongMask(o);
            if (u == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
            a[i] = (ll)u;
        }
    }
    /* The GIL is NOT released: an entry point only enqueues work (microseconds), and handing the GIL to another Python
     * thread 150 times per step invites multi-millisecond convoy stalls of the launching thread (the interpreter's switch
     * interval is 5 ms).  A launch that blocks on a full queue waits for the device only, never for another Python thread. */
    const int rc = dispatch(fn, (int)n, a);
    return PyLong_FromLong(rc);
}

static PyMethodDef methods[] = {{"call_ints", (PyCFunction)(void (*)(void))call_ints, METH_FASTCALL,
                                 "call_ints(fn_address, args) -> int: calls an all-integer/pointer C function"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastcall", "call shim for libmeshrcnn_b200", -1, methods,
                                    NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__fastcall(void) { return PyModule_Create(&module); }
