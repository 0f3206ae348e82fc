/*
 * pack_ext.c -- CPython helper for the host marshalling at the boundary ("next" row):
 * list[str] -> one contiguous UTF-8 buffer + uint64 offsets, without creating a bytes object per
 * document.  The reference gets this for free because PyO3 borrows CPython's cached UTF-8
 * (`text: &str`, src/py.rs:30); here the batch has to be contiguous for ONE native call.
 * Pure host code (no CUDA); built by __graft_entry__.build() with gcc into tiktoken_b200/_b200pack*.so.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* pack(seq_of_str) -> (blob: bytes, offsets: bytes holding uint64[n+1]) */
static PyObject *pack(PyObject *self, PyObject *arg) {
    (void)self;
    PyObject *seq = PySequence_Fast(arg, "expected a sequence of str");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject **items = PySequence_Fast_ITEMS(seq);
    PyObject *offs = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)((n + 1) * sizeof(uint64_t)));
    if (!offs) { Py_DECREF(seq); return NULL; }
    uint64_t *off = (uint64_t *)PyBytes_AS_STRING(offs);
    uint64_t total = 0;
    for (Py_ssize_t i = 0; i < n; i++) {
        Py_ssize_t len;
        if (!PyUnicode_Check(items[i])) {
            PyErr_SetString(PyExc_TypeError, "expected str");
            Py_DECREF(offs); Py_DECREF(seq); return NULL;
        }
        /* raises UnicodeEncodeError on lone surrogates, like str.encode("utf-8") */
        if (!PyUnicode_AsUTF8AndSize(items[i], &len)) { Py_DECREF(offs); Py_DECREF(seq); return NULL; }
        off[i] = total;
        total += (uint64_t)len;
    }
    off[n] = total;
    PyObject *blob = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)total);
    if (!blob) { Py_DECREF(offs); Py_DECREF(seq); return NULL; }
    char *dst = PyBytes_AS_STRING(blob);
    for (Py_ssize_t i = 0; i < n; i++) {
        Py_ssize_t len;
        const char *src = PyUnicode_AsUTF8AndSize(items[i], &len);   /* cached by the first pass */
        memcpy(dst + off[i], src, (size_t)len);
    }
    Py_DECREF(seq);
    return Py_BuildValue("(NN)", blob, offs);
}

/* unpack(tokens_addr: int, offsets_addr: int, n_docs: int) -> list[list[int]] */
static PyObject *unpack(PyObject *self, PyObject *args) {
    (void)self;
    unsigned long long ta, oa; Py_ssize_t n;
    if (!PyArg_ParseTuple(args, "KKn", &ta, &oa, &n)) return NULL;
    const uint32_t *tok = (const uint32_t *)(uintptr_t)ta;
    const uint64_t *off = (const uint64_t *)(uintptr_t)oa;
    PyObject *out = PyList_New(n);
    if (!out) return NULL;
    for (Py_ssize_t d = 0; d < n; d++) {
        const uint64_t lo = off[d], hi = off[d + 1];
        PyObject *doc = PyList_New((Py_ssize_t)(hi - lo));
        if (!doc) { Py_DECREF(out); return NULL; }
        for (uint64_t k = lo; k < hi; k++) {
            PyObject *v = PyLong_FromUnsignedLong(tok[k]);
            if (!v) { Py_DECREF(doc); Py_DECREF(out); return NULL; }
            PyList_SET_ITEM(doc, (Py_ssize_t)(k - lo), v);
        }
        PyList_SET_ITEM(out, d, doc);
    }
    return out;
}

static PyMethodDef methods[] = {
    {"pack", pack, METH_O, "list[str] -> (utf8 blob bytes, uint64 offsets bytes)"},
    {"unpack", unpack, METH_VARARGS, "(tokens_addr, offsets_addr, n_docs) -> list[list[int]]"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef mod = {PyModuleDef_HEAD_INIT, "_b200pack", "host marshalling helpers", -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__b200pack(void) { return PyModule_Create(&mod); }
