/*
 * pack_ext.c -- CPython helper for the host marshalling at the boundary ("next" row):
 * list[str] -> one contiguous UTF-8 buffer + uint64 offsets, without creating a bytes object per
 * document.  The reference gets this for free because PyO3 borrows CPython's cached UTF-8
 * (`text: &str`, src/py.rs:30); here the batch has to be contiguous for ONE native call.
 * Pure host code (no CUDA); built by __graft_entry__.build() with gcc into tiktoken_b200/_b200pack*.so.
 */
#define _GNU_SOURCE
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


/* find_first(blob, offsets, needles) -> None | (doc, needle_index, byte_pos)
 * The disallowed-special check of tiktoken/core.py:120-124 on the PACKED batch: the first document (in order)
 * that contains any of the byte strings `needles`, and the leftmost occurrence in it (longest on ties).  A match
 * never straddles two documents.  One memchr pass when all needles share their first byte (they do for the
 * published encodings: "<|"), else one memmem pass per needle. */
static PyObject *find_first(PyObject *self, PyObject *args) {
    (void)self;
    Py_buffer blob, offs;
    PyObject *needles_obj;
    if (!PyArg_ParseTuple(args, "y*y*O", &blob, &offs, &needles_obj)) return NULL;
    PyObject *seq = PySequence_Fast(needles_obj, "needles must be a sequence of bytes");
    if (!seq) { PyBuffer_Release(&blob); PyBuffer_Release(&offs); return NULL; }
    const Py_ssize_t nn = PySequence_Fast_GET_SIZE(seq);
    const char **np_ = (const char **)PyMem_Malloc((size_t)(nn + 1) * sizeof(char *));
    Py_ssize_t *nl = (Py_ssize_t *)PyMem_Malloc((size_t)(nn + 1) * sizeof(Py_ssize_t));
    PyObject *ret = NULL;
    int same_first = 1;
    for (Py_ssize_t a = 0; a < nn; a++) {
        char *p; Py_ssize_t l;
        if (PyBytes_AsStringAndSize(PySequence_Fast_GET_ITEM(seq, a), &p, &l) < 0) goto done;
        np_[a] = p; nl[a] = l;
        if (l == 0) { PyErr_SetString(PyExc_ValueError, "empty needle"); goto done; }
        if (p[0] != np_[0][0]) same_first = 0;
    }
    {
        const uint8_t *text = (const uint8_t *)blob.buf;
        const uint64_t *off = (const uint64_t *)offs.buf;
        const Py_ssize_t n_docs = offs.len / 8 - 1;
        for (Py_ssize_t d = 0; d < n_docs; d++)
            if (off[d] > off[d + 1] || off[d + 1] > (uint64_t)blob.len) {
                PyErr_SetString(PyExc_ValueError, "offsets must be non-decreasing and end inside the blob");
                goto done;
            }
        Py_ssize_t hit_doc = -1, hit_a = -1; uint64_t hit_pos = 0;
        Py_BEGIN_ALLOW_THREADS
        for (Py_ssize_t d = 0; d < n_docs && hit_doc < 0 && nn > 0; d++) {
            const uint8_t *s = text + off[d], *e = text + off[d + 1];
            const uint8_t *best = NULL; Py_ssize_t best_a = -1;
            if (same_first) {
                const uint8_t *p = s;
                while (!best && p < e && (p = (const uint8_t *)memchr(p, np_[0][0], (size_t)(e - p))) != NULL) {
                    for (Py_ssize_t a = 0; a < nn; a++)
                        if (nl[a] <= e - p && memcmp(p, np_[a], (size_t)nl[a]) == 0 && (best_a < 0 || nl[a] > nl[best_a])) { best = p; best_a = a; }
                    p++;
                }
            } else {
                for (Py_ssize_t a = 0; a < nn; a++) {
                    if (nl[a] > e - s) continue;
                    const uint8_t *f = (const uint8_t *)memmem(s, (size_t)(e - s), np_[a], (size_t)nl[a]);
                    if (f && (!best || f < best || (f == best && nl[a] > nl[best_a]))) { best = f; best_a = a; }
                }
            }
            if (best) { hit_doc = d; hit_a = best_a; hit_pos = (uint64_t)(best - text); }
        }
        Py_END_ALLOW_THREADS
        if (hit_doc < 0) { ret = Py_None; Py_INCREF(ret); }
        else ret = Py_BuildValue("(nnK)", hit_doc, hit_a, (unsigned long long)hit_pos);
    }
done:
    PyMem_Free(np_); PyMem_Free(nl);
    Py_DECREF(seq); PyBuffer_Release(&blob); PyBuffer_Release(&offs);
    return ret;
}

static PyMethodDef methods[] = {
    {"pack", pack, METH_O, "list[str] -> (utf8 blob bytes, uint64 offsets bytes)"},
    {"unpack", unpack, METH_VARARGS, "(tokens_addr, offsets_addr, n_docs) -> list[list[int]]"},
    {"find_first", find_first, METH_VARARGS, "(blob, offsets, needles) -> None | (doc, needle_index, byte_pos)"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef mod = {PyModuleDef_HEAD_INIT, "_b200pack", "host marshalling helpers", -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__b200pack(void) { return PyModule_Create(&mod); }
