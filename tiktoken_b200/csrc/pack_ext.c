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


/* parse_tiktoken(data: bytes) -> (blob: bytes, offsets: bytes uint64[n+1], ranks: bytes uint32[n])
 * The `.tiktoken` vocabulary format of tiktoken/load.py:159-171 -- one "base64(token) rank" line per token --
 * straight into the flattened arrays b200bpe_create takes, without a Python dict of 100-200 k bytes objects. */
static PyObject *parse_tiktoken(PyObject *self, PyObject *arg) {
    (void)self;
    Py_buffer in;
    if (PyObject_GetBuffer(arg, &in, PyBUF_SIMPLE) < 0) return NULL;
    const unsigned char *p = (const unsigned char *)in.buf, *end = p + in.len;
    signed char dec[256];
    memset(dec, -1, sizeof dec);
    for (int i = 0; i < 26; i++) { dec['A' + i] = (signed char)i; dec['a' + i] = (signed char)(26 + i); }
    for (int i = 0; i < 10; i++) dec['0' + i] = (signed char)(52 + i);
    dec['+'] = 62; dec['/'] = 63;
    size_t n_lines = 0;
    for (const unsigned char *q = p; q < end; q++) n_lines += (*q == '\n');
    n_lines += 1;
    unsigned char *blob = (unsigned char *)PyMem_Malloc((size_t)in.len + 4);       /* decoded bytes < input bytes */
    uint64_t *off = (uint64_t *)PyMem_Malloc((n_lines + 1) * sizeof(uint64_t));
    uint32_t *rank = (uint32_t *)PyMem_Malloc((n_lines + 1) * sizeof(uint32_t));
    PyObject *ret = NULL;
    if (!blob || !off || !rank) { PyErr_NoMemory(); goto done; }
    {
        size_t n = 0, nb = 0;
        while (p < end) {
            const unsigned char *eol = (const unsigned char *)memchr(p, '\n', (size_t)(end - p));
            if (!eol) eol = end;
            const unsigned char *le = eol;
            if (le > p && le[-1] == '\r') le--;
            if (le == p) { p = eol + 1; continue; }                                   /* empty line (load.py:165) */
            const unsigned char *sp = (const unsigned char *)memchr(p, ' ', (size_t)(le - p));
            if (!sp || sp == p || sp + 1 >= le) { PyErr_Format(PyExc_ValueError, "malformed line %zu", n + 1); goto done; }
            off[n] = nb;
            uint32_t acc = 0; int bits = 0;
            const unsigned char *q = p;
            for (; q < sp && *q != '='; q++) {
                const int v = dec[*q];
                if (v < 0) { PyErr_Format(PyExc_ValueError, "bad base64 on line %zu", n + 1); goto done; }
                acc = (acc << 6) | (uint32_t)v; bits += 6;
                if (bits >= 8) { bits -= 8; blob[nb++] = (unsigned char)(acc >> bits); acc &= (1u << bits) - 1u; }
            }
            for (; q < sp; q++) if (*q != '=') { PyErr_Format(PyExc_ValueError, "bad base64 padding on line %zu", n + 1); goto done; }
            if (nb == off[n]) { PyErr_Format(PyExc_ValueError, "empty token on line %zu", n + 1); goto done; }
            uint64_t r = 0;
            for (q = sp + 1; q < le; q++) {
                if (*q < '0' || *q > '9') { PyErr_Format(PyExc_ValueError, "bad rank on line %zu", n + 1); goto done; }
                r = r * 10 + (uint64_t)(*q - '0');
                if (r > 0xFFFFFFFFull) { PyErr_Format(PyExc_ValueError, "rank too large on line %zu", n + 1); goto done; }
            }
            rank[n++] = (uint32_t)r;
            p = eol + 1;
        }
        off[n] = nb;
        ret = Py_BuildValue("(y#y#y#)", (const char *)blob, (Py_ssize_t)nb, (const char *)off, (Py_ssize_t)((n + 1) * sizeof(uint64_t)),
                            (const char *)rank, (Py_ssize_t)(n * sizeof(uint32_t)));
    }
done:
    PyMem_Free(blob); PyMem_Free(off); PyMem_Free(rank);
    PyBuffer_Release(&in);
    return ret;
}

static PyMethodDef methods[] = {
    {"pack", pack, METH_O, "list[str] -> (utf8 blob bytes, uint64 offsets bytes)"},
    {"unpack", unpack, METH_VARARGS, "(tokens_addr, offsets_addr, n_docs) -> list[list[int]]"},
    {"parse_tiktoken", parse_tiktoken, METH_O, "(.tiktoken file bytes) -> (blob, offsets uint64, ranks uint32)"},
    {"find_first", find_first, METH_VARARGS, "(blob, offsets, needles) -> None | (doc, needle_index, byte_pos)"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef mod = {PyModuleDef_HEAD_INIT, "_b200pack", "host marshalling helpers", -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__b200pack(void) { return PyModule_Create(&mod); }
