/* latok_pypack.c -- CPython shim: list[str] -> flat UTF-8 buffer + int64 offsets, in C.
 *
 * The reference's per-string entry reads the `str` object in place (PyUnicode_READY / LENGTH / KIND / DATA,
 * latok.c:47-55); a batch engine needs the same strings as one UTF-8 buffer + offsets (include/latok_b200.h).
 * Doing that with `[t.encode() for t in texts]` + `b"".join` costs ~1 us per string in the interpreter; here the
 * loop runs in C over the list's items, chunk by chunk: sizing (fills the offsets), then encoding straight into the
 * caller's (possibly pinned) buffer.  Lone surrogates are written as three bytes, like Python's
 * 'surrogatepass' -- they are legal in the `str` the reference reads.  No tokenizer arithmetic happens here.
 *
 *   _pypack.utf8_pack(texts, offsets_addr, buf_addr, cap) -> total_bytes     offsets: int64[len(texts) + 1]
 *   _pypack.slice_tokens(texts, spans_addr, tok_offsets_addr, wide) -> list[list[str]]
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static inline Py_ssize_t utf8_len(PyObject *s)
{
    const Py_ssize_t n = PyUnicode_GET_LENGTH(s);
    if (PyUnicode_IS_COMPACT_ASCII(s)) return n;
    const int kind = PyUnicode_KIND(s);
    const void *d = PyUnicode_DATA(s);
    Py_ssize_t b = 0;
    if (kind == PyUnicode_1BYTE_KIND) {
        const uint8_t *p = (const uint8_t *)d;
        for (Py_ssize_t i = 0; i < n; ++i) b += 1 + (p[i] >> 7);
    } else if (kind == PyUnicode_2BYTE_KIND) {
        const uint16_t *p = (const uint16_t *)d;
        for (Py_ssize_t i = 0; i < n; ++i) b += 1 + (p[i] >= 0x80) + (p[i] >= 0x800);
    } else {
        const uint32_t *p = (const uint32_t *)d;
        for (Py_ssize_t i = 0; i < n; ++i) b += 1 + (p[i] >= 0x80) + (p[i] >= 0x800) + (p[i] >= 0x10000);
    }
    return b;
}

static inline uint8_t *put_cp(uint8_t *o, uint32_t c)
{
    if (c < 0x80) { *o++ = (uint8_t)c; }
    else if (c < 0x800) { *o++ = (uint8_t)(0xC0 | (c >> 6)); *o++ = (uint8_t)(0x80 | (c & 0x3F)); }
    else if (c < 0x10000) { *o++ = (uint8_t)(0xE0 | (c >> 12)); *o++ = (uint8_t)(0x80 | ((c >> 6) & 0x3F)); *o++ = (uint8_t)(0x80 | (c & 0x3F)); }
    else { *o++ = (uint8_t)(0xF0 | (c >> 18)); *o++ = (uint8_t)(0x80 | ((c >> 12) & 0x3F)); *o++ = (uint8_t)(0x80 | ((c >> 6) & 0x3F)); *o++ = (uint8_t)(0x80 | (c & 0x3F)); }
    return o;
}

static inline void utf8_write(PyObject *s, uint8_t *o)
{
    const Py_ssize_t n = PyUnicode_GET_LENGTH(s);
    const void *d = PyUnicode_DATA(s);
    if (PyUnicode_IS_COMPACT_ASCII(s)) { memcpy(o, d, (size_t)n); return; }
    const int kind = PyUnicode_KIND(s);
    if (kind == PyUnicode_1BYTE_KIND) { const uint8_t *p = (const uint8_t *)d; for (Py_ssize_t i = 0; i < n; ++i) o = put_cp(o, p[i]); }
    else if (kind == PyUnicode_2BYTE_KIND) { const uint16_t *p = (const uint16_t *)d; for (Py_ssize_t i = 0; i < n; ++i) o = put_cp(o, p[i]); }
    else { const uint32_t *p = (const uint32_t *)d; for (Py_ssize_t i = 0; i < n; ++i) o = put_cp(o, p[i]); }
}

static int get_items(PyObject *seq, PyObject ***items, Py_ssize_t *n, PyObject **fast)
{
    *fast = PySequence_Fast(seq, "texts must be a sequence of str");
    if (!*fast) return -1;
    *n = PySequence_Fast_GET_SIZE(*fast);
    *items = PySequence_Fast_ITEMS(*fast);
    return 0;
}

/* One sweep over the list in chunks of CHUNK strings: per chunk a sizing loop (fills the offsets) and, while
 * everything still fits into `cap` bytes, an encoding loop over the same -- still cached -- objects.  Returns the
 * total number of UTF-8 bytes; when that exceeds cap the buffer holds a prefix only and the caller retries. */
#define CHUNK 256
static PyObject *py_utf8_pack(PyObject *self, PyObject *args)
{
    PyObject *seq; unsigned long long off_addr, buf_addr; long long cap;
    if (!PyArg_ParseTuple(args, "OKKL", &seq, &off_addr, &buf_addr, &cap)) return NULL;
    PyObject **items, *fast; Py_ssize_t n;
    if (get_items(seq, &items, &n, &fast)) return NULL;
    int64_t *off = (int64_t *)(uintptr_t)off_addr;
    uint8_t *buf = (uint8_t *)(uintptr_t)buf_addr;
    int64_t total = 0;
    int fits = buf != NULL;
    off[0] = 0;
    for (Py_ssize_t c0 = 0; c0 < n; c0 += CHUNK) {
        const Py_ssize_t c1 = c0 + CHUNK < n ? c0 + CHUNK : n;
        for (Py_ssize_t i = c0; i < c1; ++i) {
            PyObject *s = items[i];
            if (i + 24 < n) { const char *q = (const char *)items[i + 24]; __builtin_prefetch(q); __builtin_prefetch(q + 64); __builtin_prefetch(q + 128); __builtin_prefetch(q + 192); }
            if (!PyUnicode_Check(s)) { Py_DECREF(fast); PyErr_Format(PyExc_TypeError, "texts[%zd] is not a str", i); return NULL; }
            total += utf8_len(s);
            off[i + 1] = total;
        }
        if (fits && total > cap) fits = 0;
        if (fits)
            for (Py_ssize_t i = c0; i < c1; ++i) utf8_write(items[i], buf + off[i]);
    }
    Py_DECREF(fast);
    return PyLong_FromLongLong(total);
}

/* Token texts of a batch: result[i] = [texts[i][s:e].strip() for (s, e) in spans[tok_off[i]:tok_off[i+1]] if non-empty]
 * -- the reference's per-token loop (default_tokenizer.py:151-158, 47-59 % of its CPU time) as one C loop over the span
 * array the GPU produced.  Trimming uses Python's own whitespace predicate (Py_UNICODE_ISSPACE = str.isspace). */
static PyObject *py_slice_tokens(PyObject *self, PyObject *args)
{
    PyObject *seq; unsigned long long spans_addr, off_addr; int wide;
    if (!PyArg_ParseTuple(args, "OKKi", &seq, &spans_addr, &off_addr, &wide)) return NULL;
    PyObject **items, *fast; Py_ssize_t n;
    if (get_items(seq, &items, &n, &fast)) return NULL;
    const int32_t *sp32 = (const int32_t *)(uintptr_t)spans_addr;
    const uint16_t *sp16 = (const uint16_t *)(uintptr_t)spans_addr;
    const int64_t *off = (const int64_t *)(uintptr_t)off_addr;
    PyObject *out = PyList_New(n);
    if (!out) { Py_DECREF(fast); return NULL; }
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *t = items[i];
        if (!PyUnicode_Check(t)) { PyErr_Format(PyExc_TypeError, "texts[%zd] is not a str", i); goto fail; }
        const Py_ssize_t L = PyUnicode_GET_LENGTH(t);
        const int kind = PyUnicode_KIND(t);
        const void *d = PyUnicode_DATA(t);
        const int64_t k0 = off[i], k1 = off[i + 1];
        PyObject *row = PyList_New(0);
        if (!row) goto fail;
        PyList_SET_ITEM(out, i, row);
        for (int64_t k = k0; k < k1; ++k) {
            Py_ssize_t s = wide ? sp32[2 * k] : sp16[2 * k], e = wide ? sp32[2 * k + 1] : sp16[2 * k + 1];
            if (s < 0 || e > L || s > e) { PyErr_Format(PyExc_ValueError, "span %lld of string %zd is outside the string", (long long)k, i); goto fail; }
            while (s < e && Py_UNICODE_ISSPACE(PyUnicode_READ(kind, d, s))) ++s;
            while (e > s && Py_UNICODE_ISSPACE(PyUnicode_READ(kind, d, e - 1))) --e;
            if (s == e) continue;
            PyObject *tok = PyUnicode_Substring(t, s, e);
            if (!tok || PyList_Append(row, tok) < 0) { Py_XDECREF(tok); goto fail; }
            Py_DECREF(tok);
        }
    }
    Py_DECREF(fast);
    return out;
fail:
    Py_DECREF(fast);
    Py_DECREF(out);
    return NULL;
}

static PyMethodDef methods[] = {
    {"utf8_pack", py_utf8_pack, METH_VARARGS, "utf8_pack(texts, offsets_addr, buf_addr, cap) -> total UTF-8 bytes; fills int64 offsets[len+1] and, when the total fits into cap, the buffer"},
    {"slice_tokens", py_slice_tokens, METH_VARARGS, "slice_tokens(texts, spans_addr, tok_offsets_addr, wide) -> list[list[str]] (wide: int32 spans, else uint16)"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pypack", "list[str] -> packed UTF-8 (latok_b200)", -1, methods};
PyMODINIT_FUNC PyInit__pypack(void) { return PyModule_Create(&moddef); }
