/* Host-side list assembly for the drop-in return types (CPython C API, no third-party headers).
 *
 * The reference returns nested Python lists: get_nonoverlapping_kmer_hits -> out[k] = [[start, end], ...]
 * (tg_kmer.py:173-230), one such list per read from get_tel_repeat_comp_parallel (tg_tel.py:331-389).  For 400 Mbases
 * that is ~15 million Python objects; building them from the device's span arrays with list comprehensions costs more
 * than all kernels together.  These two functions do the formatting in C:
 *
 *   hit_lists(pairs, cuts, n_slices, K) -> [[[[s, e], ...] for k in range(K)] for q in range(n_slices)]
 *       pairs: C-contiguous int32 [n_spans, 2] (start, end), ordered by (slice, k, start);
 *       cuts:  C-contiguous int64 [n_slices * K + 1], cuts[q * K + k] = first span of (slice q, k-mer k)
 *   pack_reads(reads, offsets, out, pad) -> None
 *       copies the characters of every str of `reads` (code points < 256; others become '?', like
 *       str.encode('latin-1', 'replace')) to out[offsets[i] : offsets[i] + len] and fills up to offsets[i + 1] with `pad`
 *
 * Formatting only: no part of the scan or of the alignment runs here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject *hit_lists(PyObject *self, PyObject *args)
{
    PyObject *pairs_obj, *cuts_obj;
    Py_ssize_t ns, K;
    if (!PyArg_ParseTuple(args, "OOnn", &pairs_obj, &cuts_obj, &ns, &K)) return NULL;
    Py_buffer pb, cb;
    if (PyObject_GetBuffer(pairs_obj, &pb, PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (PyObject_GetBuffer(cuts_obj, &cb, PyBUF_C_CONTIGUOUS) < 0) { PyBuffer_Release(&pb); return NULL; }
    PyObject *out = NULL;
    if (pb.itemsize != 4 || cb.itemsize != 8 || ns < 0 || K < 0 || cb.len < (Py_ssize_t)((ns * K + 1) * 8)) {
        PyErr_SetString(PyExc_ValueError, "hit_lists: pairs must be int32 [n, 2], cuts int64 [n_slices * K + 1]");
        goto done;
    }
    {
        const int32_t *pairs = (const int32_t *)pb.buf;
        const int64_t *cuts = (const int64_t *)cb.buf;
        const int64_t n_spans = (int64_t)(pb.len / 8);
        out = PyList_New(ns);
        if (!out) goto done;
        for (Py_ssize_t q = 0; q < ns; q++) {
            PyObject *per = PyList_New(K);
            if (!per) { Py_CLEAR(out); goto done; }
            PyList_SET_ITEM(out, q, per);
            for (Py_ssize_t k = 0; k < K; k++) {
                const int64_t a = cuts[q * K + k], b = cuts[q * K + k + 1];
                if (a < 0 || b < a || b > n_spans) {
                    PyErr_SetString(PyExc_ValueError, "hit_lists: cuts are not a non-decreasing partition of the spans");
                    Py_CLEAR(out);
                    goto done;
                }
                PyObject *lst = PyList_New((Py_ssize_t)(b - a));
                if (!lst) { Py_CLEAR(out); goto done; }
                PyList_SET_ITEM(per, k, lst);
                for (int64_t i = a; i < b; i++) {
                    PyObject *sp = PyList_New(2);
                    PyObject *s = PyLong_FromLong(pairs[2 * i]);
                    PyObject *e = PyLong_FromLong(pairs[2 * i + 1]);
                    if (!sp || !s || !e) { Py_XDECREF(sp); Py_XDECREF(s); Py_XDECREF(e); Py_CLEAR(out); goto done; }
                    PyList_SET_ITEM(sp, 0, s);
                    PyList_SET_ITEM(sp, 1, e);
                    PyList_SET_ITEM(lst, (Py_ssize_t)(i - a), sp);
                }
            }
        }
    }
done:
    PyBuffer_Release(&pb);
    PyBuffer_Release(&cb);
    return out;
}

static PyObject *pack_reads(PyObject *self, PyObject *args)
{
    PyObject *reads, *off_obj, *out_obj;
    int pad;
    if (!PyArg_ParseTuple(args, "OOOi", &reads, &off_obj, &out_obj, &pad)) return NULL;
    if (!PyList_Check(reads) && !PyTuple_Check(reads)) { PyErr_SetString(PyExc_TypeError, "pack_reads: reads must be a list or tuple of str"); return NULL; }
    Py_buffer ob, wb;
    if (PyObject_GetBuffer(off_obj, &ob, PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (PyObject_GetBuffer(out_obj, &wb, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) { PyBuffer_Release(&ob); return NULL; }
    PyObject *ret = NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(reads);
    if (ob.itemsize != 8 || ob.len < (n + 1) * 8 || wb.itemsize != 1) {
        PyErr_SetString(PyExc_ValueError, "pack_reads: offsets must be int64 [n + 1], out a writable byte buffer");
        goto done;
    }
    {
        const int64_t *off = (const int64_t *)ob.buf;
        uint8_t *out = (uint8_t *)wb.buf;
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *s = PySequence_Fast_GET_ITEM(reads, i);
            if (!PyUnicode_Check(s)) { PyErr_SetString(PyExc_TypeError, "pack_reads: reads must be str"); goto done; }
            const Py_ssize_t len = PyUnicode_GET_LENGTH(s);
            if (off[i] < 0 || off[i + 1] < off[i] + len || off[i + 1] > wb.len) {
                PyErr_SetString(PyExc_ValueError, "pack_reads: offsets do not leave room for the read");
                goto done;
            }
            uint8_t *dst = out + off[i];
            const int kind = PyUnicode_KIND(s);
            if (kind == PyUnicode_1BYTE_KIND) memcpy(dst, PyUnicode_1BYTE_DATA(s), (size_t)len);
            else {
                const void *data = PyUnicode_DATA(s);
                for (Py_ssize_t j = 0; j < len; j++) {
                    const Py_UCS4 c = PyUnicode_READ(kind, data, j);
                    dst[j] = c < 256 ? (uint8_t)c : (uint8_t)'?';
                }
            }
            memset(dst + len, pad, (size_t)(off[i + 1] - off[i] - len));
        }
        ret = Py_None;
        Py_INCREF(ret);
    }
done:
    PyBuffer_Release(&ob);
    PyBuffer_Release(&wb);
    return ret;
}

/* linkage_label(Z, n): scipy _hierarchy.label -- Z float64 C-contiguous [n - 1, 4], rows sorted by distance, columns 0/1 hold
 * original cluster indices of the merged pair; rewrites them as union-find cluster ids (smaller first) and fills column 3 with
 * the size of the new cluster. */
static PyObject *linkage_label(PyObject *self, PyObject *args)
{
    PyObject *z_obj;
    Py_ssize_t n;
    if (!PyArg_ParseTuple(args, "On", &z_obj, &n)) return NULL;
    Py_buffer zb;
    if (PyObject_GetBuffer(z_obj, &zb, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) return NULL;
    PyObject *ret = NULL;
    int64_t *parent = NULL, *size = NULL;
    if (zb.itemsize != 8 || n < 1 || zb.len != (Py_ssize_t)((n - 1) * 4 * 8)) {
        PyErr_SetString(PyExc_ValueError, "linkage_label: Z must be a writable float64 [n - 1, 4] array");
        goto done;
    }
    parent = (int64_t *)malloc((size_t)(2 * n) * sizeof(int64_t));
    size = (int64_t *)malloc((size_t)(2 * n) * sizeof(int64_t));
    if (!parent || !size) { PyErr_NoMemory(); goto done; }
    for (Py_ssize_t i = 0; i < 2 * n - 1; i++) { parent[i] = i; size[i] = 1; }
    {
        double *Z = (double *)zb.buf;
        int64_t next = n;
        for (Py_ssize_t i = 0; i < n - 1; i++) {
            int64_t r[2];
            for (int s = 0; s < 2; s++) {
                int64_t x = (int64_t)Z[4 * i + s];
                if (x < 0 || x >= n) { PyErr_SetString(PyExc_ValueError, "linkage_label: cluster index out of range"); goto done; }
                int64_t p = x;
                while (parent[x] != x) x = parent[x];
                while (parent[p] != x) { const int64_t q = parent[p]; parent[p] = x; p = q; }
                r[s] = x;
            }
            Z[4 * i] = (double)(r[0] < r[1] ? r[0] : r[1]);
            Z[4 * i + 1] = (double)(r[0] < r[1] ? r[1] : r[0]);
            parent[r[0]] = next;
            parent[r[1]] = next;
            size[next] = size[r[0]] + size[r[1]];
            Z[4 * i + 3] = (double)size[next];
            next++;
        }
    }
    ret = Py_None;
    Py_INCREF(ret);
done:
    free(parent);
    free(size);
    PyBuffer_Release(&zb);
    return ret;
}

static PyMethodDef methods[] = {
    {"linkage_label", linkage_label, METH_VARARGS, "scipy's union-find relabelling of a sorted merge list, in place"},
    {"hit_lists", hit_lists, METH_VARARGS, "nested [start, end] lists per (slice, k-mer) from sorted span arrays"},
    {"pack_reads", pack_reads, METH_VARARGS, "copy the characters of a list of str into a padded byte buffer"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_tg_hostlists", "host-side list assembly (formatting only)", -1, methods};

PyMODINIT_FUNC PyInit__tg_hostlists(void) { return PyModule_Create(&moddef); }
