/* fastconv.c — CPython helper for the host half of the coverage update.
 *
 * `CoverageConverter.convert_records` (boss/runs/sequences.py:678-739) walks Python objects: a dict of
 * PafLine lists and a dict of read strings. Everything numeric about the batch happens on the GPU, but picking
 * each read's record, computing the aligned slice bounds and taking the addresses of the CIGAR / read strings
 * has to touch those objects; done in Python that loop is the largest item of the end-to-end update. This module
 * does the same walk through the C API and fills caller-provided arrays. No CUDA, no NumPy C API (plain buffer
 * protocol). Semantics are those of boss_runs_b200/runs.py:CoverageConverter._convert_records_py, which the tests
 * compare it against.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *s_tname, *s_tstart, *s_tend, *s_barcode, *s_rev, *s_cigar, *s_qname, *s_qlen, *s_qstart, *s_qend;

static int attr_i64(PyObject* o, PyObject* name, long long* out) {
    PyObject* v = PyObject_GetAttr(o, name);
    if (!v) return -1;
    long long x = PyLong_AsLongLong(v);
    Py_DECREF(v);
    if (x == -1 && PyErr_Occurred()) return -1;
    *out = x;
    return 0;
}

static long long clampll(long long x, long long lo, long long hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* convert(paf_dict, seqs, contig_index, best_record, bufs, keep) -> (n_used, n_skipped)
 * bufs: tuple of 10 writable buffers sized for len(paf_dict) entries:
 *   contig i32, tstart i64, tend i64, barcode i32, rev u8, cigar_ptr u64, cigar_len i64, seq_ptr u64, seq_from i64, seq_to i64
 * keep: list that receives the str objects the pointers refer to */
static PyObject* convert(PyObject* self, PyObject* args) {
    PyObject *paf_dict, *seqs, *contig_index, *best_record, *bufs, *keep;
    if (!PyArg_ParseTuple(args, "O!O!O!OO!O!", &PyDict_Type, &paf_dict, &PyDict_Type, &seqs, &PyDict_Type, &contig_index,
                          &best_record, &PyTuple_Type, &bufs, &PyList_Type, &keep))
        return NULL;
    if (PyTuple_GET_SIZE(bufs) != 10) { PyErr_SetString(PyExc_ValueError, "expected 10 output buffers"); return NULL; }
    Py_buffer vb[10];
    static const Py_ssize_t item[10] = {4, 8, 8, 4, 1, 8, 8, 8, 8, 8};
    const Py_ssize_t n_max = PyDict_Size(paf_dict);
    int got = 0;
    for (; got < 10; ++got) {
        if (PyObject_GetBuffer(PyTuple_GET_ITEM(bufs, got), &vb[got], PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) goto fail;
        if (vb[got].len < n_max * item[got]) {
            ++got;
            PyErr_SetString(PyExc_ValueError, "output buffer too small");
            goto fail;
        }
    }
    {
        int32_t* o_contig = (int32_t*)vb[0].buf;   int64_t* o_tstart = (int64_t*)vb[1].buf;  int64_t* o_tend = (int64_t*)vb[2].buf;
        int32_t* o_bc = (int32_t*)vb[3].buf;       uint8_t* o_rev = (uint8_t*)vb[4].buf;     uint64_t* o_cp = (uint64_t*)vb[5].buf;
        int64_t* o_cl = (int64_t*)vb[6].buf;       uint64_t* o_sp = (uint64_t*)vb[7].buf;    int64_t* o_sf = (int64_t*)vb[8].buf;
        int64_t* o_st = (int64_t*)vb[9].buf;
        Py_ssize_t pos = 0, n = 0, skipped = 0;
        PyObject *key, *recs;
        while (PyDict_Next(paf_dict, &pos, &key, &recs)) {
            PyObject* rec;
            int owned = 0;
            if (PyList_Check(recs) && PyList_GET_SIZE(recs) == 1) {
                rec = PyList_GET_ITEM(recs, 0);
            } else {
                rec = PyObject_CallOneArg(best_record, recs);          /* Paf.choose_best_mapper, boss/paf.py:710-722 */
                if (!rec) goto fail;
                owned = 1;
            }
            PyObject* tname = PyObject_GetAttr(rec, s_tname);
            if (!tname) { if (owned) Py_DECREF(rec); goto fail; }
            PyObject* k = PyDict_GetItemWithError(contig_index, tname);   /* borrowed */
            Py_DECREF(tname);
            if (!k) {
                if (owned) Py_DECREF(rec);
                if (PyErr_Occurred()) goto fail;
                ++skipped;                       /* upstream files these under a key nobody reads (core.py:83-86) */
                continue;
            }
            long long ki = PyLong_AsLongLong(k), tstart, tend, qlen, qstart, qend;
            PyObject* qname = PyObject_GetAttr(rec, s_qname);
            PyObject* revo = qname ? PyObject_GetAttr(rec, s_rev) : NULL;
            PyObject* cig = revo ? PyObject_GetAttr(rec, s_cigar) : NULL;
            PyObject* bco = cig ? PyObject_GetAttr(rec, s_barcode) : NULL;
            int ok = bco && attr_i64(rec, s_tstart, &tstart) == 0 && attr_i64(rec, s_tend, &tend) == 0 &&
                     attr_i64(rec, s_qlen, &qlen) == 0 && attr_i64(rec, s_qstart, &qstart) == 0 && attr_i64(rec, s_qend, &qend) == 0;
            PyObject* s = NULL;
            if (ok) {
                s = PyDict_GetItemWithError(seqs, qname);                 /* borrowed; KeyError like seqs[rec.qname] */
                if (!s) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, qname); ok = 0; }
            }
            int rev = 0;
            if (ok) { rev = PyObject_IsTrue(revo); if (rev < 0) ok = 0; }
            if (ok && cig == Py_None) { PyErr_SetString(PyExc_AssertionError, "record without a cg:Z: CIGAR"); ok = 0; }   /* sequences.py:718 */
            Py_ssize_t clen = 0, slen = 0;
            const char *cptr = NULL, *sptr = NULL;
            if (ok) {
                cptr = PyUnicode_AsUTF8AndSize(cig, &clen);
                sptr = cptr ? PyUnicode_AsUTF8AndSize(s, &slen) : NULL;
                if (!cptr || !sptr) ok = 0;
                else if (slen != PyUnicode_GET_LENGTH(s) || clen != PyUnicode_GET_LENGTH(cig)) {
                    PyErr_SetString(PyExc_ValueError, "read and CIGAR strings must be ASCII");
                    ok = 0;
                }
            }
            long long bc = 0;
            if (ok && bco != Py_None) { bc = PyLong_AsLongLong(bco); if (bc == -1 && PyErr_Occurred()) ok = 0; }
            if (ok) {
                long long lo, hi;
                const long long len = (long long)slen;
                if (rev) {
                    /* upstream slices the reverse complement of the WHOLE string with qlen-based coordinates
                     * (sequences.py:707-711; Q12): rc[a:b] == revcomp(s[n-b:n-a]) with Python's slice clamping */
                    const long long a = qlen - qend, b = qlen - qstart;
                    lo = len - (b < len ? b : len); if (lo < 0) lo = 0;
                    hi = len - (a < len ? a : len); if (hi < 0) hi = 0;
                    if (a < 0 || b < 0) { lo = 0; hi = 0; PyErr_SetString(PyExc_ValueError, "negative query coordinates"); ok = 0; }
                } else {
                    lo = clampll(qstart, 0, len);
                    hi = clampll(qend, 0, len);
                    if (qstart < 0 || qend < 0) { PyErr_SetString(PyExc_ValueError, "negative query coordinates"); ok = 0; }
                }
                if (ok) {
                    if (hi < lo) hi = lo;
                    o_contig[n] = (int32_t)ki; o_tstart[n] = tstart; o_tend[n] = tend; o_bc[n] = (int32_t)bc; o_rev[n] = (uint8_t)rev;
                    o_cp[n] = (uint64_t)(uintptr_t)cptr; o_cl[n] = (int64_t)clen;
                    o_sp[n] = (uint64_t)(uintptr_t)sptr; o_sf[n] = lo; o_st[n] = hi;
                    if (PyList_Append(keep, cig) < 0 || PyList_Append(keep, s) < 0) ok = 0;
                    else ++n;
                }
            }
            Py_XDECREF(qname); Py_XDECREF(revo); Py_XDECREF(cig); Py_XDECREF(bco);
            if (owned) Py_DECREF(rec);
            if (!ok) goto fail;
        }
        for (int i = 0; i < 10; ++i) PyBuffer_Release(&vb[i]);
        return Py_BuildValue("nn", n, skipped);
    }
fail:
    for (int i = 0; i < got; ++i) PyBuffer_Release(&vb[i]);
    return NULL;
}

static PyMethodDef methods[] = {
    {"convert", convert, METH_VARARGS, "convert(paf_dict, seqs, contig_index, best_record, bufs, keep) -> (n_used, n_skipped)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_fastconv", "host half of the coverage update (C API walk)", -1, methods};

PyMODINIT_FUNC PyInit__fastconv(void) {
    s_tname = PyUnicode_InternFromString("tname");   s_tstart = PyUnicode_InternFromString("tstart");
    s_tend = PyUnicode_InternFromString("tend");     s_barcode = PyUnicode_InternFromString("barcode");
    s_rev = PyUnicode_InternFromString("rev");       s_cigar = PyUnicode_InternFromString("cigar");
    s_qname = PyUnicode_InternFromString("qname");   s_qlen = PyUnicode_InternFromString("qlen");
    s_qstart = PyUnicode_InternFromString("qstart"); s_qend = PyUnicode_InternFromString("qend");
    return PyModule_Create(&moduledef);
}
