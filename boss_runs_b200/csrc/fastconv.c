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
#include <string.h>

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

/* convert(paf_dict, seqs, contig_index, best_record, bufs, keep[, route]) -> (n_used, n_skipped[, n_tracked])
 * bufs: tuple of 10 writable buffers sized for len(paf_dict) entries:
 *   contig i32, tstart i64, tend i64, barcode i32, rev u8, cigar_ptr u64, cigar_len i64, seq_ptr u64, seq_from i64, seq_to i64
 * keep: list that receives the str objects the pointers refer to
 * route (one process per GPU, each holding a range of the genome): tuple (ranges i64[n_contigs][2], contig_all i32, tstart_all i64,
 *   tend_all i64, rev_all u8). Every read on a tracked contig is listed in the *_all buffers (depth totals and read starts need
 *   the whole batch); only reads whose interval overlaps [ranges[k][0], ranges[k][1]) of their contig k are converted. */
static PyObject* convert(PyObject* self, PyObject* args) {
    PyObject *paf_dict, *seqs, *contig_index, *best_record, *bufs, *keep, *route = NULL;
    if (!PyArg_ParseTuple(args, "O!O!O!OO!O!|O!", &PyDict_Type, &paf_dict, &PyDict_Type, &seqs, &PyDict_Type, &contig_index,
                          &best_record, &PyTuple_Type, &bufs, &PyList_Type, &keep, &PyTuple_Type, &route))
        return NULL;
    if (PyTuple_GET_SIZE(bufs) != 10) { PyErr_SetString(PyExc_ValueError, "expected 10 output buffers"); return NULL; }
    if (route && PyTuple_GET_SIZE(route) != 5) { PyErr_SetString(PyExc_ValueError, "route: expected 5 buffers"); return NULL; }
    Py_buffer vb[10], rb[5];
    static const Py_ssize_t item[10] = {4, 8, 8, 4, 1, 8, 8, 8, 8, 8};
    static const Py_ssize_t ritem[5] = {0, 4, 8, 8, 1};
    const Py_ssize_t n_max = PyDict_Size(paf_dict);
    int got = 0, rgot = 0;
    if (route) {
        for (; rgot < 5; ++rgot) {
            if (PyObject_GetBuffer(PyTuple_GET_ITEM(route, rgot), &rb[rgot], (rgot ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS) < 0) goto fail;
            if (rb[rgot].len < n_max * ritem[rgot]) {
                ++rgot;
                PyErr_SetString(PyExc_ValueError, "route buffer too small");
                goto fail;
            }
        }
    }
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
        Py_ssize_t pos = 0, n = 0, skipped = 0, m = 0;
        const int64_t* ranges = route ? (const int64_t*)rb[0].buf : NULL;
        const Py_ssize_t n_ranges = route ? rb[0].len / 16 : 0;
        int32_t* a_contig = route ? (int32_t*)rb[1].buf : NULL;  int64_t* a_tstart = route ? (int64_t*)rb[2].buf : NULL;
        int64_t* a_tend = route ? (int64_t*)rb[3].buf : NULL;    uint8_t* a_rev = route ? (uint8_t*)rb[4].buf : NULL;
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
            if (route) {
                /* routed batch: list the read, convert it only if this process' range of its contig sees it */
                PyObject* rv = PyObject_GetAttr(rec, s_rev);
                int r01 = rv ? PyObject_IsTrue(rv) : -1;
                Py_XDECREF(rv);
                if (r01 < 0 || attr_i64(rec, s_tstart, &tstart) < 0 || attr_i64(rec, s_tend, &tend) < 0 || ki < 0 || ki >= n_ranges) {
                    if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "contig index outside the routing table");
                    if (owned) Py_DECREF(rec);
                    goto fail;
                }
                a_contig[m] = (int32_t)ki; a_tstart[m] = tstart; a_tend[m] = tend; a_rev[m] = (uint8_t)r01; ++m;
                const long long t0 = tstart < tend ? tstart : tend, t1 = tstart < tend ? tend : tstart;
                if (t1 <= ranges[2 * ki] || t0 >= ranges[2 * ki + 1]) {
                    if (owned) Py_DECREF(rec);
                    continue;
                }
            }
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
            if (ok && cig == Py_None && qstart >= 0 && qend >= 0 && (!rev || (qlen - qend >= 0 && qlen - qstart >= 0))) {
                PyErr_SetString(PyExc_AssertionError, "record without a cg:Z: CIGAR");   /* sequences.py:718, after the slicing */
                ok = 0;
            }
            Py_ssize_t clen = 0, slen = 0;
            const char *cptr = NULL, *sptr = NULL;
            if (ok) {
                sptr = PyUnicode_AsUTF8AndSize(s, &slen);
                cptr = (sptr && cig != Py_None) ? PyUnicode_AsUTF8AndSize(cig, &clen) : "";
                if (!cptr || !sptr) ok = 0;
                else if (slen != PyUnicode_GET_LENGTH(s) || (cig != Py_None && clen != PyUnicode_GET_LENGTH(cig))) {
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
        for (int i = 0; i < rgot; ++i) PyBuffer_Release(&rb[i]);
        return route ? Py_BuildValue("nnn", n, skipped, m) : Py_BuildValue("nn", n, skipped);
    }
fail:
    for (int i = 0; i < got; ++i) PyBuffer_Release(&vb[i]);
    for (int i = 0; i < rgot; ++i) PyBuffer_Release(&rb[i]);
    return NULL;
}


/* ------------------------------------------------------------------------------------------------------------------
 * convert_text — the same batch straight from the mapper's PAF TEXT.
 *
 * Upstream goes  paf_raw (str) -> Paf.parse_PAF(StringIO(paf_raw), min_len=mu/2) -> {qname: [PafLine]} -> convert_records
 * (boss/mapper.py:63-65, boss/paf.py:18-74,653-672, boss/runs/sequences.py:694-738): one Python object with ~25
 * attributes per alignment. Here the text is tokenised in place and only the winning record of every read is turned
 * into the ten numbers the device needs; the CIGAR pointers point into the text itself. Semantics followed:
 *   - lines are split on '\n', stripped (str.strip) and split on tabs; fewer than 12 columns -> IndexError (paf.py:47-52)
 *   - every core column goes through conv_type(x, int): an all-digit query/target NAME is an int upstream and becomes its
 *     canonical decimal text again under str() ("007" -> "7", paf.py:54-56)
 *   - rev = (strand != '+') (paf.py:58); tags are split on ':' into exactly three parts (ValueError otherwise), the type
 *     letter must be one of i A f Z (KeyError otherwise), a repeated key keeps its last value (paf.py:87-99)
 *   - AS defaults to 0, primary <=> tp value == "P"; records with alignment_block_length < min_len or not primary are
 *     dropped AFTER the whole line has been parsed (paf.py:664-669)
 *   - reads keep their order of first appearance; with several records the winner is the last element of
 *     np.argsort by (mapq, AS) (paf.py:710-722): for up to 16 candidates NumPy's sort is an insertion sort, i.e. stable,
 *     so ties go to the LATER record; larger groups are handed to `best_index` (NumPy itself) to stay exact.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { const char* p; Py_ssize_t n; } span_t;

typedef struct {
    span_t tname, cigar;
    long long qlen, qstart, qend, tstart, tend, mapq, as;
    int rev, has_cigar;
    unsigned bad;             /* 1: a coordinate column is not an integer, 2: mapq is not */
    Py_ssize_t next;          /* next record of the same read, -1 at the end */
} rec_t;

typedef struct { Py_ssize_t first, last, count; PyObject* qname; } grp_t;

static int is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

static span_t strip_span(span_t s) {
    while (s.n > 0 && is_space((unsigned char)s.p[0])) { ++s.p; --s.n; }
    while (s.n > 0 && is_space((unsigned char)s.p[s.n - 1])) --s.n;
    return s;
}

/* Python's int(str) for the forms a PAF can hold: optional surrounding whitespace, optional sign, decimal digits.
 * 0 = parsed, -1 = not an integer (upstream would keep the string). Values beyond 64 bits count as "not an integer". */
static int parse_int(span_t s, long long* out) {
    s = strip_span(s);
    if (s.n == 0) return -1;
    int neg = 0;
    Py_ssize_t i = 0;
    if (s.p[0] == '+' || s.p[0] == '-') { neg = s.p[0] == '-'; i = 1; }
    if (i >= s.n) return -1;
    unsigned long long v = 0;
    int prev_digit = 0;
    for (; i < s.n; ++i) {
        if (s.p[i] == '_') {                     /* PEP 515: single underscores between digits ("1_0" == 10) */
            if (!prev_digit || i + 1 >= s.n) return -1;
            prev_digit = 0;
            continue;
        }
        unsigned c = (unsigned char)s.p[i] - '0';
        if (c > 9) return -1;
        if (v > (unsigned long long)(INT64_MAX / 10 - 1)) return -1;
        v = v * 10 + c;
        prev_digit = 1;
    }
    if (!prev_digit) return -1;
    *out = neg ? -(long long)v : (long long)v;
    return 0;
}

/* str(conv_type(x, int)) */
static PyObject* name_object(span_t s) {
    long long v;
    if (parse_int(s, &v) == 0) return PyUnicode_FromFormat("%lld", v);
    return PyUnicode_FromStringAndSize(s.p, s.n);
}

/* every surviving record of one PAF text, grouped by read in order of first appearance */
typedef struct {
    rec_t* recs; grp_t* grps; PyObject* gindex;      /* gindex: {qname: group number} */
    Py_ssize_t n_rec, n_grp;
} paf_table_t;

static void table_free(paf_table_t* T) {
    for (Py_ssize_t g = 0; g < T->n_grp; ++g) Py_XDECREF(T->grps[g].qname);
    Py_XDECREF(T->gindex);
    PyMem_Free(T->recs);
    PyMem_Free(T->grps);
    memset(T, 0, sizeof *T);
}

/* Separators of one line. The bulk of a PAF line is its cg:Z: value (kilobytes); splitting it with one memchr per question
 * (where does the line end? the field? is there a third colon?) walks that value four times. One pass over the line with
 * 16/32-byte compares records the offsets of every tab and colon up to the newline; fields and tag parts follow from those. */
typedef struct { Py_ssize_t* off; Py_ssize_t n, cap; } seps_t;

static int seps_push(seps_t* S, Py_ssize_t o) {
    if (S->n == S->cap) {
        const Py_ssize_t cap = S->cap ? 2 * S->cap : 64;
        Py_ssize_t* q = (Py_ssize_t*)PyMem_Realloc(S->off, sizeof(Py_ssize_t) * (size_t)cap);
        if (!q) { PyErr_NoMemory(); return -1; }
        S->off = q; S->cap = cap;
    }
    S->off[S->n++] = o;
    return 0;
}

/* offsets of the tabs and colons of tp[pos, e) into S, e = offset of the first newline at or after pos (tlen if none);
 * returns e, or -1 with a Python error set */
static Py_ssize_t scan_line_scalar(const char* tp, Py_ssize_t pos, Py_ssize_t tlen, seps_t* S) {
    for (; pos < tlen; ++pos) {
        const char c = tp[pos];
        if (c == '\n') return pos;
        if ((c == '\t' || c == ':') && seps_push(S, pos) < 0) return -1;
    }
    return tlen;
}

#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static Py_ssize_t scan_line_avx2(const char* tp, Py_ssize_t pos, Py_ssize_t tlen, seps_t* S) {
    const __m256i vt = _mm256_set1_epi8('\t'), vc = _mm256_set1_epi8(':'), vn = _mm256_set1_epi8('\n');
    while (pos + 32 <= tlen) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(tp + pos));
        unsigned m = (unsigned)_mm256_movemask_epi8(
            _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, vt), _mm256_cmpeq_epi8(x, vc)), _mm256_cmpeq_epi8(x, vn)));
        while (m) {
            const int k = __builtin_ctz(m);
            m &= m - 1;
            if (tp[pos + k] == '\n') return pos + k;
            if (seps_push(S, pos + k) < 0) return -1;
        }
        pos += 32;
    }
    return scan_line_scalar(tp, pos, tlen, S);
}
static Py_ssize_t scan_line_sse2(const char* tp, Py_ssize_t pos, Py_ssize_t tlen, seps_t* S) {
    const __m128i vt = _mm_set1_epi8('\t'), vc = _mm_set1_epi8(':'), vn = _mm_set1_epi8('\n');
    while (pos + 16 <= tlen) {
        const __m128i x = _mm_loadu_si128((const __m128i*)(tp + pos));
        unsigned m = (unsigned)_mm_movemask_epi8(_mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(x, vt), _mm_cmpeq_epi8(x, vc)), _mm_cmpeq_epi8(x, vn)));
        while (m) {
            const int k = __builtin_ctz(m);
            m &= m - 1;
            if (tp[pos + k] == '\n') return pos + k;
            if (seps_push(S, pos + k) < 0) return -1;
        }
        pos += 16;
    }
    return scan_line_scalar(tp, pos, tlen, S);
}
#endif

static Py_ssize_t scan_line(const char* tp, Py_ssize_t pos, Py_ssize_t tlen, seps_t* S) {
#if defined(__x86_64__)
    static int avx2 = -1;
    if (avx2 < 0) avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
    return avx2 ? scan_line_avx2(tp, pos, tlen, S) : scan_line_sse2(tp, pos, tlen, S);
#else
    return scan_line_scalar(tp, pos, tlen, S);
#endif
}

/* pass 1: tokenise every line, filter, group by read. 0 = ok, -1 = Python error set (the caller frees the table). */
static int tokenise_paf_lines(const char* tp, Py_ssize_t tlen, long long min_len, paf_table_t* T, seps_t* S);

static int tokenise_paf(PyObject* text, long long min_len, paf_table_t* T) {
    memset(T, 0, sizeof *T);
    Py_ssize_t tlen = 0;
    const char* tp = PyUnicode_AsUTF8AndSize(text, &tlen);
    if (!tp) return -1;
    if (tlen != PyUnicode_GET_LENGTH(text)) { PyErr_SetString(PyExc_ValueError, "PAF text must be ASCII"); return -1; }
    T->gindex = PyDict_New();
    if (!T->gindex) return -1;
    seps_t S = {NULL, 0, 0};
    const int rc = tokenise_paf_lines(tp, tlen, min_len, T, &S);
    PyMem_Free(S.off);
    return rc;
}

static int tokenise_paf_lines(const char* tp, Py_ssize_t tlen, long long min_len, paf_table_t* T, seps_t* S) {
    Py_ssize_t cap_rec = 0, cap_grp = 0;
    PyObject* const gindex = T->gindex;
#define n_rec (T->n_rec)
#define n_grp (T->n_grp)
    for (Py_ssize_t pos = 0; pos < tlen;) {
        S->n = 0;
        const Py_ssize_t e = scan_line(tp, pos, tlen, S);
        if (e < 0) return -1;
        span_t line = {tp + pos, e - pos};
        pos = e + 1;
        line = strip_span(line);
        const Py_ssize_t la = (Py_ssize_t)(line.p - tp), lb = la + line.n;     /* the stripped line as offsets into the text */
        span_t col[12];
        int nc = 0;
        rec_t r;
        memset(&r, 0, sizeof r);
        r.next = -1;
        long long as = 0, blocklen = 0;
        int primary = 0, has_as = 0;
        char as_typ = 'i';
        span_t as_val = {NULL, 0};
        Py_ssize_t si = 0, fs = la;
        while (si < S->n && S->off[si] < la) ++si;                               /* separators inside the stripped-off head */
        for (;;) {
            /* the field runs to the next tab inside the stripped line; its colons come first in the separator list */
            Py_ssize_t c1 = -1, c2 = -1, colons = 0;
            while (si < S->n && S->off[si] < lb && tp[S->off[si]] == ':') {
                if (colons == 0) c1 = S->off[si]; else if (colons == 1) c2 = S->off[si];
                ++colons; ++si;
            }
            const int last = !(si < S->n && S->off[si] < lb);
            const Py_ssize_t fe = last ? lb : S->off[si];
            span_t f = {tp + fs, fe - fs};
            if (nc < 12) {
                col[nc] = f;
            } else {
                /* key:type:value */
                if (colons != 2) {
                    PyErr_SetString(PyExc_ValueError, colons < 2 ? "not enough values to unpack (expected 3)" : "too many values to unpack (expected 3)");
                    return -1;
                }
                const Py_ssize_t a = c1 - fs, b = c2 - fs;
                span_t key = {f.p, a}, typ = {f.p + a + 1, b - a - 1}, val = {f.p + b + 1, f.n - b - 1};
                if (typ.n != 1 || !(typ.p[0] == 'i' || typ.p[0] == 'A' || typ.p[0] == 'f' || typ.p[0] == 'Z')) {
                    PyObject* ko = PyUnicode_FromStringAndSize(typ.p, typ.n);
                    if (ko) { PyErr_SetObject(PyExc_KeyError, ko); Py_DECREF(ko); }
                    return -1;
                }
                if (key.n == 2 && key.p[0] == 'A' && key.p[1] == 'S') {
                    as_val = val; as_typ = typ.p[0]; has_as = 1;      /* a repeated key keeps its LAST value (dict) */
                } else if (key.n == 2 && key.p[0] == 't' && key.p[1] == 'p') {
                    primary = val.n == 1 && val.p[0] == 'P';
                } else if (key.n == 2 && key.p[0] == 'c' && key.p[1] == 'g') {
                    r.cigar = val;
                    r.has_cigar = 1;
                }
            }
            ++nc;
            if (last) break;
            fs = fe + 1;
            ++si;
        }
        if (nc < 12) { PyErr_SetString(PyExc_IndexError, "list index out of range (PAF line with fewer than 12 columns)"); return -1; }
        if (has_as) {
            /* int(conv_type(val, c[type])): an 'f' value goes through float first (paf.py:62,93-99) */
            const span_t val = as_val;
            int ok = 0;
            if (as_typ == 'f') {
                span_t v = strip_span(val);
                char buf[64];
                if (v.n > 0 && v.n < (Py_ssize_t)sizeof buf) {
                    memcpy(buf, v.p, (size_t)v.n);
                    buf[v.n] = 0;
                    char* end = NULL;
                    const double d = PyOS_string_to_double(buf, &end, NULL);
                    if (!(d == -1.0 && PyErr_Occurred()) && end == buf + v.n) {
                        if (d != d) { PyErr_SetString(PyExc_ValueError, "cannot convert float NaN to integer"); return -1; }
                        if (d > 9.2e18 || d < -9.2e18) { PyErr_SetString(PyExc_OverflowError, "cannot convert float infinity to integer"); return -1; }
                        as = (long long)d;               /* truncation toward zero, like int(float) */
                        ok = 1;
                    } else {
                        PyErr_Clear();
                    }
                }
            }
            if (!ok && parse_int(val, &as) != 0) { PyErr_SetString(PyExc_ValueError, "AS tag is not an integer"); return -1; }
        }

        /* conv_type keeps a column that is not an integer as a string; upstream only trips over it where the value is used:
         * the block length in the filter below (TypeError: str < int), mapq when several records of a read compete, the
         * coordinates when the record wins. Same here: remember what is unusable, complain when it is needed. */
        if (parse_int(col[10], &blocklen)) {
            PyErr_SetString(PyExc_TypeError, "'<' not supported between instances of 'str' and 'int' (alignment block length column)");
            return -1;
        }
        if (parse_int(col[1], &r.qlen) | parse_int(col[2], &r.qstart) | parse_int(col[3], &r.qend) | parse_int(col[7], &r.tstart) |
            parse_int(col[8], &r.tend))
            r.bad |= 1;
        if (parse_int(col[11], &r.mapq)) r.bad |= 2;
        r.rev = !(col[4].n == 1 && col[4].p[0] == '+');
        r.tname = col[5];
        r.as = as;
        if (blocklen < min_len || !primary) continue;              /* paf.py:664-669 */
        PyObject* qn = name_object(col[0]);
        if (!qn) return -1;
        PyObject* gi = PyDict_GetItemWithError(gindex, qn);         /* borrowed */
        if (!gi && PyErr_Occurred()) { Py_DECREF(qn); return -1; }
        if (n_rec == cap_rec) {
            const Py_ssize_t cap = cap_rec ? 2 * cap_rec : 1024;
            rec_t* q = (rec_t*)PyMem_Realloc(T->recs, sizeof(rec_t) * (size_t)cap);
            if (!q) { Py_DECREF(qn); PyErr_NoMemory(); return -1; }
            T->recs = q; cap_rec = cap;
        }
        Py_ssize_t g;
        if (gi) {
            g = PyLong_AsSsize_t(gi);
            Py_DECREF(qn);
            T->recs[T->grps[g].last].next = n_rec;
            T->grps[g].last = n_rec;
            T->grps[g].count++;
        } else {
            if (n_grp == cap_grp) {
                const Py_ssize_t cap = cap_grp ? 2 * cap_grp : 1024;
                grp_t* q = (grp_t*)PyMem_Realloc(T->grps, sizeof(grp_t) * (size_t)cap);
                if (!q) { Py_DECREF(qn); PyErr_NoMemory(); return -1; }
                T->grps = q; cap_grp = cap;
            }
            g = n_grp++;
            PyObject* go = PyLong_FromSsize_t(g);
            if (!go || PyDict_SetItem(gindex, qn, go) < 0) { Py_XDECREF(go); Py_DECREF(qn); --n_grp; return -1; }
            Py_DECREF(go);
            T->grps[g].first = T->grps[g].last = n_rec;
            T->grps[g].count = 1;
            T->grps[g].qname = qn;                                   /* owned until `done` */
        }
        T->recs[n_rec++] = r;
    }

#undef n_rec
#undef n_grp
    return 0;
}

/* Paf.choose_best_mapper over group g (paf.py:710-722): index of the winning record, -1 = Python error set */
static Py_ssize_t pick_winner(const paf_table_t* T, Py_ssize_t g, PyObject* best_index) {
    const rec_t* recs = T->recs;
    const grp_t* grps = T->grps;
    Py_ssize_t w = grps[g].first;
    if (grps[g].count > 16) {
        PyObject* keys = PyList_New(grps[g].count);
        if (!keys) return -1;
        Py_ssize_t k = 0;
        for (Py_ssize_t x = grps[g].first; x >= 0; x = recs[x].next, ++k)
            PyList_SET_ITEM(keys, k, Py_BuildValue("(LL)", recs[x].mapq, recs[x].as));
        PyObject* bi = PyObject_CallOneArg(best_index, keys);
        Py_DECREF(keys);
        if (!bi) return -1;
        Py_ssize_t want = PyLong_AsSsize_t(bi);
        Py_DECREF(bi);
        if (want < 0 || want >= grps[g].count) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "best_index out of range"); return -1; }
        for (k = 0; k < want; ++k) w = recs[w].next;
    } else {
        for (Py_ssize_t x = recs[w].next; x >= 0; x = recs[x].next)
            if (recs[x].mapq > recs[w].mapq || (recs[x].mapq == recs[w].mapq && recs[x].as >= recs[w].as)) w = x;
    }
    if (grps[g].count > 1)
        for (Py_ssize_t x = grps[g].first; x >= 0; x = recs[x].next)
            if (recs[x].bad & 2) { PyErr_SetString(PyExc_ValueError, "mapq column is not an integer (np.array of (mapq, AS) upstream)"); return -1; }
    return w;
}

/* the ten caller-provided arrays of a batch */
typedef struct {
    Py_buffer vb[10];
    int got;
    Py_ssize_t cap, n;        /* entries the buffers hold / entries written */
} out_t;

static int out_open(PyObject* bufs, out_t* O) {
    static const Py_ssize_t item[10] = {4, 8, 8, 4, 1, 8, 8, 8, 8, 8};
    O->got = 0; O->cap = 0; O->n = 0;
    if (PyTuple_GET_SIZE(bufs) != 10) { PyErr_SetString(PyExc_ValueError, "expected 10 output buffers"); return -1; }
    for (; O->got < 10; ++O->got) {
        if (PyObject_GetBuffer(PyTuple_GET_ITEM(bufs, O->got), &O->vb[O->got], PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) return -1;
        const Py_ssize_t c = O->vb[O->got].len / item[O->got];
        if (O->got == 0 || c < O->cap) O->cap = c;
    }
    return 0;
}

static void out_close(out_t* O) {
    for (int i = 0; i < O->got; ++i) PyBuffer_Release(&O->vb[i]);
    O->got = 0;
}

static int barcode_of(PyObject* barcodes, PyObject* qname, long long* bc) {
    *bc = 0;
    if (barcodes == Py_None) return 0;
    PyObject* bo = PyDict_GetItemWithError(barcodes, qname);
    if (!bo) return PyErr_Occurred() ? -1 : 0;
    if (bo != Py_None) { *bc = PyLong_AsLongLong(bo); if (*bc == -1 && PyErr_Occurred()) return -1; }
    return 0;
}

/* One winning record -> one row of the batch (the body of upstream's convert_records loop, sequences.py:700-738, minus the
 * CIGAR walk). 1 = written, 0 = its target is not tracked (core.py:83-86), -1 = Python error set. */
static int emit_record(const rec_t* r, PyObject* qname, PyObject* seqs, PyObject* contig_index, PyObject* barcodes, out_t* O, PyObject* keep) {
    PyObject* tn = name_object(r->tname);
    if (!tn) return -1;
    PyObject* k = PyDict_GetItemWithError(contig_index, tn);      /* borrowed */
    Py_DECREF(tn);
    if (!k) return PyErr_Occurred() ? -1 : 0;
    const long long ki = PyLong_AsLongLong(k);
    if (r->bad & 1) { PyErr_SetString(PyExc_ValueError, "PAF record with a non-integer coordinate column"); return -1; }
    PyObject* s = PyDict_GetItemWithError(seqs, qname);           /* borrowed; KeyError like seqs[rec.qname] */
    if (!s) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, qname); return -1; }
    if (!PyUnicode_Check(s)) { PyErr_SetString(PyExc_TypeError, "reads must be str"); return -1; }
    Py_ssize_t slen = 0;
    const char* sptr = PyUnicode_AsUTF8AndSize(s, &slen);
    if (!sptr) return -1;
    if (slen != PyUnicode_GET_LENGTH(s)) { PyErr_SetString(PyExc_ValueError, "read and CIGAR strings must be ASCII"); return -1; }
    long long bc = 0;
    if (barcode_of(barcodes, qname, &bc) < 0) return -1;
    long long lo, hi;
    const long long len = (long long)slen;
    if (r->qstart < 0 || r->qend < 0 || (r->rev && (r->qlen - r->qend < 0 || r->qlen - r->qstart < 0))) {
        PyErr_SetString(PyExc_ValueError, "negative query coordinates");
        return -1;
    }
    if (!r->has_cigar) { PyErr_SetString(PyExc_AssertionError, "record without a cg:Z: CIGAR"); return -1; }   /* sequences.py:718, after the slicing */
    if (r->rev) {                                                   /* Q12, as in convert() above */
        const long long a = r->qlen - r->qend, b = r->qlen - r->qstart;
        lo = len - (b < len ? b : len); if (lo < 0) lo = 0;
        hi = len - (a < len ? a : len); if (hi < 0) hi = 0;
    } else {
        lo = clampll(r->qstart, 0, len);
        hi = clampll(r->qend, 0, len);
    }
    if (hi < lo) hi = lo;
    const Py_ssize_t n = O->n;
    if (n >= O->cap) { PyErr_SetString(PyExc_ValueError, "output buffer too small"); return -1; }
    ((int32_t*)O->vb[0].buf)[n] = (int32_t)ki;   ((int64_t*)O->vb[1].buf)[n] = r->tstart;   ((int64_t*)O->vb[2].buf)[n] = r->tend;
    ((int32_t*)O->vb[3].buf)[n] = (int32_t)bc;   ((uint8_t*)O->vb[4].buf)[n] = (uint8_t)r->rev;
    ((uint64_t*)O->vb[5].buf)[n] = (uint64_t)(uintptr_t)r->cigar.p;   ((int64_t*)O->vb[6].buf)[n] = (int64_t)r->cigar.n;
    ((uint64_t*)O->vb[7].buf)[n] = (uint64_t)(uintptr_t)sptr;   ((int64_t*)O->vb[8].buf)[n] = lo;   ((int64_t*)O->vb[9].buf)[n] = hi;
    if (PyList_Append(keep, s) < 0) return -1;
    O->n = n + 1;
    return 1;
}

static PyObject* convert_text(PyObject* self, PyObject* args) {
    PyObject *text, *seqs, *contig_index, *best_index, *barcodes, *bufs, *keep;
    long long min_len;
    if (!PyArg_ParseTuple(args, "UO!O!LOOO!O!", &text, &PyDict_Type, &seqs, &PyDict_Type, &contig_index, &min_len, &barcodes,
                          &best_index, &PyTuple_Type, &bufs, &PyList_Type, &keep))
        return NULL;
    if (barcodes != Py_None && !PyDict_Check(barcodes)) { PyErr_SetString(PyExc_TypeError, "barcodes must be a dict or None"); return NULL; }
    if (PyTuple_GET_SIZE(bufs) != 10) { PyErr_SetString(PyExc_ValueError, "expected 10 output buffers"); return NULL; }
    paf_table_t T;
    out_t O;
    O.got = 0;
    PyObject* result = NULL;
    Py_ssize_t skipped = 0;
    if (tokenise_paf(text, min_len, &T) < 0) goto done;
    /* pass 2: winner of every read -> the ten arrays (len(seqs) entries always suffice: every used read is a key of seqs) */
    if (out_open(bufs, &O) < 0) goto done;
    if (T.n_grp > 0 && PyList_Append(keep, text) < 0) goto done;    /* the CIGAR pointers live inside the text */
    for (Py_ssize_t g = 0; g < T.n_grp; ++g) {
        const Py_ssize_t w = pick_winner(&T, g, best_index);
        if (w < 0) goto done;
        const int rc = emit_record(&T.recs[w], T.grps[g].qname, seqs, contig_index, barcodes, &O, keep);
        if (rc < 0) goto done;
        if (rc == 0) ++skipped;
    }
    result = Py_BuildValue("nnn", O.n, skipped, T.n_grp);
done:
    out_close(&O);
    table_free(&T);
    return result;
}

/* ------------------------------------------------------------------------------------------------------------------
 * decide_text — the simulator's decision step and the batch it produces, straight from the two PAF texts.
 *
 * BossRunsSim.make_decisions + filter_paf_dict + convert_records (boss/runs/simulation.py:37-135,160-163): for every read
 * with a mu-sized mapping the winning truncated record is looked up in the current strategy
 * (strat[start // window, rev, barcode]; NumPy's index rules: negative indices wrap, out of range or no strategy for the
 * target -> reject); accepted reads contribute their winning FULL-length record, rejected ones the truncated record (and
 * only their first mu bases count later, Q12); reads without a mu-sized mapping are accepted or rejected wholesale.
 * Semantics pinned by tests/test_simulation.py against boss_runs_b200/simulation.py:make_decisions, which is itself pinned
 * against upstream's function on the reference's data.
 *
 * decide_text(paf_full, paf_trunc, seqs, contig_index, barcodes, strat, window, mu, accept_unmapped, best_index, bufs, keep,
 *             row_accepted)
 *   -> (n_rows, n_skipped, mapped: list[str], rejected: list[str], accepted_qlen: list[int], n_accepted, n_rejected)
 * row_accepted: writable u8 buffer, one flag per emitted row (1 = the row is an accepted read: record length != mu).
 * ------------------------------------------------------------------------------------------------------------------ */
static long long floordiv(long long a, long long b) { long long q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

/* strat[row, rev, bc] under NumPy's rules; 0 when anything is out of range or the array is not 3-d (reject refs hold zeros(1)) */
static int strat_bit(PyObject* arr, long long row, int rev, long long bc) {
    Py_buffer v;
    if (PyObject_GetBuffer(arr, &v, PyBUF_STRIDES | PyBUF_FORMAT) < 0) { PyErr_Clear(); return 0; }
    int d = 0;
    if (v.ndim == 3 && v.itemsize == 1 && v.shape[1] == 2) {
        const long long rows = v.shape[0], nb = v.shape[2];
        if (row < 0) row += rows;
        if (bc < 0) bc += nb;
        if (row >= 0 && row < rows && bc >= 0 && bc < nb)
            d = ((const unsigned char*)v.buf)[row * v.strides[0] + (long long)rev * v.strides[1] + bc * v.strides[2]] != 0;
    }
    PyBuffer_Release(&v);
    return d;
}

static PyObject* decide_text(PyObject* self, PyObject* args) {
    PyObject *full, *trunc, *seqs, *contig_index, *barcodes, *strat, *best_index, *bufs, *keep, *rowacc;
    long long window, mu;
    int accept_unmapped;
    if (!PyArg_ParseTuple(args, "UUO!O!O!O!LLpOO!O!O", &full, &trunc, &PyDict_Type, &seqs, &PyDict_Type, &contig_index, &PyDict_Type, &barcodes,
                          &PyDict_Type, &strat, &window, &mu, &accept_unmapped, &best_index, &PyTuple_Type, &bufs, &PyList_Type, &keep, &rowacc))
        return NULL;
    if (window <= 0) { PyErr_SetString(PyExc_ValueError, "window must be positive"); return NULL; }
    paf_table_t F, T;
    memset(&F, 0, sizeof F); memset(&T, 0, sizeof T);
    out_t O;
    O.got = 0;
    Py_buffer ra;
    int have_ra = 0;
    PyObject *result = NULL, *mapped = PyList_New(0), *rejected = PyList_New(0), *acc_qlen = PyList_New(0);
    Py_ssize_t skipped = 0, n_acc = 0, n_rej = 0;
    if (!mapped || !rejected || !acc_qlen) goto done;
    if (tokenise_paf(full, 1, &F) < 0 || tokenise_paf(trunc, 1, &T) < 0) goto done;       /* make_decisions parses with min_len = 1 */
    if (out_open(bufs, &O) < 0) goto done;
    if (PyObject_GetBuffer(rowacc, &ra, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) goto done;
    have_ra = 1;
    if (PyList_Append(keep, full) < 0 || PyList_Append(keep, trunc) < 0) goto done;       /* CIGAR pointers live inside the texts */
    for (Py_ssize_t g = 0; g < T.n_grp; ++g) {
        PyObject* qn = T.grps[g].qname;
        const Py_ssize_t w = pick_winner(&T, g, best_index);
        if (w < 0) goto done;
        const rec_t* r = &T.recs[w];
        PyObject* bo = PyDict_GetItemWithError(barcodes, qn);                              /* rec.barcode = barcodes[rec.qname] */
        if (!bo) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, qn); goto done; }
        const long long bc = PyLong_AsLongLong(bo);
        if (bc == -1 && PyErr_Occurred()) goto done;
        if (PyList_Append(mapped, qn) < 0) goto done;
        if (r->bad & 1) { PyErr_SetString(PyExc_ValueError, "PAF record with a non-integer coordinate column"); goto done; }
        int decision = 0;
        {
            PyObject* tn = name_object(r->tname);
            if (!tn) goto done;
            PyObject* arr = PyDict_GetItemWithError(strat, tn);                            /* borrowed; missing -> KeyError -> reject */
            Py_DECREF(tn);
            if (!arr && PyErr_Occurred()) goto done;
            if (arr) decision = strat_bit(arr, floordiv(r->rev ? r->tend - 1 : r->tstart, window), r->rev ? 1 : 0, bc);
        }
        const rec_t* use = r;
        if (decision) {
            PyObject* gi = PyDict_GetItemWithError(F.gindex, qn);
            if (!gi) {
                if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "index -1 is out of bounds for axis 0 with size 0 (read without a full-length record)");
                goto done;
            }
            const Py_ssize_t wf = pick_winner(&F, PyLong_AsSsize_t(gi), best_index);
            if (wf < 0) goto done;
            use = &F.recs[wf];
            ++n_acc;
        } else {
            if (PyList_Append(rejected, qn) < 0) goto done;
            ++n_rej;
        }
        const int is_acc = !(use->bad & 1) && use->qlen != mu;                             /* filter_paf_dict: record length != mu */
        if (is_acc) {
            PyObject* q = PyLong_FromLongLong(use->qlen);
            if (!q || PyList_Append(acc_qlen, q) < 0) { Py_XDECREF(q); goto done; }
            Py_DECREF(q);
        }
        const int rc = emit_record(use, qn, seqs, contig_index, barcodes, &O, keep);
        if (rc < 0) goto done;
        if (rc == 0) { ++skipped; continue; }
        if (O.n > ra.len) { PyErr_SetString(PyExc_ValueError, "row_accepted buffer too small"); goto done; }
        ((unsigned char*)ra.buf)[O.n - 1] = (unsigned char)is_acc;
    }
    if (accept_unmapped) {
        /* reads without a mu-sized mapping that do have a full-length record join the batch (no barcode: index 0) */
        PyObject *rid, *seq;
        Py_ssize_t pos = 0;
        while (PyDict_Next(seqs, &pos, &rid, &seq)) {
            int in_t = PyDict_Contains(T.gindex, rid);
            if (in_t < 0) goto done;
            if (in_t) continue;
            PyObject* gi = PyDict_GetItemWithError(F.gindex, rid);
            if (!gi) { if (PyErr_Occurred()) goto done; continue; }
            const Py_ssize_t wf = pick_winner(&F, PyLong_AsSsize_t(gi), best_index);
            if (wf < 0) goto done;
            const rec_t* use = &F.recs[wf];
            const int is_acc = !(use->bad & 1) && use->qlen != mu;
            if (is_acc) {
                PyObject* q = PyLong_FromLongLong(use->qlen);
                if (!q || PyList_Append(acc_qlen, q) < 0) { Py_XDECREF(q); goto done; }
                Py_DECREF(q);
            }
            const int rc = emit_record(use, rid, seqs, contig_index, Py_None, &O, keep);
            if (rc < 0) goto done;
            if (rc == 0) { ++skipped; continue; }
            if (O.n > ra.len) { PyErr_SetString(PyExc_ValueError, "row_accepted buffer too small"); goto done; }
            ((unsigned char*)ra.buf)[O.n - 1] = (unsigned char)is_acc;
        }
    }
    result = Py_BuildValue("nnOOOnn", O.n, skipped, mapped, rejected, acc_qlen, n_acc, n_rej);
done:
    if (have_ra) PyBuffer_Release(&ra);
    out_close(&O);
    table_free(&F);
    table_free(&T);
    Py_XDECREF(mapped); Py_XDECREF(rejected); Py_XDECREF(acc_qlen);
    return result;
}

static PyMethodDef methods[] = {
    {"convert", convert, METH_VARARGS, "convert(paf_dict, seqs, contig_index, best_record, bufs, keep) -> (n_used, n_skipped)"},
    {"convert_text", convert_text, METH_VARARGS,
     "convert_text(paf_text, seqs, contig_index, min_len, barcodes, best_index, bufs, keep) -> (n_used, n_skipped, n_reads)"},
    {"decide_text", decide_text, METH_VARARGS,
     "decide_text(paf_full, paf_trunc, seqs, contig_index, barcodes, strat, window, mu, accept_unmapped, best_index, bufs, keep, row_accepted)"
     " -> (n_rows, n_skipped, mapped, rejected, accepted_qlen, n_accepted, n_rejected)"},
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
