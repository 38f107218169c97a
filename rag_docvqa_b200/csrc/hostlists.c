/*
 * _hostlists -- CPython helper that builds the reference's nested-list view of the retrieved chunks.
 *
 * Retriever.retrieve (reference src/_modules.py:2155-2180) must return eight nested Python lists for the
 * <= k hits of every document (src/_modules.py:2144-2153).  Once the scores and the top-k come from the GPU
 * (a few microseconds per batch), building those lists in Python is the whole cost of the call (cProfile:
 * ~8 us per hit).  This module does the same walk in C with the same Python semantics, for the case every
 * shipped config uses (include_surroundings == 0: a hit is exactly its own chunk, src/_modules.py:2067-2083
 * is a no-op).  Neighbour windows and the (page, ymin, xmin) reorder stay in retriever.py.
 *
 * Per hit i of document b (rank order), following the reference line by line:
 *   label  = layout_labels_chunks[b][i]                     src/_modules.py:2019
 *   page   = page_indices[b][i]                             src/_modules.py:2020
 *   words  = list(words_text_chunks[b][i])                  the emitted words (no neighbours)
 *   boxes  = list(words_box_chunks[b][i])
 *   text   = " ".join(words)                                Chunker.compact_chunks, src/_modules.py:1102-1132
 *   bbox   = [min x0, min y0, max x1, max y1] over boxes, [0, 0, 1, 1] if empty (Python min/max semantics:
 *            first extreme element wins, original objects are returned)
 *   wlabel = [label] * len(words)                           src/_modules.py:2094-2100
 *   rect   = int(bbox * page size) with truncation, then min/max order fix    src/_modules.py:2108-2119
 *   patch  = make_patch(images[b][page], (x0, y0, x1, y1))  (page.crop, or a deferred crop)
 *
 * This is host logic only: no CUDA, no torch.  It is not a fallback for anything on the device.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#include <string.h>
#include <stdint.h>

static PyObject* s_space = NULL;     /* " " */
static PyObject* s_width = NULL;     /* "width" */
static PyObject* s_height = NULL;    /* "height" */

/* Python's `a < b` / `a > b` with a fast path for exact floats. */
static int less_than(PyObject* a, PyObject* b, int* err) {
    if (PyFloat_CheckExact(a) && PyFloat_CheckExact(b)) return PyFloat_AS_DOUBLE(a) < PyFloat_AS_DOUBLE(b);
    int r = PyObject_RichCompareBool(a, b, Py_LT);
    if (r < 0) { *err = 1; return 0; }
    return r;
}
static int greater_than(PyObject* a, PyObject* b, int* err) {
    if (PyFloat_CheckExact(a) && PyFloat_CheckExact(b)) return PyFloat_AS_DOUBLE(a) > PyFloat_AS_DOUBLE(b);
    int r = PyObject_RichCompareBool(a, b, Py_GT);
    if (r < 0) { *err = 1; return 0; }
    return r;
}

/* int(coord * size): float * int -> float -> truncation toward zero (raises like int() on nan / inf). */
static int scaled_int(PyObject* coord, PyObject* size, long* out) {
    PyObject* v = NULL;
    if (PyFloat_CheckExact(coord) && PyLong_CheckExact(size)) {
        double s = PyLong_AsDouble(size);
        if (s == -1.0 && PyErr_Occurred()) return -1;
        v = PyLong_FromDouble(PyFloat_AS_DOUBLE(coord) * s);
    } else {
        PyObject* prod = PyNumber_Multiply(coord, size);
        if (!prod) return -1;
        v = PyNumber_Long(prod);
        Py_DECREF(prod);
    }
    if (!v) return -1;
    long r = PyLong_AsLong(v);
    Py_DECREF(v);
    if (r == -1 && PyErr_Occurred()) return -1;
    *out = r;
    return 0;
}

/* The words / boxes of a hit are ~300 small heap objects (str, list, float) that were allocated long ago: the
 * walk is a chain of cache misses (list -> item array -> box list -> its item array -> float).  Touch each
 * level for the whole chunk before using it, so the misses of one level overlap instead of serialising. */
static void prefetch_chunk(PyObject* words, PyObject* boxes) {
    if (PyList_CheckExact(words)) {
        Py_ssize_t n = PyList_GET_SIZE(words);
        for (Py_ssize_t j = 0; j < n; ++j) __builtin_prefetch(PyList_GET_ITEM(words, j));
    }
    if (!PyList_CheckExact(boxes)) return;
    Py_ssize_t n = PyList_GET_SIZE(boxes);
    for (Py_ssize_t j = 0; j < n; ++j) __builtin_prefetch(PyList_GET_ITEM(boxes, j));
    for (Py_ssize_t j = 0; j < n; ++j) {
        PyObject* box = PyList_GET_ITEM(boxes, j);
        if (PyList_CheckExact(box)) __builtin_prefetch(((PyListObject*)box)->ob_item);
    }
    for (Py_ssize_t j = 0; j < n; ++j) {
        PyObject* box = PyList_GET_ITEM(boxes, j);
        if (PyList_CheckExact(box)) {
            Py_ssize_t m = PyList_GET_SIZE(box);
            for (Py_ssize_t e = 0; e < m && e < 4; ++e) __builtin_prefetch(PyList_GET_ITEM(box, e));
        } else if (PyTuple_CheckExact(box)) {
            Py_ssize_t m = PyTuple_GET_SIZE(box);
            for (Py_ssize_t e = 0; e < m && e < 4; ++e) __builtin_prefetch(PyTuple_GET_ITEM(box, e));
        }
    }
}

/* bbox of a list of boxes: returns a new 4-list, or NULL on error. */
static PyObject* bbox_of(PyObject* boxes /* list */) {
    Py_ssize_t n = PyList_GET_SIZE(boxes);
    if (n == 0) return Py_BuildValue("[iiii]", 0, 0, 1, 1);          /* src/_modules.py:1126-1127 */
    PyObject* ext[4] = {NULL, NULL, NULL, NULL};                     /* strong references */
    int err = 0;
    for (Py_ssize_t j = 0; j < n && !err; ++j) {
        PyObject* box = PyList_GET_ITEM(boxes, j);
        PyObject** it = NULL;
        if (PyList_CheckExact(box) && PyList_GET_SIZE(box) >= 4) it = ((PyListObject*)box)->ob_item;
        else if (PyTuple_CheckExact(box) && PyTuple_GET_SIZE(box) >= 4) it = ((PyTupleObject*)box)->ob_item;
        if (it && ext[0] && PyFloat_CheckExact(it[0]) && PyFloat_CheckExact(it[1]) && PyFloat_CheckExact(it[2]) &&
            PyFloat_CheckExact(it[3]) && PyFloat_CheckExact(ext[0]) && PyFloat_CheckExact(ext[1]) &&
            PyFloat_CheckExact(ext[2]) && PyFloat_CheckExact(ext[3])) {
            /* common case: plain floats -- no Python code can run, so borrowed pointers are safe and only a new
             * extreme touches a reference count */
            for (int e = 0; e < 4; ++e) {
                const double v = PyFloat_AS_DOUBLE(it[e]), cur = PyFloat_AS_DOUBLE(ext[e]);
                if (e < 2 ? v < cur : v > cur) { Py_INCREF(it[e]); Py_DECREF(ext[e]); ext[e] = it[e]; }
            }
            continue;
        }
        PyObject* c[4];
        for (int e = 0; e < 4; ++e) {
            if (it) { c[e] = it[e]; Py_INCREF(c[e]); }
            else {
                c[e] = PySequence_GetItem(box, e);
                if (!c[e]) { for (int f = 0; f < e; ++f) Py_DECREF(c[f]); err = 1; break; }
            }
        }
        if (err) break;
        for (int e = 0; e < 4; ++e) {
            if (ext[e] == NULL) { ext[e] = c[e]; continue; }                  /* first element seeds min / max */
            int better = e < 2 ? less_than(c[e], ext[e], &err) : greater_than(c[e], ext[e], &err);
            if (better) { Py_DECREF(ext[e]); ext[e] = c[e]; } else { Py_DECREF(c[e]); }
        }
    }
    if (err) { for (int e = 0; e < 4; ++e) Py_XDECREF(ext[e]); return NULL; }
    PyObject* out = PyList_New(4);
    if (!out) { for (int e = 0; e < 4; ++e) Py_XDECREF(ext[e]); return NULL; }
    for (int e = 0; e < 4; ++e) PyList_SET_ITEM(out, e, ext[e]);              /* steals */
    return out;
}

/* page.width / page.height are Python-level properties on PIL images (~0.6 us per hit for the pair): remember
 * the sizes of the last few distinct page objects of this call (they are alive for its whole duration). */
#define PAGE_CACHE 16
typedef struct { PyObject* page; PyObject* w; PyObject* h; } PageSize;

static int page_size(PageSize* cache, int* used, PyObject* page, PyObject** w, PyObject** h) {
    for (int i = 0; i < *used; ++i)
        if (cache[i].page == page) { *w = cache[i].w; *h = cache[i].h; return 0; }
    PyObject* pw = PyObject_GetAttr(page, s_width);
    if (!pw) return -1;
    PyObject* ph = PyObject_GetAttr(page, s_height);
    if (!ph) { Py_DECREF(pw); return -1; }
    int slot = *used < PAGE_CACHE ? (*used)++ : 0;
    if (cache[slot].page) { Py_DECREF(cache[slot].w); Py_DECREF(cache[slot].h); }
    cache[slot].page = page; cache[slot].w = pw; cache[slot].h = ph;     /* the cache owns the references */
    *w = pw; *h = ph;
    return 0;
}

static PyObject* gather_s0_impl(PyObject* hits, PyObject* words_all, PyObject* boxes_all, PyObject* labels_all,
                                PyObject* images_all, PyObject* pages_all, PyObject* make_patch, Py_ssize_t B);

/* gather_s0(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices, make_patch)
 *   -> 8-tuple of per-document lists */
static PyObject* gather_s0(PyObject* self, PyObject* args) {
    PyObject *hits, *words_all, *boxes_all, *labels_all, *images_all, *pages_all, *make_patch;
    if (!PyArg_ParseTuple(args, "OOOOOOO", &hits, &words_all, &boxes_all, &labels_all, &images_all, &pages_all, &make_patch))
        return NULL;
    Py_ssize_t B = PySequence_Size(hits);
    if (B < 0) return NULL;
    /* ~1300 acyclic containers are created per call; with the caller's millions of word / box objects alive, the
     * cyclic collector's periodic passes were 30-40 % of this function.  Nothing built here can form a cycle. */
    const int gc_was_enabled = PyGC_Disable();
    PyObject* result = gather_s0_impl(hits, words_all, boxes_all, labels_all, images_all, pages_all, make_patch, B);
    if (gc_was_enabled) PyGC_Enable();
    return result;
}

static PyObject* gather_s0_impl(PyObject* hits, PyObject* words_all, PyObject* boxes_all, PyObject* labels_all,
                                PyObject* images_all, PyObject* pages_all, PyObject* make_patch, Py_ssize_t B) {
    PyObject* outs[8];
    for (int o = 0; o < 8; ++o) {
        outs[o] = PyList_New(B);
        if (!outs[o]) { for (int f = 0; f < o; ++f) Py_DECREF(outs[f]); return NULL; }
    }
    PyObject *doc_hits = NULL, *words_b = NULL, *boxes_b = NULL, *labels_b = NULL, *images_b = NULL, *pages_b = NULL;
    PyObject* d[8] = {NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
    PageSize sizes[PAGE_CACHE];
    int sizes_used = 0;
    memset(sizes, 0, sizeof(sizes));
    for (Py_ssize_t b = 0; b < B; ++b) {
        doc_hits = PySequence_GetItem(hits, b);
        words_b = PySequence_GetItem(words_all, b);
        boxes_b = PySequence_GetItem(boxes_all, b);
        labels_b = PySequence_GetItem(labels_all, b);
        images_b = PySequence_GetItem(images_all, b);
        pages_b = PySequence_GetItem(pages_all, b);
        if (!doc_hits || !words_b || !boxes_b || !labels_b || !images_b || !pages_b) goto fail;
        Py_ssize_t nh = PySequence_Size(doc_hits);
        if (nh < 0) goto fail;
        for (int o = 0; o < 8; ++o) { d[o] = PyList_New(nh); if (!d[o]) goto fail; }
        for (Py_ssize_t j = 0; j < nh; ++j) {
            PyObject* idx = PySequence_GetItem(doc_hits, j);
            if (!idx) goto fail;
            PyObject* label = PyObject_GetItem(labels_b, idx);
            PyObject* page_idx = PyObject_GetItem(pages_b, idx);
            PyObject* w_src = PyObject_GetItem(words_b, idx);
            PyObject* b_src = PyObject_GetItem(boxes_b, idx);
            Py_DECREF(idx);
            PyObject *words = NULL, *boxes = NULL, *text = NULL, *bbox = NULL, *wlabels = NULL, *page = NULL, *patch = NULL;
            PyObject *pw = NULL, *ph = NULL, *rect = NULL;
            int ok = label && page_idx && w_src && b_src;
            if (ok) prefetch_chunk(w_src, b_src);
            if (ok) { words = PySequence_List(w_src); boxes = PySequence_List(b_src); ok = words && boxes; }
            if (ok) { text = PyUnicode_Join(s_space, words); bbox = bbox_of(boxes); ok = text && bbox; }
            if (ok) {
                Py_ssize_t nw = PyList_GET_SIZE(words);
                wlabels = PyList_New(nw);
                ok = wlabels != NULL;
                for (Py_ssize_t t = 0; ok && t < nw; ++t) { Py_INCREF(label); PyList_SET_ITEM(wlabels, t, label); }
            }
            if (ok) { page = PyObject_GetItem(images_b, page_idx); ok = page != NULL; }
            if (ok) ok = page_size(sizes, &sizes_used, page, &pw, &ph) == 0;       /* borrowed from the cache */
            if (ok) {
                long x0, y0, x1, y1;
                ok = scaled_int(PyList_GET_ITEM(bbox, 0), pw, &x0) == 0 && scaled_int(PyList_GET_ITEM(bbox, 1), ph, &y0) == 0 &&
                     scaled_int(PyList_GET_ITEM(bbox, 2), pw, &x1) == 0 && scaled_int(PyList_GET_ITEM(bbox, 3), ph, &y1) == 0;
                if (ok) {
                    rect = Py_BuildValue("(llll)", x0 < x1 ? x0 : x1, y0 < y1 ? y0 : y1, x0 < x1 ? x1 : x0, y0 < y1 ? y1 : y0);
                    ok = rect != NULL;
                }
            }
            if (ok) { patch = PyObject_CallFunctionObjArgs(make_patch, page, rect, NULL); ok = patch != NULL; }
            Py_XDECREF(w_src); Py_XDECREF(b_src); Py_XDECREF(page); Py_XDECREF(rect);
            if (!ok) {
                Py_XDECREF(label); Py_XDECREF(page_idx); Py_XDECREF(words); Py_XDECREF(boxes); Py_XDECREF(text);
                Py_XDECREF(bbox); Py_XDECREF(wlabels); Py_XDECREF(patch);
                goto fail;
            }
            PyList_SET_ITEM(d[0], j, text);      PyList_SET_ITEM(d[1], j, bbox);    PyList_SET_ITEM(d[2], j, label);
            PyList_SET_ITEM(d[3], j, words);     PyList_SET_ITEM(d[4], j, boxes);   PyList_SET_ITEM(d[5], j, wlabels);
            PyList_SET_ITEM(d[6], j, patch);     PyList_SET_ITEM(d[7], j, page_idx);
        }
        for (int o = 0; o < 8; ++o) { PyList_SET_ITEM(outs[o], b, d[o]); d[o] = NULL; }
        Py_CLEAR(doc_hits); Py_CLEAR(words_b); Py_CLEAR(boxes_b); Py_CLEAR(labels_b); Py_CLEAR(images_b); Py_CLEAR(pages_b);
    }
    {
        PyObject* res = PyTuple_New(8);
        if (!res) goto fail;
        for (int o = 0; o < 8; ++o) PyTuple_SET_ITEM(res, o, outs[o]);
        for (int i = 0; i < sizes_used; ++i) { Py_DECREF(sizes[i].w); Py_DECREF(sizes[i].h); }
        return res;
    }
fail:
    for (int i = 0; i < sizes_used; ++i) { Py_XDECREF(sizes[i].w); Py_XDECREF(sizes[i].h); }
    Py_XDECREF(doc_hits); Py_XDECREF(words_b); Py_XDECREF(boxes_b); Py_XDECREF(labels_b); Py_XDECREF(images_b); Py_XDECREF(pages_b);
    for (int o = 0; o < 8; ++o) { Py_XDECREF(d[o]); Py_DECREF(outs[o]); }
    return NULL;
}

/* ---- flatten_docs: the word / token / box walk of DocStore.from_lists -------------------------------------------
 * flatten_docs(words_text_chunks, words_box_chunks, tokenize, cache) ->
 *     (chunk_nwords int64[N], word_ntok int32[W], tok_ids int32[T], boxes float64[W*4])  as bytes objects
 * Every word of every chunk is visited once (2 M words for a C2 batch: seconds in Python, ~0.2 s here).  `cache` is a
 * dict word -> tuple of token ids, filled through `tokenize(word)` (ids without the trailing EOS, src/VT5.py:160). */
typedef struct { char* p; size_t n, cap; } Buf;
static int buf_push(Buf* b, const void* src, size_t bytes) {
    if (b->n + bytes > b->cap) {
        size_t cap = b->cap ? b->cap * 2 : (1u << 16);
        while (cap < b->n + bytes) cap *= 2;
        char* q = (char*)PyMem_Realloc(b->p, cap);
        if (!q) { PyErr_NoMemory(); return -1; }
        b->p = q; b->cap = cap;
    }
    memcpy(b->p + b->n, src, bytes);
    b->n += bytes;
    return 0;
}

static PyObject* flatten_docs(PyObject* self, PyObject* args) {
    PyObject *words_all, *boxes_all, *tokenize, *cache;
    if (!PyArg_ParseTuple(args, "OOOO!", &words_all, &boxes_all, &tokenize, &PyDict_Type, &cache)) return NULL;
    Buf nwords = {0}, ntok = {0}, ids = {0}, boxes = {0};
    PyObject *docs_w = PySequence_Fast(words_all, "words_text_chunks must be a sequence");
    PyObject *docs_b = docs_w ? PySequence_Fast(boxes_all, "words_box_chunks must be a sequence") : NULL;
    PyObject *doc_w = NULL, *doc_b = NULL, *chunk_w = NULL, *chunk_b = NULL, *box = NULL, *result = NULL;
    if (!docs_w || !docs_b) goto done;
    if (PySequence_Fast_GET_SIZE(docs_w) != PySequence_Fast_GET_SIZE(docs_b)) { PyErr_SetString(PyExc_ValueError, "words / boxes: different number of documents"); goto done; }
    for (Py_ssize_t b = 0; b < PySequence_Fast_GET_SIZE(docs_w); ++b) {
        doc_w = PySequence_Fast(PySequence_Fast_GET_ITEM(docs_w, b), "a document must be a sequence of chunks");
        doc_b = doc_w ? PySequence_Fast(PySequence_Fast_GET_ITEM(docs_b, b), "a document must be a sequence of chunks") : NULL;
        if (!doc_w || !doc_b) goto done;
        if (PySequence_Fast_GET_SIZE(doc_w) != PySequence_Fast_GET_SIZE(doc_b)) { PyErr_Format(PyExc_ValueError, "document %zd: words / boxes chunk counts differ", b); goto done; }
        for (Py_ssize_t c = 0; c < PySequence_Fast_GET_SIZE(doc_w); ++c) {
            chunk_w = PySequence_Fast(PySequence_Fast_GET_ITEM(doc_w, c), "a chunk must be a sequence of words");
            if (!chunk_w) goto done;
            const Py_ssize_t nw = PySequence_Fast_GET_SIZE(chunk_w);
            const int64_t nw64 = (int64_t)nw;
            if (buf_push(&nwords, &nw64, 8)) goto done;
            for (Py_ssize_t i = 0; i < nw; ++i) {
                PyObject* w = PySequence_Fast_GET_ITEM(chunk_w, i);
                PyObject* toks = PyDict_GetItemWithError(cache, w);                 /* borrowed */
                if (!toks) {
                    if (PyErr_Occurred()) goto done;
                    PyObject* raw = PyObject_CallOneArg(tokenize, w);
                    if (!raw) goto done;
                    PyObject* tup = PySequence_Tuple(raw);
                    Py_DECREF(raw);
                    if (!tup) goto done;
                    if (PyDict_SetItem(cache, w, tup) != 0) { Py_DECREF(tup); goto done; }
                    Py_DECREF(tup);
                    toks = tup;                                                     /* kept alive by the dict */
                }
                if (!PyTuple_Check(toks)) { PyErr_SetString(PyExc_TypeError, "token cache values must be tuples"); goto done; }
                const int32_t nt = (int32_t)PyTuple_GET_SIZE(toks);
                if (buf_push(&ntok, &nt, 4)) goto done;
                for (int32_t t = 0; t < nt; ++t) {
                    const long v = PyLong_AsLong(PyTuple_GET_ITEM(toks, t));
                    if (v == -1 && PyErr_Occurred()) goto done;
                    const int32_t v32 = (int32_t)v;
                    if (buf_push(&ids, &v32, 4)) goto done;
                }
            }
            if (nw > 0) {
                chunk_b = PySequence_Fast(PySequence_Fast_GET_ITEM(doc_b, c), "a chunk's boxes must be a sequence");
                if (!chunk_b) goto done;
                if (PySequence_Fast_GET_SIZE(chunk_b) != nw) { PyErr_Format(PyExc_ValueError, "document %zd chunk %zd: %zd words but %zd boxes", b, c, nw, PySequence_Fast_GET_SIZE(chunk_b)); goto done; }
                for (Py_ssize_t i = 0; i < nw; ++i) {
                    box = PySequence_Fast(PySequence_Fast_GET_ITEM(chunk_b, i), "a box must be a sequence of 4 numbers");
                    if (!box) goto done;
                    if (PySequence_Fast_GET_SIZE(box) != 4) { PyErr_SetString(PyExc_ValueError, "a box must have 4 coordinates"); goto done; }
                    double v[4];
                    for (int e = 0; e < 4; ++e) {
                        v[e] = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(box, e));
                        if (v[e] == -1.0 && PyErr_Occurred()) goto done;
                    }
                    if (buf_push(&boxes, v, 32)) goto done;
                    Py_CLEAR(box);
                }
                Py_CLEAR(chunk_b);
            }
            Py_CLEAR(chunk_w);
        }
        Py_CLEAR(doc_w); Py_CLEAR(doc_b);
    }
    result = Py_BuildValue("(y#y#y#y#)", nwords.p ? nwords.p : "", (Py_ssize_t)nwords.n, ntok.p ? ntok.p : "", (Py_ssize_t)ntok.n,
                           ids.p ? ids.p : "", (Py_ssize_t)ids.n, boxes.p ? boxes.p : "", (Py_ssize_t)boxes.n);
done:
    Py_XDECREF(docs_w); Py_XDECREF(docs_b); Py_XDECREF(doc_w); Py_XDECREF(doc_b); Py_XDECREF(chunk_w); Py_XDECREF(chunk_b); Py_XDECREF(box);
    PyMem_Free(nwords.p); PyMem_Free(ntok.p); PyMem_Free(ids.p); PyMem_Free(boxes.p);
    return result;
}

static PyMethodDef methods[] = {
    {"flatten_docs", flatten_docs, METH_VARARGS,
     "flatten_docs(words_text_chunks, words_box_chunks, tokenize, cache) -> (chunk_nwords i64, word_ntok i32, tok_ids i32, boxes f64) bytes"},
    {"gather_s0", gather_s0, METH_VARARGS,
     "gather_s0(hits, words_text_chunks, words_box_chunks, layout_labels_chunks, images, page_indices, make_patch) -> "
     "(text, bbox, labels, words, boxes, word_labels, patches, pages), each [B][k]"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_hostlists", "nested-list view of the retrieved chunks (host logic)", -1, methods};

PyMODINIT_FUNC PyInit__hostlists(void) {
    s_space = PyUnicode_InternFromString(" ");
    s_width = PyUnicode_InternFromString("width");
    s_height = PyUnicode_InternFromString("height");
    if (!s_space || !s_width || !s_height) return NULL;
    return PyModule_Create(&module);
}
