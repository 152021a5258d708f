/*
 * _hostglue — CPython extension: turns the int32 record block written by peneo_decode_resolve into the
 * Python objects pipeline/decode.py returns (ZeningLin/PEneo pipeline/decode.py:205-212, 353-378):
 *
 *   (kv_pairs, lines, LE dict, EL-head dict, EL-tail dict, LG-head dict, LG-tail dict)
 *
 * This is the only host work left on the decode path — joining token strings, merging boxes and
 * building the ordered dicts — and it needs Python objects by definition; doing it through the C API
 * instead of interpreted loops keeps it off the critical path of the GPU pipeline.
 *
 * Record layout (include/peneo_b200.h, peneo_decode_resolve):
 *   header[16]: n_le, n_lgh, n_lgt, n_elh, n_elt, n_kv ; le[2n] ; lgh[2n] ; lgt[2n] ; elh[2cap] ; elt[2cap] ;
 *   kv[4cap] = (key_head, value_head, n_key_lines, n_value_lines)
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
  PyObject** ints; /* cached PyLong objects 0..n-1 (borrowed from this array, INCREF on use) */
  Py_ssize_t n;
} IntCache;

static PyObject* cached_int(IntCache* c, long v) {
  if (v >= 0 && v < c->n) {
    if (!c->ints[v]) {
      c->ints[v] = PyLong_FromLong(v);
      if (!c->ints[v]) return NULL;
    }
    Py_INCREF(c->ints[v]);
    return c->ints[v];
  }
  return PyLong_FromLong(v);
}

/* {head: tail} in record order */
static PyObject* build_map(IntCache* c, const int32_t* pairs, int cnt) {
  PyObject* d = PyDict_New();
  if (!d) return NULL;
  for (int r = 0; r < cnt; ++r) {
    PyObject* k = cached_int(c, pairs[2 * r]);
    PyObject* v = cached_int(c, pairs[2 * r + 1]);
    if (!k || !v || PyDict_SetItem(d, k, v) < 0) {
      Py_XDECREF(k), Py_XDECREF(v), Py_DECREF(d);
      return NULL;
    }
    Py_DECREF(k), Py_DECREF(v);
  }
  return d;
}

/* {head: [tails...]} in record order (dict.setdefault(head, []).append(tail)).  Heads inside [0, n) are
 * counted first so that every list is allocated once at its final size; `slot` / `fill` are n-entry
 * scratch arrays. */
static PyObject* build_multimap(IntCache* c, const int32_t* pairs, int cnt, int n, int32_t* slot, int32_t* fill) {
  int distinct = 0;
  for (int t = 0; t < n; ++t) slot[t] = 0, fill[t] = 0;
  for (int r = 0; r < cnt; ++r) {
    const int h = pairs[2 * r];
    if (h < 0 || h >= n) {
      PyErr_SetString(PyExc_ValueError, "corrupt decode record: head outside the document");
      return NULL;
    }
    if (slot[h]++ == 0) ++distinct;
  }
  PyObject* d = PyDict_New();
  if (!d) return NULL;
  PyObject** lists = (PyObject**)calloc((size_t)n, sizeof(PyObject*)); /* borrowed (the dict owns them) */
  if (!lists) {
    Py_DECREF(d);
    PyErr_NoMemory();
    return NULL;
  }
  for (int r = 0; r < cnt; ++r) {
    const int h = pairs[2 * r];
    PyObject* v = cached_int(c, pairs[2 * r + 1]);
    if (!v) goto fail;
    if (!lists[h]) {
      PyObject* k = cached_int(c, h);
      PyObject* lst = k ? PyList_New(slot[h]) : NULL;
      if (!lst || PyDict_SetItem(d, k, lst) < 0) {
        Py_XDECREF(k), Py_XDECREF(lst), Py_DECREF(v);
        goto fail;
      }
      Py_DECREF(k), Py_DECREF(lst);
      lists[h] = lst;
      /* slots not yet filled hold NULL; they are all assigned before the dict is returned */
    }
    PyList_SET_ITEM(lists[h], fill[h]++, v);
  }
  free(lists);
  return d;
fail:
  /* lists with NULL slots must not be seen by anyone: truncate them before dropping the dict */
  for (int t = 0; t < n; ++t)
    if (lists[t]) Py_SET_SIZE((PyListObject*)lists[t], fill[t]);
  free(lists);
  Py_DECREF(d);
  return NULL;
}

/* data/data_utils.py:62-76 merge_bbox on bbox[h : t + 1]: [min x0, min y0, max x1, max y1], picking the
 * original objects the way Python's min() / max() do (first extreme element). */
static int merge_into(PyObject* box, PyObject* best[4], int first) {
  PyObject* fast = PySequence_Fast(box, "bbox entries must be sequences");
  if (!fast) return -1;
  if (PySequence_Fast_GET_SIZE(fast) != 4) {
    Py_DECREF(fast);
    PyErr_SetString(PyExc_ValueError, "bbox entries must have 4 values");
    return -1;
  }
  for (int q = 0; q < 4; ++q) {
    PyObject* v = PySequence_Fast_GET_ITEM(fast, q);
    int take = first;
    if (!first) {
      take = PyObject_RichCompareBool(v, best[q], q < 2 ? Py_LT : Py_GT);
      if (take < 0) {
        Py_DECREF(fast);
        return -1;
      }
    }
    if (take) {
      Py_INCREF(v);
      Py_XDECREF(best[q]);
      best[q] = v;
    }
  }
  Py_DECREF(fast);
  return 0;
}

static PyObject* box_from(PyObject* best[4]) {
  PyObject* out = PyList_New(4);
  if (!out) return NULL;
  for (int q = 0; q < 4; ++q) {
    Py_INCREF(best[q]);
    PyList_SET_ITEM(out, q, best[q]);
  }
  return out;
}

/* merged box over the segments (h, t) of a chain; *_segs as flat int pairs */
static PyObject* merge_segments(PyObject* bbox, const int32_t* segs, int nseg) {
  PyObject* best[4] = {NULL, NULL, NULL, NULL};
  int first = 1;
  Py_ssize_t len = PySequence_Size(bbox);
  if (len < 0) return NULL;
  for (int s = 0; s < nseg; ++s) {
    Py_ssize_t lo = segs[2 * s], hi = (Py_ssize_t)segs[2 * s + 1] + 1;
    if (hi > len) hi = len;
    if (lo >= hi) { /* merge_bbox([]) raises in the reference: zip(*[]) cannot be unpacked into 4 names */
      for (int q = 0; q < 4; ++q) Py_XDECREF(best[q]);
      PyErr_SetString(PyExc_ValueError, "not enough values to unpack (expected 4, got 0)");
      return NULL;
    }
    /* Python evaluates merge_bbox per segment and then merges the results; min/max are associative and
     * "first extreme wins" is preserved by visiting rows in order */
    for (Py_ssize_t r = lo; r < hi; ++r) {
      PyObject* box = PySequence_GetItem(bbox, r);
      if (!box || merge_into(box, best, first) < 0) {
        Py_XDECREF(box);
        for (int q = 0; q < 4; ++q) Py_XDECREF(best[q]);
        return NULL;
      }
      Py_DECREF(box);
      first = 0;
    }
  }
  PyObject* out = box_from(best);
  for (int q = 0; q < 4; ++q) Py_XDECREF(best[q]);
  return out;
}

/* "".join(text[h : t + 1]) over all segments, appended into one token list */
static int extend_tokens(PyObject* toks, PyObject* text, Py_ssize_t text_len, int h, int t) {
  Py_ssize_t hi = (Py_ssize_t)t + 1;
  if (hi > text_len) hi = text_len;
  for (Py_ssize_t r = h; r < hi; ++r) {
    PyObject* it = PyList_GET_ITEM(text, r); /* borrowed */
    if (PyList_Append(toks, it) < 0) return -1;
  }
  return 0;
}

static PyObject* join_tokens(PyObject* empty, PyObject* toks, int strip) {
  PyObject* s = PyUnicode_Join(empty, toks);
  if (!s || !strip) return s;
  PyObject* r = PyObject_CallMethod(s, "strip", NULL);
  Py_DECREF(s);
  return r;
}

static PyObject* assemble_doc(const int32_t* rec, int n, int cap, PyObject* text, PyObject* bbox, IntCache* cache,
                              int32_t* le_tail, int32_t* lg_next, int32_t* segbuf, PyObject* empty) {
  const int32_t* hdr = rec;
  const int32_t* le = rec + 16;
  const int32_t* lgh = le + 2 * (Py_ssize_t)n;
  const int32_t* lgt = lgh + 2 * (Py_ssize_t)n;
  const int32_t* elh = lgt + 2 * (Py_ssize_t)n;
  const int32_t* elt = elh + 2 * (Py_ssize_t)cap;
  const int32_t* kv = elt + 2 * (Py_ssize_t)cap;
  const int n_le = hdr[0], n_lgh = hdr[1], n_lgt = hdr[2], n_elh = hdr[3], n_elt = hdr[4], n_kv = hdr[5];
  if (n_le < 0 || n_le > n || n_lgh < 0 || n_lgh > n || n_lgt < 0 || n_lgt > n || n_elh < 0 || n_elh > cap || n_elt < 0 ||
      n_elt > cap || n_kv < 0 || n_kv > cap) {
    PyErr_SetString(PyExc_ValueError, "corrupt decode record header");
    return NULL;
  }
  if (!PyList_Check(text)) {
    PyErr_SetString(PyExc_TypeError, "text must be a list of str");
    return NULL;
  }
  const Py_ssize_t text_len = PyList_GET_SIZE(text);
  PyObject *d_le = NULL, *d_lgh = NULL, *d_lgt = NULL, *d_elh = NULL, *d_elt = NULL, *lines = NULL, *pairs = NULL;
  PyObject* toks = NULL;
  if (!(d_le = build_map(cache, le, n_le)) || !(d_lgh = build_map(cache, lgh, n_lgh)) ||
      !(d_lgt = build_map(cache, lgt, n_lgt)) || !(d_elh = build_multimap(cache, elh, n_elh, n, le_tail, lg_next)) ||
      !(d_elt = build_multimap(cache, elt, n_elt, n, le_tail, lg_next)))
    goto fail;
  for (int t = 0; t < n; ++t) le_tail[t] = -1, lg_next[t] = -1;
  for (int r = 0; r < n_le; ++r)
    if (le[2 * r] >= 0 && le[2 * r] < n) le_tail[le[2 * r]] = le[2 * r + 1];
  for (int r = 0; r < n_lgh; ++r)
    if (lgh[2 * r] >= 0 && lgh[2 * r] < n) lg_next[lgh[2 * r]] = lgh[2 * r + 1];

  /* lines: in LE dict order */
  if (!(lines = PyList_New(n_le))) goto fail;
  for (int r = 0; r < n_le; ++r) {
    if (!(toks = PyList_GetSlice(text, le[2 * r], (Py_ssize_t)le[2 * r + 1] + 1))) goto fail;
    PyObject* s = join_tokens(empty, toks, 0);
    Py_CLEAR(toks);
    if (!s) goto fail;
    if (bbox != Py_None) {
      int32_t seg[2] = {le[2 * r], le[2 * r + 1]};
      PyObject* box = merge_segments(bbox, seg, 1);
      if (!box) {
        Py_DECREF(s);
        goto fail;
      }
      PyObject* tup = PyTuple_Pack(2, s, box);
      Py_DECREF(s), Py_DECREF(box);
      if (!tup) goto fail;
      s = tup;
    }
    PyList_SET_ITEM(lines, r, s);
  }

  /* key-value pairs: re-walk the chains the kernel validated */
  if (!(pairs = PyList_New(n_kv))) goto fail;
  for (int r = 0; r < n_kv; ++r) {
    PyObject* parts[4] = {NULL, NULL, NULL, NULL}; /* ktxt, vtxt, kbox, vbox */
    for (int side = 0; side < 2; ++side) {
      int cur = kv[4 * r + side];
      const int nseg = kv[4 * r + 2 + side];
      if (nseg < 1 || nseg > 1001 || cur < 0 || cur >= n) {
        PyErr_SetString(PyExc_ValueError, "corrupt key-value record");
        goto fail_parts;
      }
      if (!(toks = PyList_New(0))) goto fail_parts;
      for (int s = 0; s < nseg; ++s) {
        if (s > 0) cur = lg_next[cur];
        if (cur < 0 || cur >= n || le_tail[cur] < 0) {
          PyErr_SetString(PyExc_ValueError, "key-value chain leaves the line map");
          goto fail_parts;
        }
        segbuf[2 * s] = cur, segbuf[2 * s + 1] = le_tail[cur];
        if (extend_tokens(toks, text, text_len, cur, le_tail[cur]) < 0) goto fail_parts;
      }
      parts[side] = join_tokens(empty, toks, 1);
      Py_CLEAR(toks);
      if (!parts[side]) goto fail_parts;
      if (bbox != Py_None) {
        parts[2 + side] = merge_segments(bbox, segbuf, nseg);
        if (!parts[2 + side]) goto fail_parts;
      }
    }
    {
      PyObject* tup = bbox != Py_None ? PyTuple_Pack(4, parts[0], parts[1], parts[2], parts[3])
                                      : PyTuple_Pack(2, parts[0], parts[1]);
      for (int q = 0; q < 4; ++q) Py_XDECREF(parts[q]);
      if (!tup) goto fail;
      PyList_SET_ITEM(pairs, r, tup);
    }
    continue;
  fail_parts:
    for (int q = 0; q < 4; ++q) Py_XDECREF(parts[q]);
    goto fail;
  }
  {
    PyObject* out = PyTuple_Pack(7, pairs, lines, d_le, d_elh, d_elt, d_lgh, d_lgt);
    Py_DECREF(pairs), Py_DECREF(lines), Py_DECREF(d_le), Py_DECREF(d_elh), Py_DECREF(d_elt), Py_DECREF(d_lgh),
        Py_DECREF(d_lgt);
    return out;
  }
fail:
  Py_XDECREF(toks);
  Py_XDECREF(pairs), Py_XDECREF(lines), Py_XDECREF(d_le), Py_XDECREF(d_elh), Py_XDECREF(d_elt), Py_XDECREF(d_lgh),
      Py_XDECREF(d_lgt);
  return NULL;
}

/* assemble(records, doc_ints, n, cap, docs, texts, bboxes) -> list of 7-tuples
 *   records : C-contiguous int32 buffer holding at least max(docs)+1 record blocks of doc_ints ints
 *   docs    : sequence of document indices (rows of `records`) to assemble
 *   texts   : sequence aligned with `docs`: list[str] per document
 *   bboxes  : None, or a sequence aligned with `docs` of (list of 4-sequences | None)            */
static PyObject* py_assemble(PyObject* self, PyObject* args) {
  Py_buffer view;
  Py_ssize_t doc_ints;
  int n, cap;
  PyObject *docs, *texts, *bboxes;
  if (!PyArg_ParseTuple(args, "y*niiOOO", &view, &doc_ints, &n, &cap, &docs, &texts, &bboxes)) return NULL;
  PyObject* result = NULL;
  PyObject *fdocs = NULL, *ftexts = NULL, *fboxes = NULL, *empty = NULL;
  IntCache cache = {NULL, 0};
  int32_t *le_tail = NULL, *lg_next = NULL, *segbuf = NULL;
  if (n < 1 || cap < 1 || doc_ints < 16 + 6 * (Py_ssize_t)n + 8 * (Py_ssize_t)cap) {
    PyErr_SetString(PyExc_ValueError, "bad record geometry");
    goto done;
  }
  if (!(fdocs = PySequence_Fast(docs, "docs must be a sequence")) ||
      !(ftexts = PySequence_Fast(texts, "texts must be a sequence")))
    goto done;
  if (bboxes != Py_None && !(fboxes = PySequence_Fast(bboxes, "bboxes must be a sequence or None"))) goto done;
  const Py_ssize_t count = PySequence_Fast_GET_SIZE(fdocs);
  if (PySequence_Fast_GET_SIZE(ftexts) != count || (fboxes && PySequence_Fast_GET_SIZE(fboxes) != count)) {
    PyErr_SetString(PyExc_ValueError, "docs / texts / bboxes length mismatch");
    goto done;
  }
  cache.n = n;
  cache.ints = (PyObject**)calloc((size_t)n, sizeof(PyObject*));
  le_tail = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  lg_next = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
  segbuf = (int32_t*)malloc(sizeof(int32_t) * 2 * 1001);
  empty = PyUnicode_FromString("");
  if (!cache.ints || !le_tail || !lg_next || !segbuf || !empty) {
    PyErr_NoMemory();
    goto done;
  }
  if (!(result = PyList_New(count))) goto done;
  for (Py_ssize_t r = 0; r < count; ++r) {
    const Py_ssize_t b = PyLong_AsSsize_t(PySequence_Fast_GET_ITEM(fdocs, r));
    if (b < 0 || (b + 1) * doc_ints * (Py_ssize_t)sizeof(int32_t) > view.len) {
      if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "document index outside the record buffer");
      Py_CLEAR(result);
      goto done;
    }
    PyObject* bbox = fboxes ? PySequence_Fast_GET_ITEM(fboxes, r) : Py_None;
    PyObject* one = assemble_doc((const int32_t*)view.buf + b * doc_ints, n, cap, PySequence_Fast_GET_ITEM(ftexts, r), bbox,
                                 &cache, le_tail, lg_next, segbuf, empty);
    if (!one) {
      Py_CLEAR(result);
      goto done;
    }
    PyList_SET_ITEM(result, r, one);
  }
done:
  if (cache.ints) {
    for (Py_ssize_t i = 0; i < cache.n; ++i) Py_XDECREF(cache.ints[i]);
    free(cache.ints);
  }
  free(le_tail), free(lg_next), free(segbuf);
  Py_XDECREF(empty), Py_XDECREF(fdocs), Py_XDECREF(ftexts), Py_XDECREF(fboxes);
  PyBuffer_Release(&view);
  return result;
}

static PyMethodDef methods[] = {
    {"assemble", py_assemble, METH_VARARGS, "decode record blocks -> reference-shaped Python results"},
    {NULL, NULL, 0, NULL},
};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostglue", "peneo_b200 decode host glue", -1, methods};
PyMODINIT_FUNC PyInit__hostglue(void) { return PyModule_Create(&moddef); }
