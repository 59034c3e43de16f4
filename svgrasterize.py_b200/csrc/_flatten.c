/* _flatten.c -- CPython extension: Scene trees -> the flat scene arrays of include/svgr_b200.h (svgr_flat).
 *
 * Reads the reference's own object model directly (duck typed like svgrasterize_b200/sceneio.py, so it takes
 * this package's classes and svgrasterize.py's alike): Scene = tuple (tag, args) (svgrasterize.py:598-647),
 * Path.subpaths = [[(tag, points), ...], ...] (:896-913), paints = ndarray(4) | GradLinear | GradRadial
 * (:1544-1575), Transform.m = 3 x 3 ndarray (:509-570).  No arithmetic happens here beyond copying doubles: the
 * scene walk itself is csrc/encode_flat.cpp.
 *
 *   flatten(jobs) -> (dict of bytes, list of job indices left to the Python encoder)
 *     jobs: sequence of (scene, (width, height), linear_rgb)
 *
 * A scene that uses something encode_flat.cpp does not cover (objectBoundingBox clip / mask units, pattern paints,
 * unknown paints) is skipped and its index returned; everything it had appended is rolled back.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* numpy arrays are read through the numpy C API (a pointer and strides: ~10 ns) -- the generic buffer protocol
 * costs ~0.4 us per array (numpy builds a format string per export), 100 arrays per icon */
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

#include "../../include/svgr_b200.h"

typedef struct {
    char *p;
    size_t n, cap;
} Buf;

static int buf_put(Buf *b, const void *src, size_t n)
{
    if (b->n + n > b->cap) {
        size_t cap = b->cap ? b->cap * 2 : 4096;
        while (cap < b->n + n)
            cap *= 2;
        char *q = (char *)realloc(b->p, cap);
        if (!q) {
            PyErr_NoMemory();
            return -1;
        }
        b->p = q, b->cap = cap;
    }
    memcpy(b->p + b->n, src, n);
    b->n += n;
    return 0;
}

typedef struct {
    Buf seg_tag, seg_data, sub_off, path_off, tr, paints, stops, nodes, children, scenes, fes, fe_inputs;
    int64_t n_seg;
    int32_t n_sub, n_path, n_tr, n_paint, n_stop, n_node, n_child, n_scene, n_fe, n_fe_input;
    PyObject *path_ids, *paint_ids; /* id(object) -> index, per scene */
    PyObject *asarray;              /* numpy.ascontiguousarray, for inputs that are not float64 buffers */
    PyObject *float64;
    int unsupported;                /* set while walking: the scene needs the Python encoder */
} Ctx;

enum { NOT_SUPPORTED = 1 };

/* ---- reading doubles out of array-likes -------------------------------------------------------------- */
static int read_doubles(Ctx *c, PyObject *obj, double *out, Py_ssize_t want)
{
    if (PyArray_Check(obj)) {
        PyArrayObject *a = (PyArrayObject *)obj;
        const int nd = PyArray_NDIM(a);
        if (PyArray_TYPE(a) == NPY_DOUBLE && PyArray_ISNOTSWAPPED(a) && nd <= 2 && PyArray_SIZE(a) == want) {
            const char *base = (const char *)PyArray_DATA(a);
            const npy_intp *sh = PyArray_DIMS(a), *st = PyArray_STRIDES(a);
            if (nd == 0) {
                out[0] = *(const double *)base;
            } else if (nd == 1) {
                for (npy_intp i = 0; i < sh[0]; i++)
                    out[i] = *(const double *)(base + i * st[0]);
            } else {
                Py_ssize_t q = 0;
                for (npy_intp i = 0; i < sh[0]; i++)
                    for (npy_intp j = 0; j < sh[1]; j++)
                        out[q++] = *(const double *)(base + i * st[0] + j * st[1]);
            }
            return 0;
        }
    }
    Py_buffer v;
    if (PyObject_CheckBuffer(obj) && PyObject_GetBuffer(obj, &v, PyBUF_STRIDES | PyBUF_FORMAT) == 0) {
        const int is_d = v.format && v.format[0] == 'd' && v.format[1] == 0 && v.itemsize == 8;
        Py_ssize_t count = 1;
        for (int k = 0; k < v.ndim; k++)
            count *= v.shape[k];
        if (is_d && count == want && v.ndim <= 2) {
            const char *base = (const char *)v.buf;
            if (v.ndim == 0) {
                out[0] = *(const double *)base;
            } else if (v.ndim == 1) {
                for (Py_ssize_t i = 0; i < v.shape[0]; i++)
                    out[i] = *(const double *)(base + i * v.strides[0]);
            } else {
                Py_ssize_t q = 0;
                for (Py_ssize_t i = 0; i < v.shape[0]; i++)
                    for (Py_ssize_t j = 0; j < v.shape[1]; j++)
                        out[q++] = *(const double *)(base + i * v.strides[0] + j * v.strides[1]);
            }
            PyBuffer_Release(&v);
            return 0;
        }
        PyBuffer_Release(&v);
    } else {
        PyErr_Clear();
    }
    /* lists, other dtypes: numpy.ascontiguousarray(obj, dtype=float64) */
    PyObject *arr = PyObject_CallFunctionObjArgs(c->asarray, obj, c->float64, NULL);
    if (!arr)
        return -1;
    int rc = -1;
    if (PyObject_GetBuffer(arr, &v, PyBUF_CONTIG_RO) == 0) {
        if ((Py_ssize_t)(v.len / 8) == want) {
            memcpy(out, v.buf, (size_t)want * 8);
            rc = 0;
        } else {
            PyErr_Format(PyExc_ValueError, "expected %zd float values, got %zd", want, (Py_ssize_t)(v.len / 8));
        }
        PyBuffer_Release(&v);
    }
    Py_DECREF(arr);
    return rc;
}

static int read_transform(Ctx *c, PyObject *tr, double *m6)
{
    PyObject *m = PyObject_GetAttrString(tr, "m");
    if (!m)
        return -1;
    double m9[9];
    const int rc = read_doubles(c, m, m9, 9);
    Py_DECREF(m);
    if (rc == 0)
        memcpy(m6, m9, 48);
    return rc;
}

static int str_is(PyObject *s, const char *lit) { return PyUnicode_Check(s) && PyUnicode_CompareWithASCIIString(s, lit) == 0; }

/* ---- paths ----------------------------------------------------------------------------------------------- */
static long cached_index(PyObject *dict, PyObject *obj)
{
    PyObject *key = PyLong_FromVoidPtr(obj);
    if (!key)
        return -2;
    PyObject *val = PyDict_GetItemWithError(dict, key); /* borrowed */
    Py_DECREF(key);
    if (!val)
        return PyErr_Occurred() ? -2 : -1;
    return PyLong_AsLong(val);
}

static int cache_index(PyObject *dict, PyObject *obj, long idx)
{
    PyObject *key = PyLong_FromVoidPtr(obj), *val = PyLong_FromLong(idx);
    int rc = (key && val) ? PyDict_SetItem(dict, key, val) : -1;
    Py_XDECREF(key);
    Py_XDECREF(val);
    return rc;
}

static int put_segment(Ctx *c, long tag, PyObject *args)
{
    double row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tag == 3) { /* PATH_ARC: (center, rx, ry, phi, eta, eta_delta) */
        PyObject *seq = PySequence_Fast(args, "arc arguments must be a sequence");
        if (!seq)
            return -1;
        if (PySequence_Fast_GET_SIZE(seq) != 6) {
            Py_DECREF(seq);
            PyErr_SetString(PyExc_ValueError, "arc segment needs (center, rx, ry, phi, eta, eta_delta)");
            return -1;
        }
        int rc = read_doubles(c, PySequence_Fast_GET_ITEM(seq, 0), row, 2);
        for (int k = 1; k < 6 && rc == 0; k++) {
            row[1 + k] = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(seq, k));
            if (row[1 + k] == -1.0 && PyErr_Occurred())
                rc = -1;
        }
        Py_DECREF(seq);
        if (rc)
            return -1;
    } else if (tag == 0 || tag == 4 || tag == 5) {
        if (read_doubles(c, args, row, 4))
            return -1;
    } else if (tag == 1) {
        if (read_doubles(c, args, row, 6))
            return -1;
    } else if (tag == 2) {
        if (read_doubles(c, args, row, 8))
            return -1;
    } else {
        PyErr_Format(PyExc_ValueError, "unsupported path type: `%ld`", tag);
        return -1;
    }
    const uint8_t t = (uint8_t)tag;
    if (buf_put(&c->seg_tag, &t, 1) || buf_put(&c->seg_data, row, 64))
        return -1;
    c->n_seg++;
    return 0;
}

static long flatten_path(Ctx *c, PyObject *path)
{
    long idx = cached_index(c->path_ids, path);
    if (idx != -1)
        return idx; /* found, or -2 on error */
    /* a path that already holds flat arrays (the native path-data reader, sceneio): copy them as they are */
    {
        PyObject *enc = PyObject_GetAttrString(path, "_enc");
        if (!enc) {
            PyErr_Clear();
        } else if (PyTuple_Check(enc) && PyTuple_GET_SIZE(enc) == 3 && PyArray_Check(PyTuple_GET_ITEM(enc, 0)) &&
                   PyArray_Check(PyTuple_GET_ITEM(enc, 1)) && PyArray_Check(PyTuple_GET_ITEM(enc, 2))) {
            PyArrayObject *t = (PyArrayObject *)PyTuple_GET_ITEM(enc, 0), *d = (PyArrayObject *)PyTuple_GET_ITEM(enc, 1),
                          *o = (PyArrayObject *)PyTuple_GET_ITEM(enc, 2);
            const npy_intp n = PyArray_SIZE(t), ns = PyArray_SIZE(o) - 1;
            if (PyArray_TYPE(t) == NPY_UINT8 && PyArray_TYPE(d) == NPY_DOUBLE && PyArray_TYPE(o) == NPY_INT32 &&
                PyArray_IS_C_CONTIGUOUS(t) && PyArray_IS_C_CONTIGUOUS(d) && PyArray_IS_C_CONTIGUOUS(o) &&
                PyArray_SIZE(d) == 8 * n && ns >= 0) {
                const int32_t *off = (const int32_t *)PyArray_DATA(o);
                int ok = buf_put(&c->seg_tag, PyArray_DATA(t), (size_t)n) == 0 &&
                         buf_put(&c->seg_data, PyArray_DATA(d), (size_t)n * 64) == 0;
                for (npy_intp k = 1; ok && k <= ns; k++) {
                    const int32_t end = (int32_t)c->n_seg + off[k];
                    ok = buf_put(&c->sub_off, &end, 4) == 0;
                }
                Py_DECREF(enc);
                if (!ok)
                    return -2;
                c->n_seg += n;
                c->n_sub += (int32_t)ns;
                const int32_t end = c->n_sub;
                if (buf_put(&c->path_off, &end, 4))
                    return -2;
                idx = c->n_path++;
                return cache_index(c->path_ids, path, idx) ? -2 : idx;
            }
            Py_DECREF(enc);
        } else {
            Py_DECREF(enc);
        }
    }
    PyObject *subs = PyObject_GetAttrString(path, "subpaths");
    if (!subs)
        return -2;
    PyObject *sub_seq = PySequence_Fast(subs, "Path.subpaths must be a sequence");
    Py_DECREF(subs);
    if (!sub_seq)
        return -2;
    for (Py_ssize_t s = 0; s < PySequence_Fast_GET_SIZE(sub_seq); s++) {
        PyObject *segs = PySequence_Fast(PySequence_Fast_GET_ITEM(sub_seq, s), "a sub-path must be a sequence");
        if (!segs)
            goto fail;
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(segs);
        for (Py_ssize_t i = 0; i < n; i++) {
            PyObject *seg = PySequence_Fast_GET_ITEM(segs, i);
            if (!PyTuple_Check(seg) || PyTuple_GET_SIZE(seg) != 2) {
                Py_DECREF(segs);
                PyErr_SetString(PyExc_ValueError, "a path segment must be a (tag, points) tuple");
                goto fail;
            }
            const long tag = PyLong_AsLong(PyTuple_GET_ITEM(seg, 0));
            if ((tag == -1 && PyErr_Occurred()) || put_segment(c, tag, PyTuple_GET_ITEM(seg, 1))) {
                Py_DECREF(segs);
                goto fail;
            }
        }
        Py_DECREF(segs);
        if (n > 0) { /* empty sub-paths are dropped (the reference skips them: :933, :1117) */
            const int32_t end = (int32_t)c->n_seg;
            if (buf_put(&c->sub_off, &end, 4))
                goto fail;
            c->n_sub++;
        }
    }
    Py_DECREF(sub_seq);
    {
        const int32_t end = c->n_sub;
        if (buf_put(&c->path_off, &end, 4))
            return -2;
    }
    idx = c->n_path++;
    if (cache_index(c->path_ids, path, idx))
        return -2;
    return idx;
fail:
    Py_DECREF(sub_seq);
    return -2;
}

/* ---- paints ---------------------------------------------------------------------------------------------- */
static int put_stops(Ctx *c, PyObject *paint, svgr_flat_paint *rec)
{
    PyObject *stops = PyObject_GetAttrString(paint, "stops");
    if (!stops)
        return -1;
    PyObject *seq = PySequence_Fast(stops, "gradient stops must be a sequence");
    Py_DECREF(stops);
    if (!seq)
        return -1;
    rec->stop_off = c->n_stop;
    rec->stop_cnt = (int32_t)PySequence_Fast_GET_SIZE(seq);
    for (Py_ssize_t i = 0; i < PySequence_Fast_GET_SIZE(seq); i++) {
        PyObject *st = PySequence_Fast(PySequence_Fast_GET_ITEM(seq, i), "a gradient stop must be (offset, color)");
        if (!st)
            goto fail;
        svgr_flat_stop s;
        int rc = PySequence_Fast_GET_SIZE(st) == 2 ? 0 : -1;
        if (rc == 0) {
            s.offset = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(st, 0));
            if (s.offset == -1.0 && PyErr_Occurred())
                rc = -1;
        } else {
            PyErr_SetString(PyExc_ValueError, "a gradient stop must be (offset, color)");
        }
        if (rc == 0)
            rc = read_doubles(c, PySequence_Fast_GET_ITEM(st, 1), s.color, 4);
        Py_DECREF(st);
        if (rc || buf_put(&c->stops, &s, sizeof s))
            goto fail;
        c->n_stop++;
    }
    Py_DECREF(seq);
    return 0;
fail:
    Py_DECREF(seq);
    return -1;
}

static int truthy_attr(PyObject *obj, const char *name, int *out)
{
    PyObject *v = PyObject_GetAttrString(obj, name);
    if (!v)
        return -1;
    *out = PyObject_IsTrue(v);
    Py_DECREF(v);
    return *out < 0 ? -1 : 0;
}

static int optional_vec2(Ctx *c, PyObject *obj, const char *name, double *out, int *given)
{
    PyObject *v = PyObject_GetAttrString(obj, name);
    if (!v)
        return -1;
    *given = v != Py_None;
    int rc = *given ? read_doubles(c, v, out, 2) : 0;
    Py_DECREF(v);
    return rc;
}

/* -> paint index, -1 for None, -2 on error, -3 unsupported */
static long flatten_paint(Ctx *c, PyObject *paint)
{
    if (paint == Py_None)
        return -1;
    long idx = cached_index(c->paint_ids, paint);
    if (idx != -1)
        return idx;
    svgr_flat_paint rec;
    memset(&rec, 0, sizeof rec);
    rec.lin = -1;
    /* a failed attribute lookup raises and clears an AttributeError (~1 us): arrays are recognised by type first */
    const int is_array = PyArray_Check(paint);
    const int has_p0 = !is_array && PyObject_HasAttrString(paint, "p0") && PyObject_HasAttrString(paint, "p1");
    const int has_center = !is_array && !has_p0 && PyObject_HasAttrString(paint, "center") && PyObject_HasAttrString(paint, "radius");
    if (!has_p0 && !has_center) {
        /* solid colour: an array of 4 floats; anything else (patterns, unknown objects) goes to the Python encoder */
        if (!is_array && (PyObject_HasAttrString(paint, "scene") || !PyObject_CheckBuffer(paint)))
            return -3;
        Py_buffer v;
        if (PyObject_GetBuffer(paint, &v, PyBUF_STRIDES) != 0) {
            PyErr_Clear();
            return -3;
        }
        const int ok = v.ndim == 1 && v.shape[0] == 4;
        PyBuffer_Release(&v);
        if (!ok)
            return -3;
        rec.kind = 1;
        if (read_doubles(c, paint, rec.p, 4))
            return -2;
    } else {
        rec.kind = has_p0 ? 2 : 3;
        int bbox = 0;
        if (truthy_attr(paint, "bbox_units", &bbox))
            return -2;
        rec.bbox_units = bbox; /* completed on the device (svgr_bbox_job) */
        PyObject *spread = PyObject_GetAttrString(paint, "spread");
        if (!spread)
            return -2;
        rec.spread = str_is(spread, "pad") ? 0 : str_is(spread, "repeat") ? 1 : str_is(spread, "reflect") ? 2 : -1;
        Py_DECREF(spread);
        PyObject *lin = PyObject_GetAttrString(paint, "linear_rgb");
        if (!lin)
            return -2;
        rec.lin = lin == Py_None ? -1 : PyObject_IsTrue(lin);
        Py_DECREF(lin);
        PyObject *tr = PyObject_GetAttrString(paint, "transform");
        if (!tr)
            return -2;
        if (tr != Py_None) {
            PyObject *inv = PyObject_GetAttrString(tr, "invert"); /* numpy's inverse, cached on the Transform */
            int rc = inv ? read_transform(c, inv, rec.inv) : -1;
            Py_XDECREF(inv);
            if (rc) {
                Py_DECREF(tr);
                return -2;
            }
            rec.has_transform = 1;
        }
        Py_DECREF(tr);
        if (has_p0) {
            PyObject *p0 = PyObject_GetAttrString(paint, "p0"), *p1 = PyObject_GetAttrString(paint, "p1");
            int rc = (p0 && p1) ? (read_doubles(c, p0, rec.p, 2) || read_doubles(c, p1, rec.p + 2, 2)) : -1;
            Py_XDECREF(p0);
            Py_XDECREF(p1);
            if (rc)
                return -2;
        } else {
            PyObject *ctr = PyObject_GetAttrString(paint, "center"), *rad = PyObject_GetAttrString(paint, "radius");
            int rc = (ctr && rad) ? read_doubles(c, ctr, rec.p, 2) : -1;
            if (rc == 0) {
                rec.p[2] = PyFloat_AsDouble(rad);
                if (rec.p[2] == -1.0 && PyErr_Occurred())
                    rc = -1;
            }
            Py_XDECREF(ctr);
            Py_XDECREF(rad);
            int given = 0;
            if (rc || optional_vec2(c, paint, "fcenter", rec.p + 3, &given))
                return -2;
            rec.focal |= given ? 1 : 0;
            PyObject *fr = PyObject_GetAttrString(paint, "fradius");
            if (!fr)
                return -2;
            if (fr != Py_None) {
                rec.focal |= 2;
                rec.p[5] = PyFloat_AsDouble(fr);
                if (rec.p[5] == -1.0 && PyErr_Occurred()) {
                    Py_DECREF(fr);
                    return -2;
                }
            }
            Py_DECREF(fr);
        }
        if (put_stops(c, paint, &rec))
            return -2;
    }
    if (buf_put(&c->paints, &rec, sizeof rec))
        return -2;
    idx = c->n_paint++;
    if (cache_index(c->paint_ids, paint, idx))
        return -2;
    return idx;
}

/* ---- scene tree ------------------------------------------------------------------------------------------ */
static long put_node(Ctx *c, const svgr_flat_node *n, const int32_t *kids)
{
    svgr_flat_node rec = *n;
    rec.child_off = c->n_child;
    if (rec.child_cnt > 0) {
        if (buf_put(&c->children, kids, (size_t)rec.child_cnt * 4))
            return -2;
        c->n_child += rec.child_cnt;
    }
    if (buf_put(&c->nodes, &rec, sizeof rec))
        return -2;
    return c->n_node++;
}

/* -> node index, -2 on error (Python exception set), -3 the scene is left to the Python encoder */
/* Filter.filters = [(tag, attrs, input slots), ...] (svgrasterize.py:1750-1799) -> svgr_flat_fe records.
 * Returns 0, -2 on a Python error. */
static long flatten_filter(Ctx *c, PyObject *flt)
{
    PyObject *list = PyObject_GetAttrString(flt, "filters");
    if (!list)
        return -2;
    PyObject *seq = PySequence_Fast(list, "Filter.filters must be a sequence");
    Py_DECREF(list);
    if (!seq)
        return -2;
    long rc = 0;
    for (Py_ssize_t i = 0; i < PySequence_Fast_GET_SIZE(seq) && rc == 0; i++) {
        PyObject *prim = PySequence_Fast(PySequence_Fast_GET_ITEM(seq, i), "a filter primitive must be (tag, attrs, inputs)");
        if (!prim) {
            rc = -2;
            break;
        }
        PyObject *attrs = NULL, *inputs = NULL;
        svgr_flat_fe fe;
        memset(&fe, 0, sizeof fe);
        if (PySequence_Fast_GET_SIZE(prim) != 3) {
            PyErr_SetString(PyExc_ValueError, "a filter primitive must be (tag, attrs, inputs)");
            rc = -2;
        }
        if (rc == 0) {
            const long tag = PyLong_AsLong(PySequence_Fast_GET_ITEM(prim, 0));
            if (tag == -1 && PyErr_Occurred())
                rc = -2;
            fe.tag = (int32_t)tag;
        }
        if (rc == 0) {
            attrs = PySequence_Fast(PySequence_Fast_GET_ITEM(prim, 1), "filter attributes must be a sequence");
            inputs = PySequence_Fast(PySequence_Fast_GET_ITEM(prim, 2), "filter inputs must be a sequence");
            if (!attrs || !inputs)
                rc = -2;
        }
        if (rc == 0) {
            fe.in_off = c->n_fe_input, fe.n_in = (int32_t)PySequence_Fast_GET_SIZE(inputs);
            for (Py_ssize_t k = 0; k < PySequence_Fast_GET_SIZE(inputs) && rc == 0; k++) {
                const long slot = PyLong_AsLong(PySequence_Fast_GET_ITEM(inputs, k));
                const int32_t s32 = (int32_t)slot;
                if ((slot == -1 && PyErr_Occurred()) || buf_put(&c->fe_inputs, &s32, 4))
                    rc = -2;
                else
                    c->n_fe_input++;
            }
        }
        if (rc == 0) {
            const Py_ssize_t nat = PySequence_Fast_GET_SIZE(attrs);
            PyObject **at = PySequence_Fast_ITEMS(attrs);
            switch (fe.tag) {
            case 11: /* offset: (dx, dy) */
            case 10: /* morphology: (rx, ry, method) */
                if (nat < 2) {
                    PyErr_SetString(PyExc_ValueError, "filter primitive: missing attributes");
                    rc = -2;
                    break;
                }
                fe.a[0] = PyFloat_AsDouble(at[0]), fe.a[1] = PyFloat_AsDouble(at[1]);
                if (PyErr_Occurred())
                    rc = -2;
                if (fe.tag == 10)
                    fe.flag = nat > 2 && str_is(at[2], "max") ? 1 : nat > 2 && str_is(at[2], "min") ? 0 : -1;
                break;
            case 8: /* gaussian blur: (std_x, std_y or None) */
                if (nat < 2) {
                    PyErr_SetString(PyExc_ValueError, "filter primitive: missing attributes");
                    rc = -2;
                    break;
                }
                fe.a[0] = PyFloat_AsDouble(at[0]);
                fe.a[1] = at[1] == Py_None ? fe.a[0] : PyFloat_AsDouble(at[1]);
                if (PyErr_Occurred())
                    rc = -2;
                break;
            case 0: /* blend: rendered as a plain over-merge, with the reference's warning (:1877) */
                if (PyErr_WarnEx(PyExc_UserWarning, "feBlend is not properly supported", 1))
                    rc = -2;
                break;
            case 3: { /* composite: (mode,) -- a COMPOSE_* code or the four arithmetic coefficients */
                PyObject *mode = nat > 0 ? at[0] : Py_None;
                fe.flag = -1;
                if (PyTuple_Check(mode) && PyTuple_GET_SIZE(mode) == 4) {
                    fe.flag = 1;
                    for (int k = 0; k < 4; k++)
                        fe.a[k] = PyFloat_AsDouble(PyTuple_GET_ITEM(mode, k));
                    if (PyErr_Occurred())
                        rc = -2;
                } else if (!PyBool_Check(mode) && (PyIndex_Check(mode) || PyFloat_Check(mode))) {
                    /* `mode in (0, 1, 2, 3, 4)`: ints of any kind, and floats that equal one of them */
                    const double m = PyFloat_Check(mode) ? PyFloat_AS_DOUBLE(mode) : (double)PyNumber_AsSsize_t(mode, NULL);
                    if (PyErr_Occurred())
                        PyErr_Clear();
                    else if (m == 0.0 || m == 1.0 || m == 2.0 || m == 3.0 || m == 4.0)
                        fe.flag = 0, fe.a[0] = m;
                }
                break;
            }
            case 1: { /* colour matrix: (matrix,) */
                PyObject *m = nat > 0 ? at[0] : Py_None;
                if (PyArray_Check(m) && PyArray_NDIM((PyArrayObject *)m) == 2 && PyArray_DIM((PyArrayObject *)m, 0) == 4 &&
                    PyArray_DIM((PyArrayObject *)m, 1) == 5) {
                    /* np.asarray(matrix, dtype=np.float32): through float64 is the same value for every real dtype */
                    if (read_doubles(c, m, fe.a, 20))
                        rc = -2;
                    fe.flag = 1;
                } else if (PyErr_WarnFormat(PyExc_UserWarning, 1, "invalid color matrix: %S", m)) {
                    rc = -2;
                }
                break;
            }
            default: /* merge (9) has no attributes; unknown tags are the walk's ValueError */
                break;
            }
        }
        if (rc == 0) {
            if (buf_put(&c->fes, &fe, sizeof fe))
                rc = -2;
            else
                c->n_fe++;
        }
        Py_XDECREF(attrs);
        Py_XDECREF(inputs);
        Py_DECREF(prim);
    }
    Py_DECREF(seq);
    return rc;
}

static long flatten_node(Ctx *c, PyObject *scene, int depth)
{
    if (depth > 2000) {
        PyErr_SetString(PyExc_RecursionError, "scene tree too deep");
        return -2;
    }
    if (!PyTuple_Check(scene) || PyTuple_GET_SIZE(scene) != 2) {
        PyErr_SetString(PyExc_ValueError, "a scene must be a (tag, args) tuple");
        return -2;
    }
    const long tag = PyLong_AsLong(PyTuple_GET_ITEM(scene, 0));
    if (tag == -1 && PyErr_Occurred())
        return -2;
    PyObject *args = PyTuple_GET_ITEM(scene, 1);
    svgr_flat_node n;
    memset(&n, 0, sizeof n);
    n.tag = (int32_t)tag;
    if (tag == 2) { /* group: args is the sequence of children */
        PyObject *seq = PySequence_Fast(args, "group children must be a sequence");
        if (!seq)
            return -2;
        const Py_ssize_t cnt = PySequence_Fast_GET_SIZE(seq);
        int32_t *kids = (int32_t *)malloc((size_t)(cnt ? cnt : 1) * 4);
        if (!kids) {
            Py_DECREF(seq);
            PyErr_NoMemory();
            return -2;
        }
        long rc = 0;
        for (Py_ssize_t i = 0; i < cnt; i++) {
            rc = flatten_node(c, PySequence_Fast_GET_ITEM(seq, i), depth + 1);
            if (rc < 0)
                break;
            kids[i] = (int32_t)rc;
        }
        Py_DECREF(seq);
        if (rc >= 0) {
            n.child_cnt = (int32_t)cnt;
            rc = put_node(c, &n, kids);
        }
        free(kids);
        return rc;
    }
    PyObject *seq = PySequence_Fast(args, "scene arguments must be a sequence");
    if (!seq)
        return -2;
    const Py_ssize_t na = PySequence_Fast_GET_SIZE(seq);
    PyObject **a = PySequence_Fast_ITEMS(seq);
    long rc = -2;
    int32_t kids[2];
    if (tag == 0 && na == 3) { /* fill: (path, paint, fill_rule) */
        const long path = flatten_path(c, a[0]);
        const long paint = path >= 0 ? flatten_paint(c, a[1]) : -2;
        if (path >= 0 && paint >= -1) {
            n.a = (int32_t)path, n.b = (int32_t)paint;
            n.c = (a[2] == Py_None || str_is(a[2], "nonzero")) ? 0 : str_is(a[2], "evenodd") ? 1 : -1;
            rc = put_node(c, &n, NULL);
        } else {
            rc = paint == -3 ? -3 : -2;
        }
    } else if (tag == 1 && na == 5) { /* stroke: (path, paint, width, linecap, linejoin) */
        const long path = flatten_path(c, a[0]);
        const long paint = path >= 0 ? flatten_paint(c, a[1]) : -2;
        const double width = PyFloat_AsDouble(a[2]);
        if (path >= 0 && paint >= -1 && !(width == -1.0 && PyErr_Occurred())) {
            n.a = (int32_t)path, n.b = (int32_t)paint, n.f[0] = width;
            n.c = (a[3] == Py_None || str_is(a[3], "butt")) ? 0 : str_is(a[3], "round") ? 1 : str_is(a[3], "square") ? 2 : -1;
            n.d = (a[4] == Py_None || str_is(a[4], "miter")) ? 0 : str_is(a[4], "round") ? 1 : str_is(a[4], "bevel") ? 2 : 3;
            rc = put_node(c, &n, NULL);
        } else {
            rc = paint == -3 ? -3 : -2;
        }
    } else if (tag == 3 && na == 2) { /* opacity: (scene, value) */
        rc = flatten_node(c, a[0], depth + 1);
        if (rc >= 0) {
            kids[0] = (int32_t)rc;
            n.child_cnt = 1;
            n.f[0] = PyFloat_AsDouble(a[1]);
            rc = (n.f[0] == -1.0 && PyErr_Occurred()) ? -2 : put_node(c, &n, kids);
        }
    } else if ((tag == 4 || tag == 5) && na == 3) { /* clip / mask: (scene, other, bbox_units) */
        const int bbox = PyObject_IsTrue(a[2]);
        if (bbox < 0) {
            rc = -2;
        } else if (bbox) {
            rc = -3;
        } else {
            rc = flatten_node(c, a[0], depth + 1);
            if (rc >= 0) {
                kids[0] = (int32_t)rc;
                rc = flatten_node(c, a[1], depth + 1);
            }
            if (rc >= 0) {
                kids[1] = (int32_t)rc;
                n.child_cnt = 2;
                rc = put_node(c, &n, kids);
            }
        }
    } else if (tag == 6 && na == 2) { /* transform: (scene, transform) */
        double m6[6];
        rc = flatten_node(c, a[0], depth + 1);
        if (rc >= 0) {
            kids[0] = (int32_t)rc;
            if (read_transform(c, a[1], m6) || buf_put(&c->tr, m6, 48)) {
                rc = -2;
            } else {
                n.a = c->n_tr++;
                n.child_cnt = 1;
                rc = put_node(c, &n, kids);
            }
        }
    } else if (tag == 7 && na == 2) { /* filter: (scene, Filter) */
        rc = flatten_node(c, a[0], depth + 1);
        if (rc >= 0) {
            kids[0] = (int32_t)rc;
            const int32_t first = c->n_fe;
            rc = flatten_filter(c, a[1]);
            if (rc >= 0) {
                n.a = first, n.b = c->n_fe - first;
                n.child_cnt = 1;
                rc = put_node(c, &n, kids);
            }
        }
    } else {
        PyErr_Format(PyExc_ValueError, "unhandled scene type: %ld", tag);
        rc = -2;
    }
    Py_DECREF(seq);
    return rc;
}

static PyObject *take_bytes(Buf *b)
{
    PyObject *o = PyBytes_FromStringAndSize(b->p ? b->p : "", (Py_ssize_t)b->n);
    free(b->p);
    b->p = NULL, b->n = b->cap = 0;
    return o;
}

static PyObject *py_flatten(PyObject *self, PyObject *args)
{
    (void)self;
    PyObject *jobs;
    if (!PyArg_ParseTuple(args, "O", &jobs))
        return NULL;
    PyObject *seq = PySequence_Fast(jobs, "jobs must be a sequence of (scene, (width, height), linear_rgb)");
    if (!seq)
        return NULL;
    Ctx c;
    memset(&c, 0, sizeof c);
    PyObject *result = NULL, *skipped = PyList_New(0), *np = PyImport_ImportModule("numpy");
    c.path_ids = PyDict_New(), c.paint_ids = PyDict_New();
    if (!skipped || !np || !c.path_ids || !c.paint_ids)
        goto done;
    c.asarray = PyObject_GetAttrString(np, "ascontiguousarray");
    c.float64 = PyObject_GetAttrString(np, "float64");
    if (!c.asarray || !c.float64)
        goto done;
    {
        const int32_t zero = 0;
        if (buf_put(&c.sub_off, &zero, 4) || buf_put(&c.path_off, &zero, 4))
            goto done;
    }
    for (Py_ssize_t j = 0; j < PySequence_Fast_GET_SIZE(seq); j++) {
        PyObject *job = PySequence_Fast(PySequence_Fast_GET_ITEM(seq, j), "a job must be (scene, (width, height), linear_rgb)");
        if (!job)
            goto done;
        if (PySequence_Fast_GET_SIZE(job) != 3) {
            Py_DECREF(job);
            PyErr_SetString(PyExc_ValueError, "a job must be (scene, (width, height), linear_rgb)");
            goto done;
        }
        PyObject *size = PySequence_Fast(PySequence_Fast_GET_ITEM(job, 1), "size must be (width, height)");
        if (!size || PySequence_Fast_GET_SIZE(size) != 2) {
            Py_XDECREF(size);
            Py_DECREF(job);
            if (!PyErr_Occurred())
                PyErr_SetString(PyExc_ValueError, "size must be (width, height)");
            goto done;
        }
        svgr_flat_scene sc;
        const double w = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(size, 0)), h = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(size, 1));
        Py_DECREF(size);
        sc.linear_rgb = PyObject_IsTrue(PySequence_Fast_GET_ITEM(job, 2));
        if (PyErr_Occurred()) {
            Py_DECREF(job);
            goto done;
        }
        sc.width = (int32_t)w, sc.height = (int32_t)h; /* int(w), int(h): truncation, as in :3860 */
        /* remember where this scene starts: an unsupported one is rolled back */
        const Ctx mark = c;
        PyDict_Clear(c.path_ids), PyDict_Clear(c.paint_ids);
        const long root = flatten_node(&c, PySequence_Fast_GET_ITEM(job, 0), 0);
        Py_DECREF(job);
        if (root == -2)
            goto done;
        if (root == -3) {
            c.seg_tag.n = mark.seg_tag.n, c.seg_data.n = mark.seg_data.n, c.sub_off.n = mark.sub_off.n;
            c.path_off.n = mark.path_off.n, c.tr.n = mark.tr.n, c.paints.n = mark.paints.n, c.stops.n = mark.stops.n;
            c.nodes.n = mark.nodes.n, c.children.n = mark.children.n, c.fes.n = mark.fes.n, c.fe_inputs.n = mark.fe_inputs.n;
            c.n_seg = mark.n_seg, c.n_sub = mark.n_sub, c.n_path = mark.n_path, c.n_tr = mark.n_tr;
            c.n_paint = mark.n_paint, c.n_stop = mark.n_stop, c.n_node = mark.n_node, c.n_child = mark.n_child;
            c.n_fe = mark.n_fe, c.n_fe_input = mark.n_fe_input;
            PyObject *idx = PyLong_FromSsize_t(j);
            const int rc = idx ? PyList_Append(skipped, idx) : -1;
            Py_XDECREF(idx);
            if (rc)
                goto done;
            continue;
        }
        sc.root = (int32_t)root;
        if (buf_put(&c.scenes, &sc, sizeof sc))
            goto done;
        c.n_scene++;
    }
    {
        PyObject *d = PyDict_New();
        if (!d)
            goto done;
        struct {
            const char *name;
            Buf *b;
        } items[] = {{"seg_tag", &c.seg_tag}, {"seg_data", &c.seg_data}, {"sub_off", &c.sub_off}, {"path_off", &c.path_off},
                     {"tr", &c.tr}, {"paints", &c.paints}, {"stops", &c.stops}, {"nodes", &c.nodes},
                     {"children", &c.children}, {"scenes", &c.scenes}, {"fes", &c.fes}, {"fe_inputs", &c.fe_inputs}};
        int ok = 1;
        for (size_t i = 0; i < sizeof items / sizeof items[0] && ok; i++) {
            PyObject *b = take_bytes(items[i].b);
            ok = b && PyDict_SetItemString(d, items[i].name, b) == 0;
            Py_XDECREF(b);
        }
        if (!ok) {
            Py_DECREF(d);
            goto done;
        }
        result = PyTuple_Pack(2, d, skipped);
        Py_DECREF(d);
    }
done:
    free(c.seg_tag.p), free(c.seg_data.p), free(c.sub_off.p), free(c.path_off.p), free(c.tr.p);
    free(c.paints.p), free(c.stops.p), free(c.nodes.p), free(c.children.p), free(c.scenes.p);
    free(c.fes.p), free(c.fe_inputs.p);
    Py_XDECREF(c.path_ids);
    Py_XDECREF(c.paint_ids);
    Py_XDECREF(c.asarray);
    Py_XDECREF(c.float64);
    Py_XDECREF(np);
    Py_XDECREF(skipped);
    Py_DECREF(seq);
    return result;
}

static PyMethodDef methods[] = {
    {"flatten", py_flatten, METH_VARARGS,
     "flatten(jobs) -> (dict of bytes in the svgr_flat layout, indices of the jobs left to the Python encoder)"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_svgr_flatten", "Scene trees -> flat scene arrays", -1, methods,
                                    NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__svgr_flatten(void)
{
    import_array();
    return PyModule_Create(&module);
}
