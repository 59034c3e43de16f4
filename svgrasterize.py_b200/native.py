"""The scene encoder in native code: Scene objects -> flat scene arrays (csrc/_flatten.c, a
CPython extension that reads the object model directly) -> scene program (csrc/encode_flat.cpp, the walk of
Scene.render, svgrasterize.py:649-752, inside libsvgr_b200.so).

    prog = native.encode_batch([(scene, (w, h), linear_rgb), ...])
    engine.render(prog, out=...)            # or render_png

Scenes that use what the native walk does not cover (objectBoundingBox units, pattern paints, filters) are
encoded by encode.Encoder and spliced into the batch at their place, so the result is always the program
encode.encode_batch would have produced (tests/test_host_logic.py::test_native_encoder_*)."""
from __future__ import annotations

import ctypes as C
import importlib.util

import numpy as np

from . import _lib, build
from .encode import Program, encode_scene

FLAT_PAINT_DT = np.dtype([("kind", "<i4"), ("spread", "<i4"), ("bbox_units", "<i4"), ("lin", "<i4"), ("has_transform", "<i4"),
                          ("stop_off", "<i4"), ("stop_cnt", "<i4"), ("focal", "<i4"), ("p", "<f8", 8), ("inv", "<f8", 6)],
                         align=True)
FLAT_STOP_DT = np.dtype([("offset", "<f8"), ("color", "<f8", 4)], align=True)
FLAT_NODE_DT = np.dtype([("tag", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4"), ("child_off", "<i4"),
                         ("child_cnt", "<i4"), ("pad", "<i4"), ("f", "<f8", 2)], align=True)
FLAT_FE_DT = np.dtype([("tag", "<i4"), ("n_in", "<i4"), ("in_off", "<i4"), ("flag", "<i4"), ("a", "<f8", 20)], align=True)
FLAT_SCENE_DT = np.dtype([("root", "<i4"), ("width", "<i4"), ("height", "<i4"), ("linear_rgb", "<i4")], align=True)


class Flat(C.Structure):
    _fields_ = [("n_seg", C.c_int64), ("seg_tag", C.c_void_p), ("seg_data", C.c_void_p),
                ("n_sub", C.c_int32), ("sub_off", C.c_void_p), ("n_path", C.c_int32), ("path_off", C.c_void_p),
                ("n_tr", C.c_int32), ("tr", C.c_void_p), ("n_paint", C.c_int32), ("paints", C.c_void_p),
                ("n_stop", C.c_int32), ("stops", C.c_void_p), ("n_node", C.c_int32), ("nodes", C.c_void_p),
                ("n_child", C.c_int32), ("children", C.c_void_p), ("n_scene", C.c_int32), ("scenes", C.c_void_p),
                ("n_fe", C.c_int32), ("fes", C.c_void_p), ("n_fe_input", C.c_int32), ("fe_inputs", C.c_void_p)]


_flatten_mod = None


def _flatten():
    """The in-tree extension module (built by svgrasterize_b200.build alongside the library)."""
    global _flatten_mod
    if _flatten_mod is None:
        path = build.flatten_path()
        spec = importlib.util.spec_from_file_location("_svgr_flatten", path)
        if spec is None:
            raise ImportError(f"{path} is missing: build it with `python -m svgrasterize_b200.build`")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _flatten_mod = mod
    return _flatten_mod


def flatten(jobs):
    """-> (dict of numpy arrays in the svgr_flat layout, indices of the jobs left to the Python encoder)"""
    raw, skipped = _flatten().flatten(jobs)
    arr = {
        "seg_tag": np.frombuffer(raw["seg_tag"], np.uint8), "seg_data": np.frombuffer(raw["seg_data"], np.float64).reshape(-1, 8),
        "sub_off": np.frombuffer(raw["sub_off"], np.int32), "path_off": np.frombuffer(raw["path_off"], np.int32),
        "tr": np.frombuffer(raw["tr"], np.float64).reshape(-1, 6), "paints": np.frombuffer(raw["paints"], FLAT_PAINT_DT),
        "stops": np.frombuffer(raw["stops"], FLAT_STOP_DT), "nodes": np.frombuffer(raw["nodes"], FLAT_NODE_DT),
        "children": np.frombuffer(raw["children"], np.int32), "scenes": np.frombuffer(raw["scenes"], FLAT_SCENE_DT),
        "fes": np.frombuffer(raw["fes"], FLAT_FE_DT), "fe_inputs": np.frombuffer(raw["fe_inputs"], np.int32),
    }
    return arr, list(skipped)


def _flat_struct(arr):
    f = Flat()
    f.n_seg, f.seg_tag, f.seg_data = len(arr["seg_tag"]), _lib.ptr(arr["seg_tag"]), _lib.ptr(arr["seg_data"])
    f.n_sub, f.sub_off = len(arr["sub_off"]) - 1, _lib.ptr(arr["sub_off"])
    f.n_path, f.path_off = len(arr["path_off"]) - 1, _lib.ptr(arr["path_off"])
    f.n_tr, f.tr = len(arr["tr"]), _lib.ptr(arr["tr"])
    f.n_paint, f.paints = len(arr["paints"]), _lib.ptr(arr["paints"])
    f.n_stop, f.stops = len(arr["stops"]), _lib.ptr(arr["stops"])
    f.n_node, f.nodes = len(arr["nodes"]), _lib.ptr(arr["nodes"])
    f.n_child, f.children = len(arr["children"]), _lib.ptr(arr["children"])
    f.n_scene, f.scenes = len(arr["scenes"]), _lib.ptr(arr["scenes"])
    f.n_fe, f.fes = len(arr["fes"]), _lib.ptr(arr["fes"])
    f.n_fe_input, f.fe_inputs = len(arr["fe_inputs"]), _lib.ptr(arr["fe_inputs"])
    return f


_ERRORS = {_lib.E_INVALID: ValueError, _lib.E_UNSUPPORTED: NotImplementedError, _lib.E_NOMEM: MemoryError}


class NativeProgram:
    """A scene program that lives in the library (svgr_encoded): Engine.render / render_png take it as it is (no
    copies); the arrays are visible as numpy views for as long as this object lives."""

    ARRAYS = Program.ARRAYS

    def __init__(self, handle):
        self._h = handle
        L = _lib.lib()
        self._cprog = _lib.Program.from_address(L.svgr_encoded_program(handle))
        cv, roots = C.c_void_p(), C.c_void_p()
        n = L.svgr_encoded_canvases(handle, C.byref(cv), C.byref(roots))
        c = np.ctypeslib.as_array(C.cast(cv, C.POINTER(C.c_int64)), shape=(n, 4)).copy() if n else np.zeros((0, 4), np.int64)
        self.canvases = [tuple(int(v) for v in row) for row in c]
        self.roots = np.ctypeslib.as_array(C.cast(roots, C.POINTER(C.c_int32)), shape=(n,)).copy().tolist() if n else []
        self.canvas_bytes = int(self._cprog.canvas_bytes)
        self.n_focal = int(self._cprog.n_focal)
        self.externals = []
        self.flatness = 0.0

    def close(self):
        if self._h:
            _lib.lib().svgr_encoded_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_c(self):
        return self._cprog, [self]

    def _view(self, ptr, count, dtype, shape=None):
        if not count or not ptr:
            return np.zeros((0,) + (shape or ()), dtype)
        dt = np.dtype(dtype)
        n_el = count * int(np.prod(shape or (1,)))
        buf = (C.c_char * (n_el * dt.itemsize)).from_address(ptr)
        a = np.frombuffer(buf, dt)
        return a.reshape((count,) + shape) if shape else a

    def __getattr__(self, name):
        p = self.__dict__.get("_cprog")
        if p is None or name.startswith("_"):
            raise AttributeError(name)
        views = {
            "seg_tag": (p.seg_tag, p.n_seg, np.uint8, None), "seg_data": (p.seg_data, p.n_seg, np.float64, (8,)),
            "seg_path": (p.seg_path, p.n_seg, np.uint32, None), "paths": (p.paths, p.n_path, _lib.PATH_DT, None),
            "strokes": (p.strokes, p.n_stroke, _lib.STROKE_DT, None),
            "stroke_sub_off": (p.stroke_sub_off, p.n_stroke_sub + 1, np.int32, None),
            "stroke_sub_job": (p.stroke_sub_job, p.n_stroke_sub, np.int32, None),
            "stroke_tag": (p.stroke_tag, p.n_stroke_seg, np.uint8, None),
            "stroke_data": (p.stroke_data, p.n_stroke_seg, np.float64, (8,)),
            "stroke_seg_job": (p.stroke_seg_job, p.n_stroke_seg, np.int32, None),
            "paints": (p.paints, p.n_paint, _lib.PAINT_DT, None), "stops": (p.stops, p.n_stop, _lib.STOP_DT, None),
            "nodes": (p.nodes, p.n_node, _lib.NODE_DT, None), "children": (p.children, p.n_child, np.int32, None),
            "bbox_jobs": (p.bbox_jobs, p.n_bbox_job, _lib.BBOX_JOB_DT, None),
            "kernels": (p.kernels, p.n_kernel, _lib.KERNEL_DT, None), "weights": (p.weights, p.n_weight, np.float32, None),
            "matrices": (p.matrices, p.n_matrix, np.float32, (20,)),
            "offset_tr": (p.offset_tr, p.n_offset_tr, np.float64, (12,)),
        }
        if name in views:
            return self._view(*views[name])
        raise AttributeError(name)

    def h2d_bytes(self) -> int:
        return int(sum(getattr(self, n).nbytes for n in self.ARRAYS))

    def to_program(self) -> Program:
        """A plain encode.Program with copies of the arrays (for Program.concat with Python-encoded scenes)."""
        out = Program()
        for n in self.ARRAYS:
            setattr(out, n, np.array(getattr(self, n)))
        if len(out.stroke_sub_off) == 0:
            out.stroke_sub_off = np.zeros(1, np.int32)
        out.n_focal, out.canvas_bytes = self.n_focal, self.canvas_bytes
        out.canvases, out.roots = list(self.canvases), list(self.roots)
        return out


def encode_flat(arr) -> NativeProgram:
    """svgr_encode_flat on arrays from flatten()."""
    L = _lib.lib()
    f = _flat_struct(arr)
    handle = C.c_void_p()
    rc = L.svgr_encode_flat(C.byref(f), C.byref(handle))
    if rc != 0:
        msg = L.svgr_encoded_error(handle).decode() if handle.value else f"error {rc}"
        if handle.value:
            L.svgr_encoded_free(handle)
        raise _ERRORS.get(rc, RuntimeError)(msg)
    return NativeProgram(handle)


def encode_batch(jobs, engine=None):
    """[(scene, (width, height), linear_rgb)] -> a program rendering all of them, in order.  All scenes covered by the
    native walk: a NativeProgram (no Python per scene at all); otherwise an encode.Program in which the others were
    encoded by encode.Encoder (`engine` is needed for objectBoundingBox units)."""
    jobs = list(jobs)
    arr, skipped = flatten(jobs)
    native = encode_flat(arr)
    if not skipped:
        return native
    # splice: the native program holds the covered scenes in order; cut it at the skipped positions
    covered = native.to_program()
    native.close()
    parts, skipped_set = [], set(skipped)
    singles = _split_scenes(covered)
    k = 0
    for j, (scene, size, lin) in enumerate(jobs):
        if j in skipped_set:
            parts.append(encode_scene(scene, size, lin, engine=engine))
        else:
            parts.append(singles[k])
            k += 1
    return Program.concat(parts)


def _split_scenes(prog: Program):
    """One Program per canvas of a program whose scenes were appended one after another (index spaces are rebased;
    the inverse of Program.concat for this layout)."""
    out = []
    node_cuts = [0] + [n + 1 for n, _off, _r, _c in prog.canvases]
    for s, (node, off, rows, cols) in enumerate(prog.canvases):
        n0, n1 = node_cuts[s], node_cuts[s + 1]
        nodes = prog.nodes[n0:n1].copy()
        leaf = nodes["tag"] == _lib.N_LEAF
        p = Program()
        paths_used = nodes["a"][leaf]
        p0 = int(paths_used.min()) if len(paths_used) else 0
        p1 = int(paths_used.max()) + 1 if len(paths_used) else 0
        paint_used = nodes["b"][leaf & (nodes["b"] >= 0)]
        q0 = int(paint_used.min()) if len(paint_used) else 0
        q1 = int(paint_used.max()) + 1 if len(paint_used) else 0
        c0 = int(nodes["child_off"][0]) if len(nodes) else 0
        c1 = int(nodes["child_off"][-1] + nodes["child_cnt"][-1]) if len(nodes) else 0
        seg = (prog.seg_path >= p0) & (prog.seg_path < p1)
        p.seg_tag, p.seg_data, p.seg_path = prog.seg_tag[seg], prog.seg_data[seg], prog.seg_path[seg] - np.uint32(p0)
        p.paths = prog.paths[p0:p1].copy()
        st = (prog.strokes["path"] >= p0) & (prog.strokes["path"] < p1)
        jobs_idx = np.nonzero(st)[0]
        if len(jobs_idx):
            j0, j1 = int(jobs_idx[0]), int(jobs_idx[-1]) + 1
            strokes = prog.strokes[j0:j1].copy()
            s0, s1 = int(strokes["sub_begin"][0]), int(strokes["sub_end"][-1])
            g0, g1 = int(prog.stroke_sub_off[s0]), int(prog.stroke_sub_off[s1])
            strokes["sub_begin"] -= s0
            strokes["sub_end"] -= s0
            strokes["path"] -= p0
            p.strokes = strokes
            p.stroke_sub_off = (prog.stroke_sub_off[s0:s1 + 1] - np.int32(g0)).astype(np.int32)
            p.stroke_sub_job = prog.stroke_sub_job[s0:s1] - np.int32(j0)
            p.stroke_tag, p.stroke_data = prog.stroke_tag[g0:g1], prog.stroke_data[g0:g1]
            p.stroke_seg_job = prog.stroke_seg_job[g0:g1] - np.int32(j0)
        paints = prog.paints[q0:q1].copy()
        grad = (paints["kind"] != _lib.PAINT_SOLID) & (paints["kind"] != _lib.PAINT_PATTERN)
        t0 = int(paints["stop_off"][grad].min()) if grad.any() else 0
        t1 = int((paints["stop_off"][grad] + paints["stop_cnt"][grad]).max()) if grad.any() else 0
        focal = paints["kind"] == _lib.PAINT_RADIAL_FOCAL
        f0 = int(paints["flag"][focal].min()) if focal.any() else 0
        paints["stop_off"][grad] -= t0
        paints["flag"][focal] -= f0
        p.paints, p.stops, p.n_focal = paints, prog.stops[t0:t1].copy(), int(focal.sum())
        nodes["child_off"] -= c0
        nodes["a"][leaf] -= p0
        nodes["b"][leaf & (nodes["b"] >= 0)] -= q0
        canvas = nodes["tag"] == _lib.N_CANVAS
        nodes["f"][canvas, 0] -= off
        p.nodes, p.children = nodes, prog.children[c0:c1] - np.int32(n0)
        p.canvas_bytes = 4 * rows * cols
        p.canvases, p.roots = [(node - n0, 0, rows, cols)], [prog.roots[s] - n0]
        out.append(p)
    return out
