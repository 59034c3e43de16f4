"""ctypes binding of libsvgr_b200.so (include/svgr_b200.h).

There is no fallback: if the library is missing or the machine has no CUDA
device the rendering entry points raise.  Loading the library itself does not
need a GPU (the symbol and layout checks run on CPU-only machines).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVGR_LIB") or os.path.join(HERE, "libsvgr_b200.so")  # SVGR_LIB: A/B builds

# ---- record layouts (must match csrc/svgr_types.h and include/svgr_b200.h) ----------------
PATH_DT = np.dtype([("m", "<f8", 6), ("viewport", "<i4", 4), ("has_viewport", "<i4"), ("fill_rule", "<i4"),
                    ("has_full", "<i4"), ("pad", "<i4"), ("full_viewport", "<i4", 4)], align=True)
STROKE_DT = np.dtype([("half_width", "<f8"), ("sub_begin", "<i4"), ("sub_end", "<i4"), ("cap", "<i4"), ("join", "<i4"),
                      ("path", "<i4"), ("pad", "<i4")], align=True)
PAINT_DT = np.dtype([("kind", "<i4"), ("spread", "<i4"), ("stop_off", "<i4"), ("stop_cnt", "<i4"), ("has_m2", "<i4"),
                     ("flag", "<i4"), ("pat_r0", "<i4"), ("pat_c0", "<i4"), ("pat_rows", "<i4"), ("pat_cols", "<i4"),
                     ("pat_node", "<i4"), ("pad", "<i4"), ("color", "<f4", 4), ("m1", "<f8", 6), ("m2", "<f8", 6),
                     ("g", "<f8", 8)], align=True)
STOP_DT = np.dtype([("offset", "<f8"), ("color", "<f4", 4), ("inv_span", "<f8")], align=True)
NODE_DT = np.dtype([("tag", "<i4"), ("a", "<i4"), ("b", "<i4"), ("c", "<i4"), ("d", "<i4"), ("child_off", "<i4"),
                    ("child_cnt", "<i4"), ("flags", "<i4"), ("f", "<f8", 4)], align=True)
KERNEL_DT = np.dtype([("rows", "<i4"), ("cols", "<i4"), ("separable", "<i4"), ("weight_off", "<i4")], align=True)
# svgr_bbox_job: an objectBoundingBox gradient the render completes on the device
BBOX_JOB_DT = np.dtype([("paint", "<i4"), ("path", "<i4"), ("has_grad_tr", "<i4"), ("pad", "<i4"), ("inv", "<f8", 6),
                        ("grad_inv", "<f8", 6), ("geom", "<f8", 6)], align=True)

PAINT_SOLID, PAINT_LINEAR, PAINT_RADIAL, PAINT_RADIAL_FOCAL, PAINT_PATTERN = range(5)
(N_EMPTY, N_LEAF, N_GROUP, N_OPACITY, N_IN, N_LUMA, N_COMPOSE, N_SRC_ALPHA, N_CONVERT, N_BLUR, N_MORPH, N_CMATRIX,
 N_OFFSET, N_MERGE_AT, N_CANVAS, N_EXTERNAL) = range(16)
STOP_NONE, STOP_STROKE, STOP_FLATTEN, STOP_COVERAGE, STOP_PLAN = range(5)
SEG_NOP = 255

E_INVALID, E_CUDA, E_NOMEM, E_UNSUPPORTED, E_STROKE, E_TYPE = -1, -2, -3, -4, -5, -6


class External(C.Structure):
    _fields_ = [("image", C.c_void_p), ("r0", C.c_int32), ("c0", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("channels", C.c_int32), ("pre_alpha", C.c_int32), ("linear_rgb", C.c_int32), ("pad", C.c_int32)]


class Program(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int64), ("seg_tag", C.c_void_p), ("seg_data", C.c_void_p), ("seg_path", C.c_void_p),
        ("n_path", C.c_int32), ("paths", C.c_void_p),
        ("n_stroke", C.c_int32), ("strokes", C.c_void_p),
        ("n_stroke_sub", C.c_int32), ("stroke_sub_off", C.c_void_p), ("stroke_sub_job", C.c_void_p),
        ("n_stroke_seg", C.c_int64), ("stroke_tag", C.c_void_p), ("stroke_data", C.c_void_p),
        ("stroke_seg_job", C.c_void_p),
        ("n_paint", C.c_int32), ("paints", C.c_void_p), ("n_stop", C.c_int32), ("stops", C.c_void_p),
        ("n_focal", C.c_int32),
        ("n_node", C.c_int32), ("nodes", C.c_void_p), ("n_child", C.c_int32), ("children", C.c_void_p),
        ("n_kernel", C.c_int32), ("kernels", C.c_void_p), ("n_weight", C.c_int32), ("weights", C.c_void_p),
        ("n_matrix", C.c_int32), ("matrices", C.c_void_p), ("n_offset_tr", C.c_int32), ("offset_tr", C.c_void_p),
        ("n_external", C.c_int32), ("externals", C.c_void_p),
        ("canvas_bytes", C.c_int64),
        ("flatness", C.c_double),
        ("n_bbox_job", C.c_int32), ("pad_bbox", C.c_int32), ("bbox_jobs", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_edges", "n_outline_segs", "n_bands", "n_cov_tiles", "n_binned", "cov_floats",
                                         "layer_floats", "n_ops", "n_levels", "n_launches", "mask_pixels",
                                         "layer_pixels", "coverage_bytes", "compose_bytes", "canvas_pixels",
                                         "n_kernels")] + \
               [(n, C.c_float) for n in ("ms_total", "ms_h2d", "ms_stroke", "ms_flatten", "ms_plan", "ms_bin",
                                         "ms_coverage", "ms_compose", "ms_canvas", "ms_d2h")] + \
               [("retries", C.c_int32), ("plan_cached", C.c_int32), ("host_plan_masks_ms", C.c_float),
                ("host_plan_nodes_ms", C.c_float), ("ms_compose_busy", C.c_float), ("pad2", C.c_float),
                ("compose_bytes_8d", C.c_int64), ("png_bytes", C.c_int64), ("ms_png", C.c_float), ("pad3", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("pad")}


# every symbol include/svgr_b200.h declares
SYMBOLS = {
    "svgr_version": (C.c_int, []),
    "svgr_sizeof": (C.c_int, [C.c_int]),
    "svgr_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "svgr_destroy": (None, [C.c_void_p]),
    "svgr_last_error": (C.c_char_p, [C.c_void_p]),
    "svgr_render": (C.c_int, [C.c_void_p, C.POINTER(Program), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.POINTER(Stats)]),
    "svgr_render_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Stats)]),
    "svgr_render_png": (C.c_int, [C.c_void_p, C.POINTER(Program), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                  C.c_int, C.POINTER(Stats)]),
    "svgr_render_resident_png": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                           C.POINTER(Stats)]),
    "svgr_png_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
    "svgr_debug_plan": (C.c_int, [C.POINTER(Program), C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  C.c_void_p]),
    "svgr_read_edges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "svgr_read_boxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgr_read_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "svgr_read_bins": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "svgr_read_outline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64)]),
    "svgr_node_info": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "svgr_read_node": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "svgr_cloud_bounds": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "svgr_arc_to_cubics": (C.c_int64, [C.c_double] * 7 + [C.c_void_p, C.c_int64]),
    "svgr_expand_arcs": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "svgr_path_from_svg": (C.c_int, [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_char_p, C.c_int32]),
    "svgr_encode_flat": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "svgr_encoded_program": (C.c_void_p, [C.c_void_p]),
    "svgr_encoded_error": (C.c_char_p, [C.c_void_p]),
    "svgr_encoded_canvases": (C.c_int64, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "svgr_encoded_free": (None, [C.c_void_p]),
    "svgr_line_signed_coverage": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "svgr_grad_pixels": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "svgr_grad_spread": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "svgr_grad_interpolate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "svgr_quantize_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "svgr_pooling": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 9 + [C.c_void_p, C.c_int32, C.c_int32]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (python -m svgrasterize_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m svgrasterize_b200.build` "
                "(there is no CPU fallback for the rasterizer core)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype, fn.argtypes = res, args
        sizes = [PATH_DT.itemsize, STROKE_DT.itemsize, PAINT_DT.itemsize, STOP_DT.itemsize, NODE_DT.itemsize,
                 KERNEL_DT.itemsize, C.sizeof(External), C.sizeof(Program), C.sizeof(Stats)]
        for what, size in list(enumerate(sizes)) + [(10, BBOX_JOB_DT.itemsize)]:
            if L.svgr_sizeof(what) != size:
                raise ImportError(f"ABI mismatch for record {what}: library {L.svgr_sizeof(what)} vs binding {size}")
        _lib = L
    return _lib


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None for empty / None)."""
    if a is None or a.size == 0:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data
