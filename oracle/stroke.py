"""Oracle: Path.stroke (svgrasterize.py:1105-1180) via oracle/svgr_oracle.c
(TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from svgrasterize_b200 import scene as S
from svgrasterize_b200.sceneio import path_arrays, path_from_arrays

from . import clib

CAPS = {None: 0, S.STROKE_CAP_BUTT: 0, S.STROKE_CAP_ROUND: 1, S.STROKE_CAP_SQUARE: 2}
JOINS = {None: 0, S.STROKE_JOIN_MITER: 0, S.STROKE_JOIN_ROUND: 1, S.STROKE_JOIN_BEVEL: 2}


def stroke_arrays(tags, data, sub_off, width, linecap=None, linejoin=None):
    if linecap not in CAPS:
        raise ValueError(f"unkown line cap type: `{linecap}`")
    fn = clib.declare(
        "orc_stroke", C.c_long,
        [C.POINTER(C.c_uint8), clib.c_double_p, C.POINTER(C.c_int32), C.c_long, C.c_double, C.c_int, C.c_int,
         C.POINTER(C.c_uint8), clib.c_double_p, C.c_long, C.POINTER(C.c_int32), C.c_long, C.POINTER(C.c_long)])
    tags = np.ascontiguousarray(tags, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.float64)
    sub_off = np.ascontiguousarray(sub_off, dtype=np.int32)
    nsub = len(sub_off) - 1
    cap = 64 + 16 * len(tags)
    while True:
        out_tag = np.zeros(cap, dtype=np.uint8)
        out_data = np.zeros((cap, 8))
        out_sub = np.zeros(2 * nsub + 2, dtype=np.int32)
        nso = C.c_long(0)
        n = fn(tags.ctypes.data_as(C.POINTER(C.c_uint8)), clib.dp(data), sub_off.ctypes.data_as(C.POINTER(C.c_int32)),
               nsub, float(width), CAPS[linecap], JOINS.get(linejoin, 3),
               out_tag.ctypes.data_as(C.POINTER(C.c_uint8)), clib.dp(out_data), cap,
               out_sub.ctypes.data_as(C.POINTER(C.c_int32)), 2 * nsub + 1, C.byref(nso))
        if n == -1:
            cap *= 4
            continue
        if n == -4:
            raise TypeError("cannot unpack non-iterable NoneType object")  # what the reference does
        if n < 0:
            raise RuntimeError(f"orc_stroke failed: {n}")
        return out_tag[:n].copy(), out_data[:n].copy(), out_sub[: nso.value + 1].copy()


def stroke_path(path, width, linecap=None, linejoin=None) -> S.Path:
    tags, data, sub_off = path_arrays(path)
    return path_from_arrays(*stroke_arrays(tags, data, sub_off, width, linecap, linejoin))
