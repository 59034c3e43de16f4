"""ctypes binding of oracle/libsvgr_oracle.so (TEST INFRASTRUCTURE ONLY).

Builds the library with ``make -C oracle`` on first use if it is missing or
older than its source.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsvgr_oracle.so")
_SRC = os.path.join(_HERE, "svgr_oracle.c")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_long_p = C.POINTER(C.c_long)


def build(force: bool = False) -> str:
    stale = not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_double_p)


def lp(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_long_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        d, l, i, p, lpp = C.c_double, C.c_long, C.c_int, c_double_p, c_long_p
        sig = {
            "orc_transform_points": (None, [p, p, l, p]),
            "orc_quad_to_cubic": (None, [p, p]),
            "orc_cubic_flatness": (d, [p]),
            "orc_cubic_split": (None, [p, p]),
            "orc_flatten_cubics": (l, [p, l, d, p, l, i]),
            "orc_arc_to_cubics": (l, [d, d, d, d, d, d, d, p, l]),
            "orc_line_coverage": (None, [p, l, l, p]),
            "orc_mask_finish": (i, [p, l, l, i]),
            "orc_mask": (i, [p, l, lpp, i, p]),
            "orc_mask_bounds": (i, [p, l, lpp, lpp]),
        }
        for name, (res, args) in sig.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def declare(name, res, args):
    fn = getattr(lib(), name)
    fn.restype, fn.argtypes = res, args
    return fn
