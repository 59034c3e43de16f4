"""Host-side colour helpers used while encoding paints (a handful of float64
values per paint; the per-pixel versions live in csrc/svgr_device.cuh).

Mirrors color_pre_to_straight_alpha / color_linear_to_srgb /
color_straight_to_pre_alpha as Path.fill and grad_stops_colorspace apply them
to a single colour (svgrasterize.py:1015-1018, :1686-1695, :471-503).
"""
from __future__ import annotations

import numpy as np


def paint_to_srgb(color) -> np.ndarray:
    """Premultiplied linear RGBA -> premultiplied sRGB RGBA (float64).  Scalar Python floats: the same IEEE
    double operations as the reference's numpy calls on a 4-vector, without their per-call overhead."""
    r, g, b, a = (float(v) for v in color)
    if a > 0.0001:
        r, g, b = r / a, g / a, b / a
    r, g, b, a = (min(max(v, 0.0), 1.0) for v in (r, g, b, a))

    def enc(v):
        return v * 12.92 if v <= 0.0031308 else 1.055 * float(np.power(v, 1.0 / 2.4)) - 0.055

    return np.array([enc(r) * a, enc(g) * a, enc(b) * a, a], dtype=np.float64)
