"""Host-side colour helpers used while encoding paints (a handful of float64
values per paint; the per-pixel versions live in csrc/svgr_device.cuh).

Mirrors color_pre_to_straight_alpha / color_linear_to_srgb /
color_straight_to_pre_alpha as Path.fill and grad_stops_colorspace apply them
to a single colour (svgrasterize.py:1015-1018, :1686-1695, :471-503).
"""
from __future__ import annotations

import numpy as np


def paint_to_srgb(color) -> np.ndarray:
    """Premultiplied linear RGBA -> premultiplied sRGB RGBA (float64)."""
    c = np.array(color, dtype=np.float64)
    a = c[3]
    if a > 0.0001:
        c[:3] = c[:3] / a
    c = np.clip(c, 0.0, 1.0)
    rgb = c[:3]
    low = rgb <= 0.0031308
    out = np.empty(3)
    out[low] = rgb[low] * 12.92
    out[~low] = 1.055 * np.power(rgb[~low], 1.0 / 2.4) - 0.055
    c[:3] = out * c[3]
    return c
