"""Oracle: numpy restatement of the small element-wise functions of the reference's call surface
(TEST INFRASTRUCTURE ONLY -- only tests/ import this).

    grad_pixels        svgrasterize.py:1653-1658
    grad_spread        svgrasterize.py:1661-1668
    grad_interpolate   svgrasterize.py:1671-1683  (+ grad_stops_colorspace :1686-1695 via oracle.render)
    pooling            svgrasterize.py:419-468

Pinned against the unmodified reference by tests/test_oracle_pins.py::test_eager_restatements (runs where
/root/reference exists).
"""
from __future__ import annotations

import numpy as np

from . import render as R


def pixel_centres(viewport) -> np.ndarray:
    r0, c0, rows, cols = viewport
    rr, cc = np.indices((rows, cols)).astype(np.float64)
    return np.stack([rr, cc], axis=-1) + [r0 + 0.5, c0 + 0.5]


def spread(t: np.ndarray, method: str) -> np.ndarray:
    if method == "pad":
        return t
    if method == "repeat":
        return np.modf(t)[0]
    if method == "reflect":
        return np.fabs(np.remainder(t + 1.0, 2.0) - 1.0)
    raise ValueError(f"invalid spread method: {method}")


def interpolate(t: np.ndarray, stops, linear_rgb: bool) -> np.ndarray:
    if not linear_rgb:
        stops = [(o, R.paint_to_srgb(np.asarray(c, dtype=np.float64))) for o, c in stops]
    out = np.zeros((*t.shape, 4))
    out[t <= stops[0][0]] = stops[0][1]
    out[t > stops[-1][0]] = stops[-1][1]
    for (o0, c0), (o1, c1) in zip(stops, stops[1:]):
        sel = (t > o0) & (t <= o1)
        w = ((t[sel] - o0) / (o1 - o0))[..., None]
        out[sel] += (1 - w) * c0 + w * c1
    return out


def pool(mat: np.ndarray, ksize, stride=None, method="max", pad=False) -> np.ndarray:
    rows, cols = mat.shape[:2]
    ky, kx = ksize
    sy, sx = (ky, kx) if stride is None else stride
    if pad:
        orows, ocols = -(-rows // sy), -(-cols // sx)
        full = np.full(((orows - 1) * sy + ky, (ocols - 1) * sx + kx) + mat.shape[2:], np.nan)
        full[:rows, :cols] = mat
    else:
        orows, ocols = (rows - ky) // sy + 1, (cols - kx) // sx + 1
        full = mat
    fn = {"max": np.nanmax, "min": np.nanmin, "mean": np.nanmean}[method]
    out = np.empty((orows, ocols) + mat.shape[2:])
    for i in range(orows):
        for j in range(ocols):
            out[i, j] = fn(full[i * sy: i * sy + ky, j * sx: j * sx + kx], axis=(0, 1))
    return out
