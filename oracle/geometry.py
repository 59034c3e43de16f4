"""Oracle: path -> presentation-space edge list -> coverage mask (TEST INFRASTRUCTURE ONLY).

Restates Path.mask (svgrasterize.py:922-993) on top of oracle/svgr_oracle.c.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from svgrasterize_b200 import scene as S
from svgrasterize_b200.sceneio import path_arrays

from . import clib

FLATNESS = 0.1  # px, literal at svgrasterize.py:955/:957
MAX_DEPTH = 0  # no bound, as in the reference


def transform_points(m6: np.ndarray, pts: np.ndarray) -> np.ndarray:
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.empty_like(pts)
    clib.lib().orc_transform_points(clib.dp(np.ascontiguousarray(m6, dtype=np.float64)), clib.dp(pts),
                                    pts.size // 2, clib.dp(out))
    return out


def arc_to_cubics(row: np.ndarray) -> np.ndarray:
    """row = [cx, cy, rx, ry, phi, eta, eta_delta, _] -> (k,4,2) (svgrasterize.py:2355)."""
    out = np.empty((64, 4, 2))
    k = clib.lib().orc_arc_to_cubics(*map(float, row[:7]), clib.dp(out), 64)
    if k < 0:
        raise ValueError("arc sweeps more than 16 pi")
    return out[:k].copy()


def path_lines_cubics(path):
    """Segment loop of Path.mask (:930-945): user-space lines (N,2,2) and cubics (M,4,2)."""
    tags, data, _sub = path_arrays(path)
    is_line = (tags == S.PATH_LINE) | (tags == S.PATH_CLOSED) | (tags == S.PATH_UNCLOSED)
    lines = data[is_line, :4].reshape(-1, 2, 2)
    cubics = []
    L = clib.lib()
    for tag, row in zip(tags, data):
        if tag == S.PATH_CUBIC:
            cubics.append(row.reshape(1, 4, 2))
        elif tag == S.PATH_QUAD:
            c = np.empty((1, 4, 2))
            L.orc_quad_to_cubic(clib.dp(np.ascontiguousarray(row[:6])), clib.dp(c))
            cubics.append(c)
        elif tag == S.PATH_ARC:
            cubics.append(arc_to_cubics(row))
    cubics = np.concatenate(cubics) if cubics else np.zeros((0, 4, 2))
    return np.ascontiguousarray(lines), np.ascontiguousarray(cubics)


def flatten_cubics(cubics: np.ndarray, tol: float = FLATNESS) -> np.ndarray:
    """bezier3_flatten_batch (:2091): (M,4,2) -> (E,2,2), reference (BFS) order."""
    cubics = np.ascontiguousarray(cubics, dtype=np.float64)
    m = len(cubics)
    cap = max(64, 16 * m)
    while True:
        out = np.empty((cap, 2, 2))
        n = clib.lib().orc_flatten_cubics(clib.dp(cubics), m, tol, clib.dp(out), cap, MAX_DEPTH)
        if n == -1:
            cap *= 4
            continue
        if n < 0:
            raise MemoryError
        return out[:n].copy()


def path_edges(path, transform) -> np.ndarray:
    """Presentation-space edge list in the reference's order (:948-957):
    transformed lines first, then the flattened cubics."""
    lines, cubics = path_lines_cubics(path)
    m6 = transform.m[:2, :].reshape(6)
    parts = []
    if len(lines):
        parts.append(transform_points(m6, lines))
    if len(cubics):
        parts.append(flatten_cubics(transform_points(m6, cubics)))
    if not parts:
        return np.zeros((0, 2, 2))
    return np.concatenate(parts)


def mask_bounds(edges: np.ndarray, viewport=None):
    bbox = np.zeros(4, dtype=np.int64)
    vp = None
    if viewport is not None:
        vp = np.asarray([int(v) for v in viewport], dtype=np.int64)
    empty = clib.lib().orc_mask_bounds(clib.dp(np.ascontiguousarray(edges)), len(edges),
                                       clib.lp(vp) if vp is not None else None, clib.lp(bbox))
    return None if empty else tuple(int(v) for v in bbox)


def fill_rule_code(fill_rule) -> int:
    if fill_rule is None or fill_rule == S.PATH_FILL_NONZERO:
        return 0
    if fill_rule == S.PATH_FILL_EVENODD:
        return 1
    raise ValueError(f"Invalid fill rule: {fill_rule}")


def edges_mask(edges: np.ndarray, bbox, fill_rule=None) -> np.ndarray:
    rule = fill_rule_code(fill_rule)
    out = np.empty((bbox[2], bbox[3]))
    bb = np.asarray(bbox, dtype=np.int64)
    clib.lib().orc_mask(clib.dp(np.ascontiguousarray(edges)), len(edges), clib.lp(bb), rule, clib.dp(out))
    return out


def path_mask(path, transform, fill_rule=None, viewport=None):
    """Path.mask (:922-993) -> (mask (rows, cols) f64, (min_row, min_col), edges) or None."""
    rule = fill_rule_code(fill_rule)  # raises like :989 (before the work, harmless)
    edges = path_edges(path, transform)
    if len(edges) == 0:
        return None
    bbox = mask_bounds(edges, viewport)
    if bbox is None:
        return None
    del rule
    return edges_mask(edges, bbox, fill_rule), (bbox[0], bbox[1]), edges
