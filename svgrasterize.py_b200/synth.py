"""Synthetic workloads named by BASELINE.json.

* :func:`icon_scene` -- config 5: one synthetic 64-unit icon rendered at
  256 px: gradients, clip, mask, strokes.  seed = icon index.
* :func:`filter_stack_scene` -- config 4: gradient-filled circle under
  feGaussianBlur -> feMorphology(dilate) -> feColorMatrix(saturate).
* :func:`feature_scenes` -- small scenes that each isolate one feature the
  demos do not cover (bbox-unit paints, spread methods, pattern, arithmetic
  composite, offsets, erode, evenodd rings, viewport clipping ...).

Scenes are built directly as Scene objects (no SVG text, no parser): the same
objects are rendered by the CUDA core, by the oracle, and -- converted by
tools/make_golden.py -- by the unmodified reference to produce the golden
fixtures.
"""
from __future__ import annotations

import math

import numpy as np

from . import scene as S


class PathBuilder:
    """Builds ``Path.subpaths`` with the segment structure the reference's
    path-data reader produces (svgrasterize.py:1296-1428): every sub-path ends
    with a PATH_CLOSED or PATH_UNCLOSED segment back to its start."""

    def __init__(self):
        self.subpaths = []
        self.cur = []
        self.pos = (0.0, 0.0)
        self.start = (0.0, 0.0)

    def _end(self, tag):
        if self.cur or tag == S.PATH_CLOSED:
            self.cur.append((tag, np.array([self.pos, self.start], dtype=np.float64)))
            self.subpaths.append(self.cur)
            self.cur = []

    def move_to(self, x, y):
        if self.cur:
            self._end(S.PATH_UNCLOSED)
        self.pos = self.start = (float(x), float(y))
        return self

    def line_to(self, x, y):
        dst = (float(x), float(y))
        self.cur.append((S.PATH_LINE, np.array([self.pos, dst])))
        self.pos = dst
        return self

    def quad_to(self, cx, cy, x, y):
        dst = (float(x), float(y))
        self.cur.append((S.PATH_QUAD, np.array([self.pos, (cx, cy), dst], dtype=np.float64)))
        self.pos = dst
        return self

    def cubic_to(self, c0x, c0y, c1x, c1y, x, y):
        dst = (float(x), float(y))
        self.cur.append((S.PATH_CUBIC, np.array([self.pos, (c0x, c0y), (c1x, c1y), dst], dtype=np.float64)))
        self.pos = dst
        return self

    def arc(self, cx, cy, rx, ry, phi, eta, eta_delta):
        """Parametric arc (the form PATH_ARC stores, svgrasterize.py:903-904);
        the pen moves to the arc's end point."""
        self.cur.append((S.PATH_ARC, (np.array([cx, cy], dtype=np.float64), float(rx), float(ry), float(phi),
                                      float(eta), float(eta_delta))))
        a = eta + eta_delta
        c, s = math.cos(phi), math.sin(phi)
        ex, ey = rx * math.cos(a), ry * math.sin(a)
        self.pos = (cx + c * ex - s * ey, cy + s * ex + c * ey)
        return self

    def close(self):
        self._end(S.PATH_CLOSED)
        self.pos = self.start
        return self

    def path(self) -> S.Path:
        if self.cur:
            self._end(S.PATH_UNCLOSED)
        return S.Path(self.subpaths)


def rect_path(x, y, w, h, rx=0.0, ry=None) -> S.Path:
    """Same outline order as the reference's rect element (svgrasterize.py:3365-3393)."""
    ry = rx if ry is None else ry
    b = PathBuilder().move_to(x + rx, y).line_to(x + w - rx, y)
    q = math.pi / 2
    if rx > 0 and ry > 0:
        b.arc(x + w - rx, y + ry, rx, ry, 0.0, -q, q)
    b.line_to(x + w, y + h - ry)
    if rx > 0 and ry > 0:
        b.arc(x + w - rx, y + h - ry, rx, ry, 0.0, 0.0, q)
    b.line_to(x + rx, y + h)
    if rx > 0 and ry > 0:
        b.arc(x + rx, y + h - ry, rx, ry, 0.0, q, q)
    b.line_to(x, y + ry)
    if rx > 0 and ry > 0:
        b.arc(x + rx, y + ry, rx, ry, 0.0, 2 * q, q)
    return b.close().path()


def ellipse_path(cx, cy, rx, ry=None) -> S.Path:
    """Four quarter arcs, as the reference's circle/ellipse elements (:3396-3413)."""
    ry = rx if ry is None else ry
    b = PathBuilder().move_to(cx + rx, cy)
    q = math.pi / 2
    for k in range(4):
        b.arc(cx, cy, rx, ry, 0.0, k * q, q)
    return b.close().path()


def color(r, g, b, a=1.0) -> np.ndarray:
    """Premultiplied linear RGBA from straight sRGB components in [0, 1]
    (what the reference's colour reader stores, svgrasterize.py:3611-3618)."""
    rgb = np.array([r, g, b], dtype=np.float64)
    lin = np.where(rgb <= 0.04045, rgb / 12.92, np.power((rgb + 0.055) / 1.055, 2.4))
    return np.array([*(lin * a), a], dtype=np.float64)


def _rand_color(rng, a=1.0):
    r, g, b = rng.uniform(0.05, 0.95, 3)
    return color(r, g, b, a)


def _blob(rng, cx, cy, r) -> S.Path:
    """Closed shape made of two cubics."""
    ang = rng.uniform(0, 2 * math.pi)
    p0 = (cx + r * math.cos(ang), cy + r * math.sin(ang))
    p1 = (cx - r * math.cos(ang), cy - r * math.sin(ang))
    k = rng.uniform(0.8, 1.6, 4) * r
    nx, ny = -math.sin(ang), math.cos(ang)
    b = PathBuilder().move_to(*p0)
    b.cubic_to(p0[0] + nx * k[0], p0[1] + ny * k[0], p1[0] + nx * k[1], p1[1] + ny * k[1], *p1)
    b.cubic_to(p1[0] - nx * k[2], p1[1] - ny * k[2], p0[0] - nx * k[3], p0[1] - ny * k[3], *p0)
    return b.close().path()


ICON_UNITS = 64.0
ICON_PX = 256


def icon_scene(seed: int) -> S.Scene:
    """Config-5 icon: 12 masks, 3 strokes, 3 gradients, clip + mask."""
    rng = np.random.default_rng(seed)
    U = ICON_UNITS
    parts = []

    # 1. rounded rect, 3-stop linear gradient, one stop with opacity
    stops = [(0.0, _rand_color(rng)), (float(rng.uniform(0.3, 0.7)), _rand_color(rng, rng.uniform(0.4, 0.9))),
             (1.0, _rand_color(rng))]
    grad = S.GradLinear(np.array([rng.uniform(0, 20), rng.uniform(0, 20)]),
                        np.array([rng.uniform(44, 64), rng.uniform(44, 64)]), stops, None, "pad", False, None)
    parts.append(S.Scene.fill(rect_path(4, 4, U - 8, U - 8, rng.uniform(4, 12)), grad))

    # 2. four blobs clipped by a circle: 2 solid, 2 focal radial, one evenodd, fill-opacity
    blobs = []
    for k in range(4):
        cx, cy, r = rng.uniform(18, 46), rng.uniform(18, 46), rng.uniform(8, 18)
        path = _blob(rng, cx, cy, r)
        if k < 2:
            paint = _rand_color(rng)
        else:
            gstops = [(0.0, _rand_color(rng)), (1.0, _rand_color(rng, rng.uniform(0.5, 1.0)))]
            paint = S.GradRadial(np.array([cx, cy]), float(r * 1.2),
                                 np.array([cx + rng.uniform(-0.3, 0.3) * r, cy + rng.uniform(-0.3, 0.3) * r]), None,
                                 gstops, None, ("pad", "reflect")[k - 2], False, None)
        node = S.Scene.fill(path, paint, S.PATH_FILL_EVENODD if k == 1 else None)
        blobs.append(node.opacity(float(rng.uniform(0.5, 0.99))))
    clip = S.Scene.fill(ellipse_path(U / 2, U / 2, rng.uniform(18, 26)), np.ones(4))
    parts.append(S.Scene.group(blobs).clip(clip))

    # 3. quad stroke (round cap, round join) under a luminance mask (rounded rect minus circle)
    b = PathBuilder().move_to(rng.uniform(6, 16), rng.uniform(40, 58))
    for _ in range(3):
        b.quad_to(rng.uniform(8, 56), rng.uniform(8, 56), rng.uniform(8, 56), rng.uniform(8, 56))
    stroke = S.Scene.stroke(b.path(), _rand_color(rng), float(rng.uniform(2, 5)), "round", "round")
    mask = S.Scene.group([
        S.Scene.fill(rect_path(8, 8, U - 16, U - 16, 6.0), color(1, 1, 1)),
        S.Scene.fill(ellipse_path(rng.uniform(24, 40), rng.uniform(24, 40), rng.uniform(5, 10)), color(0, 0, 0)),
    ])
    parts.append(stroke.mask(mask))

    # 4. polyline stroke, miter join, square cap
    b = PathBuilder().move_to(rng.uniform(6, 20), rng.uniform(6, 20))
    for _ in range(4):
        b.line_to(rng.uniform(6, 58), rng.uniform(6, 58))
    parts.append(S.Scene.stroke(b.path(), _rand_color(rng), float(rng.uniform(1, 3)), "square", "miter"))

    # 5. filled and stroked circle
    circle = ellipse_path(rng.uniform(20, 44), rng.uniform(20, 44), rng.uniform(5, 12))
    parts.append(S.Scene.fill(circle, _rand_color(rng, rng.uniform(0.6, 1.0))))
    parts.append(S.Scene.stroke(circle, _rand_color(rng), float(rng.uniform(0.8, 2.5))))

    scale = ICON_PX / U
    return S.Scene.group(parts).transform(S.Transform().scale(scale))


def icon_size():
    return (ICON_PX, ICON_PX)


def saturate_matrix(value: float) -> np.ndarray:
    """feColorMatrix type=saturate (svgrasterize.py:1954-1957 with :1740-1747)."""
    hue = np.array([
        [[0.213, 0.715, 0.072], [0.213, 0.715, 0.072], [0.213, 0.715, 0.072]],
        [[0.787, -0.715, -0.072], [-0.213, 0.285, -0.072], [-0.213, -0.715, 0.928]],
        [[-0.213, -0.715, 0.928], [0.143, 0.140, -0.283], [-0.787, 0.715, 0.072]],
    ])
    m = np.eye(4, 5)
    m[:3, :3] = np.dot(hue.T, [1, value, 0]).T
    return m


def filter_stack_scene(n: int, sigma: float = 4.0, radius: float = 3.0, saturate: float = 0.5) -> S.Scene:
    """Config 4: circle r = 0.45 n with a 3-stop linear gradient, under
    blur(sigma) -> dilate(radius) -> saturate, identity-scale viewBox."""
    # offsets / end points chosen off the pixel lattice so that no pixel lands on an exact
    # x*255 = k + 0.5 rounding tie (a 1e-16 wobble would flip the byte)
    stops = [(0.0, color(0.88, 0.2, 0.12)), (0.47, color(0.12, 0.8, 0.32, 0.83)), (1.0, color(0.12, 0.2, 0.88))]
    grad = S.GradLinear(np.array([0.093 * n, 0.0]), np.array([0.931 * n, 0.817 * n]), stops, None, "pad", False, None)
    flt = S.Filter.empty().blur(sigma, sigma).morphology(radius, radius, "max", None).color_matrix(
        None, saturate_matrix(saturate))
    return S.Scene.fill(ellipse_path(n / 2, n / 2, 0.45 * n), grad).filter(flt)


# feature scenes rendered with linear_rgb=True (the CLI's --linear-rgb, svgrasterize.py:3810)
FEATURES_LINEAR_RGB = {"filter_blur_lin", "linear_bbox_repeat", "mask_bbox_units"}


def feature_scenes() -> dict:
    """name -> (scene, (w, h)) ; each isolates features the demos do not reach."""
    out = {}
    sq = rect_path(8, 8, 48, 40)
    two = [(0.0, color(1, 0, 0)), (1.0, color(0, 0, 1, 0.48))]
    three = [(0.0, color(1, 1, 0)), (0.4, color(0, 1, 0, 0.68)), (1.0, color(0, 0, 1))]

    def grp(*nodes):
        return S.Scene.group(list(nodes))

    # objectBoundingBox gradients, spread methods, gradientTransform
    for spread in ("pad", "repeat", "reflect"):
        g = S.GradLinear(np.array([0.25, 0.0]), np.array([0.6, 0.3]), three, None, spread, True, None)
        out[f"linear_bbox_{spread}"] = (S.Scene.fill(sq, g).transform(S.Transform().scale(2).rotate(0.2)), (160, 140))
    gt = S.Transform().translate(3, -2).rotate(0.4).scale(1.5, 0.7)
    g = S.GradRadial(np.array([30.0, 28.0]), 20.0, None, None, three, gt, "reflect", False, None)
    out["radial_simple_transform"] = (S.Scene.fill(sq, g).transform(S.Transform().scale(2)), (128, 112))
    g = S.GradRadial(np.array([0.5, 0.5]), 0.5, np.array([0.3, 0.35]), 0.05, two, None, "pad", True, None)
    out["radial_focal_bbox"] = (S.Scene.fill(ellipse_path(32, 30, 26, 20), g).transform(S.Transform().scale(2)), (128, 120))
    # focus outside the end circle: det < 0 region -> transparent, t <= fr/(fr-r) masked
    g = S.GradRadial(np.array([32.0, 30.0]), 10.0, np.array([50.0, 30.0]), None, two, None, "repeat", False, None)
    out["radial_focal_outside"] = (S.Scene.fill(sq, g).transform(S.Transform().scale(2)), (128, 112))
    g = S.GradLinear(np.array([8.0, 0.0]), np.array([56.0, 0.0]), two, None, "pad", False, True)
    out["linear_interp_linear_rgb"] = (S.Scene.fill(sq, g).transform(S.Transform().scale(2)), (128, 112))

    # evenodd ring pair + nonzero overlap, clipped by the viewport on all four sides
    ring = PathBuilder()
    for r in (30, 20, 10):
        ring.move_to(32 + r, 32)
        for k in range(4):
            ring.arc(32, 32, r, r, 0.0, k * math.pi / 2, math.pi / 2)
        ring.close()
    ring = ring.path()
    out["evenodd_rings_clipped"] = (
        grp(S.Scene.fill(ring, color(0.12, 0.48, 0.88), "evenodd"),
            S.Scene.fill(ring, color(0.88, 0.32, 0.12, 0.48), "nonzero").transform(S.Transform().translate(20, 14)))
        .transform(S.Transform().translate(-8, -6).scale(1.7)), (90, 84))

    # group opacity, nested clip with bbox units, luminance mask with bbox units
    unit_circle = S.Scene.fill(ellipse_path(0.5, 0.5, 0.45, 0.4), np.ones(4))
    inner = grp(S.Scene.fill(sq, color(0.2, 0.68, 0.32)), S.Scene.fill(ellipse_path(40, 30, 20), color(0.8, 0.12, 0.6, 0.8)))
    out["clip_bbox_units_opacity"] = (inner.opacity(0.6).clip(unit_circle, True).transform(S.Transform().scale(2)), (128, 112))
    lum = grp(S.Scene.fill(rect_path(0, 0, 1, 1), color(1, 1, 1)),
              S.Scene.fill(ellipse_path(0.5, 0.5, 0.3), color(0.2, 0.2, 0.2, 0.88)))
    out["mask_bbox_units"] = (inner.mask(lum, True).transform(S.Transform().scale(2).rotate(-0.1)), (128, 112))

    # strokes: caps x joins matrix on an open zig-zag and a closed triangle with a cusp-y cubic
    cells = []
    for i, cap in enumerate((None, "round", "square")):
        for j, join in enumerate((None, "round", "bevel")):
            b = PathBuilder().move_to(4, 20).line_to(12, 4).line_to(20, 20).cubic_to(28, 40, 8, 40, 24, 8)
            b.move_to(30, 6).line_to(44, 10).quad_to(50, 24, 34, 22).close()
            node = S.Scene.stroke(b.path(), color(0.1 + 0.4 * i, 0.2, 0.9 - 0.4 * j, 0.88), 1.5 + i + 0.5 * j, cap, join)
            cells.append(node.transform(S.Transform().translate(52 * i, 44 * j)))
    out["stroke_caps_joins"] = (grp(*cells).transform(S.Transform().scale(1.5)), (240, 200))

    # filters: offset + merge (drop shadow), arithmetic composite, erode, hueRotate-like matrix, rotated blur
    base = grp(S.Scene.fill(sq, color(0.88, 0.48, 0.12)), S.Scene.fill(ellipse_path(36, 30, 14), color(0.12, 0.32, 0.8, 0.68)))
    shadow = (S.Filter.empty().blur(2.0, None, S.FE_SOURCE_ALPHA, "blur").offset(3, 4, "blur", "off")
              .merge(["off", S.FE_SOURCE_GRAPHIC]))
    out["filter_drop_shadow"] = (base.filter(shadow).transform(S.Transform().scale(2)), (140, 124))
    arith = (S.Filter.empty().blur(1.5, 3.0, S.FE_SOURCE_GRAPHIC, "b")
             .composite(S.FE_SOURCE_GRAPHIC, "b", (0.5, 0.6, 0.4, 0.02)))
    out["filter_arithmetic"] = (base.filter(arith).transform(S.Transform().scale(2)), (140, 124))
    for name, mode in (("xor", S.COMPOSE_XOR), ("atop", S.COMPOSE_ATOP), ("out", S.COMPOSE_OUT), ("in", S.COMPOSE_IN)):
        f = S.Filter.empty().offset(6, 5, S.FE_SOURCE_GRAPHIC, "o").composite(S.FE_SOURCE_GRAPHIC, "o", mode)
        out[f"filter_composite_{name}"] = (base.filter(f).transform(S.Transform().scale(2)), (140, 124))
    erode = S.Filter.empty().morphology(1.5, 1.0, "min", None).color_matrix(None, saturate_matrix(1.8))
    out["filter_erode_matrix"] = (base.filter(erode).transform(S.Transform().scale(2)), (140, 124))
    m = np.eye(4, 5)
    m[0, 4], m[3, 4], m[1, 0] = 0.2, 0.1, 0.5
    out["filter_matrix_bias"] = (base.filter(S.Filter.empty().color_matrix(None, m)).transform(S.Transform().scale(2)), (140, 124))
    rot = S.Filter.empty().blur(3.0, 1.0)
    out["filter_blur_rotated"] = (base.filter(rot).transform(S.Transform().translate(40, -10).rotate(0.5).scale(2)), (170, 170))
    out["filter_blur_lin"] = (base.filter(S.Filter.empty().blur(2.5)).transform(S.Transform().scale(2)), (140, 124))

    # pattern paint (userSpaceOnUse tile with a transform)
    tile = grp(S.Scene.fill(rect_path(0, 0, 6, 6), color(0.88, 0.12, 0.12)),
               S.Scene.fill(ellipse_path(8, 8, 3), color(0.12, 0.12, 0.88, 0.8)))
    pat = S.Pattern(tile, False, None, 0.0, 0.0, 12.0, 12.0, S.Transform().rotate(0.3), False)
    out["pattern_user_space"] = (S.Scene.fill(sq, pat).transform(S.Transform().scale(2)), (128, 112))
    return out
