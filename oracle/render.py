"""Oracle: layers, paint, Porter-Duff compose, filters and the scene walk
(TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

A float64 CPU restatement of the reference's render model, eager and
node-by-node like the reference so that every intermediate layer can be
compared.  Heavy loops are in oracle/svgr_oracle.c; element-wise float64 numpy
is used where the reference itself is element-wise numpy (same roundings).

Coordinates follow the reference: after the canvas transform component 0 of a
point is the image ROW, component 1 the COLUMN; ``offset = (row, col)``.
"""
from __future__ import annotations

import math
import warnings
from typing import NamedTuple

import numpy as np

from svgrasterize_b200 import scene as S

from . import clib, geometry
from .stroke import stroke_path

F64 = np.float64

# When set to a dict {"leaves": [], "strokes": []}, every path-mask and stroke
# call is recorded in call order (same order as the reference's Path.mask /
# Path.stroke calls), for the stage-level parity tests.
TAPS = None


def _path_mask(path, transform, fill_rule, viewport):
    res = geometry.path_mask(path, transform, fill_rule, viewport)
    if TAPS is not None:
        TAPS["leaves"].append(None if res is None else
                              ((res[1][0], res[1][1], res[0].shape[0], res[0].shape[1]), res[2], res[0]))
    return res


def _stroke(path, width, cap, join):
    out = stroke_path(path, width, cap, join)
    if TAPS is not None:
        TAPS["strokes"].append(out)
    return out


class OLayer(NamedTuple):
    """Layer (svgrasterize.py:61-65): image (rows, cols, 1|4) f64, offset (row, col)."""

    image: np.ndarray
    offset: tuple
    pre_alpha: bool
    linear_rgb: bool

    @property
    def bbox(self):
        return (self.offset[0], self.offset[1], self.image.shape[0], self.image.shape[1])


class Cloud:
    """Stand-in for ConvexHull (svgrasterize.py:1963-2029): only the user-space
    bounding box of the hull is ever consumed (:2002-2023), and the extremes of an
    affine image of a point set are attained on its hull, so the raw end points
    are kept instead of running the Graham scan."""

    def __init__(self, points):
        self.points = np.asarray(points, dtype=F64).reshape(-1, 2)

    @classmethod
    def merge(cls, clouds):
        return cls(np.concatenate([c.points for c in clouds]))

    def bbox(self, transform):
        pts = transform.invert(self.points)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        return [lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]]

    def bbox_transform(self, transform):
        x, y, w, h = self.bbox(transform)
        if w <= 0 and h <= 0:
            return transform
        return transform.translate(x, y).scale(w, h)


# ------------------------------------------------------------------------------
# colour space / alpha (svgrasterize.py:471-503); all return new arrays
# ------------------------------------------------------------------------------
def unpremultiply(rgba):
    out = rgba.copy()
    a = out[..., 3:]
    np.divide(out[..., :3], a, out=out[..., :3], where=a > 0.0001)
    return np.clip(out, 0, 1)


def premultiply(rgba):
    out = rgba.copy()
    out[..., :3] *= out[..., 3:]
    return out


def linear_to_srgb(rgba):
    out = rgba.copy()
    rgb = out[..., :3]
    lo = rgb <= 0.0031308
    rgb[lo] = rgb[lo] * 12.92
    rgb[~lo] = 1.055 * np.power(rgb[~lo], 1.0 / 2.4) - 0.055
    return out


def srgb_to_linear(rgba):
    out = rgba.copy()
    rgb = out[..., :3]
    lo = rgb <= 0.04045
    rgb[lo] = rgb[lo] / 12.92
    rgb[~lo] = np.power((rgb[~lo] + 0.055) / 1.055, 2.4)
    return out


def paint_to_srgb(color):
    """Premultiplied linear colour -> premultiplied sRGB (:1015-1018, :1690-1693)."""
    return premultiply(linear_to_srgb(unpremultiply(np.asarray(color, dtype=F64))))


def convert(layer: OLayer, pre_alpha=None, linear_rgb=None) -> OLayer:
    """Layer.convert (:129-164)."""
    pre_alpha = layer.pre_alpha if pre_alpha is None else pre_alpha
    linear_rgb = layer.linear_rgb if linear_rgb is None else linear_rgb
    if layer.image.shape[2] == 1:
        return OLayer(layer.image, layer.offset, pre_alpha, linear_rgb)
    image, pre, lin = layer.image, layer.pre_alpha, layer.linear_rgb
    if lin != linear_rgb:
        if pre:
            image, pre = unpremultiply(image), False
        image = srgb_to_linear(image) if linear_rgb else linear_to_srgb(image)
        lin = linear_rgb
    if pre != pre_alpha:
        image = premultiply(image) if pre_alpha else unpremultiply(image)
        pre = pre_alpha
    if image is layer.image:
        return layer
    return OLayer(image, layer.offset, pre, lin)


# ------------------------------------------------------------------------------
# Porter-Duff (svgrasterize.py:277-298) and the bbox-aligned mergers (:304-416)
# ------------------------------------------------------------------------------
def blend(mode, dst, src):
    sa = src[..., -1:]
    da = dst[..., -1:]
    if mode == S.COMPOSE_OVER:
        return src + dst * (1 - sa)
    if mode == S.COMPOSE_OUT:
        return src * (1 - da)
    if mode == S.COMPOSE_IN:
        return src * da
    if mode == S.COMPOSE_ATOP:
        return src * da + dst * (1 - sa)
    if mode == S.COMPOSE_XOR:
        return src * (1 - da) + dst * (1 - sa)
    if isinstance(mode, tuple) and len(mode) == 4:
        k1, k2, k3, k4 = mode
        return (k1 * src * dst + k2 * src + k3 * dst + k4).clip(0, 1)
    raise ValueError(f"invalid compose mode: {mode}")


def _union(boxes):
    r0 = min(b[0] for b in boxes)
    c0 = min(b[1] for b in boxes)
    r1 = max(b[0] + b[2] for b in boxes)
    c1 = max(b[1] + b[3] for b in boxes)
    return r0, c0, r1 - r0, c1 - c0


def _intersection(boxes):
    r0 = max(b[0] for b in boxes)
    c0 = max(b[1] for b in boxes)
    r1 = min(b[0] + b[2] for b in boxes)
    c1 = min(b[1] + b[3] for b in boxes)
    return r0, c0, r1 - r0, c1 - c0


def _window(image, offset, box):
    r, c = box[0] - offset[0], box[1] - offset[1]
    return image[r: r + box[2], c: c + box[3]]


def merge_over(items, mode):
    """canvas_merge_union(full=False) (:366-377): first layer copied, the rest
    blended inside their own sub-rectangles."""
    box = _union([(*o, *im.shape[:2]) for im, o in items])
    out = np.zeros((box[2], box[3], 4))
    for k, (im, o) in enumerate(items):
        win = _window(out, box[:2], (*o, *im.shape[:2]))
        win[...] = im if k == 0 else blend(mode, win, im)
    return out, box[:2]


def merge_full(items, mode):
    """canvas_merge_union(full=True) (:353-364): every layer padded to the union."""
    box = _union([(*o, *im.shape[:2]) for im, o in items])
    out = None
    for im, o in items:
        full = np.zeros((box[2], box[3], 4))
        _window(full, box[:2], (*o, *im.shape[:2]))[...] = im
        out = full if out is None else blend(mode, out, full)
    return out, box[:2]


def merge_intersect(items, mode):
    """canvas_merge_intersect (:382-416)."""
    box = _intersection([(*o, *im.shape[:2]) for im, o in items])
    if box[2] <= 0 or box[3] <= 0:
        return None
    (first, fo), *rest = items
    out = _window(first, fo, box)
    if out.shape[2] == 1:
        out = np.broadcast_to(out, (box[2], box[3], 4))
    out = out.copy()
    for im, o in rest:
        out[...] = blend(mode, out, _window(im, o, box))
    return out, box[:2]


def compose(layers, method=S.COMPOSE_OVER, linear_rgb=False):
    """Layer.compose (:178-207)."""
    if not layers:
        return None
    if len(layers) == 1:
        return layers[0]
    pre = method in S.COMPOSE_PRE_ALPHA
    items = []
    for layer in layers:
        layer = convert(layer, pre_alpha=pre, linear_rgb=linear_rgb)
        items.append((layer.image, layer.offset))
    if method == S.COMPOSE_IN:
        res = merge_intersect(items, method)
    elif method == S.COMPOSE_OVER:
        res = merge_over(items, method)
    else:
        res = merge_full(items, method)
    if res is None:
        return None
    return OLayer(res[0], tuple(res[1]), pre, linear_rgb)


def merge_at(base, overlay, offset, mode=S.COMPOSE_OVER):
    """canvas_merge_at (:304-327): clipped in-place blend, result clipped to [0, 1]."""
    box = _intersection([(0, 0, *base.shape[:2]), (*offset, *overlay.shape[:2])])
    if box[2] <= 0 or box[3] <= 0:
        return base
    win = _window(base, (0, 0), box)
    win[...] = blend(mode, win, _window(overlay, offset, box)).clip(0, 1)
    return base


def opacity(layer, value, linear_rgb=False):
    layer = convert(layer, pre_alpha=True, linear_rgb=linear_rgb)
    return OLayer(layer.image * value, layer.offset, True, linear_rgb)


# ------------------------------------------------------------------------------
# paint (svgrasterize.py:995-1103, :1544-1695)
# ------------------------------------------------------------------------------
SPREAD = {"pad": 0, "repeat": 1, "reflect": 2}


def _m6(transform):
    return np.ascontiguousarray(transform.m[:2, :], dtype=F64).reshape(6)


def gradient_image(paint, bbox, user_tr, linear_rgb):
    """paint.fill(user_tr(grad_pixels(bbox))) (:1027-1031) -> (rows, cols, 4)."""
    L = clib.lib()
    r0, c0, rows, cols = bbox
    kind = S.paint_kind(paint)
    if paint.spread not in SPREAD:
        raise ValueError(f"invalid spread method: {paint.spread}")
    m1 = _m6(user_tr)
    m2 = None if paint.transform is None else _m6(paint.transform.invert)
    m2p = None if m2 is None else clib.dp(m2)
    t = np.empty((rows, cols))
    valid = None
    if kind == "linear":
        L2 = clib.declare("orc_grad_linear_t", None,
                          [clib.C.c_long] * 4 + [clib.c_double_p] * 5)
        L2(rows, cols, r0, c0, clib.dp(m1), m2p, clib.dp(np.asarray(paint.p0, dtype=F64)),
           clib.dp(np.asarray(paint.p1, dtype=F64)), clib.dp(t))
    else:
        has_focal = not (paint.fcenter is None and paint.fradius is None)
        fcenter = np.asarray(paint.center if paint.fcenter is None else paint.fcenter, dtype=F64)
        fradius = float(paint.fradius or 0)
        valid_buf = np.ones((rows, cols), dtype=np.uint8)
        fn = clib.declare(
            "orc_grad_radial_t", clib.C.c_int,
            [clib.C.c_long] * 4 + [clib.c_double_p] * 3 + [clib.C.c_double, clib.C.c_int, clib.c_double_p,
                                                            clib.C.c_double, clib.c_double_p,
                                                            clib.C.POINTER(clib.C.c_uint8)])
        any_neg = fn(rows, cols, r0, c0, clib.dp(m1), m2p, clib.dp(np.asarray(paint.center, dtype=F64)),
                     float(paint.radius), int(has_focal), clib.dp(fcenter), fradius, clib.dp(t),
                     valid_buf.ctypes.data_as(clib.C.POINTER(clib.C.c_uint8)))
        if any_neg:
            valid = valid_buf
    stops = paint.stops
    offs = np.asarray([o for o, _ in stops], dtype=F64)
    cols_ = np.asarray([c if linear_rgb else paint_to_srgb(c) for _, c in stops], dtype=F64)
    out = np.empty((rows, cols, 4))
    fn = clib.declare("orc_grad_colors", None,
                      [clib.C.c_long, clib.c_double_p, clib.C.POINTER(clib.C.c_uint8), clib.C.c_int, clib.C.c_long,
                       clib.c_double_p, clib.c_double_p, clib.c_double_p])
    fn(rows * cols, clib.dp(t), None if valid is None else valid.ctypes.data_as(clib.C.POINTER(clib.C.c_uint8)),
       SPREAD[paint.spread], len(stops), clib.dp(offs), clib.dp(np.ascontiguousarray(cols_)), clib.dp(out))
    return out


def viewbox_transform(bbox, viewbox):
    """svg_viewbox_transform (:3116-3133), needed by pattern paint (:1059)."""
    vx, vy, vw, vh = viewbox
    x, y, w, h = bbox
    if h is None and w is None:
        h, w = vh, vw
    elif h is None:
        h = vh * w / vw
    elif w is None:
        w = vw * h / vh
    scale = min(w / vw, h / vh)
    tx = -vx + (w / scale - vw) / 2 + x / scale
    ty = -vy + (h / scale - vh) / 2 + y / scale
    return S.Transform().scale(scale).translate(tx, ty)


def pixel_centres(bbox):
    r0, c0, rows, cols = bbox
    rr, cc = np.indices((rows, cols)).astype(F64)
    return np.stack([rr, cc], axis=2) + [r0 + 0.5, c0 + 0.5]


def fill_path(path, transform, paint, fill_rule=None, viewport=None, linear_rgb=True):
    """Path.fill (:995-1103) -> (OLayer, Cloud) or None."""
    if paint is None:
        return None
    res = _path_mask(path, transform, fill_rule, viewport)
    if res is None:
        return None
    mask, offset, edges = res
    cloud = Cloud(edges)
    kind = S.paint_kind(paint)
    bbox = (offset[0], offset[1], mask.shape[0], mask.shape[1])
    cov = mask[..., None]
    if kind == "solid":
        color = np.asarray(paint, dtype=F64)
        if not linear_rgb:
            color = paint_to_srgb(color)
        return OLayer(cov * color, offset, True, linear_rgb), cloud
    if kind in ("linear", "radial"):
        user_tr = (cloud.bbox_transform(transform) if paint.bbox_units else transform).invert
        if paint.linear_rgb is not None:
            linear_rgb = paint.linear_rgb
        image = gradient_image(paint, bbox, user_tr, linear_rgb)
        return OLayer(image * cov, offset, True, linear_rgb), cloud
    if kind == "pattern":
        pat_tr = transform.no_translate()
        if paint.scene_view_box:
            if paint.bbox_units:
                px, py, pw, ph = paint.bbox()
                _hx, _hy, hw, hh = cloud.bbox(transform)
                box = (px * hw, py * hh, pw * hw, ph * hh)
            else:
                box = paint.bbox()
            pat_tr = pat_tr @ viewbox_transform(box, paint.scene_view_box)
        elif paint.scene_bbox_units:
            pat_tr = cloud.bbox_transform(pat_tr)
        pat_tr = pat_tr @ paint.transform
        tile = render(paint.scene, pat_tr, linear_rgb=linear_rgb)
        if tile is None:
            return None
        tile = tile[0]
        rep = transform
        if paint.bbox_units:
            rep = cloud.bbox_transform(rep)
        rep = (rep @ paint.transform).no_translate()
        offs = rep.invert(pixel_centres(bbox))
        offs = rep(np.remainder(offs - [paint.x, paint.y], [paint.width, paint.height]))
        offs = offs.astype(int)
        corners = rep(np.array([[0, 0], [paint.width, 0], [0, paint.height], [paint.width, paint.height]], dtype=F64))
        hi = corners.max(axis=0).astype(int)
        lo = corners.min(axis=0).astype(int)
        offs -= lo
        pat = np.zeros((hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, 4))
        box = _intersection([(0, 0, *pat.shape[:2]), (tile.offset[0] - lo[0], tile.offset[1] - lo[1], *tile.image.shape[:2])])
        if box[2] <= 0 or box[3] <= 0:
            # canvas_merge_at returns None when the tile misses the repeat cell (:315-322) and the gather at :1094
            # subscripts it: the reference's failure mode for such a pattern
            raise TypeError("'NoneType' object is not subscriptable")
        pat = merge_at(pat, tile.image, (tile.offset[0] - lo[0], tile.offset[1] - lo[1]))
        image = pat[offs[..., 0], offs[..., 1]] * cov
        return OLayer(image, offset, tile.pre_alpha, tile.linear_rgb), cloud
    warnings.warn(f"fill method is not implemented: {paint}")
    return None


def mask_path(path, transform, fill_rule=None, viewport=None):
    res = _path_mask(path, transform, fill_rule, viewport)
    if res is None:
        return None
    mask, offset, edges = res
    return OLayer(mask[..., None], offset, True, True), Cloud(edges)


# ------------------------------------------------------------------------------
# filters (svgrasterize.py:1801-1957)
# ------------------------------------------------------------------------------
def blur_kernel(transform, sigma):
    """blur_kernel (:1903-1944) -> (kw, kh) weights or None (no-op blur)."""
    sx, sy = sigma
    basis = transform(np.eye(2)) - transform(np.zeros(2))
    scale_x, scale_y = np.linalg.norm(basis, axis=1)
    if scale_x * sx < 0.5 and scale_y * sy < 0.5:
        return None
    if scale_x * sx < 0.5:
        sx = 0.5 / scale_x
    elif scale_y * sy < 0.5:
        sy = 0.5 / scale_y
    ext = 2.5
    corners = np.array([[-ext * sx, -ext * sy], [-ext * sx, ext * sy], [ext * sx, ext * sy], [ext * sx, -ext * sy]])
    box = transform(corners) - transform(np.zeros(2))
    lo = box.min(axis=0).astype(int)
    hi = box.max(axis=0).astype(int)
    kw, kh = (int(v) for v in (hi - lo))
    kw += 1 - (kw & 1)
    kh += 1 - (kh & 1)
    inv = transform.invert
    pts = inv(pixel_centres((-kw / 2, -kh / 2, kw, kh))) - inv(np.zeros(2))
    w = np.exp(-np.square(pts) / (2 * np.square(np.array([sx, sy])))).prod(axis=-1)
    return w / w.sum()


def convolve(layer, kernel):
    """Layer.convolve (:106-115)."""
    layer = convert(layer, pre_alpha=False, linear_rgb=True)
    kw, kh = kernel.shape
    img = np.ascontiguousarray(layer.image)
    rows, cols, ch = img.shape
    out = np.empty((rows + kw - 1, cols + kh - 1, ch))
    fn = clib.declare("orc_convolve_full", None,
                      [clib.c_double_p, clib.C.c_long, clib.C.c_long, clib.C.c_long, clib.c_double_p, clib.C.c_long,
                       clib.C.c_long, clib.c_double_p])
    fn(clib.dp(img), rows, cols, ch, clib.dp(np.ascontiguousarray(kernel)), kw, kh, clib.dp(out))
    off = (int(layer.offset[0] - kw / 2), int(layer.offset[1] - kh / 2))
    return OLayer(out, off, False, True)


def morphology(layer, k0, k1, method):
    """Layer.morphology (:120-127) + pooling (:419-468), stride 1, no padding."""
    if method not in ("max", "min"):
        raise ValueError(f"invalid poll method: {method}")
    layer = convert(layer, pre_alpha=True, linear_rgb=True)
    img = np.ascontiguousarray(layer.image)
    rows, cols, ch = img.shape
    out = np.empty((max(rows - k0 + 1, 0), max(cols - k1 + 1, 0), ch))
    fn = clib.declare("orc_pool", None,
                      [clib.c_double_p] + [clib.C.c_long] * 5 + [clib.C.c_int, clib.c_double_p])
    fn(clib.dp(img), rows, cols, ch, k0, k1, int(method == "max"), clib.dp(out))
    return OLayer(out, layer.offset, True, True)


def color_matrix(layer, matrix):
    """Layer.color_matrix (:95-104)."""
    if not isinstance(matrix, np.ndarray) or matrix.shape != (4, 5):
        raise ValueError("expected 4x5 matrix")
    layer = convert(layer, pre_alpha=False, linear_rgb=True)
    image = np.matmul(layer.image, matrix[:, :4].T) + matrix[:, 4]
    return OLayer(np.clip(image, 0, 1), layer.offset, False, True)


def apply_filter(flt, transform, source):
    """Filter.__call__ (:1801-1831)."""
    alpha = OLayer(source.image[..., -1:] * np.array([0.0, 0.0, 0.0, 1.0]), source.offset, True, True)
    stack = [alpha, convert(source, pre_alpha=False, linear_rgb=True)]
    for tag, attrs, inputs in flt.filters:
        args = [stack[i] for i in inputs]
        if tag == S.FE_OFFSET:
            dx, dy = attrs
            (lay,) = args
            r, c = lay.offset
            tr, tc = transform(transform.invert(np.array([r, c], dtype=F64)) + [dx, dy])
            out = OLayer(lay.image, (r + (int(tr) - r), c + (int(tc) - c)), lay.pre_alpha, lay.linear_rgb)
        elif tag == S.FE_MERGE:
            out = compose(args, linear_rgb=True)
        elif tag == S.FE_BLEND:
            warnings.warn("feBlend is not properly supported")
            out = compose([args[1], args[0]], linear_rgb=True)
        elif tag == S.FE_COMPOSITE:
            out = compose([args[1], args[0]], attrs[0], linear_rgb=True)
        elif tag == S.FE_GAUSSIAN_BLUR:
            sx, sy = attrs
            kernel = blur_kernel(transform, (sx, sx if sy is None else sy))
            out = args[0] if kernel is None else convolve(args[0], kernel)
        elif tag == S.FE_COLOR_MATRIX:
            (matrix,) = attrs
            if not isinstance(matrix, np.ndarray) or matrix.shape != (4, 5):
                warnings.warn(f"invalid color matrix: {matrix}")
                out = args[0]
            else:
                out = color_matrix(args[0], matrix)
        elif tag == S.FE_MORPHOLOGY:
            rx, ry, method = attrs
            u = transform(np.array([[rx, 0], [0, ry]], dtype=F64)) - transform(np.zeros((2, 2)))
            k0 = int(np.linalg.norm(u[0]) * 2)
            k1 = int(np.linalg.norm(u[1]) * 2)
            out = args[0] if (k0 < 1 or k1 < 1) else morphology(args[0], k0, k1, method)
        else:
            raise ValueError(f"unsupported filter type: {tag}")
        stack.append(out)
    return stack[-1]


# ------------------------------------------------------------------------------
# scene walk (svgrasterize.py:649-752)
# ------------------------------------------------------------------------------
LUMA = [0.2125, 0.7154, 0.072]  # :735 (not the 0.0721 of COLOR_MATRIX_LUM)


def render(scene, transform, mask_only=False, viewport=None, linear_rgb=False):
    tag, args = scene
    if tag == S.RENDER_FILL:
        path, paint, rule = args
        if mask_only:
            return mask_path(path, transform, rule, viewport)
        return fill_path(path, transform, paint, rule, viewport, linear_rgb)
    if tag == S.RENDER_STROKE:
        path, paint, width, cap, join = args
        outline = _stroke(path, width, cap, join)
        if mask_only:
            return mask_path(outline, transform, None, viewport)
        return fill_path(outline, transform, paint, None, viewport, linear_rgb)
    if tag == S.RENDER_GROUP:
        layers, clouds = [], []
        for child in args:
            res = render(child, transform, mask_only, viewport, linear_rgb)
            if res is not None:
                layers.append(res[0])
                clouds.append(res[1])
        group = compose(layers, S.COMPOSE_OVER, linear_rgb)
        if group is None:
            return None
        return group, Cloud.merge(clouds)
    if tag == S.RENDER_OPACITY:
        res = render(args[0], transform, mask_only, viewport, linear_rgb)
        if res is None:
            return None
        return opacity(res[0], args[1], linear_rgb), res[1]
    if tag == S.RENDER_TRANSFORM:
        return render(args[0], transform @ args[1], mask_only, viewport, linear_rgb)
    if tag in (S.RENDER_CLIP, S.RENDER_MASK):
        target, other, bbox_units = args
        res = render(target, transform, mask_only, viewport, linear_rgb)
        if res is None:
            return None
        image, cloud = res
        if bbox_units:
            transform = cloud.bbox_transform(transform)
        if tag == S.RENDER_CLIP:
            sub = render(other, transform, True, viewport, linear_rgb)
            if sub is None:
                return None
            stencil = sub[0]
        else:
            sub = render(other, transform, mask_only, viewport, linear_rgb)
            if sub is None:
                return None
            m = convert(sub[0], pre_alpha=False, linear_rgb=linear_rgb)
            luma = m.image[..., :3] @ LUMA * m.image[..., 3]
            stencil = OLayer(luma[..., None], m.offset, False, linear_rgb)
        out = compose([stencil, image], S.COMPOSE_IN, linear_rgb)
        if out is None:
            return None
        return out, cloud
    if tag == S.RENDER_FILTER:
        res = render(args[0], transform, mask_only, viewport, linear_rgb)
        if res is None:
            return None
        return apply_filter(args[1], transform, res[0]), res[1]
    raise ValueError(f"unhandled scene type: {tag}")


def canvas_transform():
    """The x/y swap every render starts from (svgrasterize.py:246, :3823)."""
    return S.Transform().matrix(0, 1, 0, 1, 0, 0)


def quantize(image):
    """canvas_to_png's float -> uint8 step (:263): round-half-even of x*255."""
    return np.round(image * 255.0).astype(np.uint8)


def render_canvas(scene, size, linear_rgb=False, bg=None, transform=None):
    """main() :3854-3881 up to the uint8 array handed to zlib: render with the
    canvas viewport, blit onto a zero canvas (clipped to [0,1]), optional
    background, convert to straight-alpha sRGB, quantize.  size = (w, h) as in
    the reference; returns (h, w, 4) uint8, or None when nothing rendered."""
    w, h = size
    tr = canvas_transform() if transform is None else transform
    res = render(scene, tr, viewport=[0, 0, int(h), int(w)], linear_rgb=linear_rgb)
    if res is None:
        return None
    out = convert(res[0], pre_alpha=True, linear_rgb=linear_rgb)
    base = np.zeros((int(h), int(w), 4))
    if out.image.shape[2] == 1:
        raise ValueError("Only RGBA layers are supported")
    merge_at(base, out.image, out.offset)
    layer = OLayer(base, (0, 0), True, linear_rgb)
    if bg is not None:
        layer = convert(layer, pre_alpha=True, linear_rgb=True)
        layer = OLayer(blend(S.COMPOSE_OVER, np.asarray(bg, dtype=F64)[None, None, :], layer.image), (0, 0), True, True)
    layer = convert(layer, pre_alpha=False, linear_rgb=False)
    return quantize(layer.image)
