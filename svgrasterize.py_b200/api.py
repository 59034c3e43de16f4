"""The reference's Python call surface for the hot path, backed by the CUDA core.

Same names, argument meaning, return types and None / exception behaviour as
svgrasterize.py, so that its Scene tree, CLI and helper scripts can call into
this core unchanged:

    Path.mask / Path.fill / Path.stroke           svgrasterize.py:922 / :995 / :1105
    Layer + its methods, Layer.compose            svgrasterize.py:61-232
    canvas_create / compose / merge_* / to_png    svgrasterize.py:235-416
    pooling                                       svgrasterize.py:419
    Scene.render                                  svgrasterize.py:649
    Filter.__call__                               svgrasterize.py:1801
    bezier3_flatten_batch, blur_kernel            svgrasterize.py:2091, :1903

Every call encodes a (tiny) scene program and runs it through the same engine
the batch path uses; images come back as float32 numpy arrays (`Layer.image`
stays a plain ndarray, so callers such as font_speciment.py can keep mutating
it).  There is no CPU implementation behind any of these: without the library
or a GPU they raise.
"""
from __future__ import annotations

import io
import struct
import zlib
from typing import NamedTuple

import numpy as np

from . import _lib
from . import scene as S
from .encode import Encoder, blur_kernel, canvas_transform, device_path  # noqa: F401  (re-exported)
from .engine import default_engine
from .sceneio import path_from_arrays

COMPOSE_OVER, COMPOSE_OUT, COMPOSE_IN, COMPOSE_ATOP, COMPOSE_XOR = 0, 1, 2, 3, 4
COMPOSE_PRE_ALPHA = S.COMPOSE_PRE_ALPHA


class Hull:
    """Stand-in for ConvexHull (svgrasterize.py:1963-2029).  Callers only ever consume the user-space
    bounding box of the hull (:2002-2023); the extremes of an affine image of a point set are attained on
    its hull, so the flattened end points are kept as they are."""

    __slots__ = ("points",)

    def __init__(self, points):
        self.points = np.asarray(points, dtype=np.float64).reshape(-1, 2)

    @classmethod
    def merge(cls, hulls):
        return cls(np.concatenate([h.points for h in hulls]))

    def bbox(self, transform):
        pts = transform.invert(self.points)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        return [lo[0], lo[1], hi[0] - lo[0], hi[1] - lo[1]]

    def bbox_transform(self, transform):
        x, y, w, h = self.bbox(transform)
        if w <= 0 and h <= 0:
            return transform
        return transform.translate(x, y).scale(w, h)


class Layer(NamedTuple):
    image: np.ndarray
    offset: tuple
    pre_alpha: bool
    linear_rgb: bool

    @property
    def x(self):
        return self.offset[0]

    @property
    def y(self):
        return self.offset[1]

    @property
    def width(self):
        return self.image.shape[1]

    @property
    def height(self):
        return self.image.shape[0]

    @property
    def channels(self):
        return self.image.shape[2]

    @property
    def bbox(self):
        return (*self.offset, *self.image.shape[:2])

    def translate(self, x, y):
        return Layer(self.image, (self.x + x, self.y + y), self.pre_alpha, self.linear_rgb)

    # -- device ops ---------------------------------------------------------------------------
    def _run(self, build):
        """build(encoder, node of self) -> node; returns the resulting Layer (or None)."""
        eng = default_engine()
        enc = Encoder(eng)
        src = enc.add_external(self.image, self.offset, self.pre_alpha, self.linear_rgb)
        node = build(enc, src)
        return _read(eng, enc, node)

    def color_matrix(self, matrix):
        if not isinstance(matrix, np.ndarray) or matrix.shape != (4, 5):
            raise ValueError("expected 4x5 matrix")

        def build(enc, src):
            enc.matrices.append(np.asarray(matrix, dtype=np.float32).reshape(20))
            return enc._node(_lib.N_CMATRIX, len(enc.matrices) - 1, children=[src])

        return self._run(build)

    def convolve(self, kernel):
        kernel = np.asarray(kernel, dtype=np.float64)
        return self._run(lambda enc, src: enc._blur(src, kernel))

    def morphology(self, x, y, method):
        if method not in ("max", "min"):
            raise ValueError(f"invalid poll method: {method}")
        return self._run(lambda enc, src: enc._node(_lib.N_MORPH, int(x), int(y), int(method == "max"), children=[src]))

    def convert(self, pre_alpha=None, linear_rgb=None):
        pre = self.pre_alpha if pre_alpha is None else pre_alpha
        lin = self.linear_rgb if linear_rgb is None else linear_rgb
        if self.channels == 1:
            return Layer(self.image, self.offset, pre, lin)
        if pre == self.pre_alpha and lin == self.linear_rgb:
            return self
        return self._run(lambda enc, src: enc._node(_lib.N_CONVERT, int(pre), int(lin), children=[src]))

    def background(self, color):
        layer = self.convert(pre_alpha=True, linear_rgb=True)
        bg = np.broadcast_to(np.asarray(color, dtype=np.float32), (*layer.image.shape[:2], 4))
        return Layer.compose([Layer(np.ascontiguousarray(bg), layer.offset, True, True), layer], COMPOSE_OVER, True)

    def opacity(self, opacity, linear_rgb=False):
        return self._run(lambda enc, src: enc._node(_lib.N_OPACITY, children=[src], flags=int(bool(linear_rgb)),
                                                    f=(opacity, 0, 0, 0)))

    @staticmethod
    def compose(layers, method=COMPOSE_OVER, linear_rgb=False):
        """Layer.compose (svgrasterize.py:178-207)."""
        layers = list(layers)
        if not layers:
            return None
        if len(layers) == 1:
            return layers[0]
        eng = default_engine()
        enc = Encoder(eng)
        nodes = [enc.add_external(l.image, l.offset, l.pre_alpha, l.linear_rgb) for l in layers]
        return _read(eng, enc, enc._compose(nodes, method, linear_rgb))

    def write_png(self, output=None):
        if self.channels != 4:
            raise ValueError("Only RGBA layers are supported")
        layer = self.convert(pre_alpha=False, linear_rgb=False)
        return canvas_to_png(layer.image, output)


def _read(eng, enc, node):
    """Materialise `node`, run the program, read the layer back."""
    if enc._is_empty(node):
        return None
    root = enc._node(_lib.N_CONVERT, -1, -1, children=[node], flags=2)
    prog = enc.finish()
    eng.render(prog)
    res = eng.node(root)
    if res is None:
        return None
    img, offset, pre, lin = res
    return Layer(img, offset, pre, lin)


def _hull(eng, paths):
    edges, edge_path = eng.edges()
    boxes = eng.boxes()
    alive = [p for p in paths if boxes[p][2] > 0 and boxes[p][3] > 0]
    sel = np.isin(edge_path, np.asarray(alive, dtype=np.uint32))
    return Hull(edges[sel].reshape(-1, 2))


# ---------------------------------------------------------------------------------------------
# Path
# ---------------------------------------------------------------------------------------------
def path_mask(path, transform, fill_rule=None, viewport=None):
    """Path.mask (svgrasterize.py:922-993) -> (Layer, Hull) or None."""
    eng = default_engine()
    enc = Encoder(eng)
    node = enc.encode(S.Scene.fill(path, None, fill_rule), transform, True, viewport, True)
    layer = _read(eng, enc, node)
    if layer is None:
        return None
    return Layer(layer.image, layer.offset, True, True), _hull(eng, [0])


def path_fill(path, transform, paint, fill_rule=None, viewport=None, linear_rgb=True):
    """Path.fill (svgrasterize.py:995-1103) -> (Layer, Hull) or None."""
    if paint is None:
        return None
    eng = default_engine()
    enc = Encoder(eng)
    node = enc.encode(S.Scene.fill(path, paint, fill_rule), transform, False, viewport, linear_rgb)
    layer = _read(eng, enc, node)
    if layer is None:
        return None
    return layer, _hull(eng, [0])


def path_stroke(path, width, linecap=None, linejoin=None):
    """Path.stroke (svgrasterize.py:1105-1180) -> outline Path (LINE / QUAD / CUBIC segments)."""
    eng = default_engine()
    enc = Encoder(eng)
    enc.add_stroke_path(path, S.Transform(), width, linecap, linejoin, None)
    eng.render(enc.finish(), stop=_lib.STOP_STROKE)
    tag, data, _path, sub = eng.outline()
    bounds = np.concatenate([[0], np.nonzero(np.diff(sub))[0] + 1, [len(sub)]]) if len(sub) else np.zeros(1, int)
    out = path_from_arrays(tag, data, bounds.astype(np.int32))
    # hand back the caller's own Path class (the reference's, once install() has rebound its methods)
    return out if type(path) is S.Path else type(path)(out.subpaths)


def bezier3_flatten_batch(batch, flatness=0.1):
    """bezier3_flatten_batch (svgrasterize.py:2091-2098): (M, 4, 2) cubics -> (E, 2, 2) lines.  The device emits
    the same set of lines as the reference, in a different order."""
    flatness = float(flatness)
    if not flatness > 0.0:
        raise ValueError("flatness must be positive (the reference never terminates otherwise)")
    batch = np.asarray(batch, dtype=np.float64).reshape(-1, 4, 2)
    path = S.Path([[(S.PATH_CUBIC, c) for c in batch]]) if len(batch) else S.Path([])
    eng = default_engine()
    enc = Encoder(eng)
    enc.add_fill_path(path, S.Transform(), None, None)
    prog = enc.finish()
    prog.flatness = flatness
    eng.render(prog, stop=_lib.STOP_FLATTEN)
    edges, _ = eng.edges()
    return edges.reshape(-1, 2, 2)


def line_signed_coverage(canvas, line):
    """line_signed_coverage(canvas, line) (svgrasterize.py:2213-2304): adds the signed-area deltas of `line`
    ((r0, c0), (r1, c1)) -- or of an (n, 2, 2) batch of lines -- to the 2-D trace `canvas` in place and returns
    it.  The device accumulates in float32 (like the tile kernel of Path.mask)."""
    if canvas.ndim != 2:
        raise ValueError("canvas must be a 2-D trace")
    lines = np.asarray(line, dtype=np.float64).reshape(-1, 4)
    trace = np.ascontiguousarray(canvas, dtype=np.float32)
    if trace is canvas or np.shares_memory(trace, canvas):
        default_engine().line_signed_coverage(trace, lines)
    else:
        canvas[...] = default_engine().line_signed_coverage(trace, lines)
    return canvas


def grad_pixels(viewport):
    """grad_pixels (svgrasterize.py:1653-1658): (width, height, 2) float64 pixel centres of a viewport."""
    off_x, off_y, width, height = viewport
    return default_engine().grad_pixels(off_x, off_y, width, height)


def grad_spread(offsets, spread):
    """grad_spread (svgrasterize.py:1661-1668)."""
    from .encode import SPREAD

    if spread not in SPREAD:
        raise ValueError(f"invalid spread method: {spread}")
    return default_engine().grad_spread(np.asarray(offsets, dtype=np.float64), SPREAD[spread])


def grad_interpolate(offset, stops, linear_rgb):
    """grad_interpolate (svgrasterize.py:1671-1683): offsets (...) -> premultiplied RGBA (..., 4) float32; the
    stop colours are converted to sRGB on the host when not linear_rgb (grad_stops_colorspace, :1686)."""
    from .color import paint_to_srgb

    if not stops:
        raise ValueError("gradient without stops")
    rec = np.zeros(len(stops), _lib.STOP_DT)
    offs = [float(o) for o, _ in stops]
    for k, (o, c) in enumerate(stops):
        rec["offset"][k] = float(o)
        rec["color"][k] = np.asarray(c, dtype=np.float64) if linear_rgb else paint_to_srgb(c)
        with np.errstate(divide="ignore"):
            rec["inv_span"][k] = np.float64(1.0) / np.float64(offs[k + 1] - offs[k]) if k + 1 < len(offs) else 0.0
    return default_engine().grad_interpolate(np.asarray(offset, dtype=np.float64), rec)


def path_from_svg(input: str):
    """Path.from_svg (svgrasterize.py:1252-1430) by the native path-data reader (csrc/pathdata.cpp): the path holds
    flat segment arrays (what the encoder consumes) and unpacks them into ``subpaths`` only on demand."""
    import ctypes as C

    raw = input.encode("utf-8") if isinstance(input, str) else bytes(input)
    cap = len(raw) // 2 + 8  # every segment and every sub-path costs at least two characters
    tags = np.empty(cap, np.uint8)
    data = np.empty((cap, 8), np.float64)
    sub_off = np.empty(cap + 1, np.int32)
    n_seg, n_sub = C.c_int64(), C.c_int64()
    err = C.create_string_buffer(256)
    rc = _lib.lib().svgr_path_from_svg(raw, len(raw), tags.ctypes.data, data.ctypes.data, cap, sub_off.ctypes.data, cap,
                                      C.byref(n_seg), C.byref(n_sub), err, 256)
    if rc != 0:
        raise ValueError(err.value.decode() or "malformed path data")
    return path_from_arrays(tags[: n_seg.value].copy(), data[: n_seg.value].copy(), sub_off[: n_sub.value + 1].copy())


S.Path.mask = path_mask
S.Path.fill = path_fill
S.Path.stroke = path_stroke
S.Path.from_svg = staticmethod(path_from_svg)


# ---------------------------------------------------------------------------------------------
# Scene / Filter
# ---------------------------------------------------------------------------------------------
def scene_render(scene, transform, mask_only=False, viewport=None, linear_rgb=False):
    """Scene.render (svgrasterize.py:649-752) -> (Layer, Hull) or None; the whole tree is one program."""
    eng = default_engine()
    enc = Encoder(eng)
    node = enc.encode(scene, transform, mask_only, viewport, linear_rgb)
    cloud = enc.cloud.get(node, [])
    layer = _read(eng, enc, node)
    if layer is None:
        return None
    alive = {}

    def live(n):
        if n not in alive:
            alive[n] = eng.node_info(n)[0] != 0
        return alive[n]

    return layer, _hull(eng, sorted({p for n, p in cloud if live(n)}))


S.Scene.render = scene_render


def filter_call(flt, transform, source):
    """Filter.__call__ (svgrasterize.py:1801-1831)."""
    return source._run(lambda enc, src: enc._filter(flt, transform, src))


S.Filter.__call__ = filter_call


def render_canvas(scene, size, linear_rgb=False):
    """main() of the reference up to the uint8 array handed to the PNG encoder (svgrasterize.py:3854-3881)."""
    eng = default_engine()
    enc = Encoder(eng)
    enc.add_scene(scene, size, linear_rgb)
    prog = enc.finish()
    res = eng.render(prog)
    return eng.canvas(prog, res["canvas"]).copy()


def render_png(scene, size, linear_rgb=False) -> bytes:
    """main() of the reference from the Scene to the PNG file (svgrasterize.py:3854-3881), with the file made on the
    device (Paeth filter + dynamic-Huffman deflate, csrc/k_png.cu): decodes to the same pixels as render_canvas,
    but is not the reference's byte stream (that is canvas_to_png's default path)."""
    eng = default_engine()
    enc = Encoder(eng)
    enc.add_scene(scene, size, linear_rgb)
    res = eng.render_png(enc.finish())
    return res["png"].tobytes()


def render_png_batch(jobs, processes: int = 0):
    """[(scene, size, linear_rgb)] -> list of PNG files, all rendered and encoded as one batch."""
    from .encode import encode_batch

    eng = default_engine()
    prog = encode_batch(jobs, processes)
    res = eng.render_png(prog)
    off = res["offsets"]
    return [res["png"][off[i]: off[i + 1]].tobytes() for i in range(len(off) - 1)]


# ---------------------------------------------------------------------------------------------
# canvas helpers (svgrasterize.py:235-468)
# ---------------------------------------------------------------------------------------------
def canvas_create(width, height, bg=None):
    """-> (canvas (height, width, 4), canvas transform) (svgrasterize.py:235-246)."""
    if bg is None:
        canvas = np.zeros((height, width, 4), dtype=np.float32)
    else:
        canvas = np.broadcast_to(np.asarray(bg, dtype=np.float32), (height, width, 4)).copy()
    return canvas, canvas_transform()


def _as_layer(image, offset=(0, 0)):
    image = np.asarray(image, dtype=np.float32)
    if image.ndim == 2:
        image = image[..., None]
    return Layer(np.ascontiguousarray(image), tuple(int(v) for v in offset), True, True)


def _blend_mode(blend, default=COMPOSE_OVER):
    """The reference passes blends around as functools.partial(canvas_compose, mode) (svgrasterize.py:301, :198);
    a bare mode is accepted too.  Arbitrary Python callables cannot run on the device."""
    import functools

    if blend is None:
        return default
    if isinstance(blend, functools.partial) and blend.args and getattr(blend.func, "__name__", "") == "canvas_compose":
        return blend.args[0]
    if isinstance(blend, (int, tuple)) and not isinstance(blend, bool):
        return blend
    raise TypeError("blend must be a compose mode or functools.partial(canvas_compose, mode)")


def _raw_compose(layers, mode, intersect=False, clip_box=None):
    """canvas_compose folded over plain (image, offset) arrays: no Layer.convert, union box (zero padded) or
    intersection box; clip_box = (r0, c0, rows, cols): the result is placed on zeros of that box and clipped to
    [0, 1] (canvas_merge_at).  -> Layer or None"""
    eng = default_engine()
    enc = Encoder(eng)
    nodes = [enc.add_external(l.image, l.offset, True, True) for l in layers]
    node = enc._compose(nodes, mode, True, intersect=intersect, raw=True)
    if clip_box is not None:
        node = enc._node(_lib.N_MERGE_AT, *clip_box, children=[node])
    return _read(eng, enc, node)


def canvas_compose(mode, dst, src):
    """canvas_compose (svgrasterize.py:277-298): blends two broadcast-compatible premultiplied images as they
    are (no colour conversion); the alpha of an image is its last channel, a 2-D or one-channel image is alpha."""
    if not (isinstance(mode, tuple) and len(mode) == 4) and (isinstance(mode, bool) or mode not in (0, 1, 2, 3, 4)):
        raise ValueError(f"invalid compose mode: {mode}")
    dst, src = np.asarray(dst, dtype=np.float32), np.asarray(src, dtype=np.float32)
    flat = dst.ndim == 2 and src.ndim == 2
    d3 = dst[..., None] if dst.ndim == 2 else dst
    s3 = src[..., None] if src.ndim == 2 else src
    one = d3.shape[-1] == 1 and s3.shape[-1] == 1
    shape = np.broadcast_shapes(d3.shape[:2], s3.shape[:2])
    d3 = np.broadcast_to(d3, (*shape, d3.shape[-1]))
    s3 = np.broadcast_to(s3, (*shape, s3.shape[-1]))
    out = _raw_compose([_as_layer(d3), _as_layer(s3)], mode).image
    if one:
        out = out[..., 3:]
    return out[..., 0] if flat else out


def canvas_merge_at(base, overlay, offset, blend=None):
    """canvas_merge_at (svgrasterize.py:304-327): `overlay` is blended onto `base` at `offset` in place; only the
    affected rectangle changes and is clipped to [0, 1].  Returns `base`, or None when the overlay misses it."""
    mode = _blend_mode(blend)
    x, y = (int(v) for v in offset)
    b_h, b_w = base.shape[:2]
    o_h, o_w = overlay.shape[:2]
    clip = lambda v, lo, hi: lo if v < lo else hi if v > hi else v  # noqa: E731
    b_x_low, b_x_high = clip(x, 0, b_h), clip(x + o_h, 0, b_h)
    b_y_low, b_y_high = clip(y, 0, b_w), clip(y + o_w, 0, b_w)
    effected = base[b_x_low:b_x_high, b_y_low:b_y_high]
    if effected.size == 0:
        return None
    o_x_low, o_x_high = clip(-x, 0, o_h), clip(b_h - x, 0, o_h)
    o_y_low, o_y_high = clip(-y, 0, o_w), clip(b_w - y, 0, o_w)
    overlay = overlay[o_x_low:o_x_high, o_y_low:o_y_high]
    if overlay.size == 0:
        return None
    out = _raw_compose([_as_layer(effected), _as_layer(overlay)], mode, clip_box=(0, 0, *effected.shape[:2]))
    effected[...] = out.image if effected.ndim == 3 and effected.shape[2] == 4 else out.image[..., 3:].reshape(effected.shape)
    return base


def canvas_merge_union(layers, full=True, blend=None):
    """canvas_merge_union (svgrasterize.py:330-379): layers = [(image, offset)] -> (image, offset) over the union
    of the boxes, every layer zero padded to it (`full`); for the over blend the sub-rectangle fast path
    (full=False, the only way the reference calls it, :201) gives the same pixels."""
    mode = _blend_mode(blend)
    layers = list(layers)
    if not layers:
        raise ValueError("can not blend zero layers")
    if len(layers) == 1:
        return layers[0]
    if not full and mode != COMPOSE_OVER:
        raise ValueError("the sub-rectangle fast path (full=False) is only defined for the over blend "
                         "(svgrasterize.py:199-203 never calls it otherwise)")
    out = _raw_compose([_as_layer(im, off) for im, off in layers], mode)
    return out.image, out.offset


def canvas_merge_intersect(layers, blend=None):
    """canvas_merge_intersect (svgrasterize.py:382-416): blend on the intersection of the boxes -> (image, offset),
    or None when the intersection is empty."""
    mode = _blend_mode(blend)
    layers = list(layers)
    if not layers:
        raise ValueError("can not blend zero layers")
    if len(layers) == 1:
        return layers[0]
    out = _raw_compose([_as_layer(im, off) for im, off in layers], mode, intersect=True)
    return None if out is None else (out.image, out.offset)


def pooling(mat, ksize, stride=None, method="max", pad=False):
    """pooling (svgrasterize.py:419-468): overlapping pooling of a 2-D or 3-D array, NaN-ignoring max / min /
    mean, optional NaN padding to ceil(n / stride) outputs."""
    methods = {"max": 0, "min": 1, "mean": 2}
    if method not in methods:
        raise ValueError(f"invalid poll method: {method}")
    ky, kx = (int(v) for v in ksize)
    sy, sx = (ky, kx) if stride is None else (int(v) for v in stride)
    mat = np.asarray(mat, dtype=np.float32)
    if mat.ndim not in (2, 3):
        raise ValueError("pooling expects a 2-D or 3-D array")
    img = mat[..., None] if mat.ndim == 2 else mat
    out = default_engine().pooling(img, (ky, kx), (sy, sx), methods[method], bool(pad))
    return out[..., 0] if mat.ndim == 2 else out


def _deflate_parallel(raw: bytes, level: int, threads: int, block: int = 1 << 20) -> bytes:
    """One zlib stream made of independently compressed blocks (the pigz construction): every block but the
    last ends on a sync flush (byte aligned, not final), the last one finishes the stream; the Adler-32 trailer
    covers the whole input.  zlib releases the GIL, so the blocks run on `threads` host cores."""
    from concurrent.futures import ThreadPoolExecutor

    n = max(1, (len(raw) + block - 1) // block)
    view = memoryview(raw)

    def pack(i):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)  # raw deflate, no header / trailer
        body = c.compress(view[i * block: (i + 1) * block])
        return body + (c.flush(zlib.Z_FINISH) if i == n - 1 else c.flush(zlib.Z_FULL_FLUSH))

    with ThreadPoolExecutor(max_workers=threads) as pool:
        parts = list(pool.map(pack, range(n)))
    head = b"\x78\xda" if level >= 7 else (b"\x78\x9c" if level >= 6 else b"\x78\x01" if level < 2 else b"\x78\x5e")
    return head + b"".join(parts) + struct.pack(">I", zlib.adler32(raw) & 0xFFFFFFFF)


def canvas_to_png(canvas, output=None, threads=1, level=9, device=False):
    """canvas_to_png (svgrasterize.py:249-274): straight-alpha sRGB float image -> PNG bytes.  The
    float -> uint8 quantisation (:263) is the in-scope part; deflate stays on the host like the reference.

    Returns `output` (a fresh io.BytesIO when None, like the reference).  With the defaults the bytes are the reference's (filter 0 rows, one zlib stream at level 9).  `threads`
    > 1 compresses 1 MiB blocks of the same filtered rows on that many host cores (level 9 on
    one core takes 6.7 s for a 4096 x 4096 canvas): the file decodes to the same pixels but is not
    byte-identical and a few percent larger."""
    canvas = np.asarray(canvas)
    if canvas.dtype != np.uint8:
        canvas = default_engine().quantize_u8(canvas)  # np.round(canvas * 255).astype(uint8) on the device (:263)
    if device:
        # the whole file on the device (csrc/k_png.cu): same pixels, a different (Paeth + dynamic Huffman) stream
        png = default_engine().png_encode([canvas])[0]
        output = io.BytesIO() if output is None else output
        output.write(png)
        return output
    height, width = canvas.shape[:2]
    rows = np.zeros((height, 1 + width * 4), dtype=np.uint8)  # filter type 0 in front of every row
    rows[:, 1:] = canvas.reshape(height, width * 4)
    raw = rows.tobytes()

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    idat = zlib.compress(raw, level) if threads <= 1 else _deflate_parallel(raw, level, int(threads))
    png = b"".join([b"\x89PNG\r\n\x1a\n", chunk(b"IHDR", struct.pack(">2I5B", width, height, 8, 6, 0, 0, 0)),
                    chunk(b"IDAT", idat), chunk(b"IEND", b"")])
    output = io.BytesIO() if output is None else output  # svgrasterize.py:267
    output.write(png)
    return output


# ---------------------------------------------------------------------------------------------
# drop-in installer
# ---------------------------------------------------------------------------------------------
_MODULE_FUNCTIONS = ("canvas_create", "canvas_to_png", "canvas_compose", "canvas_merge_at", "canvas_merge_union",
                     "canvas_merge_intersect", "pooling", "bezier3_flatten_batch", "line_signed_coverage",
                     "grad_pixels", "grad_spread", "grad_interpolate", "blur_kernel")
_MISSING = object()


def install(module):
    """Rebind the hot-path entry points of a reference-shaped module -- the reference's own `svgrasterize`
    module, or anything with the same names -- to this core, so that its parser, Scene builders, CLI `main()`
    (svgrasterize.py:3795-3881) and helper scripts run on the GPU unmodified:

        Path.mask / fill / stroke (:922 / :995 / :1105), Scene.render (:649), Filter.__call__ (:1801),
        Layer (:61-232) and the module functions canvas_* / pooling (:235-468), bezier3_flatten_batch (:2091),
        line_signed_coverage (:2213), grad_pixels / grad_spread / grad_interpolate (:1653-1683), blur_kernel (:1903).

    Path.from_svg (:1252) is rebound to the native path-data reader, which the module's parser and fonts then use for
    every `d` attribute and glyph.  Everything else (Transform, paints, the SVG parser, fonts) stays the module's own.  Returns a token for
    uninstall()."""
    import functools

    saved = {}

    def bind(owner, name, value):
        saved[(owner, name)] = owner.__dict__.get(name, _MISSING) if isinstance(owner, type) else getattr(owner, name, _MISSING)
        setattr(owner, name, value)

    bind(module.Path, "mask", path_mask)
    bind(module.Path, "fill", path_fill)
    bind(module.Path, "stroke", path_stroke)
    if isinstance(module.Path, type) and issubclass(module.Path, S.Path):
        bind(module.Path, "from_svg", staticmethod(path_from_svg))  # array-backed paths need this package's Path
    elif callable(getattr(module.Path, "from_svg", None)):
        # the module's own Path class: the native reader tokenises (svgrasterize.py:1252-1430, bit-identical
        # segments), the sub-path lists the class expects are built from its arrays
        def from_svg(input, _Path=module.Path):
            return _Path(path_from_svg(input).subpaths)

        bind(module.Path, "from_svg", staticmethod(from_svg))
    bind(module.Scene, "render", scene_render)
    bind(module.Filter, "__call__", filter_call)
    bind(module, "Layer", Layer)
    g = globals()
    for name in _MODULE_FUNCTIONS:
        bind(module, name, g[name])
    bind(module, "CANVAS_COMPOSE_OVER", functools.partial(canvas_compose, COMPOSE_OVER))
    return saved


def uninstall(saved):
    """Undo install()."""
    for (owner, name), value in saved.items():
        if value is _MISSING:
            try:
                delattr(owner, name)
            except AttributeError:
                pass
        else:
            setattr(owner, name, value)
