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
    return path_from_arrays(tag, data, bounds.astype(np.int32))


def bezier3_flatten_batch(batch, flatness=0.1):
    """bezier3_flatten_batch (svgrasterize.py:2091): (M, 4, 2) cubics -> (E, 2, 2) lines.  The device emits
    the same set of lines as the reference, in a different order.  Only the reference's own flatness (0.1)
    is wired through the C-ABI."""
    if flatness != 0.1:
        raise NotImplementedError("the device flattener uses the reference's literal flatness of 0.1")
    batch = np.asarray(batch, dtype=np.float64).reshape(-1, 4, 2)
    path = S.Path([[(S.PATH_CUBIC, c) for c in batch]]) if len(batch) else S.Path([])
    eng = default_engine()
    enc = Encoder(eng)
    enc.add_fill_path(path, S.Transform(), None, None)
    eng.render(enc.finish(), stop=_lib.STOP_FLATTEN)
    edges, _ = eng.edges()
    return edges.reshape(-1, 2, 2)


S.Path.mask = path_mask
S.Path.fill = path_fill
S.Path.stroke = path_stroke


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
    return Layer(np.ascontiguousarray(image), tuple(offset), True, True)


def canvas_compose(mode, dst, src):
    """canvas_compose (svgrasterize.py:277-298) on two same-sized premultiplied images."""
    out = Layer.compose([_as_layer(dst), _as_layer(src)], mode, True)
    return out.image


def canvas_merge_at(base, overlay, offset, blend=None):
    """canvas_merge_at (svgrasterize.py:304-327) for the over blend (its only use in the reference, :3873):
    the overlay is blended onto `base` in place and the result is clipped to [0, 1]."""
    if blend not in (None, COMPOSE_OVER):
        raise NotImplementedError("canvas_merge_at is wired for the over blend")
    eng = default_engine()
    enc = Encoder(eng)
    b = enc.add_external(np.asarray(base, dtype=np.float32), (0, 0), True, True)
    o = enc.add_external(np.asarray(overlay, dtype=np.float32), offset, True, True)
    rows, cols = base.shape[:2]
    node = enc._node(_lib.N_MERGE_AT, 0, 0, rows, cols, children=[enc._compose([b, o], COMPOSE_OVER, True)])
    out = _read(eng, enc, node)
    base[...] = out.image
    return base


def canvas_merge_union(layers, full=True, blend=None):
    """canvas_merge_union (svgrasterize.py:330-379): layers = [(image, offset)] -> (image, offset)."""
    mode = COMPOSE_OVER if blend is None else blend
    if not full and mode != COMPOSE_OVER:
        raise ValueError("the sub-rectangle fast path is only defined for the over blend")
    out = Layer.compose([_as_layer(im, off) for im, off in layers], mode, True)
    return out.image, out.offset


def canvas_merge_intersect(layers, blend=None):
    """canvas_merge_intersect (svgrasterize.py:382-416) -> (image, offset) or None."""
    mode = COMPOSE_IN if blend is None else blend
    if mode != COMPOSE_IN:
        raise NotImplementedError("intersection merge is wired for the `in` blend (its only use in the reference)")
    out = Layer.compose([_as_layer(im, off) for im, off in layers], COMPOSE_IN, True)
    return None if out is None else (out.image, out.offset)


def pooling(mat, ksize, stride=None, method="max", pad=False):
    """pooling (svgrasterize.py:419-468) for the case the reference uses: stride 1, no padding."""
    if method not in ("max", "min"):
        raise ValueError(f"invalid poll method: {method}")
    if (stride not in (None, (1, 1))) or pad:
        raise NotImplementedError("only stride (1, 1) without padding is on the hot path (Layer.morphology)")
    mat = np.asarray(mat, dtype=np.float32)
    squeeze = mat.ndim == 2
    img = mat[..., None] if squeeze else mat
    if img.shape[2] not in (1, 4):
        raise ValueError("pooling expects 1 or 4 channels")
    if img.shape[2] == 1:
        img = np.repeat(img, 4, axis=2)
    out = Layer(np.ascontiguousarray(img), (0, 0), True, True).morphology(ksize[0], ksize[1], method).image
    out = out[..., :1] if mat.ndim == 2 or mat.shape[-1] == 1 else out
    return out[..., 0] if squeeze else out


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


def canvas_to_png(canvas, output=None, threads=1, level=9):
    """canvas_to_png (svgrasterize.py:249-274): straight-alpha sRGB float image -> PNG bytes.  The
    float -> uint8 quantisation (:263) is the in-scope part; deflate stays on the host like the reference.

    With the defaults the bytes are the reference's (filter 0 rows, one zlib stream at level 9).  `threads`
    > 1 compresses 1 MiB blocks of the same filtered rows on that many host cores (level 9 on
    one core takes 6.7 s for a 4096 x 4096 canvas): the file decodes to the same pixels but is not
    byte-identical and a few percent larger."""
    canvas = np.asarray(canvas)
    if canvas.dtype != np.uint8:
        canvas = np.round(np.clip(canvas, 0, 1) * 255.0).astype(np.uint8)
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
    if output is not None:
        output.write(png)
        return output
    return png
