"""Host-side scene model: the data types the hot path is called with.

These mirror the reference's call surface (names, fields, argument meaning) so
that the reference's own ``Scene`` tree, parser and CLI can hand their objects
to this core unchanged (duck typing: a reference ``Path`` only needs
``.subpaths``, a reference ``Scene`` is a ``(type, args)`` tuple, paints are
4-vectors or objects with the gradient fields below).

Reference: svgrasterize.py:576-647 (scene node tags and builders), :865-892 (path
segment tags), :509-570 (Transform), :1544-1575 (gradients), :1698-1710
(Pattern), :1718-1799 (filter program builder).

Nothing here touches the GPU; rendering entry points live in ``api.py``
and are attached to these classes there.
"""
from __future__ import annotations

import math
from typing import Any, NamedTuple

import numpy as np

# scene node tags (svgrasterize.py:576-583)
RENDER_FILL = 0
RENDER_STROKE = 1
RENDER_GROUP = 2
RENDER_OPACITY = 3
RENDER_CLIP = 4
RENDER_MASK = 5
RENDER_TRANSFORM = 6
RENDER_FILTER = 7

# path segment tags (svgrasterize.py:865-873)
PATH_LINE = 0
PATH_QUAD = 1
PATH_CUBIC = 2
PATH_ARC = 3
PATH_CLOSED = 4
PATH_UNCLOSED = 5

PATH_FILL_NONZERO = "nonzero"
PATH_FILL_EVENODD = "evenodd"
STROKE_JOIN_MITER = "miter"
STROKE_JOIN_ROUND = "round"
STROKE_JOIN_BEVEL = "bevel"
STROKE_CAP_BUTT = "butt"
STROKE_CAP_ROUND = "round"
STROKE_CAP_SQUARE = "square"

# compose modes (svgrasterize.py:47-52); "arithmetic" is a 4-tuple (k1..k4)
COMPOSE_OVER = 0
COMPOSE_OUT = 1
COMPOSE_IN = 2
COMPOSE_ATOP = 3
COMPOSE_XOR = 4
COMPOSE_PRE_ALPHA = {COMPOSE_OVER, COMPOSE_OUT, COMPOSE_IN, COMPOSE_ATOP, COMPOSE_XOR}

# filter primitive tags (svgrasterize.py:1718-1732)
FE_BLEND = 0
FE_COLOR_MATRIX = 1
FE_COMPOSITE = 3
FE_GAUSSIAN_BLUR = 8
FE_MERGE = 9
FE_MORPHOLOGY = 10
FE_OFFSET = 11
FE_SOURCE_ALPHA = "SourceAlpha"
FE_SOURCE_GRAPHIC = "SourceGraphic"


class Transform:
    """3x3 affine matrix with a lazily computed inverse (svgrasterize.py:509-570).

    The composition ``a @ b`` is a float64 numpy matmul, exactly as in the
    reference, because flattening is bit-exact only if the matrix handed to the
    device has the same bits.
    """

    __slots__ = ("m", "_inv")

    def __init__(self, matrix=None, matrix_inv=None):
        if matrix is None:
            matrix = np.identity(3)
            matrix_inv = matrix
        self.m = matrix
        self._inv = matrix_inv

    def __matmul__(self, other: "Transform") -> "Transform":
        return Transform(self.m @ other.m)

    @property
    def invert(self) -> "Transform":
        if self._inv is None:
            self._inv = np.linalg.inv(self.m)
        return Transform(self._inv, self.m)

    def __call__(self, points):
        if len(points) == 0:
            return points
        return points @ self.m[:2, :2].T + self.m[:2, 2]

    def _post(self, rows) -> "Transform":
        return Transform(self.m @ np.array(rows))

    def matrix(self, m00, m01, m02, m10, m11, m12) -> "Transform":
        return self._post([[m00, m01, m02], [m10, m11, m12], [0, 0, 1]])

    def translate(self, tx, ty) -> "Transform":
        return self._post([[1, 0, tx], [0, 1, ty], [0, 0, 1]])

    def scale(self, sx, sy=None) -> "Transform":
        if sy is None:
            sy = sx
        return self._post([[sx, 0, 0], [0, sy, 0], [0, 0, 1]])

    def rotate(self, angle) -> "Transform":
        c, s = math.cos(angle), math.sin(angle)
        return self._post([[c, -s, 0], [s, c, 0], [0, 0, 1]])

    def skew(self, ax, ay) -> "Transform":
        return self._post([[1, math.tan(ax), 0], [math.tan(ay), 1, 0], [0, 0, 1]])

    def no_translate(self) -> "Transform":
        m = self.m.copy()
        m[:2, 2] = 0
        return Transform(m)

    def coeffs(self) -> np.ndarray:
        """Row-major 2x3 part as 6 doubles: what the C-ABI takes per path."""
        return np.ascontiguousarray(self.m[:2, :], dtype=np.float64).reshape(6)

    def __repr__(self) -> str:
        return str(np.around(self.m, 4).tolist()[:2])


class Path:
    """List of sub-paths, each a list of ``(tag, points)`` segments
    (svgrasterize.py:896-913).  ``mask`` / ``fill`` / ``stroke`` are attached
    by ``api.py`` and run on the device.

    A path read by the native path-data reader (``Path.from_svg``) holds flat
    arrays (``_enc``: tags, 8 doubles per segment, sub-path offsets) and only
    builds the nested Python lists when somebody asks for ``subpaths``."""

    __slots__ = ("_subpaths", "_enc", "_flat")

    def __init__(self, subpaths):
        self._subpaths = subpaths
        self._enc = None
        self._flat = None

    @classmethod
    def from_arrays(cls, seg_tag, seg_data, sub_off) -> "Path":
        """A path over flat segment arrays (the layout of sceneio.path_arrays); nothing is copied or unpacked."""
        path = cls.__new__(cls)
        path._subpaths = None
        path._enc = (seg_tag, seg_data, sub_off)
        path._flat = None
        return path

    @property
    def subpaths(self):
        if self._subpaths is None:
            from .sceneio import subpaths_from_arrays

            self._subpaths = subpaths_from_arrays(*self._enc)
        return self._subpaths

    @subpaths.setter
    def subpaths(self, value):
        self._subpaths, self._enc, self._flat = value, None, None

    def __iter__(self):
        return iter(self.subpaths)

    def __bool__(self) -> bool:
        return bool(len(self._enc[2]) - 1) if self._subpaths is None else bool(self._subpaths)

    def is_empty(self) -> bool:
        return not bool(self)


class GradLinear(NamedTuple):
    p0: np.ndarray
    p1: np.ndarray
    stops: list
    transform: Any
    spread: str
    bbox_units: bool
    linear_rgb: Any


class GradRadial(NamedTuple):
    center: np.ndarray
    radius: float
    fcenter: Any
    fradius: Any
    stops: list
    transform: Any
    spread: str
    bbox_units: bool
    linear_rgb: Any


class Pattern(NamedTuple):
    scene: Any
    scene_bbox_units: bool
    scene_view_box: Any
    x: float
    y: float
    width: float
    height: float
    transform: Any
    bbox_units: bool

    def bbox(self):
        return (self.x, self.y, self.width, self.height)


def paint_kind(paint) -> str:
    """Classify a paint object by duck typing: works for this module's types
    and for the reference's (svgrasterize.py:1014, :1021, :1049)."""
    if paint is None:
        return "none"
    if isinstance(paint, np.ndarray):
        return "solid" if paint.shape == (4,) else "unknown"
    if hasattr(paint, "p0") and hasattr(paint, "p1") and hasattr(paint, "stops"):
        return "linear"
    if hasattr(paint, "center") and hasattr(paint, "radius") and hasattr(paint, "stops"):
        return "radial"
    if hasattr(paint, "scene") and hasattr(paint, "scene_view_box"):
        return "pattern"
    return "unknown"


class Filter(NamedTuple):
    """Filter program: ``filters[i] = (tag, attrs, input_slots)``; slot 0 is
    SourceAlpha, slot 1 SourceGraphic, slot i+2 the result of primitive i
    (svgrasterize.py:1750-1799)."""

    names: dict
    filters: list

    @classmethod
    def empty(cls) -> "Filter":
        return cls({FE_SOURCE_ALPHA: 0, FE_SOURCE_GRAPHIC: 1}, [])

    def add_filter(self, tag, attrs, inputs, result):
        prev = len(self.filters) + 1
        slots = []
        for name in inputs:
            slot = prev if name is None else self.names.get(name)
            if slot is None:
                import warnings

                warnings.warn(f"unknown filter result name: {name}")
                slot = prev
            slots.append(slot)
        names = dict(self.names)
        if result is not None:
            names[result] = len(self.filters) + 2
        return Filter(names, [*self.filters, (tag, attrs, slots)])

    def offset(self, dx, dy, input=None, result=None):
        return self.add_filter(FE_OFFSET, (dx, dy), [input], result)

    def merge(self, inputs, result=None):
        return self.add_filter(FE_MERGE, tuple(), inputs, result)

    def blur(self, std_x, std_y=None, input=None, result=None):
        return self.add_filter(FE_GAUSSIAN_BLUR, (std_x, std_y), [input], result)

    def blend(self, in1, in2, mode=None, result=None):
        return self.add_filter(FE_BLEND, (mode,), [in1, in2], result)

    def composite(self, in1, in2, mode=None, result=None):
        return self.add_filter(FE_COMPOSITE, (mode,), [in1, in2], result)

    def color_matrix(self, input, matrix, result=None):
        return self.add_filter(FE_COLOR_MATRIX, (matrix,), [input], result)

    def morphology(self, rx, ry, method, input, result=None):
        return self.add_filter(FE_MORPHOLOGY, (rx, ry, method), [input], result)


class Scene(tuple):
    """Tagged-tuple scene graph node ``(tag, args)`` (svgrasterize.py:598-647).
    ``render`` is attached by ``api.py`` and runs on the device."""

    __slots__ = ()

    def __new__(cls, tag, args):
        return tuple.__new__(cls, (tag, args))

    def __reduce__(self):  # tuple subclasses with a custom __new__ need this to cross process boundaries
        return (Scene, (self[0], self[1]))

    @classmethod
    def fill(cls, path, paint, fill_rule=None):
        return cls(RENDER_FILL, (path, paint, fill_rule))

    @classmethod
    def stroke(cls, path, paint, width, linecap=None, linejoin=None):
        return cls(RENDER_STROKE, (path, paint, width, linecap, linejoin))

    @classmethod
    def group(cls, children):
        children = tuple(children)
        if not children:
            raise ValueError("group have to contain at least one child")
        if len(children) == 1:
            return children[0]
        return cls(RENDER_GROUP, children)

    def opacity(self, opacity):
        if opacity > 0.999:
            return self
        return Scene(RENDER_OPACITY, (self, opacity))

    def clip(self, clip, bbox_units=False):
        return Scene(RENDER_CLIP, (self, clip, bbox_units))

    def mask(self, mask, bbox_units=False):
        return Scene(RENDER_MASK, (self, mask, bbox_units))

    def transform(self, transform):
        tag, args = self
        if tag == RENDER_TRANSFORM:
            target, inner = args
            return Scene(RENDER_TRANSFORM, (target, transform @ inner))
        return Scene(RENDER_TRANSFORM, (self, transform))

    def filter(self, flt):
        return Scene(RENDER_FILTER, (self, flt))
