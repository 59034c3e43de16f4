"""Scene -> scene program: the host side of the hot path.

The reference renders eagerly, one numpy call per node (Scene.render,
svgrasterize.py:649-752).  Here the same tree walk only *records* what has to be
done -- flat geometry, a path table, a paint table and a post-order list of layer
ops -- and the device executes the whole batch in a handful of launches
(csrc/engine.cu).  Many scenes concatenate into one program
(:meth:`Program.concat`), which is how icon batches are rendered.

Everything that depends on pixels stays on the device.  The only device round
trip at encode time is the user-space bounding box needed by objectBoundingBox
paints / clips / masks (ConvexHull.bbox, svgrasterize.py:2002-2023): the encoder
asks the engine to flatten the target and reduce its end points
(:meth:`Encoder._cloud_bbox`).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib
from . import scene as S
from .color import paint_to_srgb
from .sceneio import path_arrays

CAPS = {None: 0, S.STROKE_CAP_BUTT: 0, S.STROKE_CAP_ROUND: 1, S.STROKE_CAP_SQUARE: 2}
JOINS = {None: 0, S.STROKE_JOIN_MITER: 0, S.STROKE_JOIN_ROUND: 1, S.STROKE_JOIN_BEVEL: 2}
SPREAD = {"pad": 0, "repeat": 1, "reflect": 2}


def fill_rule_code(fill_rule) -> int:
    if fill_rule is None or fill_rule == S.PATH_FILL_NONZERO:
        return 0
    if fill_rule == S.PATH_FILL_EVENODD:
        return 1
    raise ValueError(f"Invalid fill rule: {fill_rule}")  # svgrasterize.py:989


def m6(transform) -> np.ndarray:
    return np.ascontiguousarray(transform.m[:2, :], dtype=np.float64).reshape(6)


def expand_arcs(tags, data, sub_off):
    """Replace PATH_ARC rows by their cubic pieces (arc_to_bezier3, svgrasterize.py:2355, evaluated on
    the host with the C library's libm so the control points have the reference's bits)."""
    n_arc = int(np.count_nonzero(tags == S.PATH_ARC))
    if n_arc == 0:
        return tags, data, sub_off
    n = len(tags)
    cap = n + 64 * n_arc
    out_t = np.empty(cap, dtype=np.uint8)
    out_d = np.empty((cap, 8), dtype=np.float64)
    new_index = np.empty(n + 1, dtype=np.int64)
    tags = np.ascontiguousarray(tags, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.float64)
    m = _lib.lib().svgr_expand_arcs(tags.ctypes.data, data.ctypes.data, n, out_t.ctypes.data, out_d.ctypes.data, cap,
                                    new_index.ctypes.data)
    if m < 0:
        raise ValueError("arc sweeps more than 16 pi")
    return out_t[:m].copy(), out_d[:m].copy(), new_index[np.asarray(sub_off, dtype=np.int64)].astype(np.int32)


def device_path(path):
    """(tags, data, sub_off) of a path with arcs expanded; cached on the path object."""
    enc = getattr(path, "_flat", None)
    if enc is None:
        enc = expand_arcs(*path_arrays(path))
        try:
            path._flat = enc
        except AttributeError:
            pass
    return enc


def blur_kernel(transform, sigma):
    """Gaussian kernel of feGaussianBlur under `transform` (blur_kernel, svgrasterize.py:1903-1944):
    (rows, cols) float64 weights, or None when the blur is a no-op (< half a pixel both ways)."""
    sx, sy = sigma
    origin = transform(np.zeros(2))
    basis = transform(np.eye(2)) - origin
    scale_x, scale_y = np.linalg.norm(basis, axis=1)
    if scale_x * sx < 0.5 and scale_y * sy < 0.5:
        return None
    if scale_x * sx < 0.5:
        sx = 0.5 / scale_x
    elif scale_y * sy < 0.5:
        sy = 0.5 / scale_y
    reach = 2.5
    box = np.array([[-reach * sx, -reach * sy], [-reach * sx, reach * sy], [reach * sx, reach * sy],
                    [reach * sx, -reach * sy]])
    box = transform(box) - origin
    lo, hi = box.min(axis=0).astype(int), box.max(axis=0).astype(int)
    kw, kh = int(hi[0] - lo[0]), int(hi[1] - lo[1])
    kw += 1 - (kw & 1)
    kh += 1 - (kh & 1)
    rr, cc = np.indices((kw, kh)).astype(np.float64)
    centres = np.stack([rr, cc], axis=2) + [-kw / 2 + 0.5, -kh / 2 + 0.5]
    inv = transform.invert
    pts = inv(centres) - inv(np.zeros(2))
    w = np.exp(-np.square(pts) / (2 * np.square(np.array([sx, sy])))).prod(axis=-1)
    return w / w.sum()


def viewbox_transform(bbox, viewbox):
    """svg_viewbox_transform (svgrasterize.py:3116-3133), used by pattern paint (:1059)."""
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


class Program:
    """Flat host arrays of one batch, in the layout of include/svgr_b200.h."""

    ARRAYS = ("seg_tag", "seg_data", "seg_path", "paths", "strokes", "stroke_sub_off", "stroke_sub_job", "stroke_tag",
              "stroke_data", "stroke_seg_job", "paints", "stops", "nodes", "children", "kernels", "weights", "matrices",
              "offset_tr", "bbox_jobs")

    def __init__(self):
        self.seg_tag = np.zeros(0, np.uint8)
        self.seg_data = np.zeros((0, 8), np.float64)
        self.seg_path = np.zeros(0, np.uint32)
        self.paths = np.zeros(0, _lib.PATH_DT)
        self.strokes = np.zeros(0, _lib.STROKE_DT)
        self.stroke_sub_off = np.zeros(1, np.int32)
        self.stroke_sub_job = np.zeros(0, np.int32)
        self.stroke_tag = np.zeros(0, np.uint8)
        self.stroke_data = np.zeros((0, 8), np.float64)
        self.stroke_seg_job = np.zeros(0, np.int32)
        self.paints = np.zeros(0, _lib.PAINT_DT)
        self.stops = np.zeros(0, _lib.STOP_DT)
        self.n_focal = 0
        self.nodes = np.zeros(0, _lib.NODE_DT)
        self.children = np.zeros(0, np.int32)
        self.kernels = np.zeros(0, _lib.KERNEL_DT)
        self.weights = np.zeros(0, np.float32)
        self.matrices = np.zeros((0, 20), np.float32)
        self.offset_tr = np.zeros((0, 12), np.float64)
        self.bbox_jobs = np.zeros(0, _lib.BBOX_JOB_DT)  # objectBoundingBox gradients completed on the device
        self.externals = []  # (image float32, r0, c0, pre_alpha, linear_rgb)
        self.canvas_bytes = 0
        self.canvases = []  # (node, byte offset, rows, cols)
        self.roots = []  # node index of every scene added with add_scene / add_root
        self.flatness = 0.0  # 0: Path.mask's literal 0.1 (svgrasterize.py:955)

    # -- C view ------------------------------------------------------------------------------
    def to_c(self):
        """-> (_lib.Program, keepalive list)"""
        p = _lib.Program()
        keep = []

        def put(name, arr, count_name=None, count=None):
            arr = np.ascontiguousarray(arr)
            keep.append(arr)
            setattr(p, name, _lib.ptr(arr))
            if count_name:
                setattr(p, count_name, len(arr) if count is None else count)

        put("seg_tag", self.seg_tag, "n_seg")
        put("seg_data", self.seg_data)
        put("seg_path", self.seg_path)
        put("paths", self.paths, "n_path")
        put("strokes", self.strokes, "n_stroke")
        put("stroke_sub_off", self.stroke_sub_off, "n_stroke_sub", len(self.stroke_sub_off) - 1)
        put("stroke_sub_job", self.stroke_sub_job)
        put("stroke_tag", self.stroke_tag, "n_stroke_seg")
        put("stroke_data", self.stroke_data)
        put("stroke_seg_job", self.stroke_seg_job)
        put("paints", self.paints, "n_paint")
        put("stops", self.stops, "n_stop")
        p.n_focal = int(self.n_focal)
        put("nodes", self.nodes, "n_node")
        put("children", self.children, "n_child")
        put("kernels", self.kernels, "n_kernel")
        put("weights", self.weights, "n_weight")
        put("matrices", self.matrices, "n_matrix")
        put("offset_tr", self.offset_tr, "n_offset_tr")
        put("bbox_jobs", self.bbox_jobs, "n_bbox_job")
        ext = (_lib.External * max(len(self.externals), 1))()
        for i, (img, r0, c0, pre, lin) in enumerate(self.externals):
            img = np.ascontiguousarray(img, dtype=np.float32)
            keep.append(img)
            ext[i] = _lib.External(img.ctypes.data, int(r0), int(c0), img.shape[0], img.shape[1], img.shape[2],
                                   int(bool(pre)), int(bool(lin)), 0)
        keep.append(ext)
        p.externals = C.cast(ext, C.c_void_p)
        p.n_external = len(self.externals)
        p.canvas_bytes = int(self.canvas_bytes)
        p.flatness = float(self.flatness)
        return p, keep

    def h2d_bytes(self) -> int:
        return int(sum(getattr(self, n).nbytes for n in self.ARRAYS))

    # -- batching ----------------------------------------------------------------------------
    @staticmethod
    def concat(programs):
        """One program rendering all the given programs (index spaces are shifted, nothing is shared)."""
        out = Program()
        if not programs:
            return out
        base = dict(path=0, stroke=0, sub=0, sseg=0, paint=0, stop=0, focal=0, node=0, child=0, kernel=0, weight=0,
                    matrix=0, offtr=0, ext=0, canvas=0)
        parts = {n: [] for n in Program.ARRAYS}
        sub_off_parts = [np.zeros(1, np.int32)]
        for p in programs:
            parts["seg_tag"].append(p.seg_tag)
            parts["seg_data"].append(p.seg_data)
            parts["seg_path"].append(p.seg_path + np.uint32(base["path"]))
            parts["paths"].append(p.paths)
            st = p.strokes.copy()
            st["sub_begin"] += base["sub"]
            st["sub_end"] += base["sub"]
            st["path"] += base["path"]
            parts["strokes"].append(st)
            sub_off_parts.append(p.stroke_sub_off[1:] + np.int32(base["sseg"]))
            parts["stroke_sub_job"].append(p.stroke_sub_job + np.int32(base["stroke"]))
            parts["stroke_tag"].append(p.stroke_tag)
            parts["stroke_data"].append(p.stroke_data)
            parts["stroke_seg_job"].append(p.stroke_seg_job + np.int32(base["stroke"]))
            pa = p.paints.copy()
            grad = (pa["kind"] != _lib.PAINT_SOLID) & (pa["kind"] != _lib.PAINT_PATTERN)
            pa["stop_off"][grad] += base["stop"]
            pa["flag"][pa["kind"] == _lib.PAINT_RADIAL_FOCAL] += base["focal"]
            pa["pat_node"] = np.where(pa["pat_node"] >= 0, pa["pat_node"] + base["node"], pa["pat_node"])
            parts["paints"].append(pa)
            parts["stops"].append(p.stops)
            nd = p.nodes.copy()
            tag = nd["tag"]
            nd["child_off"] += base["child"]
            leaf = tag == _lib.N_LEAF
            nd["a"][leaf] += base["path"]
            nd["b"][leaf & (nd["b"] >= 0)] += base["paint"]
            nd["d"][leaf & (nd["d"] >= 0)] += base["node"]
            nd["a"][tag == _lib.N_BLUR] += base["kernel"]
            nd["a"][tag == _lib.N_CMATRIX] += base["matrix"]
            nd["a"][tag == _lib.N_OFFSET] += base["offtr"]
            nd["a"][tag == _lib.N_EXTERNAL] += base["ext"]
            nd["f"][tag == _lib.N_CANVAS, 0] += base["canvas"]
            parts["nodes"].append(nd)
            parts["children"].append(p.children + np.int32(base["node"]))
            kn = p.kernels.copy()
            kn["weight_off"] += base["weight"]
            parts["kernels"].append(kn)
            parts["weights"].append(p.weights)
            parts["matrices"].append(p.matrices)
            parts["offset_tr"].append(p.offset_tr)
            bj = p.bbox_jobs.copy()
            bj["paint"] += base["paint"]
            bj["path"] += base["path"]
            parts["bbox_jobs"].append(bj)
            out.externals.extend(p.externals)
            out.canvases.extend((n + base["node"], off + base["canvas"], r, c) for n, off, r, c in p.canvases)
            out.roots.extend(n + base["node"] for n in p.roots)
            base["path"] += len(p.paths)
            base["stroke"] += len(p.strokes)
            base["sub"] += len(p.stroke_sub_off) - 1
            base["sseg"] += len(p.stroke_tag)
            base["paint"] += len(p.paints)
            base["stop"] += len(p.stops)
            base["focal"] += p.n_focal
            base["node"] += len(p.nodes)
            base["child"] += len(p.children)
            base["kernel"] += len(p.kernels)
            base["weight"] += len(p.weights)
            base["matrix"] += len(p.matrices)
            base["offtr"] += len(p.offset_tr)
            base["ext"] += len(p.externals)
            base["canvas"] += p.canvas_bytes
        for n in Program.ARRAYS:
            if n == "stroke_sub_off":
                continue
            setattr(out, n, np.concatenate(parts[n]))
        out.stroke_sub_off = np.concatenate(sub_off_parts).astype(np.int32)
        out.n_focal = base["focal"]
        out.canvas_bytes = base["canvas"]
        return out


class Encoder:
    """Walks Scene trees the way Scene.render does and records a Program.

    `engine` is only needed for scenes that use objectBoundingBox units."""

    def __init__(self, engine=None):
        self.engine = engine
        self.seg_tag, self.seg_data, self.seg_path = [], [], []
        self.paths = []  # (m6, viewport or None, fill_rule code)
        self.strokes = []  # (half_width, sub_begin, sub_end, cap, join, path)
        self.s_sub_off, self.s_sub_job = [0], []
        self.s_tag, self.s_data, self.s_seg_job = [], [], []
        self.n_sseg = 0
        self.paints, self.stops, self.n_focal = [], [], 0
        self.bbox_jobs = []  # (paint, path, has_grad_tr, inv6, grad_inv6, geom6)
        self.nodes, self.children = [], []
        self.kernels, self.weights, self.n_weight = [], [], 0
        self.matrices, self.offset_tr = [], []
        self.externals = []
        self.canvases, self.canvas_bytes = [], 0
        self.roots = []
        self.cloud = {}  # node -> list of (node, path) pairs contributing end points when the node is non-empty
        # row-band renders clip masks to the band but must resolve objectBoundingBox units against the
        # whole canvas (a path outside the band still belongs to its group's bounding box)
        self.cloud_viewport = None

    def _cv(self, viewport):
        return viewport if self.cloud_viewport is None else self.cloud_viewport

    # -- tables ------------------------------------------------------------------------------
    def _node(self, tag, a=0, b=0, c=0, d=0, children=(), flags=0, f=(0.0, 0.0, 0.0, 0.0)) -> int:
        off = len(self.children)
        self.children.extend(int(k) for k in children)
        self.nodes.append((tag, int(a), int(b), int(c), int(d), off, len(children), int(flags), tuple(float(v) for v in f)))
        return len(self.nodes) - 1

    def _empty(self) -> int:
        return self._node(_lib.N_EMPTY)

    def _is_empty(self, node) -> bool:
        return self.nodes[node][0] == _lib.N_EMPTY

    def _viewport(self, viewport):
        return None if viewport is None else tuple(int(v) for v in viewport)

    def add_fill_path(self, path, transform, fill_rule, viewport) -> int:
        rule = fill_rule_code(fill_rule)
        tags, data, _sub = device_path(path)
        pid = len(self.paths)
        self.paths.append((m6(transform), self._viewport(viewport), rule, self._viewport(self.cloud_viewport)))
        if len(tags):
            self.seg_tag.append(tags)
            self.seg_data.append(data)
            self.seg_path.append(np.full(len(tags), pid, dtype=np.uint32))
        return pid

    def add_stroke_path(self, path, transform, width, linecap, linejoin, viewport) -> int:
        if linecap not in CAPS:
            raise ValueError(f"unkown line cap type: `{linecap}`")  # svgrasterize.py:1492
        tags, data, sub_off = device_path(path)
        pid = len(self.paths)
        self.paths.append((m6(transform), self._viewport(viewport), 0, self._viewport(self.cloud_viewport)))
        job = len(self.strokes)
        sub_begin = len(self.s_sub_job)
        nsub = len(sub_off) - 1
        self.s_sub_off.extend(int(o) + self.n_sseg for o in sub_off[1:])
        self.s_sub_job.extend([job] * nsub)
        if len(tags):
            self.s_tag.append(tags)
            self.s_data.append(data)
            self.s_seg_job.append(np.full(len(tags), job, dtype=np.int32))
            self.n_sseg += len(tags)
        self.strokes.append((float(width) / 2, sub_begin, sub_begin + nsub, CAPS[linecap], JOINS.get(linejoin, 3), pid))
        return pid

    def _stops(self, paint, lin) -> tuple:
        off = len(self.stops)
        if not paint.stops:
            raise ValueError("gradient without stops")
        offs = [float(o) for o, _ in paint.stops]
        for k, (o, c) in enumerate(paint.stops):
            col = np.asarray(c, dtype=np.float64) if lin else paint_to_srgb(c)
            with np.errstate(divide="ignore"):
                inv = float(np.float64(1.0) / np.float64(offs[k + 1] - offs[k])) if k + 1 < len(offs) else 0.0
            self.stops.append((float(o), tuple(float(v) for v in col), inv))
        return off, len(paint.stops)

    def _paint_record(self, **kw) -> int:
        rec = dict(kind=0, spread=0, stop_off=0, stop_cnt=0, has_m2=0, flag=0, pat_r0=0, pat_c0=0, pat_rows=0,
                   pat_cols=0, pat_node=-1, color=(0, 0, 0, 0), m1=np.zeros(6), m2=np.zeros(6), g=np.zeros(8))
        rec.update(kw)
        self.paints.append(rec)
        return len(self.paints) - 1

    # -- objectBoundingBox support ---------------------------------------------------------------
    def _cloud_bbox(self, build, transform):
        """ConvexHull.bbox(transform) of what `build(encoder)` renders (svgrasterize.py:2002-2007), or
        None when it renders nothing.  Runs flatten + the host plan on the device engine."""
        if self.engine is None:
            raise RuntimeError("objectBoundingBox units need an Engine (Encoder(engine=...))")
        sub = Encoder(self.engine)
        sub.cloud_viewport = self.cloud_viewport
        root = build(sub)
        if sub._is_empty(root):
            return None
        prog = sub.finish()
        self.engine.render(prog, stop=_lib.STOP_PLAN)
        alive = {}

        def live(n):
            if n not in alive:
                alive[n] = self.engine.node_info(n)[0] != 0
            return alive[n]

        if not live(root):
            return None
        paths = sorted({p for n, p in sub.cloud.get(root, []) if live(n)})
        if not paths:
            return None
        mm = self.engine.cloud_bounds([paths], [m6(transform.invert)])[0]
        return [mm[0], mm[1], mm[2] - mm[0], mm[3] - mm[1]]

    @staticmethod
    def _bbox_transform(bbox, transform):
        """ConvexHull.bbox_transform (svgrasterize.py:2009-2023)."""
        x, y, w, h = bbox
        if w <= 0 and h <= 0:
            return transform
        return transform.translate(x, y).scale(w, h)

    # -- leaves ------------------------------------------------------------------------------
    def _leaf(self, pid, paint, transform, mask_only, linear_rgb, self_cloud):
        """Path.fill / Path.mask on path `pid` (svgrasterize.py:995-1103).  self_cloud() -> bbox of this
        leaf's own end points in user space (lazy, device)."""
        if mask_only:
            node = self._node(_lib.N_LEAF, pid, -1, 1, -1)
            self.cloud[node] = [(node, pid)]
            return node
        kind = S.paint_kind(paint)
        if kind == "none":
            return self._empty()
        if kind == "solid":
            col = np.asarray(paint, dtype=np.float64) if linear_rgb else paint_to_srgb(paint)
            pidx = self._paint_record(kind=_lib.PAINT_SOLID, color=tuple(float(v) for v in col))
            node = self._node(_lib.N_LEAF, pid, pidx, int(bool(linear_rgb)), -1)
        elif kind in ("linear", "radial"):
            if paint.spread not in SPREAD:
                raise ValueError(f"invalid spread method: {paint.spread}")  # svgrasterize.py:1668
            lin = linear_rgb if paint.linear_rgb is None else bool(paint.linear_rgb)
            stop_off, stop_cnt = self._stops(paint, lin)
            rec = dict(spread=SPREAD[paint.spread], stop_off=stop_off, stop_cnt=stop_cnt)
            # the gradient's own geometry (what does not depend on the transform)
            g, m1, geom = np.zeros(8), np.zeros(6), np.zeros(6)
            if kind == "linear":
                p0, vec = np.asarray(paint.p0, dtype=np.float64), np.asarray(paint.p1, dtype=np.float64) - paint.p0
                geom[0:2], geom[2:4], geom[4] = p0, vec, float(np.dot(vec, vec))
                rec.update(kind=_lib.PAINT_LINEAR)
            elif paint.fcenter is None and paint.fradius is None:
                geom[0:2], geom[2] = np.asarray(paint.center, dtype=np.float64), float(paint.radius)
                rec.update(kind=_lib.PAINT_RADIAL)
            else:
                c, r = np.asarray(paint.center, dtype=np.float64), float(paint.radius)
                f = c if paint.fcenter is None else np.asarray(paint.fcenter, dtype=np.float64)
                fr = float(paint.fradius or 0)
                cd, rd = c - f, r - fr
                a = float((cd ** 2).sum() - rd ** 2)
                geom[0:2] = f
                with np.errstate(divide="ignore", invalid="ignore"):
                    g[:] = (cd[0], cd[1], fr * rd, a, fr * fr, np.float64(fr) / np.float64(fr - r), float(fr != r),
                            np.float64(1.0) / np.float64(a))
                rec.update(kind=_lib.PAINT_RADIAL_FOCAL, flag=self.n_focal)
                self.n_focal += 1
            # pixel centre -> gradient space: transform.invert, then the inverse gradientTransform
            # (svgrasterize.py:1022-1031, :1558, :1602); composed here so that the device evaluates one
            # affine expression per pixel
            to_user = transform.invert.m
            if paint.bbox_units:
                # hull.bbox_transform(transform) (:1023-1026) needs the flattened leaf: the map is completed on the
                # device between flattening and compositing (svgr_bbox_job), no round trip here
                rec.update(g=g, m1=m1)
                pidx = self._paint_record(**rec)
                grad_inv = np.zeros(6) if paint.transform is None else m6(paint.transform.invert)
                self.bbox_jobs.append((pidx, pid, int(paint.transform is not None), to_user[:2, :].reshape(6).copy(),
                                       grad_inv, geom))
                node = self._node(_lib.N_LEAF, pid, pidx, int(lin), -1)
                self.cloud[node] = [(node, pid)]
                return node
            if paint.transform is not None:
                to_user = paint.transform.invert.m @ to_user
            A, T = to_user[:2, :2], to_user[:2, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                if kind == "linear":
                    vec, vv = geom[2:4], geom[4]
                    g[0:2] = (vec @ A) / vv
                    g[2] = np.dot(T - geom[0:2], vec) / vv
                elif rec["kind"] == _lib.PAINT_RADIAL:
                    r = geom[2]
                    m1[0:2], m1[3:5] = A[0] / r, A[1] / r
                    m1[2], m1[5] = (T - geom[0:2]) / r
                else:
                    m1[0:2], m1[3:5] = A[0], A[1]
                    m1[2], m1[5] = T - geom[0:2]
            rec.update(g=g, m1=m1)
            pidx = self._paint_record(**rec)
            node = self._node(_lib.N_LEAF, pid, pidx, int(lin), -1)
        elif kind == "pattern":
            node = self._pattern_leaf(pid, paint, transform, linear_rgb, self_cloud)
            if node is None:
                return self._empty()
        else:
            warnings.warn(f"fill method is not implemented: {paint}")  # svgrasterize.py:1099-1101
            return self._empty()
        self.cloud[node] = [(node, pid)]
        return node

    def _pattern_leaf(self, pid, paint, transform, linear_rgb, self_cloud):
        """Pattern branch of Path.fill (svgrasterize.py:1049-1097)."""
        bbox = None
        if paint.bbox_units or (paint.scene_bbox_units and not paint.scene_view_box):
            bbox = self_cloud()
            if bbox is None:
                return None
        pat_tr = transform.no_translate()
        if paint.scene_view_box:
            if paint.bbox_units:
                px, py, pw, ph = paint.bbox()
                _hx, _hy, hw, hh = bbox
                box = (px * hw, py * hh, pw * hw, ph * hh)
            else:
                box = paint.bbox()
            pat_tr = pat_tr @ viewbox_transform(box, paint.scene_view_box)
        elif paint.scene_bbox_units:
            pat_tr = self._bbox_transform(bbox, pat_tr)
        pat_tr = pat_tr @ paint.transform
        tile = self.encode(paint.scene, pat_tr, False, None, linear_rgb)
        if self._is_empty(tile):
            return None
        rep = transform
        if paint.bbox_units:
            rep = self._bbox_transform(bbox, rep)
        rep = (rep @ paint.transform).no_translate()
        corners = rep(np.array([[0, 0], [paint.width, 0], [0, paint.height], [paint.width, paint.height]],
                               dtype=np.float64))
        hi = corners.max(axis=0).astype(int)
        lo = corners.min(axis=0).astype(int)
        rows, cols = int(hi[0] - lo[0] + 1), int(hi[1] - lo[1] + 1)
        pat = self._node(_lib.N_MERGE_AT, int(lo[0]), int(lo[1]), rows, cols, children=[tile])
        g = np.zeros(8)
        g[0:4] = (paint.x, paint.y, paint.width, paint.height)
        pidx = self._paint_record(kind=_lib.PAINT_PATTERN, m1=m6(rep.invert), m2=m6(rep), g=g, pat_r0=int(lo[0]),
                                  pat_c0=int(lo[1]), pat_rows=rows, pat_cols=cols, pat_node=pat)
        return self._node(_lib.N_LEAF, pid, pidx, int(bool(linear_rgb)), pat)

    # -- filters -------------------------------------------------------------------------------
    def _filter(self, flt, transform, source) -> int:
        """Filter.__call__ (svgrasterize.py:1801-1831) lowered to nodes."""
        stack = [self._node(_lib.N_SRC_ALPHA, children=[source]),
                 self._node(_lib.N_CONVERT, 0, 1, children=[source])]
        for tag, attrs, inputs in flt.filters:
            args = [stack[i] for i in inputs]
            if tag == S.FE_OFFSET:
                dx, dy = attrs
                self.offset_tr.append(np.concatenate([m6(transform), m6(transform.invert)]))
                out = self._node(_lib.N_OFFSET, len(self.offset_tr) - 1, children=[args[0]], f=(dx, dy, 0, 0))
            elif tag == S.FE_MERGE:
                out = self._compose(args, S.COMPOSE_OVER, True)
            elif tag == S.FE_BLEND:
                warnings.warn("feBlend is not properly supported")  # svgrasterize.py:1877
                out = self._compose([args[1], args[0]], S.COMPOSE_OVER, True)
            elif tag == S.FE_COMPOSITE:
                out = self._compose([args[1], args[0]], attrs[0], True)
            elif tag == S.FE_GAUSSIAN_BLUR:
                sx, sy = attrs
                kernel = blur_kernel(transform, (sx, sx if sy is None else sy))
                out = args[0] if kernel is None else self._blur(args[0], kernel)
            elif tag == S.FE_COLOR_MATRIX:
                (matrix,) = attrs
                if not isinstance(matrix, np.ndarray) or matrix.shape != (4, 5):
                    warnings.warn(f"invalid color matrix: {matrix}")
                    out = args[0]
                else:
                    self.matrices.append(np.asarray(matrix, dtype=np.float32).reshape(20))
                    out = self._node(_lib.N_CMATRIX, len(self.matrices) - 1, children=[args[0]])
            elif tag == S.FE_MORPHOLOGY:
                rx, ry, method = attrs
                u = transform(np.array([[rx, 0], [0, ry]], dtype=np.float64)) - transform(np.zeros((2, 2)))
                k0 = int(np.linalg.norm(u[0]) * 2)
                k1 = int(np.linalg.norm(u[1]) * 2)
                if k0 < 1 or k1 < 1:
                    out = args[0]
                else:
                    if method not in ("max", "min"):
                        raise ValueError(f"invalid poll method: {method}")  # svgrasterize.py:466
                    out = self._node(_lib.N_MORPH, k0, k1, int(method == "max"), children=[args[0]])
            else:
                raise ValueError(f"unsupported filter type: {tag}")  # svgrasterize.py:1828
            stack.append(out)
        return stack[-1]

    def _compose(self, nodes, mode, lin, intersect=False, raw=False) -> int:
        """Layer.compose (svgrasterize.py:178-207).  intersect: blend on the intersection of the boxes whatever
        the mode (canvas_merge_intersect); raw: plain arrays, no Layer.convert (canvas_compose)."""
        flags = int(bool(lin)) | (4 if intersect else 0) | (8 if raw else 0)
        if isinstance(mode, tuple) and len(mode) == 4:
            return self._node(_lib.N_COMPOSE, 5, children=nodes, flags=flags, f=mode)
        if isinstance(mode, bool) or mode not in (0, 1, 2, 3, 4):
            raise ValueError(f"invalid compose mode: {mode}")  # svgrasterize.py:298
        return self._node(_lib.N_COMPOSE, int(mode), children=nodes, flags=flags)

    def _blur(self, source, kernel) -> int:
        rows, cols = kernel.shape
        a, b = kernel.sum(axis=1), kernel.sum(axis=0)
        separable = float(np.abs(np.outer(a, b) - kernel).max()) < 1e-12
        if separable:
            w = np.concatenate([a, b])
        else:
            w = kernel.reshape(-1)
        self.kernels.append((rows, cols, int(separable), self.n_weight))
        self.weights.append(np.asarray(w, dtype=np.float32))
        self.n_weight += len(w)
        return self._node(_lib.N_BLUR, len(self.kernels) - 1, children=[source])

    # -- the scene walk (Scene.render, svgrasterize.py:649-752) ----------------------------------------
    def encode(self, scene, transform, mask_only=False, viewport=None, linear_rgb=False) -> int:
        tag, args = scene
        lin = int(bool(linear_rgb))
        if tag == S.RENDER_FILL:
            path, paint, rule = args
            if not mask_only and paint is None:
                return self._empty()
            pid = self.add_fill_path(path, transform, rule, viewport)
            own = lambda: self._cloud_bbox(  # noqa: E731
                lambda e: e._leaf(e.add_fill_path(path, transform, rule, self._cv(viewport)), None, transform, True,
                                  linear_rgb, None), transform)
            return self._leaf(pid, paint, transform, mask_only, linear_rgb, own)
        if tag == S.RENDER_STROKE:
            path, paint, width, cap, join = args
            if not mask_only and paint is None:
                if cap not in CAPS:
                    raise ValueError(f"unkown line cap type: `{cap}`")
                return self._empty()
            pid = self.add_stroke_path(path, transform, width, cap, join, viewport)
            own = lambda: self._cloud_bbox(  # noqa: E731
                lambda e: e._leaf(e.add_stroke_path(path, transform, width, cap, join, self._cv(viewport)), None,
                                  transform, True, linear_rgb, None), transform)
            return self._leaf(pid, paint, transform, mask_only, linear_rgb, own)
        if tag == S.RENDER_GROUP:
            kids = [self.encode(c, transform, mask_only, viewport, linear_rgb) for c in args]
            kids = [k for k in kids if not self._is_empty(k)]
            if not kids:
                return self._empty()
            if len(kids) == 1:
                return kids[0]
            node = self._node(_lib.N_GROUP, children=kids, flags=lin)
            self.cloud[node] = [pair for k in kids for pair in self.cloud.get(k, [])]
            return node
        if tag == S.RENDER_OPACITY:
            target, value = args
            t = self.encode(target, transform, mask_only, viewport, linear_rgb)
            if self._is_empty(t):
                return t
            node = self._node(_lib.N_OPACITY, children=[t], flags=lin, f=(value, 0, 0, 0))
            self.cloud[node] = self.cloud.get(t, [])
            return node
        if tag == S.RENDER_TRANSFORM:
            target, tr = args
            return self.encode(target, transform @ tr, mask_only, viewport, linear_rgb)
        if tag in (S.RENDER_CLIP, S.RENDER_MASK):
            target, other, bbox_units = args
            t = self.encode(target, transform, mask_only, viewport, linear_rgb)
            if self._is_empty(t):
                return t
            if bbox_units:
                bbox = self._cloud_bbox(
                    lambda e: e.encode(target, transform, mask_only, self._cv(viewport), linear_rgb), transform)
                if bbox is None:
                    return self._empty()
                transform = self._bbox_transform(bbox, transform)
            if tag == S.RENDER_CLIP:
                stencil = self.encode(other, transform, True, viewport, linear_rgb)
                if self._is_empty(stencil):
                    return stencil
            else:
                sub = self.encode(other, transform, mask_only, viewport, linear_rgb)
                if self._is_empty(sub):
                    return sub
                stencil = self._node(_lib.N_LUMA, children=[sub], flags=lin)
            node = self._node(_lib.N_IN, children=[stencil, t], flags=lin)
            self.cloud[node] = self.cloud.get(t, [])
            return node
        if tag == S.RENDER_FILTER:
            target, flt = args
            t = self.encode(target, transform, mask_only, viewport, linear_rgb)
            if self._is_empty(t):
                return t
            node = self._filter(flt, transform, t)
            if node != t:
                self.cloud[node] = self.cloud.get(t, [])
            return node
        raise ValueError(f"unhandled scene type: {tag}")  # svgrasterize.py:752

    # -- entry points --------------------------------------------------------------------------------
    def add_root(self, scene, transform, mask_only=False, viewport=None, linear_rgb=False, materialize=True) -> int:
        """Scene.render(transform, mask_only, viewport, linear_rgb): returns the root node, whose layer can
        be read back after a render (Engine.node)."""
        node = self.encode(scene, transform, mask_only, viewport, linear_rgb)
        if materialize and not self._is_empty(node):
            rec = list(self.nodes[node])
            if not rec[7] & 2:
                # wrap instead of flagging in place: the node may be shared with a group
                node2 = self._node(_lib.N_CONVERT, -1, -1, children=[node], flags=2)
                self.cloud[node2] = self.cloud.get(node, [])
                node = node2
        self.roots.append(node)
        return node

    def add_scene(self, scene, size, linear_rgb=False, transform=None) -> int:
        """main() of the reference (svgrasterize.py:3854-3881): render with the canvas viewport, merge onto a
        transparent canvas, convert to straight-alpha sRGB, quantise.  size = (width, height).  Returns the
        index into Program.canvases."""
        w, h = size
        tr = canvas_transform() if transform is None else transform
        root = self.encode(scene, tr, False, [0, 0, int(h), int(w)], linear_rgb)
        kids = [] if self._is_empty(root) else [root]
        node = self._node(_lib.N_CANVAS, int(h), int(w), children=kids, flags=int(bool(linear_rgb)),
                          f=(self.canvas_bytes, 0, 0, 0))
        self.canvases.append((node, self.canvas_bytes, int(h), int(w)))
        self.canvas_bytes += 4 * int(h) * int(w)
        self.roots.append(root)
        return len(self.canvases) - 1

    def add_scene_band(self, scene, size, rows, halo=0, linear_rgb=False, transform=None) -> int:
        """One row band [rows[0], rows[1]) of add_scene's canvas: masks are clipped to the band
        plus `halo` rows (what filters inside the scene reach across), the canvas node writes only the band."""
        w, h = int(size[0]), int(size[1])
        a, b = max(0, int(rows[0])), min(h, int(rows[1]))
        tr = canvas_transform() if transform is None else transform
        lo, hi = max(0, a - int(halo)), min(h, b + int(halo))
        self.cloud_viewport = [0, 0, h, w]
        root = self.encode(scene, tr, False, [lo, 0, hi - lo, w], linear_rgb)
        self.cloud_viewport = None
        kids = [] if self._is_empty(root) else [root]
        node = self._node(_lib.N_CANVAS, b - a, w, a, 0, children=kids, flags=int(bool(linear_rgb)),
                          f=(self.canvas_bytes, 0, 0, 0))
        self.canvases.append((node, self.canvas_bytes, b - a, w))
        self.canvas_bytes += 4 * (b - a) * w
        self.roots.append(root)
        return len(self.canvases) - 1

    def filter_reach(self) -> int:
        """Upper bound of how many rows the filters recorded so far can carry a pixel: blur kernel rows,
        morphology windows and feOffset shifts, summed over all filter nodes (conservative)."""
        reach = 0
        for tag, a, b, c, d, off, cnt, flags, f in self.nodes:
            if tag == _lib.N_BLUR:
                reach += self.kernels[a][0]
            elif tag == _lib.N_MORPH:
                reach += a
            elif tag == _lib.N_OFFSET:
                tr = self.offset_tr[a]
                reach += int(abs(tr[0] * f[0]) + abs(tr[1] * f[1])) + 2
        return reach

    def add_external(self, image, offset, pre_alpha, linear_rgb) -> int:
        img = np.ascontiguousarray(image, dtype=np.float32)
        if img.ndim != 3 or img.shape[2] not in (1, 4):
            raise ValueError("layer image must be (rows, cols, 1|4)")
        self.externals.append((img, int(offset[0]), int(offset[1]), bool(pre_alpha), bool(linear_rgb)))
        return self._node(_lib.N_EXTERNAL, len(self.externals) - 1)

    def finish(self) -> Program:
        p = Program()
        if self.seg_tag:
            p.seg_tag = np.concatenate(self.seg_tag)
            p.seg_data = np.ascontiguousarray(np.concatenate(self.seg_data).reshape(-1, 8))
            p.seg_path = np.concatenate(self.seg_path)
        # records are built column-wise: one numpy call per field instead of one per record
        n = len(self.paths)
        p.paths = np.zeros(n, _lib.PATH_DT)
        if n:
            p.paths["m"] = np.asarray([rec[0] for rec in self.paths], dtype=np.float64).reshape(n, 6)
            p.paths["has_viewport"] = [rec[1] is not None for rec in self.paths]
            p.paths["viewport"] = [rec[1] if rec[1] is not None else (0, 0, 0, 0) for rec in self.paths]
            p.paths["fill_rule"] = [rec[2] for rec in self.paths]
            # row-band renders: the whole canvas' viewport, against which filter origins are computed
            p.paths["has_full"] = [rec[3] is not None for rec in self.paths]
            p.paths["full_viewport"] = [rec[3] if rec[3] is not None else (0, 0, 0, 0) for rec in self.paths]
        n = len(self.strokes)
        p.strokes = np.zeros(n, _lib.STROKE_DT)
        if n:
            cols = list(zip(*self.strokes))
            for name, col in zip(("half_width", "sub_begin", "sub_end", "cap", "join", "path"), cols):
                p.strokes[name] = col
        p.stroke_sub_off = np.asarray(self.s_sub_off, dtype=np.int32)
        p.stroke_sub_job = np.asarray(self.s_sub_job, dtype=np.int32)
        if self.s_tag:
            p.stroke_tag = np.concatenate(self.s_tag)
            p.stroke_data = np.ascontiguousarray(np.concatenate(self.s_data).reshape(-1, 8))
            p.stroke_seg_job = np.concatenate(self.s_seg_job)
        n = len(self.paints)
        p.paints = np.zeros(n, _lib.PAINT_DT)
        if n:
            for name in ("kind", "spread", "stop_off", "stop_cnt", "has_m2", "flag", "pat_r0", "pat_c0", "pat_rows",
                         "pat_cols", "pat_node", "color", "m1", "m2", "g"):
                p.paints[name] = [rec[name] for rec in self.paints]
        n = len(self.stops)
        p.stops = np.zeros(n, _lib.STOP_DT)
        if n:
            p.stops["offset"] = [o for o, _c, _i in self.stops]
            p.stops["color"] = [c for _o, c, _i in self.stops]
            p.stops["inv_span"] = [i for _o, _c, i in self.stops]
        p.n_focal = self.n_focal
        n = len(self.nodes)
        p.nodes = np.zeros(n, _lib.NODE_DT)
        if n:
            cols = list(zip(*self.nodes))
            for name, col in zip(("tag", "a", "b", "c", "d", "child_off", "child_cnt", "flags", "f"), cols):
                p.nodes[name] = col
        p.children = np.asarray(self.children, dtype=np.int32)
        p.kernels = np.zeros(len(self.kernels), _lib.KERNEL_DT)
        for i, k in enumerate(self.kernels):
            p.kernels[i] = k
        if self.weights:
            p.weights = np.concatenate(self.weights).astype(np.float32)
        if self.matrices:
            p.matrices = np.stack(self.matrices).astype(np.float32)
        if self.offset_tr:
            p.offset_tr = np.stack(self.offset_tr).astype(np.float64)
        p.bbox_jobs = np.zeros(len(self.bbox_jobs), _lib.BBOX_JOB_DT)
        for i, (paint, path, has_tr, inv, grad_inv, geom) in enumerate(self.bbox_jobs):
            p.bbox_jobs[i] = (paint, path, has_tr, 0, inv, grad_inv, geom)
        p.externals = list(self.externals)
        p.canvas_bytes = self.canvas_bytes
        p.canvases = list(self.canvases)
        p.roots = list(self.roots)
        return p


def canvas_transform():
    """The x/y swap every render starts from (svgrasterize.py:246, :3823)."""
    return S.Transform().matrix(0, 1, 0, 1, 0, 0)


def encode_scene(scene, size, linear_rgb=False, engine=None) -> Program:
    enc = Encoder(engine)
    enc.add_scene(scene, size, linear_rgb)
    return enc.finish()


def _encode_job(job):
    scene, size, linear_rgb = job
    return encode_scene(scene, size, linear_rgb)


def encode_batch(jobs, processes: int = 0, engine=None, native: bool = True):
    """Encode many (scene, size, linear_rgb) jobs into one program.

    native (default): the scene walk runs in the library (native.encode_batch: csrc/_flatten.c reads the Scene
    objects, csrc/encode_flat.cpp records the program, ~50 us per icon on one core against ~1.5 ms for this module);
    scenes it does not cover (objectBoundingBox units, patterns, filters) are encoded by Encoder and spliced in.
    native=False: every scene through Encoder, optionally spread over `processes` worker processes (scenes that use
    objectBoundingBox units need the device at encode time and must be encoded in the rendering process)."""
    jobs = list(jobs)
    if native:
        from . import native as N

        return N.encode_batch(jobs, engine=engine)
    if processes and processes > 1 and len(jobs) > 1:
        import multiprocessing as mp

        with mp.get_context("spawn").Pool(processes) as pool:
            progs = pool.map(_encode_job, jobs, chunksize=max(1, len(jobs) // (processes * 4)))
    else:
        progs = [encode_scene(scene, size, lin, engine=engine) for scene, size, lin in jobs]
    return Program.concat(progs)
