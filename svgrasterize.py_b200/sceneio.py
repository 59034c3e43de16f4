"""Scene (de)serialisation: a Scene DAG <-> flat numpy arrays + a JSON node table.

Used (a) to carry scenes parsed by the reference's SVG front-end into test
fixtures (``tests/golden/*.npz``, produced by ``tools/make_golden.py`` in the
build container where ``/root/reference`` exists) and (b) as the wire format
between the parse workers and the render process of a batch job.

The writer is duck-typed, so it accepts the reference's own ``Scene`` /
``Path`` / ``GradLinear`` / ``GradRadial`` / ``Pattern`` / ``Filter`` objects
(svgrasterize.py:598, :896, :1544, :1566, :1698, :1750) as well as this
package's.  float64 values survive bit-exactly: geometry is stored as float64
arrays, everything else through ``repr``-round-tripping JSON floats.
"""
from __future__ import annotations

import json

import numpy as np

from . import scene as S

SEG_WIDTH = 8  # doubles per segment row: up to 4 points, or the 7 arc parameters


def path_arrays(path):
    """Flatten ``path.subpaths`` into (seg_tag u8[S], seg_data f64[S,8], sub_off i32[NS+1]).
    Empty sub-paths are dropped (the reference skips them: svgrasterize.py:933, :1117)."""
    enc = getattr(path, "_enc", None)
    if enc is not None:
        return enc
    tags, rows, sub_off = [], [], [0]
    for sub in path.subpaths:
        if not sub:
            continue
        for tag, args in sub:
            row = np.zeros(SEG_WIDTH)
            if tag == S.PATH_ARC:
                center, rx, ry, phi, eta, eta_delta = args
                row[0:2] = np.asarray(center, dtype=np.float64)
                row[2:7] = (rx, ry, phi, eta, eta_delta)
            elif tag in (S.PATH_LINE, S.PATH_CLOSED, S.PATH_UNCLOSED):
                row[0:4] = np.asarray(args, dtype=np.float64).reshape(4)
            elif tag == S.PATH_QUAD:
                row[0:6] = np.asarray(args, dtype=np.float64).reshape(6)
            elif tag == S.PATH_CUBIC:
                row[0:8] = np.asarray(args, dtype=np.float64).reshape(8)
            else:
                raise ValueError(f"unsupported path type: `{tag}`")
            tags.append(tag)
            rows.append(row)
        sub_off.append(len(tags))
    enc = (
        np.asarray(tags, dtype=np.uint8),
        np.asarray(rows, dtype=np.float64).reshape(len(rows), SEG_WIDTH),
        np.asarray(sub_off, dtype=np.int32),
    )
    try:
        path._enc = enc
    except AttributeError:  # a reference Path has __slots__ without _enc
        pass
    return enc


def subpaths_from_arrays(seg_tag, seg_data, sub_off):
    """The nested ``[[(tag, points), ...], ...]`` form of flat segment arrays."""
    npts = {S.PATH_LINE: 2, S.PATH_CLOSED: 2, S.PATH_UNCLOSED: 2, S.PATH_QUAD: 3, S.PATH_CUBIC: 4}
    subpaths = []
    for a, b in zip(sub_off[:-1], sub_off[1:]):
        sub = []
        for i in range(a, b):
            tag = int(seg_tag[i])
            row = seg_data[i]
            if tag == S.PATH_ARC:
                sub.append((tag, (row[0:2].copy(), *map(float, row[2:7]))))
            else:
                sub.append((tag, row[: 2 * npts[tag]].reshape(-1, 2).copy()))
        subpaths.append(sub)
    return subpaths


def path_from_arrays(seg_tag, seg_data, sub_off) -> S.Path:
    return S.Path.from_arrays(
        np.ascontiguousarray(seg_tag, dtype=np.uint8),
        np.ascontiguousarray(seg_data, dtype=np.float64).reshape(-1, SEG_WIDTH),
        np.ascontiguousarray(sub_off, dtype=np.int32),
    )


def _f(v):
    return None if v is None else float(v)


def _vec(v):
    return None if v is None else [float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)]


def _tr(t):
    return None if t is None else _vec(t.m)


class _Writer:
    def __init__(self):
        self.nodes, self.node_ids = [], {}
        self.paints, self.paint_ids = [], {}
        self.path_ids = {}
        self.seg_tag, self.seg_data, self.sub_off, self.path_off = [], [], [0], [0]
        self._keep = []  # keep objects alive so id() stays unique

    def path(self, path) -> int:
        pid = self.path_ids.get(id(path))
        if pid is not None:
            return pid
        self._keep.append(path)
        tags, data, sub_off = path_arrays(path)
        base = sum(len(t) for t in self.seg_tag)
        self.seg_tag.append(tags)
        self.seg_data.append(data)
        self.sub_off.extend(int(o) + base for o in sub_off[1:])
        self.path_off.append(len(self.sub_off) - 1)
        pid = len(self.path_off) - 2
        self.path_ids[id(path)] = pid
        return pid

    def paint(self, paint):
        if paint is None:
            return None
        idx = self.paint_ids.get(id(paint))
        if idx is not None:
            return idx
        self._keep.append(paint)
        kind = S.paint_kind(paint)
        if kind == "solid":
            rec = {"k": kind, "c": _vec(paint)}
        elif kind in ("linear", "radial"):
            rec = {
                "k": kind,
                "stops": [[float(o), _vec(c)] for o, c in paint.stops],
                "tr": _tr(paint.transform),
                "spread": paint.spread,
                "bb": bool(paint.bbox_units),
                "lin": paint.linear_rgb,
            }
            if kind == "linear":
                rec.update(p0=_vec(paint.p0), p1=_vec(paint.p1))
            else:
                rec.update(c=_vec(paint.center), r=_f(paint.radius), fc=_vec(paint.fcenter), fr=_f(paint.fradius))
        elif kind == "pattern":
            rec = {
                "k": kind,
                "scene": self.node(paint.scene),
                "sbb": bool(paint.scene_bbox_units),
                "vb": _vec(paint.scene_view_box),
                "x": _f(paint.x), "y": _f(paint.y), "w": _f(paint.width), "h": _f(paint.height),
                "tr": _tr(paint.transform),
                "bb": bool(paint.bbox_units),
            }
        else:
            rec = {"k": "unknown", "repr": repr(paint)}
        self.paints.append(rec)
        self.paint_ids[id(paint)] = len(self.paints) - 1
        return len(self.paints) - 1

    def filter(self, flt):
        out = []
        for tag, attrs, inputs in flt.filters:
            enc = []
            for a in attrs:
                if isinstance(a, np.ndarray):
                    enc.append({"nd": _vec(a), "shape": list(a.shape)})
                elif isinstance(a, (tuple, list)):
                    enc.append({"tuple": [float(x) for x in a]})
                elif isinstance(a, (int, float, np.floating, np.integer)) and not isinstance(a, bool):
                    enc.append(float(a) if not isinstance(a, (int, np.integer)) else int(a))
                else:
                    enc.append(a)  # None / str
            out.append([int(tag), enc, [int(i) for i in inputs]])
        return out

    def node(self, scene) -> int:
        idx = self.node_ids.get(id(scene))
        if idx is not None:
            return idx
        self._keep.append(scene)
        tag, args = scene
        if tag == S.RENDER_FILL:
            path, paint, rule = args
            rec = {"t": tag, "path": self.path(path), "paint": self.paint(paint), "rule": rule}
        elif tag == S.RENDER_STROKE:
            path, paint, width, cap, join = args
            rec = {"t": tag, "path": self.path(path), "paint": self.paint(paint), "w": float(width),
                   "cap": cap, "join": join}
        elif tag == S.RENDER_GROUP:
            rec = {"t": tag, "c": [self.node(c) for c in args]}
        elif tag == S.RENDER_OPACITY:
            rec = {"t": tag, "s": self.node(args[0]), "o": float(args[1])}
        elif tag in (S.RENDER_CLIP, S.RENDER_MASK):
            rec = {"t": tag, "s": self.node(args[0]), "m": self.node(args[1]), "bb": bool(args[2])}
        elif tag == S.RENDER_TRANSFORM:
            rec = {"t": tag, "s": self.node(args[0]), "tr": _tr(args[1])}
        elif tag == S.RENDER_FILTER:
            rec = {"t": tag, "s": self.node(args[0]), "f": self.filter(args[1])}
        else:
            raise ValueError(f"unhandled scene type: {tag}")
        self.nodes.append(rec)
        self.node_ids[id(scene)] = len(self.nodes) - 1
        return len(self.nodes) - 1


def dump_scene(scene) -> dict:
    """Scene -> dict of numpy arrays (ready for ``np.savez_compressed``)."""
    w = _Writer()
    root = w.node(scene)
    meta = {"root": root, "nodes": w.nodes, "paints": w.paints}
    return {
        "meta": np.array(json.dumps(meta)),
        "seg_tag": np.concatenate(w.seg_tag) if w.seg_tag else np.zeros(0, np.uint8),
        "seg_data": np.concatenate(w.seg_data) if w.seg_data else np.zeros((0, SEG_WIDTH)),
        "sub_off": np.asarray(w.sub_off, dtype=np.int32),
        "path_off": np.asarray(w.path_off, dtype=np.int32),
    }


def load_scene(blob) -> S.Scene:
    """Inverse of :func:`dump_scene`; builds this package's scene objects."""
    meta = json.loads(str(blob["meta"]))
    seg_tag, seg_data = blob["seg_tag"], blob["seg_data"]
    sub_off, path_off = blob["sub_off"], blob["path_off"]
    paths, paints, nodes = {}, {}, {}

    def tr(v):
        return None if v is None else S.Transform(np.array(v, dtype=np.float64).reshape(3, 3))

    def arr(v):
        return None if v is None else np.array(v, dtype=np.float64)

    def path(pid):
        if pid not in paths:
            subs = sub_off[path_off[pid]: path_off[pid + 1] + 1]
            a, b = int(subs[0]), int(subs[-1])
            paths[pid] = path_from_arrays(seg_tag[a:b], seg_data[a:b], subs - a)
        return paths[pid]

    def paint(idx):
        if idx is None:
            return None
        if idx in paints:
            return paints[idx]
        rec = meta["paints"][idx]
        k = rec["k"]
        if k == "solid":
            p = arr(rec["c"])
        elif k == "linear":
            stops = [(o, arr(c)) for o, c in rec["stops"]]
            p = S.GradLinear(arr(rec["p0"]), arr(rec["p1"]), stops, tr(rec["tr"]), rec["spread"], rec["bb"], rec["lin"])
        elif k == "radial":
            stops = [(o, arr(c)) for o, c in rec["stops"]]
            p = S.GradRadial(arr(rec["c"]), rec["r"], arr(rec["fc"]), rec["fr"], stops, tr(rec["tr"]),
                             rec["spread"], rec["bb"], rec["lin"])
        elif k == "pattern":
            vb = None if rec["vb"] is None else tuple(rec["vb"])
            p = S.Pattern(node(rec["scene"]), rec["sbb"], vb, rec["x"], rec["y"], rec["w"], rec["h"],
                          tr(rec["tr"]), rec["bb"])
        else:
            p = rec  # unknown paint: the renderer warns and skips (svgrasterize.py:1099-1101)
        paints[idx] = p
        return p

    def filt(prims):
        flt = []
        for tag, attrs, inputs in prims:
            dec = []
            for a in attrs:
                if isinstance(a, dict) and "nd" in a:
                    dec.append(np.array(a["nd"], dtype=np.float64).reshape(a["shape"]))
                elif isinstance(a, dict) and "tuple" in a:
                    dec.append(tuple(a["tuple"]))
                else:
                    dec.append(a)
            flt.append((tag, tuple(dec), list(inputs)))
        return S.Filter({S.FE_SOURCE_ALPHA: 0, S.FE_SOURCE_GRAPHIC: 1}, flt)

    def node(idx):
        if idx in nodes:
            return nodes[idx]
        rec = meta["nodes"][idx]
        t = rec["t"]
        if t == S.RENDER_FILL:
            n = S.Scene(t, (path(rec["path"]), paint(rec["paint"]), rec["rule"]))
        elif t == S.RENDER_STROKE:
            n = S.Scene(t, (path(rec["path"]), paint(rec["paint"]), rec["w"], rec["cap"], rec["join"]))
        elif t == S.RENDER_GROUP:
            n = S.Scene(t, tuple(node(c) for c in rec["c"]))
        elif t == S.RENDER_OPACITY:
            n = S.Scene(t, (node(rec["s"]), rec["o"]))
        elif t in (S.RENDER_CLIP, S.RENDER_MASK):
            n = S.Scene(t, (node(rec["s"]), node(rec["m"]), rec["bb"]))
        elif t == S.RENDER_TRANSFORM:
            n = S.Scene(t, (node(rec["s"]), tr(rec["tr"])))
        elif t == S.RENDER_FILTER:
            n = S.Scene(t, (node(rec["s"]), filt(rec["f"])))
        else:
            raise ValueError(f"unhandled scene type: {t}")
        nodes[idx] = n
        return n

    return node(meta["root"])


def save_scene(file, scene, **extra) -> None:
    np.savez_compressed(file, **dump_scene(scene), **extra)
