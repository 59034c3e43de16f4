#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  For every fixture
scene it stores the scene itself (svgrasterize_b200.sceneio format) and what
the reference computed for it:

  canvas_u8      final (h, w, 4) uint8 array handed to the PNG encoder
                 (svgrasterize.py:3870-3881, :263)
  root_*         the Layer returned by Scene.render (float32 copy, offset, flags)
  leaf_* / edges / masks     per Path.mask call, in call order: bbox, the
                 flattened edge list (rows sorted, float64 exact) and the mask
                 (float32)                                (svgrasterize.py:922-993)
  stroke_*       per Path.stroke call: the outline path arrays (:1105-1180)

Usage: python tools/make_golden.py [--only NAME ...]
"""
import argparse
import glob
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

import svgrasterize as R  # noqa: E402  (the reference)
import svgrasterize_b200 as B  # noqa: E402
from svgrasterize_b200 import sceneio, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def to_reference(scene):
    """Rebuild a svgrasterize_b200 scene with the reference's own classes."""
    memo = {}

    def tr(t):
        return None if t is None else R.Transform(t.m.copy())

    def path(p):
        if id(p) not in memo:
            memo[id(p)] = R.Path([[(tag, args) for tag, args in sub] for sub in p.subpaths])
        return memo[id(p)]

    def paint(p):
        kind = B.paint_kind(p)
        if kind == "linear":
            return R.GradLinear(p.p0, p.p1, p.stops, tr(p.transform), p.spread, p.bbox_units, p.linear_rgb)
        if kind == "radial":
            return R.GradRadial(p.center, p.radius, p.fcenter, p.fradius, p.stops, tr(p.transform), p.spread,
                                p.bbox_units, p.linear_rgb)
        if kind == "pattern":
            return R.Pattern(node(p.scene), p.scene_bbox_units, p.scene_view_box, p.x, p.y, p.width, p.height,
                             tr(p.transform), p.bbox_units)
        return p

    def node(s):
        tag, args = s
        if tag == B.RENDER_FILL:
            return R.Scene(tag, (path(args[0]), paint(args[1]), args[2]))
        if tag == B.RENDER_STROKE:
            return R.Scene(tag, (path(args[0]), paint(args[1]), *args[2:]))
        if tag == B.RENDER_GROUP:
            return R.Scene(tag, tuple(node(c) for c in args))
        if tag == B.RENDER_OPACITY:
            return R.Scene(tag, (node(args[0]), args[1]))
        if tag in (B.RENDER_CLIP, B.RENDER_MASK):
            return R.Scene(tag, (node(args[0]), node(args[1]), args[2]))
        if tag == B.RENDER_TRANSFORM:
            return R.Scene(tag, (node(args[0]), tr(args[1])))
        if tag == B.RENDER_FILTER:
            return R.Scene(tag, (node(args[0]), R.Filter(dict(args[1].names), list(args[1].filters))))
        raise ValueError(tag)

    return node(scene)


class Taps:
    """Records every Path.mask / Path.stroke call of the reference."""

    def __init__(self):
        self.leaves, self.strokes, self._edges = [], [], None
        self._mask, self._stroke, self._hull = R.Path.mask, R.Path.stroke, R.ConvexHull.__init__

    def __enter__(self):
        taps = self

        def hull_init(self, points):
            if isinstance(points, np.ndarray) and points.ndim == 3:
                taps._edges = points
            taps._hull(self, points)

        def mask(self, transform, fill_rule=None, viewport=None):
            taps._edges = None
            res = taps._mask(self, transform, fill_rule, viewport)
            taps.leaves.append(None if res is None else (res[0].bbox, taps._edges, res[0].image[..., 0]))
            return res

        def stroke(self, width, linecap=None, linejoin=None):
            out = taps._stroke(self, width, linecap, linejoin)
            taps.strokes.append(sceneio.path_arrays(out))
            return out

        R.ConvexHull.__init__, R.Path.mask, R.Path.stroke = hull_init, mask, stroke
        return self

    def __exit__(self, *exc):
        R.ConvexHull.__init__, R.Path.mask, R.Path.stroke = self._hull, self._mask, self._stroke

    def arrays(self):
        bbox = np.full((len(self.leaves), 4), -1, dtype=np.int64)
        edge_off, mask_off, edges, masks = [0], [0], [], []
        for i, leaf in enumerate(self.leaves):
            if leaf is not None:
                bbox[i] = leaf[0]
                e = leaf[1].reshape(-1, 4)
                edges.append(e[np.lexsort(e.T[::-1])])
                masks.append(leaf[2].astype(np.float32).reshape(-1))
            edge_off.append(edge_off[-1] + (0 if leaf is None else len(edges[-1])))
            mask_off.append(mask_off[-1] + (0 if leaf is None else len(masks[-1])))
        out = {
            "leaf_bbox": bbox,
            "leaf_edge_off": np.asarray(edge_off, dtype=np.int64),
            "leaf_mask_off": np.asarray(mask_off, dtype=np.int64),
            "edges": np.concatenate(edges) if edges else np.zeros((0, 4)),
            "masks": np.concatenate(masks) if masks else np.zeros(0, np.float32),
        }
        if self.strokes:
            so, sso = [0], [0]
            for t, d, s in self.strokes:
                so.append(so[-1] + len(t))
                sso.append(sso[-1] + len(s))
            out.update(
                stroke_off=np.asarray(so, dtype=np.int64), stroke_sub_idx=np.asarray(sso, dtype=np.int64),
                stroke_tag=np.concatenate([t for t, _, _ in self.strokes]),
                stroke_data=np.concatenate([d for _, d, _ in self.strokes]),
                stroke_sub_off=np.concatenate([s for _, _, s in self.strokes]))
        return out


def reference_canvas(scene, size, linear_rgb):
    """main() :3854-3881 up to the uint8 array (no background)."""
    w, h = size
    tr = R.Transform().matrix(0, 1, 0, 1, 0, 0)
    res = scene.render(tr, viewport=[0, 0, int(h), int(w)], linear_rgb=linear_rgb)
    if res is None:
        return None, None
    root = res[0]
    out = root.convert(pre_alpha=True, linear_rgb=linear_rgb)
    base = np.zeros((int(h), int(w), 4))
    image = R.canvas_merge_at(base, out.image, out.offset)
    layer = R.Layer(image, (0, 0), True, linear_rgb).convert(pre_alpha=False, linear_rgb=False)
    return np.round(layer.image * 255.0).astype(np.uint8), root


def emit(name, ref_scene, size, linear_rgb=False, stages=False, root=False, golden_dir=None, sample_masks=0):
    """sample_masks > 0: besides the boxes keep the masks of that many evenly spaced non-empty leaves (full-size
    fixtures, where all masks together would be hundreds of megabytes)."""
    with Taps() as taps:
        canvas, root_layer = reference_canvas(ref_scene, size, linear_rgb)
    blob = sceneio.dump_scene(ref_scene)
    extra = {"size": np.asarray(size, dtype=np.float64), "linear_rgb": np.asarray(linear_rgb),
             "canvas_u8": canvas}
    if root and root_layer is not None:
        extra.update(root_image=root_layer.image.astype(np.float32), root_offset=np.asarray(root_layer.offset),
                     root_flags=np.asarray([root_layer.pre_alpha, root_layer.linear_rgb]))
    if stages:
        extra.update(taps.arrays())
    elif sample_masks:
        live = [i for i, leaf in enumerate(taps.leaves) if leaf is not None]
        pick = sorted({live[int(k)] for k in np.linspace(0, len(live) - 1, sample_masks)})
        bbox = np.full((len(taps.leaves), 4), -1, dtype=np.int64)
        for i in live:
            bbox[i] = taps.leaves[i][0]
        masks = [taps.leaves[i][2].astype(np.float32).reshape(-1) for i in pick]
        extra.update(leaf_bbox=bbox, sample_leaf=np.asarray(pick, dtype=np.int64),
                     sample_mask_off=np.concatenate([[0], np.cumsum([len(m) for m in masks])]).astype(np.int64),
                     sample_masks=np.concatenate(masks))
    else:
        extra.update(leaf_bbox=taps.arrays()["leaf_bbox"])
    path = os.path.join(golden_dir or GOLDEN, f"{name}.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **blob, **extra)
    print(f"{name:40s} {canvas.shape}  leaves {len(taps.leaves):5d}  strokes {len(taps.strokes):3d}  "
          f"{os.path.getsize(path) / 1024:8.1f} KiB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    opts = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    fonts = R.FontsDB()
    fonts.register_file("/root/reference/fonts.svgz")
    jobs = []

    def svg(name, file, width, **kw):
        def run():
            scene, _ids, size = R.svg_scene_from_filepath(file, width=width, fonts=fonts)
            emit(name, scene, size, **kw)
        jobs.append((name, run))

    def synth_scene(name, scene, size, **kw):
        jobs.append((name, lambda: emit(name, to_reference(scene), size, **kw)))

    demo = "/root/reference/demo"
    svg("demo_prompt", f"{demo}/prompt.svg", None, stages=True, root=True)
    svg("demo_icons_w512", f"{demo}/icons.svg", 512, stages=True)
    svg("demo_icons_native", f"{demo}/icons.svg", None)
    svg("demo_material_w1024", f"{demo}/material-design.svg", 1024)
    for f in sorted(glob.glob(f"{demo}/icons/*.svg")):
        svg("icon_" + os.path.basename(f)[:-4].replace("-", "_"), f, 256)
    for seed in range(6):
        synth_scene(f"synth_icon_{seed}", synth.icon_scene(seed), synth.icon_size(), stages=seed < 2, root=True)
    synth_scene("synth_filter_stack_192", synth.filter_stack_scene(192), (192, 192), root=True)
    for name, (scene, size) in synth.feature_scenes().items():
        synth_scene("feat_" + name, scene, size, linear_rgb=name in synth.FEATURES_LINEAR_RGB, stages=True, root=True)

    # BASELINE.json's configurations at their stated sizes (kept apart: the parametrised stage tests iterate over
    # tests/golden, these are compared once each by tests/test_gpu_fullsize.py).  Only made when asked for by name.
    big = os.path.join(ROOT, "tests", "golden_big")
    svg("demo_material_w4096", f"{demo}/material-design.svg", 4096, golden_dir=big, sample_masks=24)
    synth_scene("synth_filter_stack_2048", synth.filter_stack_scene(2048), (2048, 2048), golden_dir=big)
    big_names = {"demo_material_w4096", "synth_filter_stack_2048"}

    for name, run in jobs:
        if opts.only and name not in opts.only:
            continue
        if not opts.only and name in big_names:
            continue
        run()


if __name__ == "__main__":
    main()
