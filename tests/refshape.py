"""A reference-shaped module for the drop-in tests: the names svgrasterize.py defines around its hot path
(Path, Scene, Transform, Filter, Layer, canvas_* ..., FLOAT) plus a `main_flow` that restates the render part
of the reference's CLI entry point (svgrasterize.py:3854-3881) and, like it, resolves `Layer` and
`canvas_merge_at` through the module's globals at call time.  The reference itself cannot travel to the GPU
box; where it exists (/root/reference) tests/test_host_logic.py installs the core on the real module too."""
import types

import numpy as np

FLOW = '''
def main_flow(scene, size, linear_rgb, bg, output, transform=None):
    """svgrasterize.py:3823, :3857-3881 with opts.* as arguments"""
    transform = Transform().matrix(0, 1, 0, 1, 0, 0) if transform is None else transform
    if size is not None:
        w, h = size
        result = scene.render(transform, viewport=[0, 0, int(h), int(w)], linear_rgb=linear_rgb)
    else:
        result = scene.render(transform, linear_rgb=linear_rgb)
    if result is None:
        return 1
    out, _convex_hull = result
    if size is not None:
        w, h = size
        out = out.convert(pre_alpha=True, linear_rgb=linear_rgb)
        base = np.zeros((int(h), int(w), 4), dtype=FLOAT)
        image = canvas_merge_at(base, out.image, out.offset)
        out = Layer(image, (0, 0), pre_alpha=True, linear_rgb=linear_rgb)
    if bg is not None:
        out = out.background(bg)
    out.write_png(output)
    return 0
'''


def make_module():
    """A fresh module whose hot-path entry points are unbound stubs: nothing renders until install()."""
    from svgrasterize_b200 import scene as S

    mod = types.ModuleType("svgrasterize_like")

    def stub(name):
        def fn(*_a, **_k):
            raise RuntimeError(f"{name}: the reference's numpy implementation is not part of this module")
        fn.__name__ = name
        return fn

    # fresh subclasses, so that binding methods on them does not touch the package's own classes
    mod.Path = type("Path", (S.Path,), {"mask": stub("Path.mask"), "fill": stub("Path.fill"), "stroke": stub("Path.stroke")})
    mod.Scene = type("Scene", (S.Scene,), {"render": stub("Scene.render"), "__slots__": ()})
    mod.Filter = type("Filter", (S.Filter,), {"__call__": stub("Filter.__call__")})
    mod.Transform = S.Transform
    mod.Layer = None
    mod.np, mod.FLOAT = np, np.float64
    for name in ("canvas_create", "canvas_to_png", "canvas_compose", "canvas_merge_at", "canvas_merge_union",
                 "canvas_merge_intersect", "pooling", "bezier3_flatten_batch", "line_signed_coverage", "grad_pixels",
                 "grad_spread", "grad_interpolate", "blur_kernel"):
        setattr(mod, name, stub(name))
    exec(FLOW, mod.__dict__)
    return mod


def rebuild(mod, scene):
    """The same scene tree with the module's own Scene / Path / Filter classes."""
    from svgrasterize_b200 import scene as S

    memo = {}

    def path(p):
        if id(p) not in memo:
            memo[id(p)] = mod.Path(p.subpaths)
        return memo[id(p)]

    def paint(p):
        if S.paint_kind(p) == "pattern":
            return p._replace(scene=node(p.scene))
        return p

    def node(s):
        tag, args = s
        if tag == S.RENDER_FILL:
            p, pt, rule = args
            return mod.Scene(tag, (path(p), paint(pt), rule))
        if tag == S.RENDER_STROKE:
            p, pt, *rest = args
            return mod.Scene(tag, (path(p), paint(pt), *rest))
        if tag == S.RENDER_GROUP:
            return mod.Scene(tag, tuple(node(c) for c in args))
        if tag in (S.RENDER_CLIP, S.RENDER_MASK):
            t, o, units = args
            return mod.Scene(tag, (node(t), node(o), units))
        if tag == S.RENDER_FILTER:
            t, flt = args
            return mod.Scene(tag, (node(t), mod.Filter(*flt)))
        t, *rest = args  # opacity, transform
        return mod.Scene(tag, (node(t), *rest))

    return node(scene)
