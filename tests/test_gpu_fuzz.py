"""Randomised parity: seeded random scenes (paths of every segment kind, transforms, every paint, opacity,
clip, luminance mask, strokes with every cap / join, filter chains) rendered by the CUDA core and by the CPU
oracle; final RGBA8 within +-1 LSB (BASELINE.json's bar), root layers within 2e-5."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_path(rng, S, synth, n_sub=None):
    b = synth.PathBuilder()
    for _ in range(n_sub or int(rng.integers(1, 4))):
        x, y = rng.uniform(4, 60, 2)
        b.move_to(x, y)
        for _ in range(int(rng.integers(2, 6))):
            kind = rng.integers(0, 4)
            px, py = rng.uniform(2, 62, 2)
            if kind == 0:
                b.line_to(px, py)
            elif kind == 1:
                b.quad_to(*rng.uniform(0, 64, 2), px, py)
            elif kind == 2:
                b.cubic_to(*rng.uniform(-8, 72, 4), px, py)
            else:
                b.arc(*rng.uniform(10, 54, 2), rng.uniform(2, 20), rng.uniform(2, 20), rng.uniform(0, 3.0),
                      rng.uniform(0, 6.0), rng.uniform(-5.0, 5.0))
        if rng.random() < 0.7:
            b.close()
    if b.cur:
        b._end(S.PATH_UNCLOSED)
    return S.Path(b.subpaths)


def _rand_paint(rng, S, synth):
    k = rng.integers(0, 5)
    if k <= 1:
        return synth.color(*rng.uniform(0, 1, 3), rng.uniform(0.2, 1.0))
    n = int(rng.integers(2, 5))
    offs = np.sort(rng.uniform(0, 1, n))
    offs[0] = 0.0 if rng.random() < 0.5 else offs[0]
    stops = [(float(o), synth.color(*rng.uniform(0, 1, 3), rng.uniform(0.3, 1.0))) for o in offs]
    spread = ("pad", "repeat", "reflect")[int(rng.integers(0, 3))]
    tr = None if rng.random() < 0.5 else S.Transform().rotate(rng.uniform(-1, 1)).scale(*rng.uniform(0.5, 1.5, 2))
    lin = (None, None, True, False)[int(rng.integers(0, 4))]
    if k == 2:
        return S.GradLinear(rng.uniform(0, 64, 2), rng.uniform(0, 64, 2), stops, tr, spread, False, lin)
    c = rng.uniform(16, 48, 2)
    r = float(rng.uniform(8, 30))
    if k == 3:
        return S.GradRadial(c, r, None, None, stops, tr, spread, False, lin)
    f = c + rng.uniform(-1.2, 1.2, 2) * r  # sometimes outside the end circle
    return S.GradRadial(c, r, f, float(rng.uniform(0, 4)) if rng.random() < 0.5 else None, stops, tr, spread, False, lin)


def _rand_leaf(rng, S, synth):
    path = _rand_path(rng, S, synth)
    paint = _rand_paint(rng, S, synth)
    if rng.random() < 0.3:
        cap = (None, "butt", "round", "square")[int(rng.integers(0, 4))]
        join = (None, "miter", "round", "bevel")[int(rng.integers(0, 4))]
        return S.Scene.stroke(path, paint, float(rng.uniform(0.5, 6)), cap, join)
    return S.Scene.fill(path, paint, (None, "nonzero", "evenodd")[int(rng.integers(0, 3))])


def _rand_filter(rng, S, synth):
    f = S.Filter.empty()
    for _ in range(int(rng.integers(1, 4))):
        k = rng.integers(0, 6)
        if k == 0:
            f = f.blur(float(rng.uniform(0.3, 2.5)), None if rng.random() < 0.5 else float(rng.uniform(0.3, 2.5)))
        elif k == 1:
            f = f.offset(float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5)))
        elif k == 2:
            f = f.morphology(float(rng.uniform(0.3, 1.5)), float(rng.uniform(0.3, 1.5)), ("max", "min")[int(rng.integers(0, 2))], None)
        elif k == 3:
            f = f.color_matrix(None, synth.saturate_matrix(float(rng.uniform(0, 2))))
        elif k == 4:
            mode = (0, 1, 2, 3, 4, tuple(rng.uniform(0, 0.7, 4)))[int(rng.integers(0, 6))]
            f = f.composite(S.FE_SOURCE_GRAPHIC, None, mode)
        else:
            f = f.merge([None, S.FE_SOURCE_GRAPHIC])
    return f


def _rand_scene(rng, S, synth, depth=0):
    kids = []
    for _ in range(int(rng.integers(1, 4))):
        r = rng.random()
        node = _rand_scene(rng, S, synth, depth + 1) if (r < 0.25 and depth < 2) else _rand_leaf(rng, S, synth)
        r = rng.random()
        if r < 0.2:
            node = node.opacity(float(rng.uniform(0.2, 0.95)))
        elif r < 0.35:
            node = node.clip(S.Scene.fill(_rand_path(rng, S, synth, 1), np.ones(4)))
        elif r < 0.45:
            node = node.mask(S.Scene.group([_rand_leaf(rng, S, synth), _rand_leaf(rng, S, synth)]))
        elif r < 0.55:
            node = node.filter(_rand_filter(rng, S, synth))
        elif r < 0.7:
            node = node.transform(S.Transform().translate(*rng.uniform(-6, 6, 2)).rotate(rng.uniform(-0.4, 0.4)))
        kids.append(node)
    return S.Scene.group(kids)


@pytest.mark.parametrize("seed", range(48))
def test_random_scene_matches_oracle(seed):
    import warnings

    import svgrasterize_b200 as B
    from oracle import render as O
    from svgrasterize_b200 import scene as S, synth

    rng = np.random.default_rng(7000 + seed)
    size = (int(rng.integers(48, 140)), int(rng.integers(48, 140)))
    scale = float(rng.uniform(0.8, 2.2))
    scene = _rand_scene(rng, S, synth).transform(S.Transform().scale(scale).rotate(float(rng.uniform(-0.2, 0.2))))
    linear_rgb = bool(seed % 3 == 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref = O.render_canvas(scene, size, linear_rgb)
        except TypeError:  # the reference's own failure mode on a degenerate stroke (bezier3_offset, :2157)
            with pytest.raises(TypeError):
                B.render_canvas(scene, size, linear_rgb)
            return
        got = B.render_canvas(scene, size, linear_rgb)
    if ref is None:
        assert not got.any()
        return
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1, (seed, int(diff.max()), int((diff > 1).sum()))  # 400 seeds measured: worst 1 LSB


@pytest.mark.parametrize("seed", range(24))
def test_random_geometry_is_bit_exact(seed):
    """Edge lists (per-path multisets) and stroke outlines of random paths under random transforms are
    bit-identical to the oracle's (which is pinned bit-exactly to the reference)."""
    from oracle import geometry
    from oracle.stroke import stroke_path
    from svgrasterize_b200 import _lib, encode, scene as S, synth
    from svgrasterize_b200.engine import default_engine

    rng = np.random.default_rng(9000 + seed)
    eng = default_engine()
    paths = [_rand_path(rng, S, synth) for _ in range(6)]
    trs = [S.Transform().matrix(0, 1, 0, 1, 0, 0) @ S.Transform().translate(*rng.uniform(-20, 20, 2))
           .rotate(rng.uniform(-3, 3)).scale(*rng.uniform(0.3, 9.0, 2)).skew(rng.uniform(-0.3, 0.3), 0.0) for _ in paths]
    enc = encode.Encoder(eng)
    for p, t in zip(paths, trs):
        enc.add_fill_path(p, t, None, None)
    strokes = [(paths[i], float(rng.uniform(0.2, 8)), (None, "round", "square")[i % 3], (None, "round", "bevel")[(i // 2) % 3])
               for i in range(3)]
    for p, w, cap, join in strokes:
        enc.add_stroke_path(p, S.Transform(), w, cap, join, None)
    prog = enc.finish()
    try:
        eng.render(prog, stop=_lib.STOP_FLATTEN)
    except TypeError:
        for p, w, cap, join in strokes:  # at least one of them must be the reference's degenerate case
            try:
                stroke_path(p, w, cap, join)
            except TypeError:
                return
        raise
    edges, edge_path = eng.edges()
    for i, (p, t) in enumerate(zip(paths, trs)):
        mine = edges[edge_path == i]
        ref = geometry.path_edges(p, t).reshape(-1, 4)
        mine = mine[np.lexsort(mine.T[::-1])]
        ref = ref[np.lexsort(ref.T[::-1])]
        assert mine.shape == ref.shape and np.array_equal(mine.view(np.uint64), ref.view(np.uint64)), (seed, i)
    tag, data, path, sub = eng.outline()
    for k, (p, w, cap, join) in enumerate(strokes):
        sel = path == len(paths) + k
        rt, rd, _rs = encode.path_arrays(stroke_path(p, w, cap, join))
        assert np.array_equal(tag[sel], rt) and np.array_equal(data[sel].view(np.uint64), rd.view(np.uint64)), (seed, k)


@pytest.mark.parametrize("closed", [False, True])
def test_long_polyline_stroke_is_bit_exact(closed):
    """A sub-path of 1 500 segments (lines, quads, repeated points) is assembled by several warps, each a range of
    256 segments: the outline must still be the oracle's, segment for segment, in order."""
    from oracle.stroke import stroke_path
    from svgrasterize_b200 import _lib, encode, scene as S, synth
    from svgrasterize_b200.engine import default_engine

    rng = np.random.default_rng(77)
    n = 1500
    xs = np.linspace(5, 250, n)
    ys = 128 + 100 * np.sin(xs / 7.0) * rng.uniform(0.5, 1.0, n)
    pb = synth.PathBuilder().move_to(float(xs[0]), float(ys[0]))
    for k in range(1, n):
        if k % 97 == 0:
            pb.line_to(float(xs[k - 1]), float(ys[k - 1]))  # a zero-length segment
        if k % 3:
            pb.line_to(float(xs[k]), float(ys[k]))
        else:
            pb.quad_to(float(xs[k] - 0.1), float(ys[k] + 3), float(xs[k]), float(ys[k]))
    if closed:
        pb.close()
    path = pb.path()
    eng = default_engine()
    for width, cap, join in ((1.5, "round", "round"), (3.0, None, None), (0.7, "square", "bevel")):
        enc = encode.Encoder(eng)
        enc.add_stroke_path(path, S.Transform(), width, cap, join, None)
        eng.render(enc.finish(), stop=_lib.STOP_STROKE)
        tag, data, _path, sub = eng.outline()
        want = stroke_path(path, width, cap, join)
        rt, rd, rs = encode.path_arrays(want)
        assert np.array_equal(tag, rt) and np.array_equal(data.view(np.uint64), rd.view(np.uint64)), (closed, cap, join)
        # the sub-path structure (forward / backward outlines of a closed path are separate sub-paths)
        assert len(np.unique(sub)) == len(rs) - 1


def _rand_pattern(rng, S, synth):
    """Pattern paint (svgrasterize.py:1049-1097): a small tile scene repeated under its own transform, in user
    space or objectBoundingBox units, with or without a viewBox."""
    tile = S.Scene.group([
        S.Scene.fill(synth.rect_path(*rng.uniform(0, 3, 2), *rng.uniform(3, 7, 2)), synth.color(*rng.uniform(0, 1, 3), rng.uniform(0.4, 1))),
        S.Scene.fill(synth.ellipse_path(*rng.uniform(4, 9, 2), rng.uniform(1.5, 4)), synth.color(*rng.uniform(0, 1, 3), rng.uniform(0.4, 1))),
    ])
    tr = S.Transform() if rng.random() < 0.4 else S.Transform().rotate(rng.uniform(-0.6, 0.6)).scale(*rng.uniform(0.7, 1.6, 2))
    mode = int(rng.integers(0, 4))
    if mode == 0:  # everything in user space
        return S.Pattern(tile, False, None, float(rng.uniform(0, 4)), float(rng.uniform(0, 4)), float(rng.uniform(8, 16)),
                         float(rng.uniform(8, 16)), tr, False)
    if mode == 1:  # user-space tile with a viewBox
        return S.Pattern(tile, False, (0.0, 0.0, 12.0, 12.0), 0.0, 0.0, float(rng.uniform(8, 20)), float(rng.uniform(8, 20)), tr, False)
    if mode == 2:  # tile rectangle in objectBoundingBox units, content through a viewBox
        return S.Pattern(tile, False, (0.0, 0.0, 12.0, 12.0), 0.0, 0.0, float(rng.uniform(0.15, 0.5)), float(rng.uniform(0.15, 0.5)),
                         tr, True)
    # content in objectBoundingBox units (patternContentUnits): a unit-square tile scene
    unit = S.Scene.group([
        S.Scene.fill(synth.rect_path(0.0, 0.0, 0.12, 0.12), synth.color(*rng.uniform(0, 1, 3))),
        S.Scene.fill(synth.ellipse_path(0.16, 0.16, 0.06), synth.color(*rng.uniform(0, 1, 3), 0.7)),
    ])
    return S.Pattern(unit, True, None, 0.0, 0.0, float(rng.uniform(0.2, 0.4)), float(rng.uniform(0.2, 0.4)), S.Transform(), True)


def _rand_bbox_paint(rng, S, synth):
    """Gradients in objectBoundingBox units (the SVG default for gradientUnits)."""
    stops = [(0.0, synth.color(*rng.uniform(0, 1, 3))), (float(rng.uniform(0.3, 0.7)), synth.color(*rng.uniform(0, 1, 3), 0.8)),
             (1.0, synth.color(*rng.uniform(0, 1, 3)))]
    spread = ("pad", "repeat", "reflect")[int(rng.integers(0, 3))]
    tr = None if rng.random() < 0.6 else S.Transform().rotate(rng.uniform(-0.5, 0.5))
    if rng.random() < 0.5:
        return S.GradLinear(rng.uniform(0, 0.4, 2), rng.uniform(0.6, 1.0, 2), stops, tr, spread, True, None)
    c = rng.uniform(0.3, 0.7, 2)
    f = None if rng.random() < 0.5 else c + rng.uniform(-0.2, 0.2, 2)
    return S.GradRadial(c, float(rng.uniform(0.3, 0.6)), f, None, stops, tr, spread, True, None)


@pytest.mark.parametrize("seed", range(32))
def test_random_pattern_and_bbox_unit_scenes_match_oracle(seed):
    """Pattern fills and objectBoundingBox units everywhere they can appear (paints, clip paths, masks): the
    cases the encoder resolves with a device round trip per node (ConvexHull.bbox, svgrasterize.py:2002-2023)."""
    import warnings

    import svgrasterize_b200 as B
    from oracle import render as O
    from svgrasterize_b200 import scene as S, synth

    rng = np.random.default_rng(91000 + seed)
    size = (int(rng.integers(64, 150)), int(rng.integers(64, 150)))
    kids = []
    for _ in range(int(rng.integers(2, 5))):
        path = _rand_path(rng, S, synth, 1) if rng.random() < 0.6 else synth.rect_path(*rng.uniform(4, 20, 2), *rng.uniform(20, 40, 2), 4.0)
        r = rng.random()
        paint = _rand_pattern(rng, S, synth) if r < 0.5 else _rand_bbox_paint(rng, S, synth) if r < 0.85 else _rand_paint(rng, S, synth)
        if rng.random() < 0.25:
            node = S.Scene.stroke(path, paint, float(rng.uniform(2, 7)), "round", "round")
        else:
            node = S.Scene.fill(path, paint, (None, "evenodd")[int(rng.integers(0, 2))])
        r = rng.random()
        if r < 0.2:  # clipPathUnits = objectBoundingBox
            node = node.clip(S.Scene.fill(synth.ellipse_path(0.5, 0.5, float(rng.uniform(0.3, 0.5))), np.ones(4)), bbox_units=True)
        elif r < 0.35:  # maskContentUnits = objectBoundingBox
            node = node.mask(S.Scene.fill(synth.rect_path(0.1, 0.1, 0.8, 0.8, 0.2), _rand_bbox_paint(rng, S, synth)), bbox_units=True)
        elif r < 0.5:
            node = node.opacity(float(rng.uniform(0.3, 0.9)))
        kids.append(node)
    scene = S.Scene.group(kids).transform(S.Transform().scale(float(rng.uniform(0.9, 2.0))))
    linear_rgb = bool(seed % 4 == 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ref = O.render_canvas(scene, size, linear_rgb)
        except TypeError:  # the reference's failure mode when a pattern tile misses its repeat cell (:1093-1094)
            with pytest.raises(TypeError):
                B.render_canvas(scene, size, linear_rgb)
            return
        got = B.render_canvas(scene, size, linear_rgb)
    if ref is None:
        assert not got.any()
        return
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1, (seed, int(diff.max()), int((diff > 1).sum()))


@pytest.mark.parametrize("seed", range(24))
def test_demand_boxes_do_not_change_a_byte(seed, monkeypatch):
    """The planner folds a group under a clip / mask only where its consumer can observe it (engine.cu demand_range).
    Coverage is deterministic, so the claim "same arithmetic per visible pixel" is checked exactly: random scenes with
    nested groups, clips, masks, filters and opacities render to the same bytes with the demand boxes switched off."""
    import warnings

    import svgrasterize_b200 as B
    from svgrasterize_b200 import scene as S, synth

    rng = np.random.default_rng(9100 + seed)
    size = (int(rng.integers(64, 160)), int(rng.integers(64, 160)))
    scene = _rand_scene(rng, S, synth).transform(S.Transform().scale(float(rng.uniform(1.0, 2.4))))
    # make sure a group sits under a clip and under a mask whatever the random tree looks like
    inner = S.Scene.group([_rand_leaf(rng, S, synth), _rand_leaf(rng, S, synth), _rand_leaf(rng, S, synth)])
    clipped = inner.opacity(0.8).clip(S.Scene.fill(synth.ellipse_path(30, 28, 14, 11), np.ones(4)))
    masked = S.Scene.group([_rand_leaf(rng, S, synth), _rand_leaf(rng, S, synth)]).mask(
        S.Scene.group([_rand_leaf(rng, S, synth), _rand_leaf(rng, S, synth)]))
    scene = S.Scene.group([scene, clipped, masked])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            monkeypatch.delenv("SVGR_NO_DEMAND", raising=False)
            with_demand = B.render_canvas(scene, size, bool(seed % 2))
        except TypeError:  # degenerate stroke (the reference's own failure mode): nothing to compare
            return
        monkeypatch.setenv("SVGR_NO_DEMAND", "1")
        without = B.render_canvas(scene, size, bool(seed % 2))
    assert np.array_equal(with_demand, without)
