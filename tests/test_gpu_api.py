"""The reference's call surface (svgrasterize_b200.api) on the GPU against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import svgrasterize_b200 as B

    return B


@pytest.fixture(scope="module")
def O():
    from oracle import render as O

    return O


def _tr(B):
    return B.Transform().matrix(0, 1, 0, 1, 0, 0) @ B.Transform().scale(3.0).rotate(0.15).translate(2, 1)


def _olayer(O, layer):
    return O.OLayer(layer.image.astype(np.float64), tuple(layer.offset), layer.pre_alpha, layer.linear_rgb)


def test_path_mask_fill_stroke(B, O):
    from oracle import geometry
    from oracle.stroke import stroke_path
    from svgrasterize_b200 import synth

    path = synth._blob(np.random.default_rng(5), 30, 28, 14)
    tr = _tr(B)
    for rule in (None, "evenodd"):
        layer, hull = path.mask(tr, rule, viewport=[0, 0, 120, 150])
        ref_mask, ref_off, ref_edges = geometry.path_mask(path, tr, rule, [0, 0, 120, 150])
        assert tuple(layer.offset) == tuple(ref_off) and layer.image.shape == (*ref_mask.shape, 1)
        assert (layer.pre_alpha, layer.linear_rgb) == (True, True)
        assert np.abs(layer.image[..., 0] - ref_mask).max() <= 1e-5
        assert np.allclose(hull.bbox(tr), O.Cloud(ref_edges).bbox(tr), rtol=0, atol=0)
    assert path.mask(tr, None, viewport=[500, 500, 10, 10]) is None
    with pytest.raises(ValueError):
        path.mask(tr, "bogus")
    # gradient fill
    stops = [(0.0, synth.color(1, 0, 0)), (1.0, synth.color(0, 0, 1, 0.5))]
    grad = B.GradLinear(np.array([10.0, 10.0]), np.array([50.0, 40.0]), stops, None, "reflect", False, None)
    got, _ = path.fill(tr, grad, None, None, linear_rgb=False)
    ref, _ = O.fill_path(path, tr, grad, None, None, False)
    assert tuple(got.offset) == tuple(ref.offset) and np.abs(got.image - ref.image).max() <= 1e-5
    assert path.fill(tr, None) is None
    # stroke outline: bit-exact
    out = path.stroke(2.5, "round", "round")
    ref = stroke_path(path, 2.5, "round", "round")
    assert len(out.subpaths) == len(ref.subpaths)
    for a, b in zip(out.subpaths, ref.subpaths):
        assert len(a) == len(b)
        for (ta, pa), (tb, pb) in zip(a, b):
            assert ta == tb and np.array_equal(np.asarray(pa), np.asarray(pb))
    with pytest.raises(ValueError):
        path.stroke(1.0, "bogus")


def test_flatten_batch_is_the_reference_set(B):
    from oracle import geometry

    rng = np.random.default_rng(11)
    cubics = rng.uniform(0, 300, size=(200, 4, 2))
    got = B.bezier3_flatten_batch(cubics, 0.1).reshape(-1, 4)
    ref = geometry.flatten_cubics(cubics).reshape(-1, 4)
    got = got[np.lexsort(got.T[::-1])]
    ref = ref[np.lexsort(ref.T[::-1])]
    assert got.shape == ref.shape and np.array_equal(got.view(np.uint64), ref.view(np.uint64))
    assert B.bezier3_flatten_batch(np.zeros((0, 4, 2))).shape == (0, 2, 2)


def test_layer_ops(B, O):
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, size=(40, 50, 4))
    a[..., :3] *= a[..., 3:]
    b = rng.uniform(0, 1, size=(30, 45, 4))
    b[..., :3] *= b[..., 3:]
    m = rng.uniform(0, 1, size=(35, 35, 1))
    la, lb = B.Layer(a.astype(np.float32), (3, 4), True, False), B.Layer(b.astype(np.float32), (10, -5), True, False)
    lm = B.Layer(m.astype(np.float32), (0, 0), True, True)
    oa, ob, om = _olayer(O, la), _olayer(O, lb), _olayer(O, lm)
    for mode in (0, 1, 2, 3, 4, (0.3, 0.5, 0.4, 0.05)):
        got = B.Layer.compose([la, lb], mode, linear_rgb=False)
        ref = O.compose([oa, ob], mode, False)
        assert tuple(got.offset) == tuple(ref.offset) and got.image.shape == ref.image.shape
        assert (got.pre_alpha, got.linear_rgb) == (ref.pre_alpha, ref.linear_rgb)
        assert np.abs(got.image - ref.image).max() <= 2e-5, mode
    got, ref = B.Layer.compose([lm, la], 2, True), O.compose([om, oa], 2, True)
    assert tuple(got.offset) == tuple(ref.offset) and np.abs(got.image - ref.image).max() <= 2e-5
    assert B.Layer.compose([la, B.Layer(b.astype(np.float32), (500, 500), True, False)], 2) is None
    assert B.Layer.compose([la]) is la and B.Layer.compose([]) is None
    with pytest.raises(ValueError):
        B.Layer.compose([la, lb], 9)
    # convert round trip, opacity
    for pre, lin in ((False, True), (True, True), (False, False)):
        got, ref = la.convert(pre, lin), O.convert(oa, pre, lin)
        assert (got.pre_alpha, got.linear_rgb) == (pre, lin) and np.abs(got.image - ref.image).max() <= 2e-5
    got, ref = la.opacity(0.4, True), O.opacity(oa, 0.4, True)
    assert np.abs(got.image - ref.image).max() <= 2e-5
    # filters
    mat = np.eye(4, 5)
    mat[0, 1], mat[2, 4], mat[3, 3] = 0.4, 0.1, 0.8
    got, ref = la.color_matrix(mat), O.color_matrix(oa, mat)
    assert (got.pre_alpha, got.linear_rgb) == (False, True) and np.abs(got.image - ref.image).max() <= 2e-5
    with pytest.raises(ValueError):
        la.color_matrix(np.eye(4))
    for k0, k1, method in ((3, 5, "max"), (4, 2, "min")):
        got, ref = la.morphology(k0, k1, method), O.morphology(oa, k0, k1, method)
        assert got.image.shape == ref.image.shape and tuple(got.offset) == tuple(ref.offset)
        assert np.abs(got.image - ref.image).max() <= 2e-5
    with pytest.raises(ValueError):
        la.morphology(2, 2, "avg")
    sep = B.blur_kernel(B.Transform().matrix(0, 1, 0, 1, 0, 0).scale(2.0), (1.5, 2.5))
    rot = B.blur_kernel(B.Transform().rotate(0.5).scale(2.0, 1.3), (2.0, 0.8))
    assert np.allclose(sep, O.blur_kernel(B.Transform().matrix(0, 1, 0, 1, 0, 0).scale(2.0), (1.5, 2.5)), atol=0, rtol=0)
    for kern in (sep, rot):
        got, ref = la.convolve(kern), O.convolve(oa, kern)
        assert got.image.shape == ref.image.shape and tuple(got.offset) == tuple(ref.offset)
        assert np.abs(got.image - ref.image).max() <= 2e-5


def test_scene_render_and_canvas(B, O):
    from svgrasterize_b200 import synth

    scene, size = synth.icon_scene(4), synth.icon_size()
    tr = B.Transform().matrix(0, 1, 0, 1, 0, 0)
    vp = [0, 0, size[1], size[0]]
    got, hull = scene.render(tr, viewport=vp)
    ref, cloud = O.render(scene, tr, viewport=vp)
    assert tuple(got.offset) == tuple(ref.offset) and got.image.shape == ref.image.shape
    assert np.abs(got.image - ref.image).max() <= 2e-5
    assert np.array_equal(np.asarray(hull.bbox(tr)), np.asarray(cloud.bbox(tr)))
    mask, _ = scene.render(tr, mask_only=True, viewport=vp)
    mref, _ = O.render(scene, tr, True, vp)
    assert mask.image.shape == mref.image.shape and np.abs(mask.image - mref.image).max() <= 2e-5
    u8 = B.render_canvas(scene, size)
    assert np.abs(u8.astype(int) - O.render_canvas(scene, size).astype(int)).max() <= 1
    png = B.canvas_to_png(u8).getvalue()  # a BytesIO like the reference's (svgrasterize.py:267)
    assert png[:8] == b"\x89PNG\r\n\x1a\n"


def test_canvas_helpers(B, O):
    rng = np.random.default_rng(8)
    base = rng.uniform(0, 1, size=(20, 30, 4)).astype(np.float32)
    base[..., :3] *= base[..., 3:]
    over = rng.uniform(0, 1, size=(12, 40, 4)).astype(np.float32)
    over[..., :3] *= over[..., 3:]
    ref = O.merge_at(base.astype(np.float64).copy(), over.astype(np.float64), (15, -4))
    got = B.canvas_merge_at(base.copy(), over, (15, -4))
    assert np.abs(got - ref).max() <= 2e-5
    pooled = B.pooling(base, (3, 2), stride=(1, 1), method="max")
    want = O.morphology(O.OLayer(base.astype(np.float64), (0, 0), True, True), 3, 2, "max").image
    assert np.abs(pooled - want).max() <= 1e-6
    img, off = B.canvas_merge_union([(base, (0, 0)), (over, (5, 5))], full=False)
    want, woff = O.merge_over([(base.astype(np.float64), (0, 0)), (over.astype(np.float64), (5, 5))], 0)
    assert tuple(off) == tuple(woff) and np.abs(img - want).max() <= 2e-5


def test_row_bands_reassemble_to_the_full_render(B, O):
    """Single huge render sharded by row bands: each band rendered on its own must equal
    the same rows of the full render -- including under a blur + dilate filter stack (halo by redundant
    compute) -- so that the NCCL gather of bands is the only multi-GPU step."""
    from svgrasterize_b200 import parallel as P, synth
    from svgrasterize_b200.engine import default_engine

    eng = default_engine()
    for scene, size in ((synth.filter_stack_scene(320), (320, 320)), (synth.icon_scene(1), synth.icon_size())):
        full = B.render_canvas(scene, size)
        for world in (2, 3):
            bands = [P.render_band(eng, scene, size, world, r) for r in range(world)]
            got = np.concatenate(bands, axis=0)
            assert got.shape == full.shape
            assert int(np.abs(got.astype(int) - full.astype(int)).max()) <= 1, (world, size)
    dev = P.render_band(eng, synth.icon_scene(1), synth.icon_size(), 2, 1, device_out=True)
    assert dev.is_cuda and tuple(dev.shape) == (128, 256, 4)


def test_row_bands_place_filter_results_where_the_full_render_does(B, O):
    """Layer.convolve puts its result at int(r0 - kw / 2) (svgrasterize.py:114) and feOffset shifts by int(...)
    (:1849): truncation toward ZERO, i.e. a ceiling for a layer that starts within half a kernel of row 0 and a
    floor for one that starts lower.  A band sees the layer clipped to its own rows, so it takes the truncation from
    the box the whole-canvas render gives the layer (planner shadow boxes): the bands of a blurred / offset shape
    that touches the canvas top must join up with no seam, and the two-circle gradient's any(det < 0) rule (:1621)
    must be decided over the whole mask."""
    from svgrasterize_b200 import parallel as P, scene as S, synth
    from svgrasterize_b200.engine import default_engine

    eng = default_engine()
    w = h = 240
    stops = [(0.0, synth.color(0.9, 0.2, 0.1)), (1.0, synth.color(0.1, 0.3, 0.9, 0.7))]
    focal = S.GradRadial(np.array([120.0, 60.0]), 70.0, np.array([120.0 + 90.0, 60.0]), None, stops, None, "pad", False, None)
    tall = synth.rect_path(30, 1, 180, 230, 12.0)  # starts at row 1: int(r0 - kw / 2) is negative in the full render
    flt = S.Filter.empty().blur(3.0, 3.0).offset(2.6, -3.4)
    scenes = [
        S.Scene.fill(tall, focal).filter(flt),
        S.Scene.group([S.Scene.fill(tall, synth.color(0.2, 0.6, 0.3)),
                       S.Scene.fill(synth.ellipse_path(120, 200, 30), synth.color(0.9, 0.8, 0.1))]).filter(
                           S.Filter.empty().blur(5.0, 2.0).morphology(2.0, 2.0, "max", None)),
    ]
    for scene in scenes:
        full = B.render_canvas(scene, (w, h))
        ref = O.render_canvas(scene, (w, h))
        assert int(np.abs(full.astype(int) - ref.astype(int)).max()) <= 1
        for world in (2, 3, 5):
            got = np.concatenate([P.render_band(eng, scene, (w, h), world, r) for r in range(world)], axis=0)
            assert got.shape == full.shape
            assert int(np.abs(got.astype(int) - full.astype(int)).max()) <= 1, world


def test_empty_and_degenerate_inputs(B, O):
    """What the reference answers with None / an all-transparent canvas (svgrasterize.py:958-959, :974-975,
    :1004-1010, :686-687, :403-404)."""
    from svgrasterize_b200 import synth

    tr = B.Transform().matrix(0, 1, 0, 1, 0, 0)
    assert B.Path([]).mask(tr) is None                      # no segments
    assert B.Path([[]]).mask(tr) is None                    # an empty sub-path
    dot = synth.PathBuilder().move_to(5, 5).line_to(5, 5).close().path()
    res = dot.mask(tr)                                       # zero-area path: a (tiny) mask of zeros, like the reference
    ref = O.mask_path(dot, tr)
    assert (res is None) == (ref is None)
    if res is not None:
        assert res[0].image.shape == ref[0].image.shape and float(np.abs(res[0].image).max()) == 0.0
    sq = synth.rect_path(10, 10, 20, 20)
    assert sq.mask(tr, viewport=[100, 100, 10, 10]) is None  # clipped away by the viewport
    nothing = B.Scene.fill(sq, None)                          # fill without paint renders nothing
    assert nothing.render(tr) is None
    u8 = B.render_canvas(B.Scene.group([nothing, B.Scene.fill(sq, None)]), (32, 24))
    assert u8.shape == (24, 32, 4) and not u8.any()
    # a clip that misses its target: compose IN with an empty intersection -> None
    far = synth.rect_path(200, 200, 5, 5)
    clipped = B.Scene.fill(sq, synth.color(1, 0, 0)).clip(B.Scene.fill(far, np.ones(4)))
    assert clipped.render(tr) is None and O.render(clipped, tr) is None
    # a huge coordinate range is clipped by the viewport, not allocated
    big = synth.rect_path(-1e6, -1e6, 2e6, 2e6)
    layer, _ = big.mask(tr, viewport=[0, 0, 16, 16])
    assert layer.image.shape == (16, 16, 1) and float(layer.image.min()) == 1.0


def test_background_intersect_and_png(B, O):
    import io
    import zlib

    rng = np.random.default_rng(21)
    a = rng.uniform(0, 1, size=(20, 24, 4)).astype(np.float32)
    a[..., :3] *= a[..., 3:]
    la = B.Layer(a, (2, 3), True, True)
    bg = np.array([0.2, 0.4, 0.1, 1.0])
    got = la.background(bg)
    want = O.blend(0, np.broadcast_to(bg, a.shape).astype(np.float64), a.astype(np.float64))
    assert got.image.shape == a.shape and np.abs(got.image - want).max() <= 2e-5
    m = rng.uniform(0, 1, size=(30, 30, 1)).astype(np.float32)
    res = B.canvas_merge_intersect([(m, (0, 0)), (a, (2, 3))], blend=2)
    ref = O.merge_intersect([(m.astype(np.float64), (0, 0)), (a.astype(np.float64), (2, 3))], 2)
    assert tuple(res[1]) == tuple(ref[1]) and np.abs(res[0] - ref[0]).max() <= 2e-5
    assert B.canvas_merge_intersect([(m, (0, 0)), (a, (100, 100))]) is None
    buf = io.BytesIO()
    la.convert(pre_alpha=True, linear_rgb=False).write_png(buf)
    png = buf.getvalue()
    assert png[:8] == b"\x89PNG\r\n\x1a\n" and png[12:16] == b"IHDR"
    width, height = int.from_bytes(png[16:20], "big"), int.from_bytes(png[20:24], "big")
    assert (width, height) == (24, 20)
    idat = png[png.index(b"IDAT") + 4: png.index(b"IEND") - 8]
    raw = zlib.decompress(idat)
    assert len(raw) == height * (1 + 4 * width)


def test_engine_recovers_after_a_failed_call(B, O):
    """A degenerate control-polygon leg makes the reference's stroker raise TypeError (bezier3_offset :2157,
    line_offset returns None); the call fails half way through the pipeline, and the same engine must render
    the next scene correctly (nothing of the failed call may still be in flight)."""
    from oracle.stroke import stroke_path
    from svgrasterize_b200 import synth

    bad = synth.PathBuilder().move_to(0, 0).cubic_to(0, 1.2e-8, 30, 40, 50, 10).path()
    with pytest.raises(TypeError):
        stroke_path(bad, 3.0)
    for _ in range(3):
        with pytest.raises(TypeError):
            bad.stroke(3.0)
        scene = B.Scene.group([synth.icon_scene(5), B.Scene.stroke(bad, synth.color(0, 0, 1), 3.0)])
        with pytest.raises(TypeError):
            B.render_canvas(scene, (64, 64))
        got = B.render_canvas(synth.icon_scene(7), synth.icon_size())
        want = O.render_canvas(synth.icon_scene(7), synth.icon_size())
        assert np.abs(got.astype(np.int16) - want.astype(np.int16)).max() <= 1
