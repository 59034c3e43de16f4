"""Size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes):
batch == singles, scale consistency of a 4096 px render, and algebraic identities of the filter stencils
on 8192 x 8192 layers."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from svgrasterize_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def test_batch_equals_singles(eng):
    """c5: rendering many SVGs as one program gives the same bytes as rendering each on its own."""
    from svgrasterize_b200 import encode, synth

    n = 192
    progs = [encode.encode_scene(synth.icon_scene(1000 + i), synth.icon_size()) for i in range(n)]
    batch = encode.Program.concat(progs)
    res = eng.render(batch)
    assert res["n_launches"] <= 4  # the whole batch is a handful of compose launches, not n of them
    for i in (0, 1, 63, 100, n - 1):
        single = eng.canvas(progs[i], eng.render(progs[i])["canvas"])
        got = eng.canvas(batch, res["canvas"], i)
        assert got.shape == single.shape
        assert int(np.abs(got.astype(int) - single.astype(int)).max()) <= 1


def test_material_design_4096_is_the_1024_render_scaled(eng):
    """c2 at its full size (4096 x 4096, 1 924 masks, a 935-layer group): the reference at this size takes
    ~6 s, so the check is scale consistency -- 4 x 4 box-averaging the 4096 render must reproduce the
    1024 render (whose bytes are pinned to the reference) up to anti-aliasing differences on edges."""
    from svgrasterize_b200 import encode
    from svgrasterize_b200.scene import Transform

    scene, size, lin, z = load_golden("demo_material_w1024")
    big = scene.transform(Transform().scale(4.0))
    prog = encode.encode_scene(big, (4096, 4096), lin, engine=eng)
    res = eng.render(prog, timing=True)
    img = eng.canvas(prog, res["canvas"]).astype(np.float64)
    assert img.shape == (4096, 4096, 4)
    small = img.reshape(1024, 4, 1024, 4, 4).mean(axis=(1, 3))
    ref = z["canvas_u8"].astype(np.float64)
    alpha_err = np.abs(small[..., 3] - ref[..., 3])
    assert alpha_err.mean() < 0.6 and np.percentile(alpha_err, 99.9) < 48
    assert abs(small[..., 3].sum() / ref[..., 3].sum() - 1.0) < 2e-3  # covered area is preserved
    interior, outside = ref[..., 3] == 255, ref[..., 3] == 0
    assert np.percentile(np.abs(small[interior][:, 3] - 255), 99.9) <= 16
    assert np.percentile(small[outside][:, 3], 99.9) <= 16


def test_filter_identities_8192(eng):
    """c4-sized layers (8192 x 8192 RGBA float32 = 1 GiB each):
    * full convolution with a normalised kernel preserves the sum of every channel,
    * a k1 window max followed by a k2 window max is the (k1 + k2 - 1) window max, exactly,
    * a colour matrix that only permutes channels is a permutation, exactly."""
    import svgrasterize_b200 as B

    n = 8192
    rng = np.random.default_rng(0)
    tile = rng.uniform(0.0, 1.0, size=(256, 256, 4)).astype(np.float32)
    img = np.tile(tile, (n // 256, n // 256, 1))
    layer = B.Layer(img, (0, 0), False, True)  # straight alpha, linear: convolve / color_matrix do not convert
    # --- blur: 21 x 21 separable Gaussian of sigma 4 at identity scale
    kern = B.blur_kernel(B.Transform().matrix(0, 1, 0, 1, 0, 0), (4.0, 4.0))
    assert kern.shape == (21, 21)
    out = layer.convolve(kern)
    assert out.image.shape == (n + 20, n + 20, 4) and tuple(out.offset) == (-10, -10)
    want = img.sum(axis=(0, 1), dtype=np.float64)
    got = out.image.sum(axis=(0, 1), dtype=np.float64)
    assert np.abs(got / want - 1.0).max() < 1e-5
    del out
    # --- morphology on the premultiplied-linear flagged copy (no conversion)
    pre = B.Layer(img, (0, 0), True, True)
    a = pre.morphology(4, 3, "max").morphology(3, 5, "max")
    b = pre.morphology(6, 7, "max")
    assert a.image.shape == b.image.shape == (n - 5, n - 6, 4)
    assert np.array_equal(a.image, b.image)
    del a, b
    # --- colour matrix: swap red and blue, keep the rest
    m = np.zeros((4, 5))
    m[0, 2] = m[2, 0] = m[1, 1] = m[3, 3] = 1.0
    sw = layer.color_matrix(m)
    assert np.array_equal(sw.image[..., 0], img[..., 2]) and np.array_equal(sw.image[..., 2], img[..., 0])
    assert np.array_equal(sw.image[..., 3], img[..., 3])


def test_plan_cache_reuses_tables_only_for_the_same_program_and_boxes(eng):
    """A resident re-render that finds the same path boxes repeats the launches on the tables of the render before
    (no host planning, uploads or culls); a new program, or SVGR_NO_PLAN_CACHE, plans again.  Same pixels either way."""
    import os

    import torch

    from svgrasterize_b200 import encode, synth

    prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(50 + i), synth.icon_size()) for i in range(40)])
    other = encode.encode_scene(synth.filter_stack_scene(200), (200, 200))
    out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
    first = eng.render(prog, out=out)
    assert first["plan_cached"] == 0
    a = out.cpu().numpy().copy()
    st = eng.render_resident(out)
    assert st["plan_cached"] == 1 and st["host_plan_nodes_ms"] == 0.0 and st["n_launches"] == first["n_launches"]
    b = out.cpu().numpy().copy()
    assert int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) <= 1
    os.environ["SVGR_NO_PLAN_CACHE"] = "1"
    try:
        st = eng.render_resident(out)
        assert st["plan_cached"] == 0 and st["host_plan_nodes_ms"] > 0.0
    finally:
        del os.environ["SVGR_NO_PLAN_CACHE"]
    c = out.cpu().numpy()
    assert int(np.abs(a.astype(np.int16) - c.astype(np.int16)).max()) <= 1
    # filters (tensor maps) through the cache too
    small = torch.empty(other.canvas_bytes, dtype=torch.uint8, device="cuda")
    assert eng.render(other, out=small)["plan_cached"] == 0  # a new program invalidates
    x = small.cpu().numpy().copy()
    assert eng.render_resident(small)["plan_cached"] == 1
    assert int(np.abs(x.astype(np.int16) - small.cpu().numpy().astype(np.int16)).max()) <= 1
    assert eng.render(prog, out=out)["plan_cached"] == 0


def test_renders_are_bit_reproducible():
    """Coverage accumulates in fixed point (integer shared-memory atomics commute exactly), so nothing in the pipeline
    depends on the order in which threads arrive any more: the same program renders to the same bytes, and to the
    same PNG files, every time."""
    from svgrasterize_b200 import native, synth
    from svgrasterize_b200.engine import Engine

    eng = Engine(0)
    try:
        jobs = [(synth.icon_scene(8100 + i), synth.icon_size(), bool(i % 3 == 0)) for i in range(192)]
        prog = native.encode_batch(jobs)
        first = eng.render(prog)["canvas"].copy()
        png = eng.render_png(prog)
        files = png["png"][: png["offsets"][-1]].copy()
        for _ in range(3):
            assert np.array_equal(eng.render(prog)["canvas"], first)
        again = eng.render_png(prog)
        assert np.array_equal(again["png"][: again["offsets"][-1]], files)
        other = Engine(0)  # a second context: different buffers, same bytes
        try:
            assert np.array_equal(other.render(prog)["canvas"], first)
        finally:
            other.close()
        prog.close()
    finally:
        eng.close()
