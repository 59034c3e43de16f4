"""Pins the CPU oracle (oracle/) to the unmodified reference.

The vectors under tests/golden/ were produced by tools/make_golden.py, which
imports /root/reference/svgrasterize.py and records what it computes.  These
tests need no GPU and no reference checkout.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import render as O
from svgrasterize_b200 import sceneio

ALL = golden_names()
STAGED = [n for n in ALL if "leaf_edge_off" in load_golden(n)[3].files]
ROOTED = [n for n in ALL if "root_image" in load_golden(n)[3].files]


@pytest.mark.parametrize("name", ALL)
def test_canvas_matches_reference_exactly(name):
    """Final uint8 image (svgrasterize.py:3870-3881, :263): the float64 oracle
    reproduces the reference's bytes."""
    scene, size, linear_rgb, z = load_golden(name)
    got = O.render_canvas(scene, size, linear_rgb=linear_rgb)
    assert got.shape == z["canvas_u8"].shape
    assert np.array_equal(got, z["canvas_u8"])


@pytest.mark.parametrize("name", STAGED)
def test_leaf_stages(name):
    """Per Path.mask call: bbox exact, edge list bit-exact (as a sorted multiset),
    mask equal to the float32-rounded reference mask."""
    scene, size, linear_rgb, z = load_golden(name)
    O.TAPS = {"leaves": [], "strokes": []}
    try:
        O.render(scene, O.canvas_transform(), viewport=[0, 0, int(size[1]), int(size[0])], linear_rgb=linear_rgb)
        taps = O.TAPS
    finally:
        O.TAPS = None
    bbox, eoff, moff = z["leaf_bbox"], z["leaf_edge_off"], z["leaf_mask_off"]
    assert len(taps["leaves"]) == len(bbox)
    for i, leaf in enumerate(taps["leaves"]):
        if leaf is None:
            assert tuple(bbox[i]) == (-1, -1, -1, -1)
            continue
        bb, edges, mask = leaf
        assert tuple(bbox[i]) == tuple(bb)
        e = edges.reshape(-1, 4)
        e = e[np.lexsort(e.T[::-1])]
        ref_e = z["edges"][eoff[i]: eoff[i + 1]]
        assert e.shape == ref_e.shape and np.array_equal(e.view(np.uint64), ref_e.view(np.uint64))
        ref_m = z["masks"][moff[i]: moff[i + 1]].reshape(mask.shape)
        assert np.array_equal(mask.astype(np.float32), ref_m)
    if "stroke_off" in z.files:
        so, ss = z["stroke_off"], z["stroke_sub_idx"]
        assert len(taps["strokes"]) == len(so) - 1
        for k, path in enumerate(taps["strokes"]):
            t, d, s = sceneio.path_arrays(path)
            assert np.array_equal(t, z["stroke_tag"][so[k]: so[k + 1]])
            assert np.array_equal(s, z["stroke_sub_off"][ss[k]: ss[k + 1]])
            assert np.array_equal(d.view(np.uint64), z["stroke_data"][so[k]: so[k + 1]].view(np.uint64))


@pytest.mark.parametrize("name", ROOTED)
def test_root_layer(name):
    """The Layer returned by Scene.render: offset, shape, flags exact; pixels to float32 rounding."""
    scene, size, linear_rgb, z = load_golden(name)
    res = O.render(scene, O.canvas_transform(), viewport=[0, 0, int(size[1]), int(size[0])], linear_rgb=linear_rgb)
    layer = res[0]
    assert tuple(z["root_offset"]) == tuple(layer.offset)
    assert tuple(z["root_flags"]) == (layer.pre_alpha, layer.linear_rgb)
    assert z["root_image"].shape == layer.image.shape
    assert np.abs(layer.image - z["root_image"]).max() < 1e-6


def test_scene_roundtrip_is_lossless():
    scene, _size, _lin, z = load_golden("demo_icons_w512")
    again = sceneio.dump_scene(scene)
    for key in ("seg_tag", "seg_data", "sub_off", "path_off"):
        assert np.array_equal(again[key], z[key])


def _eager():
    import os

    from conftest import ROOT

    return np.load(os.path.join(ROOT, "tests", "golden_eager", "eager.npz"), allow_pickle=False)


def test_eager_restatements():
    """The oracle's restatements of the reference's small module-level functions (oracle/eager.py, the C
    line coverage / flattener with a free flatness, the blend / merge helpers with every mode) against
    vectors recorded from the unmodified reference by tools/make_golden_eager.py."""
    from oracle import clib, eager, geometry

    z = _eager()
    # line_signed_coverage (:2213): exact, the C restatement uses the same float64 operations
    trace = np.zeros_like(z["cov_trace"])
    for ln in z["cov_lines"]:
        clib.lib().orc_line_coverage(clib.dp(trace), trace.shape[0], trace.shape[1],
                                     clib.dp(np.ascontiguousarray(ln.reshape(4))))
    assert np.array_equal(trace, z["cov_trace"])
    # bezier3_flatten_batch (:2091) with other flatness values: same set of lines, bit for bit
    for i in range(4):
        got = geometry.flatten_cubics(z["flat_cubics"], float(z[f"flat_tol_{i}"])).reshape(-1, 4)
        assert np.array_equal(got[np.lexsort(got.T[::-1])], z[f"flat_lines_{i}"])
    assert np.array_equal(eager.pixel_centres(z["gp_viewport"]), z["gp_out"])
    for m in ("pad", "repeat", "reflect"):
        assert np.array_equal(eager.spread(z["gs_in"], m), z[f"gs_{m}"])
    stops = list(zip(z["gi_stop_off"], z["gi_stop_col"]))
    for lin in (0, 1):
        assert np.array_equal(eager.interpolate(z["gs_in"], stops, bool(lin)), z[f"gi_out_{lin}"])
    for i in range(int(z["pool_n"])):
        s = tuple(int(v) for v in z[f"pool_{i}_s"])
        got = eager.pool(z["pool_in"], tuple(z[f"pool_{i}_k"]), s if s != (0, 0) else None, str(z[f"pool_{i}_m"]),
                         bool(z[f"pool_{i}_pad"]))
        assert np.array_equal(got, z[f"pool_{i}_out"], equal_nan=True)
    assert np.array_equal(eager.pool(z["pool2_in"], (3, 3), (2, 1), "max", False), z["pool2_out"])
    modes = [0, 1, 2, 3, 4, tuple(z["cc_arith"])]
    for i, mode in enumerate(modes):
        assert np.array_equal(O.blend(mode, z["cc_dst"], z["cc_src"]), z[f"cc_out_{i}"])
    for i, mode in enumerate((0, 4, 3)):
        got = O.merge_at(z["ma_base"].copy(), z["ma_over"], tuple(z["ma_off"]), mode)
        assert np.array_equal(got, z[f"ma_out_{i}"])
    layers = [(z["cc_dst"], (0, 0)), (z["ma_over"], (5, -3)), (z["mu_l3"], (-4, 20))]
    for i, (full, mode) in enumerate(((False, 0), (True, 0), (True, 4), (True, 1))):
        img, off = (O.merge_full if full else O.merge_over)(layers, mode)
        assert tuple(off) == tuple(z[f"mu_off_{i}"]) and np.array_equal(img, z[f"mu_out_{i}"])
    ilayers = [(z["cc_m1"], (0, 0)), (z["ma_over"], (5, -3)), (z["mu_l3"], (2, 6))]
    for i, mode in enumerate((0, 2, 4)):
        img, off = O.merge_intersect(ilayers, mode)
        assert tuple(off) == tuple(z[f"mi_off_{i}"]) and np.array_equal(img, z[f"mi_out_{i}"])


@pytest.mark.parametrize("name", ["demo_material_w4096", "synth_filter_stack_2048"])
def test_full_size_canvases_match_reference_exactly(name):
    """BASELINE.json's c2 at 4096 x 4096 and c4 at 2048 x 2048 (tests/golden_big, recorded from the unmodified
    reference): the oracle reproduces the reference's bytes at the stated sizes too."""
    import os

    from conftest import ROOT

    z = np.load(os.path.join(ROOT, "tests", "golden_big", name + ".npz"), allow_pickle=False)
    scene, size = sceneio.load_scene(z), tuple(float(v) for v in z["size"])
    got = O.render_canvas(scene, size, linear_rgb=bool(z["linear_rgb"]))
    assert np.array_equal(got, z["canvas_u8"])
