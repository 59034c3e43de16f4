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
