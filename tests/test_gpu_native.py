"""The native scene encoder end to end on the device: a program recorded by csrc/encode_flat.cpp renders to the
pixels the Python-encoded program renders to, and to the reference's goldens."""
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


def test_native_batch_renders_like_the_python_encoded_batch(eng):
    from svgrasterize_b200 import encode, native, synth

    jobs = [(synth.icon_scene(900 + i), synth.icon_size(), False) for i in range(64)]
    nat = native.encode_batch(jobs)
    assert isinstance(nat, native.NativeProgram)
    a = eng.render(nat)["canvas"].copy()
    ref = encode.Program.concat([encode.encode_scene(s, size, lin) for s, size, lin in jobs])
    b = eng.render(ref)["canvas"]
    assert a.shape == b.shape
    assert np.array_equal(a, b)  # the two encoders record the same program; rendering is deterministic
    res = eng.render_png(nat)
    assert len(res["offsets"]) == 65 and res["png_bytes"] > 0
    nat.close()


@pytest.mark.parametrize("name", ["demo_material_w1024", "demo_prompt", "icon_tiger", "icon_rust", "synth_icon_4",
                                  "feat_radial_focal_outside", "feat_stroke_caps_joins",
                                  # filters (blur incl. rotated, offset, merge, composite, colour matrix, morphology)
                                  # and objectBoundingBox gradients through the native walk
                                  "demo_icons_w512", "icon_inkscape", "icon_office", "feat_filter_blur_rotated",
                                  "feat_filter_drop_shadow", "feat_filter_arithmetic", "feat_filter_erode_matrix",
                                  "synth_filter_stack_192", "feat_linear_bbox_reflect", "feat_radial_focal_bbox"])
def test_native_encoder_against_the_reference_bytes(eng, name):
    from svgrasterize_b200 import native

    scene, size, lin, z = load_golden(name)
    prog = native.encode_batch([(scene, size, lin)])
    assert isinstance(prog, native.NativeProgram)
    got = eng.canvas(prog, eng.render(prog)["canvas"])
    ref = z["canvas_u8"]
    assert got.shape == ref.shape
    assert int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1
    prog.close()


def _bbox_gradient_scenes():
    """Gradients in objectBoundingBox units under rotated / skewed transforms, with and without a gradientTransform,
    on fills and strokes: the pixel -> gradient map is completed on the device (svgr_bbox_job)."""
    from svgrasterize_b200 import scene as S, synth

    red, blue, green = synth.color(0.9, 0.1, 0.1), synth.color(0.1, 0.2, 0.9, 0.8), synth.color(0.1, 0.8, 0.2)
    stops = [(0.0, red), (0.4, green), (1.0, blue)]
    gt = S.Transform().rotate(0.3).scale(1.2, 0.8)
    lin = S.GradLinear(np.array([0.1, 0.0]), np.array([0.9, 1.0]), stops, None, "reflect", True, None)
    lin_t = S.GradLinear(np.array([0.0, 0.0]), np.array([1.0, 0.5]), stops, gt, "repeat", True, None)
    rad = S.GradRadial(np.array([0.5, 0.5]), 0.5, None, None, stops, None, "pad", True, None)
    foc = S.GradRadial(np.array([0.5, 0.5]), 0.55, np.array([0.35, 0.4]), 0.05, stops, gt, "pad", True, True)
    blob = synth.ellipse_path(40, 36, 28, 18)
    rect = synth.rect_path(10, 12, 70, 50, 9.0)
    parts = [
        S.Scene.fill(rect, lin),
        S.Scene.fill(blob, foc).opacity(0.8),
        S.Scene.stroke(rect, lin_t, 6.0, "round", "round"),
        S.Scene.fill(synth.ellipse_path(66, 60, 14, 22), rad),
    ]
    tr = S.Transform().translate(12, -6).rotate(0.2).skew(0.15, 0.0).scale(1.1)
    return {"group": S.Scene.group(parts).transform(tr), "single": S.Scene.fill(blob, lin_t).transform(tr)}


@pytest.mark.parametrize("name", ["group", "single"])
@pytest.mark.parametrize("linear_rgb", [False, True])
def test_bbox_unit_gradients_resolved_on_the_device(eng, name, linear_rgb):
    """ConvexHull.bbox_transform (:2002-2023) without an encode-time round trip: native program, no engine at encode
    time, pixels equal to the oracle's (which follows the reference's host computation) within 1 LSB."""
    from oracle import render as O
    from svgrasterize_b200 import encode, native

    scene = _bbox_gradient_scenes()[name]
    size = (112, 100)
    prog = native.encode_batch([(scene, size, linear_rgb)])
    assert isinstance(prog, native.NativeProgram) and len(prog.bbox_jobs) == (4 if name == "group" else 1)
    got = eng.canvas(prog, eng.render(prog)["canvas"]).copy()
    py = encode.encode_scene(scene, size, linear_rgb)  # no engine needed any more
    assert len(py.bbox_jobs) == len(prog.bbox_jobs)
    again = eng.canvas(py, eng.render(py)["canvas"])
    assert int(np.abs(got.astype(np.int16) - again.astype(np.int16)).max()) <= 1
    ref = O.render_canvas(scene, size, linear_rgb=linear_rgb)
    assert ref.shape == got.shape
    assert int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1
    # a resident re-render resolves again from the same inputs
    import torch

    buf = torch.zeros(py.canvas_bytes, dtype=torch.uint8, device="cuda")
    eng.render_resident(buf)
    assert int(np.abs(eng.canvas(py, buf.cpu().numpy()).astype(np.int16) - ref.astype(np.int16)).max()) <= 1
    prog.close()


def test_bbox_unit_gradient_on_a_degenerate_box_raises_like_numpy(eng):
    """A leaf whose end points span no width (but some height): bbox_transform scales by (0, h) and .invert raises
    numpy.linalg.LinAlgError (a ValueError) in the reference; the device resolver reports the same."""
    from svgrasterize_b200 import encode, scene as S, synth

    stops = [(0.0, synth.color(1, 0, 0)), (1.0, synth.color(0, 0, 1))]
    grad = S.GradLinear(np.zeros(2), np.ones(2), stops, None, "pad", True, None)
    b = synth.PathBuilder().move_to(20.0, 10.0).line_to(20.0, 40.0).line_to(20.0, 25.0)
    scene = S.Scene.fill(b.path(), grad)
    prog = encode.encode_scene(scene, (48, 48), False)
    assert len(prog.bbox_jobs) == 1
    with pytest.raises(ValueError, match="Singular matrix"):
        eng.render(prog)
