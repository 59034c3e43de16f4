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
    assert int(np.abs(a.astype(np.int16) - b.astype(np.int16)).max()) <= 1  # coverage float atomics, run to run
    res = eng.render_png(nat)
    assert len(res["offsets"]) == 65 and res["png_bytes"] > 0
    nat.close()


@pytest.mark.parametrize("name", ["demo_material_w1024", "demo_prompt", "icon_tiger", "icon_rust", "synth_icon_4",
                                  "feat_radial_focal_outside", "feat_stroke_caps_joins"])
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
