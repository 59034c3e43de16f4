"""The drop-in claim itself: install() rebinds the hot-path entry points on a reference-shaped
module, and the render part of the reference's main() (svgrasterize.py:3854-3881: Scene.render -> Layer.convert
-> canvas_merge_at -> Layer.background -> Layer.write_png) then runs on the CUDA core without being edited.
The PNGs it writes decode to the reference's own pixels (golden canvases, +-1 LSB)."""
import io
import struct
import zlib

import numpy as np
import pytest

import refshape
from conftest import load_golden

pytestmark = pytest.mark.gpu


def decode_png(png: bytes) -> np.ndarray:
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, size = 8, b"", None
    while pos < len(png):
        n = struct.unpack(">I", png[pos:pos + 4])[0]
        tag, data = png[pos + 4:pos + 8], png[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + data) & 0xFFFFFFFF == struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0]
        if tag == b"IHDR":
            size = struct.unpack(">2I", data[:8])
        if tag == b"IDAT":
            idat += data
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(size[1], 1 + 4 * size[0])
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(size[1], size[0], 4)


@pytest.fixture(scope="module")
def mod():
    import svgrasterize_b200 as B

    m = refshape.make_module()
    with pytest.raises(RuntimeError):  # nothing behind the module before the core is installed
        m.main_flow(m.Scene.fill(m.Path([]), np.ones(4)), (4, 4), False, None, io.BytesIO())
    token = B.install(m)
    yield m
    B.uninstall(token)
    with pytest.raises(RuntimeError):
        m.canvas_create(2, 2)


@pytest.mark.parametrize("name", ["demo_prompt", "demo_icons_w512", "icon_tiger", "synth_icon_3", "feat_pattern_user_space",
                                  "feat_filter_drop_shadow", "feat_mask_bbox_units", "feat_linear_interp_linear_rgb"])
def test_main_flow_writes_the_reference_png(mod, name):
    from conftest import golden_names

    if name not in golden_names():
        pytest.skip(f"no golden fixture called {name}")
    scene, size, lin, z = load_golden(name)
    out = io.BytesIO()
    assert mod.main_flow(refshape.rebuild(mod, scene), size, lin, None, out) == 0
    got = decode_png(out.getvalue())
    ref = z["canvas_u8"]
    assert got.shape == ref.shape
    assert int(np.abs(got.astype(np.int16) - ref.astype(np.int16)).max()) <= 1


def test_main_flow_background_and_no_size(mod):
    """opts.bg (:3876-3877) and the `-id` branch without a canvas size (:3864, layer written as rendered)."""
    from oracle import render as O
    from svgrasterize_b200 import synth

    scene, size = synth.icon_scene(11), synth.icon_size()
    bg = np.array([0.2, 0.3, 0.4, 1.0])
    out = io.BytesIO()
    assert mod.main_flow(refshape.rebuild(mod, scene), size, False, bg, out) == 0
    want = O.render_canvas(scene, size, bg=bg)
    assert int(np.abs(decode_png(out.getvalue()).astype(np.int16) - want.astype(np.int16)).max()) <= 1
    out = io.BytesIO()
    assert mod.main_flow(refshape.rebuild(mod, scene), None, False, None, out) == 0
    layer, _ = O.render(scene, O.canvas_transform())
    want = O.quantize(O.convert(layer, pre_alpha=False, linear_rgb=False).image)
    got = decode_png(out.getvalue())
    assert got.shape == want.shape and int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 1


def test_installed_methods_return_the_modules_own_types(mod):
    from svgrasterize_b200 import synth

    path = mod.Path(synth.rect_path(2, 3, 20, 10, 3).subpaths)
    outline = path.stroke(2.0, None, None)
    assert type(outline) is mod.Path
    layer, hull = path.fill(mod.Transform().matrix(0, 1, 0, 1, 0, 0), np.array([1.0, 0, 0, 1.0]))
    assert type(layer) is mod.Layer and layer.image.shape[2] == 4
    assert len(hull.bbox(mod.Transform())) == 4
