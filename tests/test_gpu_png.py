"""PNG files made on the device (csrc/k_png.cu): every file is a standard PNG (PIL and zlib
decode it, so the chunk CRCs, the Adler-32 and every deflate block are valid) that decodes to EXACTLY the RGBA8
canvas svgr_render returns; checked on icon batches, the demos, a multi-segment 4096-wide canvas, noise (the worst
case for the slot bound) and degenerate shapes."""
import io
import struct
import zlib

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def decode(png: bytes) -> np.ndarray:
    """Chunk walk + zlib + un-filter in numpy (no PIL needed); checks every CRC."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, size = 8, b"", None
    while pos < len(png):
        n = struct.unpack(">I", png[pos:pos + 4])[0]
        tag, data = png[pos + 4:pos + 8], png[pos + 8:pos + 8 + n]
        assert zlib.crc32(tag + data) & 0xFFFFFFFF == struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])[0], tag
        if tag == b"IHDR":
            size = struct.unpack(">2I", data[:8])
            assert data[8:] == bytes([8, 6, 0, 0, 0])
        if tag == b"IDAT":
            idat += data
        pos += 12 + n
    assert pos == len(png) and tag == b"IEND"
    w, h = size
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    out = np.zeros((h, w, 4), dtype=np.int32)
    for r in range(h):
        ft, line = int(raw[r, 0]), raw[r, 1:].reshape(w, 4).astype(np.int32)
        up = out[r - 1] if r else np.zeros((w, 4), np.int32)
        if ft == 0:
            out[r] = line
        elif ft == 2:
            out[r] = (line + up) & 255
        elif ft in (1, 4):
            left = np.zeros(4, np.int32)
            ul = np.zeros(4, np.int32)
            for c in range(w):
                if ft == 1:
                    pred = left
                else:
                    p = left + up[c] - ul
                    pa, pb, pc = np.abs(p - left), np.abs(p - up[c]), np.abs(p - ul)
                    pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up[c], ul))
                out[r, c] = (line[c] + pred) & 255
                left, ul = out[r, c], up[c]
        else:
            raise AssertionError(f"filter {ft}")
    return out.astype(np.uint8)


def decode_fast(png: bytes) -> np.ndarray:
    """PIL when present (C speed), else the numpy decoder above."""
    try:
        from PIL import Image
    except ImportError:
        return decode(png)
    im = Image.open(io.BytesIO(png))
    im.load()
    assert im.mode == "RGBA"
    return np.asarray(im)


@pytest.fixture(scope="module")
def eng():
    from svgrasterize_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def test_icon_batch_files_decode_to_the_canvases(eng):
    from svgrasterize_b200 import encode, synth

    n = 96
    prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(300 + i), synth.icon_size()) for i in range(n)])
    raw = eng.render(prog)["canvas"].copy()
    res = eng.render_png(prog, timing=True)
    off, png = res["offsets"], res["png"]
    assert len(off) == n + 1 and off[-1] == len(png) == res["png_bytes"]
    assert res["png_bytes"] < prog.canvas_bytes / 4  # flat vector art: well below a quarter of the raw bytes
    for i in range(n):
        got = decode_fast(png[off[i]: off[i + 1]].tobytes())
        assert np.array_equal(got, eng.canvas(prog, raw, i)), i
    # the pure-numpy decoder (explicit CRC / Adler / filter checks) on a few of them
    for i in (0, 17, n - 1):
        assert np.array_equal(decode(png[off[i]: off[i + 1]].tobytes()), eng.canvas(prog, raw, i))
    # resident re-encode gives the same bytes
    again = eng.render_resident_png()
    assert np.array_equal(again["png"], png)


@pytest.mark.parametrize("name", ["demo_prompt", "demo_icons_w512", "demo_icons_native", "demo_material_w1024"])
def test_demo_files(eng, name):
    import svgrasterize_b200 as B

    scene, size, lin, z = load_golden(name)
    png = B.render_png(scene, size, lin)
    got = decode_fast(png)
    assert got.shape == z["canvas_u8"].shape
    assert int(np.abs(got.astype(np.int16) - z["canvas_u8"].astype(np.int16)).max()) <= 1
    again = B.render_canvas(scene, size, lin)  # a second render: coverage is deterministic, the pixels are the same
    assert np.array_equal(got, again)


def test_encode_host_images_of_odd_shapes(eng):
    """Multi-segment canvases (4096 and 5000 columns: 16 and 20 work items per row), a single pixel, a single
    row / column, and uniform noise (no matches, 256 equally likely literals: the worst case)."""
    rng = np.random.default_rng(12)
    images = [
        rng.integers(0, 256, (37, 4096, 4), dtype=np.uint8),
        np.zeros((300, 5000, 4), dtype=np.uint8),
        np.full((1, 1, 4), 200, dtype=np.uint8),
        rng.integers(0, 256, (1, 700, 4), dtype=np.uint8),
        rng.integers(0, 2, (513, 1, 4), dtype=np.uint8) * 255,
        np.tile(np.arange(256, dtype=np.uint8)[None, :, None], (256, 1, 4)),
        rng.integers(0, 256, (256, 256, 4), dtype=np.uint8),
    ]
    images[1][100:200, 1000:3000] = (10, 200, 30, 255)
    files = eng.png_encode(images)
    assert len(files) == len(images)
    for im, f in zip(images, files):
        assert np.array_equal(decode_fast(f), im)
    assert len(files[1]) < 40000  # 6 MB of flat colour
    assert len(files[6]) < 256 * 256 * 4 * 1.02 + 400  # noise does not grow by more than the block overhead
    import svgrasterize_b200 as B

    out = B.canvas_to_png(images[5], device=True).getvalue()
    assert np.array_equal(decode(out), images[5])
