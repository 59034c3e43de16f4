"""Filter stencils beyond the sizes of the golden scenes: kernels longer than any shared-memory stage (tap groups
of the TMA-fed kernels, row groups of the direct 2-D convolution), large morphology windows, sources that are not
RGBA layers, and the TMA-fed path against the generic one.  Layer.convolve is scipy.signal.convolve in the reference
(svgrasterize.py:106-118), so scipy is the checker here; morphology goes against the oracle's pooling."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import svgrasterize_b200 as B

    return B


def _premult(rng, shape):
    a = rng.uniform(0, 1, size=shape).astype(np.float32)
    a[..., :3] *= a[..., 3:]
    return a


def _swap():
    from svgrasterize_b200.scene import Transform

    return Transform().matrix(0, 1, 0, 1, 0, 0)


@pytest.mark.parametrize("sigma", [(200.0, 200.0), (3.0, 150.0), (70.0, 0.2)])
def test_very_long_separable_blur(B, sigma):
    """stdDeviation x scale up to 200 px: 1001 taps per axis (svgrasterize.py:1903-1944 gives 2.5 sigma each way)."""
    from scipy.signal import convolve

    rng = np.random.default_rng(5)
    img = rng.uniform(0, 1, size=(83, 97, 4)).astype(np.float32)
    kernel = B.blur_kernel(_swap(), sigma)
    assert max(kernel.shape) >= 351
    layer = B.Layer(img, (13, -7), False, True)  # straight alpha, linear: convolve does not convert
    out = layer.convolve(kernel)
    want = convolve(img.astype(np.float64), kernel[..., None])
    assert out.image.shape == want.shape
    assert tuple(out.offset) == (int(13 - kernel.shape[0] / 2), int(-7 - kernel.shape[1] / 2))
    assert np.abs(out.image - want).max() <= 2e-6


def test_large_rotated_blur_is_a_direct_2d_convolution_in_row_groups(B):
    """An anisotropic Gaussian under a rotation is not separable: ~150 x 150 taps do not fit one staging buffer."""
    from scipy.signal import convolve
    from svgrasterize_b200.scene import Transform

    rng = np.random.default_rng(6)
    img = rng.uniform(0, 1, size=(61, 45, 4)).astype(np.float32)
    kernel = B.blur_kernel(_swap().rotate(0.5), (30.0, 4.0))
    assert (8 + kernel.shape[0] - 1) * (32 + kernel.shape[1] - 1) * 16 > 200 * 1024  # more than one staging buffer
    a, b = kernel.sum(axis=1), kernel.sum(axis=0)
    assert np.abs(np.outer(a, b) - kernel).max() > 1e-9  # really not separable
    out = B.Layer(img, (0, 0), False, True).convolve(kernel)
    want = convolve(img.astype(np.float64), kernel[..., None])
    assert out.image.shape == want.shape and np.abs(out.image - want).max() <= 2e-6


def test_large_morphology_windows(B):
    from oracle import render as O

    rng = np.random.default_rng(7)
    img = _premult(rng, (300, 411, 4))
    layer = B.Layer(img, (5, 9), True, True)
    ref = O.OLayer(img.astype(np.float64), (5, 9), True, True)
    for k0, k1, method in ((80, 3, "max"), (2, 131, "min"), (97, 260, "max")):
        got = layer.morphology(k0, k1, method)
        want = O.morphology(ref, k0, k1, method)
        assert got.image.shape == want.image.shape and tuple(got.offset) == tuple(want.offset)
        assert np.array_equal(got.image, want.image.astype(np.float32))


def test_long_blur_of_a_one_channel_layer(B):
    """A coverage-like one-channel source under a kernel that the generic stencil cannot stage: the planner writes
    the source out as RGBA and runs the TMA-fed passes (every channel carries the value)."""
    from scipy.signal import convolve

    rng = np.random.default_rng(8)
    img = rng.uniform(0, 1, size=(40, 52, 1)).astype(np.float32)
    kernel = B.blur_kernel(_swap(), (160.0, 1.0))
    out = B.Layer(img, (0, 0), True, True).convolve(kernel)
    want = convolve(img.astype(np.float64), kernel[..., None])
    assert out.image.shape[:2] == want.shape[:2]
    for ch in range(out.image.shape[2]):
        assert np.abs(out.image[..., ch] - want[..., 0]).max() <= 2e-6


def test_blurred_scene_with_a_long_kernel_matches_the_oracle_structure(B):
    """feGaussianBlur with stdDeviation 40 on an icon-sized scene (201 x 201 taps) through Scene.render: the same
    pixels as blurring the source graphic with scipy (the oracle's direct C convolution would take minutes)."""
    from oracle import render as O
    from scipy.signal import convolve
    from svgrasterize_b200 import scene as S, synth

    src = synth.icon_scene(2)
    tr, vp = _swap().scale(1.5), [0, 0, 200, 200]
    layer, _ = O.render(src, tr, viewport=vp)
    straight = O.convert(layer, pre_alpha=False, linear_rgb=True)
    kernel = B.blur_kernel(tr, (40.0, 40.0))
    want = convolve(straight.image, kernel[..., None])
    got, _hull = src.filter(S.Filter.empty().blur(40.0, 40.0)).render(tr, viewport=vp)
    assert got.image.shape == want.shape and (got.pre_alpha, got.linear_rgb) == (False, True)
    assert tuple(got.offset) == (int(straight.offset[0] - kernel.shape[0] / 2), int(straight.offset[1] - kernel.shape[1] / 2))
    assert np.abs(got.image - want).max() <= 5e-6


def test_tma_path_equals_the_generic_path():
    """The same filter scenes rendered by a process with SVGR_NO_TMA=1 (generic fetch_src stencils) and by the default
    TMA-fed kernels: equal to float32 summation order."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "import svgrasterize_b200 as B\n"
        "from svgrasterize_b200 import synth\n"
        "out = {}\n"
        "for n in (192, 700):\n"
        "    out[f's{n}'] = B.render_canvas(synth.filter_stack_scene(n), (n, n))\n"
        "np.savez(sys.argv[1], **out)\n" % ROOT)
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        files = []
        for tag, env in (("tma", {}), ("generic", {"SVGR_NO_TMA": "1"})):
            f = os.path.join(d, tag + ".npz")
            subprocess.run([sys.executable, "-c", code, f], check=True, env={**os.environ, **env})
            files.append(np.load(f))
        for key in files[0].files:
            diff = np.abs(files[0][key].astype(np.int16) - files[1][key].astype(np.int16))
            assert int(diff.max()) <= 1 and float((diff > 0).mean()) < 1e-3
