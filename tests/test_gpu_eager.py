"""GPU parity of the module-level call surface outside Scene.render: line_signed_coverage,
bezier3_flatten_batch with a free flatness, grad_pixels / grad_spread / grad_interpolate, pooling with stride /
padding / mean, canvas_compose and the canvas_merge_* helpers with every blend, canvas_to_png's quantisation --
against vectors recorded from the unmodified reference (tests/golden_eager, tools/make_golden_eager.py)."""
import os
from functools import partial

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import svgrasterize_b200 as B

    return B


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(ROOT, "tests", "golden_eager", "eager.npz"), allow_pickle=False)


def test_line_signed_coverage(B, z):
    """svgrasterize.py:2213-2304: one line at a time (the reference's signature) and a batch; float32 trace."""
    trace = np.zeros(z["cov_trace"].shape, dtype=np.float32)
    for ln in z["cov_lines"][:50]:
        assert B.line_signed_coverage(trace, ln) is trace
    B.line_signed_coverage(trace, z["cov_lines"][50:])
    assert np.abs(trace - z["cov_trace"]).max() <= 1e-5
    t64 = np.zeros(z["cov_trace"].shape)  # a float64 canvas like the reference's is updated in place too
    B.line_signed_coverage(t64, z["cov_lines"])
    assert np.abs(t64 - z["cov_trace"]).max() <= 1e-5


def test_flatten_batch_with_other_flatness(B, z):
    """svgrasterize.py:2091-2098: the same set of lines, bit for bit, for every flatness."""
    for i in range(4):
        got = B.bezier3_flatten_batch(z["flat_cubics"], float(z[f"flat_tol_{i}"])).reshape(-1, 4)
        assert np.array_equal(got[np.lexsort(got.T[::-1])], z[f"flat_lines_{i}"])
    with pytest.raises(ValueError):
        B.bezier3_flatten_batch(z["flat_cubics"], 0.0)


def test_grad_functions(B, z):
    """svgrasterize.py:1653-1683: pixel centres and spreads are float64 and exact, colours float32."""
    assert np.array_equal(B.grad_pixels(tuple(int(v) for v in z["gp_viewport"])), z["gp_out"])
    for m in ("pad", "repeat", "reflect"):
        assert np.array_equal(B.grad_spread(z["gs_in"], m), z[f"gs_{m}"])
    with pytest.raises(ValueError):
        B.grad_spread(z["gs_in"], "mirror")
    stops = list(zip(z["gi_stop_off"], z["gi_stop_col"]))
    for lin in (0, 1):
        got = B.grad_interpolate(z["gs_in"], stops, bool(lin))
        assert got.shape == z[f"gi_out_{lin}"].shape and np.abs(got - z[f"gi_out_{lin}"]).max() <= 2e-6


def test_pooling_general(B, z):
    """svgrasterize.py:419-468: strides, NaN padding, mean, NaN-ignoring reductions, 2-D input."""
    for i in range(int(z["pool_n"])):
        s = tuple(int(v) for v in z[f"pool_{i}_s"])
        got = B.pooling(z["pool_in"], tuple(int(v) for v in z[f"pool_{i}_k"]), s if s != (0, 0) else None,
                        str(z[f"pool_{i}_m"]), bool(z[f"pool_{i}_pad"]))
        want = z[f"pool_{i}_out"]
        assert got.shape == want.shape
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) <= 1e-6
    got = B.pooling(z["pool2_in"], (3, 3), (2, 1), "max", False)
    assert got.shape == z["pool2_out"].shape and np.abs(got - z["pool2_out"]).max() <= 1e-7
    with pytest.raises(ValueError):
        B.pooling(z["pool_in"], (2, 2), method="median")


def test_canvas_compose_every_mode(B, z):
    """svgrasterize.py:277-298 on raw arrays: no colour conversion, arithmetic included."""
    modes = [0, 1, 2, 3, 4, tuple(float(v) for v in z["cc_arith"])]
    for i, mode in enumerate(modes):
        got = B.canvas_compose(mode, z["cc_dst"], z["cc_src"])
        assert np.abs(got - z[f"cc_out_{i}"]).max() <= 2e-6, mode
    got = B.canvas_compose(B.COMPOSE_IN, z["cc_m1"], z["cc_src"])
    assert np.abs(got - z["cc_in_mask"]).max() <= 2e-6
    got = B.canvas_compose(B.COMPOSE_OVER, z["cc_m1"], z["cc_m1"][::-1].copy())
    assert got.shape == z["cc_over_11"].shape and np.abs(got - z["cc_over_11"]).max() <= 2e-6
    with pytest.raises(ValueError):
        B.canvas_compose(7, z["cc_dst"], z["cc_src"])


def test_canvas_merge_helpers_every_blend(B, z):
    """svgrasterize.py:304-416 with blends other than their defaults, passed the way the reference passes them
    (functools.partial(canvas_compose, mode)) or as bare modes."""
    for i, mode in enumerate((0, 4, 3)):
        base = z["ma_base"].copy()  # float64 base, values above 1 outside the affected rectangle stay untouched
        out = B.canvas_merge_at(base, z["ma_over"], tuple(int(v) for v in z["ma_off"]), partial(B.canvas_compose, mode))
        assert out is base and np.abs(base - z[f"ma_out_{i}"]).max() <= 2e-6
    assert B.canvas_merge_at(z["ma_base"].copy(), z["ma_over"], (100, 100)) is None and bool(z["ma_miss"])
    layers = [(z["cc_dst"], (0, 0)), (z["ma_over"], (5, -3)), (z["mu_l3"], (-4, 20))]
    for i, (full, mode) in enumerate(((False, 0), (True, 0), (True, 4), (True, 1))):
        img, off = B.canvas_merge_union(layers, full, mode)
        assert tuple(off) == tuple(z[f"mu_off_{i}"]) and np.abs(img - z[f"mu_out_{i}"]).max() <= 2e-6
    with pytest.raises(ValueError):
        B.canvas_merge_union(layers, False, 4)
    ilayers = [(z["cc_m1"], (0, 0)), (z["ma_over"], (5, -3)), (z["mu_l3"], (2, 6))]
    for i, mode in enumerate((0, 2, 4)):
        img, off = B.canvas_merge_intersect(ilayers, partial(B.canvas_compose, mode))
        assert tuple(off) == tuple(z[f"mi_off_{i}"]) and np.abs(img - z[f"mi_out_{i}"]).max() <= 2e-6
    assert B.canvas_merge_intersect([(z["cc_m1"], (0, 0)), (z["ma_over"], (500, 500))]) is None
    with pytest.raises(TypeError):
        B.canvas_merge_at(z["ma_base"].copy(), z["ma_over"], (0, 0), lambda d, s: s)


def test_canvas_to_png_quantises_on_the_device(B, z):
    """svgrasterize.py:263: round-half-even of image x 255, then the reference's exact PNG bytes."""
    from svgrasterize_b200.engine import default_engine

    assert np.array_equal(default_engine().quantize_u8(z["q_in"]), z["q_out"])
    assert B.canvas_to_png(z["q_in"]).getvalue() == z["q_png"].tobytes()
