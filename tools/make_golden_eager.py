#!/usr/bin/env python
"""Generate tests/golden_eager/eager.npz by calling the UNMODIFIED reference's small module-level functions
(the SURVEY.md 8(b) call surface outside Scene.render) on seeded random inputs.

Runs only in the build container (needs /root/reference).  Stored per case: the inputs and what the reference
returned (float64).  tests/test_oracle_pins.py pins oracle/ against these vectors on the CPU, and
tests/test_gpu_eager.py compares the CUDA entry points with them on the GPU box.

Usage: python tools/make_golden_eager.py
"""
import os
import sys
import warnings
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

import svgrasterize as R  # noqa: E402  (the reference)

OUT = os.path.join(ROOT, "tests", "golden_eager")


def premult(rng, shape):
    a = rng.uniform(0, 1, size=shape)
    a[..., :3] *= a[..., 3:]
    return a


def main():
    rng = np.random.default_rng(20261018)
    z = {}
    # ---- line_signed_coverage (:2213): lines in and around a 37 x 53 trace, incl. horizontal / vertical / outside
    lines = rng.uniform(-8, 60, size=(400, 2, 2))
    lines[:40, 0, 0] = lines[:40, 1, 0]  # horizontal in row space: skipped (:2232)
    lines[40:80, 0, 1] = lines[40:80, 1, 1]  # constant column
    lines[80:120] = np.round(lines[80:120])  # integer end points
    trace = np.zeros((37, 53))
    for ln in lines:
        R.line_signed_coverage(trace, ln)
    z["cov_lines"], z["cov_trace"] = lines, trace
    # ---- bezier3_flatten_batch (:2091) at other flatness values
    cubics = rng.uniform(0, 200, size=(60, 4, 2))
    z["flat_cubics"] = cubics
    for i, tol in enumerate((0.02, 0.1, 0.75, 3.0)):
        z[f"flat_tol_{i}"] = np.float64(tol)
        out = R.bezier3_flatten_batch(cubics, tol).reshape(-1, 4)
        z[f"flat_lines_{i}"] = out[np.lexsort(out.T[::-1])]
    # ---- grad_pixels / grad_spread / grad_interpolate (:1653-1683)
    z["gp_viewport"] = np.array([-3, 7, 11, 6])
    z["gp_out"] = R.grad_pixels((-3, 7, 11, 6))
    t = rng.uniform(-2.5, 3.5, size=(41, 29))
    t[0, :6] = [0.0, 1.0, -1.0, 2.0, 0.25, 0.5]
    z["gs_in"] = t
    for m in ("pad", "repeat", "reflect"):
        z[f"gs_{m}"] = R.grad_spread(t, m)
    stops = [(0.0, premult(rng, (4,))), (0.25, premult(rng, (4,))), (0.25, premult(rng, (4,))), (0.8, premult(rng, (4,))),
             (1.0, premult(rng, (4,)))]
    z["gi_stop_off"] = np.array([o for o, _ in stops])
    z["gi_stop_col"] = np.stack([c for _, c in stops])
    for lin in (0, 1):
        z[f"gi_out_{lin}"] = R.grad_interpolate(t, stops, bool(lin))
    # ---- pooling (:419): strides, padding, mean, NaN
    mat = rng.uniform(0, 1, size=(23, 31, 4))
    mat[5, 7, 1] = np.nan
    z["pool_in"] = mat
    cases = [((3, 2), (1, 1), "max", False), ((3, 2), None, "min", False), ((4, 5), (2, 3), "max", True),
             ((4, 5), (2, 3), "mean", True), ((2, 2), (2, 2), "mean", False), ((5, 2), (1, 2), "min", True)]
    z["pool_n"] = np.int64(len(cases))
    for i, (k, s, m, pad) in enumerate(cases):
        z[f"pool_{i}_k"] = np.array(k)
        z[f"pool_{i}_s"] = np.array(s if s is not None else (0, 0))
        z[f"pool_{i}_m"] = np.array(m)
        z[f"pool_{i}_pad"] = np.bool_(pad)
        z[f"pool_{i}_out"] = R.pooling(mat, k, s, m, pad)
    mat2 = rng.uniform(0, 1, size=(17, 9))
    z["pool2_in"], z["pool2_out"] = mat2, R.pooling(mat2, (3, 3), (2, 1), "max", False)
    # ---- canvas_compose (:277): every mode incl. arithmetic, RGBA and one-channel operands
    dst, src = premult(rng, (19, 27, 4)), premult(rng, (19, 27, 4))
    m1 = rng.uniform(0, 1, size=(19, 27, 1))
    z["cc_dst"], z["cc_src"], z["cc_m1"] = dst, src, m1
    modes = [0, 1, 2, 3, 4, (0.3, 0.5, -0.2, 0.1)]
    for i, mode in enumerate(modes):
        z[f"cc_out_{i}"] = R.canvas_compose(mode, dst, src)
    z["cc_arith"] = np.array(modes[5])
    z["cc_in_mask"] = R.canvas_compose(R.COMPOSE_IN, m1, src)  # clip: image x mask alpha
    z["cc_over_11"] = R.canvas_compose(R.COMPOSE_OVER, m1, m1[::-1].copy())
    # ---- canvas_merge_at (:304) with over / xor, overlay partly outside; untouched pixels stay unclipped
    base = premult(rng, (20, 30, 4)) * 1.5
    over = premult(rng, (12, 40, 4))
    z["ma_base"], z["ma_over"], z["ma_off"] = base, over, np.array([15, -4])
    for i, mode in enumerate((0, 4, 3)):
        b = base.copy()
        R.canvas_merge_at(b, over, (15, -4), partial(R.canvas_compose, mode))
        z[f"ma_out_{i}"] = b
    z["ma_miss"] = np.bool_(R.canvas_merge_at(base.copy(), over, (100, 100)) is None)
    # ---- canvas_merge_union (:330) full / fast path, canvas_merge_intersect (:382) with over / in / xor
    l3 = premult(rng, (9, 14, 4))
    layers = [(dst, (0, 0)), (over, (5, -3)), (l3, (-4, 20))]
    z["mu_l3"] = l3
    for i, (full, mode) in enumerate(((False, 0), (True, 0), (True, 4), (True, 1))):
        img, off = R.canvas_merge_union(layers, full, partial(R.canvas_compose, mode))
        z[f"mu_out_{i}"], z[f"mu_off_{i}"] = img, np.array(off)
    ilayers = [(m1, (0, 0)), (over, (5, -3)), (l3, (2, 6))]
    for i, mode in enumerate((0, 2, 4)):
        img, off = R.canvas_merge_intersect(ilayers, partial(R.canvas_compose, mode))
        z[f"mi_out_{i}"], z[f"mi_off_{i}"] = img, np.array(off)
    z["mi_empty"] = np.bool_(R.canvas_merge_intersect([(m1, (0, 0)), (over, (500, 500))]) is None)
    # ---- canvas_to_png quantisation (:263) of a float image
    img = rng.uniform(0, 1, size=(13, 21, 4)).astype(np.float32)
    img[0, :8, 0] = [0.5 / 255, 1.5 / 255, 2.5 / 255, 0.0, 1.0, 127.5 / 255, 128.5 / 255, 254.5 / 255]
    z["q_in"] = img
    z["q_out"] = np.round(img * 255.0).astype(np.uint8)
    z["q_png"] = np.frombuffer(R.canvas_to_png(img).getvalue(), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "eager.npz"), **z)
    print("wrote", os.path.join(OUT, "eager.npz"), len(z), "arrays")
    pathdata()


def pathdata():
    """Path.from_svg (:1252-1430) on real path data: every `d` attribute of the three demo files, every 16th glyph
    outline of fonts.svgz, hand-written strings for the corners of the grammar, and random arc commands (the
    conversion to the parametric form, :2397-2450).  Stored: the strings and the reference's segments, flattened."""
    import gzip
    import re

    sys.path.insert(0, ROOT)
    from svgrasterize_b200 import sceneio  # the flat segment layout (tags, 8 doubles, sub-path offsets)

    strings = []
    for f in ("icons.svg", "material-design.svg", "prompt.svg"):
        strings += re.findall(r'\sd="([^"]+)"', open(f"/root/reference/demo/{f}").read())
    strings += re.findall(r'\sd="([^"]+)"', gzip.open("/root/reference/fonts.svgz").read().decode())[::16]
    strings = sorted(set(s.replace("&#10;", " ") for s in strings))
    strings += [
        "", "M1 2", "1 2 3 M4 5 6 7z", "m1,2 3,4-5-6", "M0 0h10v10H0V0z", "M.5.5l1.5e1-2E-1,3.,-4", "M0 0 1 1 2 0zm5 5 1 1z",
        "M0 0C1 1 2 2 3 0S5 -1 6 0s1 1 2 0", "M0 0S1 1 2 0", "M0 0Q1 2 3 0T6 0t3 0", "M0 0T3 3", "M0 0L1 1zL2 2", "M0 0zz",
        "M10 10A5 5 0 0 1 20 10a5,3 30 1,0 10,0", "M0 0A0 5 0 0 1 9 9", "M0 0A5 5 0 1 1 0 0", "M0 0A1 1 45 0 0 100 100",
        "M 0,0 a 25,25 -30 0,1 50,-25 l 50,-25", "M0 0L+1-1L-.5+.5", "M0 0 L 1e2 1E+2 1.e-1 .1E1",
    ]
    rng = np.random.default_rng(77)
    for _ in range(300):
        rx, ry, rot = rng.uniform(0.1, 80), rng.uniform(0.1, 80), rng.uniform(-400, 400)
        strings.append("M%r %r%s%r %r %r %d %d %r %r" % (float(rng.uniform(-50, 50)), float(rng.uniform(-50, 50)),
                                                         "Aa"[int(rng.integers(0, 2))], float(rx), float(ry), float(rot),
                                                         int(rng.integers(0, 2)), int(rng.integers(0, 2)),
                                                         float(rng.uniform(-90, 90)), float(rng.uniform(-90, 90))))
    tags, data, sub_off, seg_end, sub_end = [], [], [], [], []
    n_seg = n_sub = 0
    for s in strings:
        t, d, o = sceneio.path_arrays(R.Path.from_svg(s))
        tags.append(t), data.append(d), sub_off.append(o[1:] if len(o) > 1 else np.zeros(0, np.int32))
        n_seg += len(t)
        n_sub += len(o) - 1
        seg_end.append(n_seg), sub_end.append(n_sub)
    blob = "\n".join(strings).encode()
    bad = []
    for s in ("M0 0L1", "M0 0Z1", "M0 0X", "M0 0h", "M", "M0 0C1 2 3 4 5", "M0 0 L 1 . 2"):
        try:
            R.Path.from_svg(s)
        except ValueError:
            bad.append(s)
    np.savez_compressed(os.path.join(OUT, "pathdata.npz"), strings=np.frombuffer(blob, np.uint8),
                        tags=np.concatenate(tags), data=np.concatenate(data).reshape(-1, 8),
                        sub_off=np.concatenate(sub_off).astype(np.int32), seg_end=np.asarray(seg_end, np.int64),
                        sub_end=np.asarray(sub_end, np.int64), invalid=np.frombuffer("\n".join(bad).encode(), np.uint8))
    print("wrote pathdata.npz:", len(strings), "strings,", n_seg, "segments,", len(bad), "invalid strings")


if __name__ == "__main__":
    main()
