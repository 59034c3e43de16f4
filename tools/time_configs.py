#!/usr/bin/env python
"""Times BASELINE.json's other configurations (they are parity-test cases, not bench lines) on the GPU:
stage breakdown from svgr_render's CUDA events, resident re-render time, and checks against the golden bytes.

    python tools/time_configs.py [--filter-n 8192]

The CPU oracle's time on the same scenes is measured by tests/time_oracle_configs.py (only tests/, smoke() and
bench.py's CPU legs touch oracle/).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import svgrasterize_b200  # noqa: E402,F401
from conftest import load_golden  # noqa: E402
from svgrasterize_b200 import encode, synth  # noqa: E402
from svgrasterize_b200.engine import Engine  # noqa: E402


def time_program(eng, prog, reps=5):
    import torch

    out = torch.empty(max(prog.canvas_bytes, 4), dtype=torch.uint8, device="cuda")
    eng.render(prog, out=out)
    eng.render_resident(out)
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng.render_resident(out, timing=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if best is None or dt < best[0]:
            best = (dt, st)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filter-n", type=int, default=8192)
    opts = ap.parse_args()
    eng = Engine(0)
    rows = []
    cases = [("c1 demo/icons.svg -w 512", "demo_icons_w512"), ("c1' demo/icons.svg native", "demo_icons_native"),
             ("c2' demo/material-design.svg -w 1024", "demo_material_w1024"), ("c3 demo/prompt.svg", "demo_prompt")]
    for label, name in cases:
        scene, size, lin, z = load_golden(name)
        t0 = time.perf_counter()
        prog = encode.encode_scene(scene, size, lin, engine=eng)
        t_enc = time.perf_counter() - t0
        (ms, st), out = time_program(eng, prog)
        got = out.cpu().numpy()[: prog.canvas_bytes].reshape(int(size[1]), int(size[0]), 4)
        diff = int(np.abs(got.astype(int) - z["canvas_u8"].astype(int)).max())
        row = {"config": label, "canvas": f"{int(size[0])}x{int(size[1])}", "masks": len(prog.paths),
               "edges": st["n_edges"], "mask_px": st["mask_pixels"], "wall_ms": round(ms, 3),
               "mpx_s": round(size[0] * size[1] / ms / 1e3, 1), "encode_s": round(t_enc, 3),
               "max_lsb_vs_reference": diff,
               "stages_ms": {k[3:]: round(v, 3) for k, v in st.items() if k.startswith("ms_") and v > 0.0005}}
        rows.append(row)
        print(json.dumps(row), flush=True)
    n = opts.filter_n
    scene, size = synth.filter_stack_scene(n), (n, n)
    t0 = time.perf_counter()
    prog = encode.encode_scene(scene, size, False, engine=eng)
    t_enc = time.perf_counter() - t0
    (ms, st), out = time_program(eng, prog, reps=3)
    row = {"config": f"c4 filter stack blur 4 -> dilate 3 -> saturate, {n}x{n}", "canvas": f"{n}x{n}",
           "wall_ms": round(ms, 3), "mpx_s": round(n * n / ms / 1e3, 1), "encode_s": round(t_enc, 3),
           "compose_bytes": st["compose_bytes"], "layer_GiB": round(st["layer_floats"] * 4 / 2**30, 2),
           "stages_ms": {k[3:]: round(v, 3) for k, v in st.items() if k.startswith("ms_") and v > 0.0005}}
    if st["ms_compose"] > 0:
        row["compose_GBps"] = round(st["compose_bytes"] / st["ms_compose"] / 1e6, 1)
    rows.append(row)
    print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
