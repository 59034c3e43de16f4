#!/usr/bin/env python
"""One-core wall time of the CPU oracle on the demo configurations tools/time_configs.py times on the GPU
(c1, c1', c2', c3): the "reference path on this box" column beside those rows.  Test infrastructure: it runs
the oracle, so it lives under tests/.

    python tests/time_oracle_configs.py
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from conftest import load_golden  # noqa: E402
from oracle import render as O  # noqa: E402

CASES = [("c1 demo/icons.svg -w 512", "demo_icons_w512"), ("c1' demo/icons.svg native", "demo_icons_native"),
         ("c2' demo/material-design.svg -w 1024", "demo_material_w1024"), ("c3 demo/prompt.svg", "demo_prompt")]

if __name__ == "__main__":
    for label, name in CASES:
        scene, size, lin, _z = load_golden(name)
        t0 = time.perf_counter()
        O.render_canvas(scene, size, lin)
        print(json.dumps({"config": label, "oracle_1core_s": round(time.perf_counter() - t0, 3)}), flush=True)
