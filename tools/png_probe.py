#!/usr/bin/env python
"""Times the device PNG encoder on a resident icon batch: python tools/png_probe.py [icons]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import svgrasterize_b200  # noqa: E402,F401
from svgrasterize_b200 import encode, synth  # noqa: E402
from svgrasterize_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = Engine(0)
prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(n)])
out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
eng.render_png(prog, out=out)
rows = []
for _ in range(5):
    st = eng.render_resident_png(out=out, timing=True)
    rows.append({k: round(v, 3) for k, v in st.items() if k in ("ms_png", "ms_total", "ms_compose", "ms_coverage")} | {"png_bytes": int(st["png_bytes"])})
print(json.dumps({"icons": n, "raw_bytes": prog.canvas_bytes, "runs": rows}))
