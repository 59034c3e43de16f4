"""Diagnostic: device time of stroking + filling one long polyline / curve chain (tail behaviour of the
per-sub-path stroke assembly and the per-warp flatten)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import svgrasterize_b200 as B
from svgrasterize_b200 import encode, synth
from svgrasterize_b200.engine import Engine

eng = Engine(0)
for n in (100, 1000, 10000, 50000):
    rng = np.random.default_rng(1)
    pb = synth.PathBuilder().move_to(5, 128)
    xs = np.linspace(5, 250, n)
    ys = 128 + 100 * np.sin(xs / 7.0) * rng.uniform(0.5, 1.0, n)
    for k in range(1, n):
        if k % 3:
            pb.line_to(float(xs[k]), float(ys[k]))
        else:
            pb.quad_to(float(xs[k] - 0.1), float(ys[k] + 3), float(xs[k]), float(ys[k]))
    scene = B.Scene.stroke(pb.path(), synth.color(0.1, 0.2, 0.8), 1.5, "round", "round")
    prog = encode.encode_scene(scene, (256, 256))
    eng.render(prog)
    st = eng.render(prog, timing=True)
    print(n, {k: round(v, 3) for k, v in st.items() if k.startswith("ms_") and v > 0.002}, "edges", st["n_edges"], flush=True)
