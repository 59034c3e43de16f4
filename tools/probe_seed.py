"""Diagnostic: which icon of a batch makes flatten_kernel slow (bisects on the stage time)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import svgrasterize_b200
from svgrasterize_b200 import encode, synth, _lib
from svgrasterize_b200.engine import Engine

lo, hi = int(sys.argv[1]), int(sys.argv[2])
eng = Engine(0)
progs = {i: encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(lo, hi)}

def t_flatten(a, b):
    prog = encode.Program.concat([progs[i] for i in range(a, b)])
    out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
    eng.render(prog, out=out)
    best = 1e9
    for _ in range(3):
        best = min(best, eng.render_resident(out, timing=True)["ms_flatten"])
    return best

print("all", t_flatten(lo, hi))
a, b = lo, hi
while b - a > 1:
    m = (a + b) // 2
    ta, tb = t_flatten(a, m), t_flatten(m, b)
    print(a, m, b, round(ta, 3), round(tb, 3))
    if ta > tb:
        b = m
    else:
        a = m
print("icon", a)
prog = progs[a]
eng.render(prog, stop=_lib.STOP_FLATTEN)
edges, edge_path = eng.edges()
cnt = np.bincount(edge_path, minlength=len(prog.paths))
print("edges per path", cnt.tolist())
k = int(cnt.argmax())
e = edges[edge_path == k]
print("path", k, "edges", len(e), "extent", e.min(0), e.max(0))
