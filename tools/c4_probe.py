import sys, os, json
sys.path.insert(0, os.getcwd())
import svgrasterize_b200
from svgrasterize_b200 import encode, synth
from svgrasterize_b200.engine import Engine
import torch
eng = Engine(0)
n = 8192
prog = encode.encode_scene(synth.filter_stack_scene(n), (n, n), False, engine=eng)
out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
eng.render(prog, out=out)
for _ in range(2):
    st = eng.render_resident(out, timing=True)
print({k: v for k, v in st.items() if k.startswith("ms_") or k in ("compose_bytes", "n_launches")})
