"""Renders the bench's default batch (c5, N icons) once from host arrays and then `steps` more times resident:
the program ncu is pointed at for per-step DRAM traffic (tools/ncu_summary.py traffic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import svgrasterize_b200  # noqa: E402,F401
from svgrasterize_b200 import encode, synth  # noqa: E402
from svgrasterize_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(0)
prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(n)])
out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
st = eng.render(prog, out=out)
print("kernels per render:", st["n_kernels"], flush=True)
for _ in range(steps):
    st = eng.render_resident(out, timing=True)
print({k: round(v, 3) for k, v in st.items() if k.startswith("ms_") or k.startswith("host_")}, st["compose_bytes"],
      st["coverage_bytes"])
