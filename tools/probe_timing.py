import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import svgrasterize_b200
from svgrasterize_b200 import encode, synth
from svgrasterize_b200.engine import Engine
n = 2048
eng = Engine(0)
prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(n)])
out = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
eng.render(prog, out=out)
for timing in (True, False, True, False):
    for _ in range(3):
        eng.render_resident(out, timing=timing)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(10):
        st = eng.render_resident(out, timing=timing)
    e1.record()
    torch.cuda.synchronize()
    print("timing", timing, "ms/step events", e0.elapsed_time(e1) / 10, "wall", (time.perf_counter() - t0) * 100, {k: round(v, 3) for k, v in st.items() if k.startswith("ms_") or k.startswith("host")})
for ch in (1, 2, 4, 8, 16):
    os.environ["SVGR_CHUNKS"] = str(ch)
    for _ in range(3):
        eng.render_resident(out, timing=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        st = eng.render_resident(out, timing=False)
    torch.cuda.synchronize()
    print("chunks", ch, "wall ms/step", (time.perf_counter() - t0) * 100)
