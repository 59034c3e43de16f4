"""Renders one synthetic icon (diagnostics under ncu): python tools/one_icon.py SEED"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svgrasterize_b200  # noqa: F401
from svgrasterize_b200 import encode, synth
from svgrasterize_b200.engine import Engine

eng = Engine(0)
prog = encode.encode_scene(synth.icon_scene(int(sys.argv[1])), synth.icon_size())
st = eng.render(prog)
st = eng.render(prog, timing=True)
print({k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, st["n_edges"])
