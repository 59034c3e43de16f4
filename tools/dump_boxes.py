"""Dumps the path boxes of the c5 bench batch (needed to profile the host planner on a CPU-only machine)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svgrasterize_b200  # noqa: E402,F401
from svgrasterize_b200 import _lib, encode, synth  # noqa: E402
from svgrasterize_b200.engine import Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
eng = Engine(0)
prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(n)])
eng.render(prog, stop=_lib.STOP_FLATTEN)
np.save(f"gpurun_out/boxes_c5_{n}.npy", eng.boxes())
print("saved", n)
