"""Times the host planner on the c5 batch without a GPU (boxes from tools/dump_boxes.py)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import svgrasterize_b200  # noqa: E402,F401
from svgrasterize_b200 import _lib, encode, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
boxes = np.ascontiguousarray(np.load(f"gpurun_out/boxes_c5_{n}.npy").astype(np.int32))
prog = encode.Program.concat([encode.encode_scene(synth.icon_scene(i), synth.icon_size()) for i in range(n)])
cprog, keep = prog.to_c()
a, b = C.c_float(), C.c_float()
info = (C.c_int64 * 8)()
rc = _lib.lib().svgr_debug_plan(C.byref(cprog), boxes.ctypes.data, 20, C.byref(a), C.byref(b), info)
print("rc", rc, "masks ms", a.value, "nodes ms", b.value, "info", list(info))
