#!/usr/bin/env python
"""BASELINE.json's fifth configuration at its full size: 100 000 synthetic icon SVGs (seed = icon index) at
256 x 256, encoded on the host cores and rendered through the public call (Engine.render, host buffers in and
out) in batches of 2048.  Prints one JSON line: wall times, Mpx/s with and without the host encoding, a CRC of
all result bytes, and a check that does not depend on the size of the run -- sampled icons of the big run are
byte-identical to the same icons rendered alone (whose parity with the oracle is what tests/ establish).

    python tools/run_c5_full.py [icons] [batch]
"""
import json
import multiprocessing as mp
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SUB = 128  # icons per encoding job


def encode_range(span):
    """Worker: generate the scenes (the stand-in for parsing) and encode them with the native walk."""
    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200 import native, synth

    lo, hi = span
    t0 = time.perf_counter()
    jobs = [(synth.icon_scene(i), synth.icon_size(), False) for i in range(lo, hi)]
    t1 = time.perf_counter()
    prog = native.encode_batch(jobs)
    out = prog.to_program() if isinstance(prog, native.NativeProgram) else prog
    out.t_generate, out.t_encode = t1 - t0, time.perf_counter() - t1
    return out


def main():
    import io

    import torch

    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200 import encode, synth
    from svgrasterize_b200.engine import Engine

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    procs = max(1, (os.cpu_count() or 2) - 2)
    px = synth.icon_size()[0] * synth.icon_size()[1]
    eng = Engine(0)
    out = torch.empty(batch * px * 2, dtype=torch.uint8, pin_memory=True).numpy()  # PNG files: far below the raw bytes
    rng = np.random.default_rng(0)
    sample = sorted(set(int(v) for v in rng.integers(0, n, 48)))
    kept = {}
    spans = [(lo, min(lo + SUB, n)) for lo in range(0, n, SUB)]
    crc, t_render, done, png_bytes = 0, 0.0, 0, 0
    t_gen = t_enc = 0.0
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        pending = []
        for sub in pool.imap(encode_range, spans, chunksize=1):  # results arrive in order, encoded ahead
            pending.append(sub)
            t_gen += sub.t_generate
            t_enc += sub.t_encode
            if len(pending) * SUB < batch and done + sum(len(p.canvases) for p in pending) < n:
                continue
            prog = encode.Program.concat(pending)
            pending = []
            k = len(prog.canvases)
            t1 = time.perf_counter()
            res = eng.render_png(prog, out=out)  # the result of the job: PNG files in host memory
            t_render += time.perf_counter() - t1
            off = res["offsets"]
            crc = zlib.crc32(out[: off[-1]], crc)
            png_bytes += int(off[-1])
            for i in sample:
                if done <= i < done + k:
                    kept[i] = out[off[i - done]: off[i - done + 1]].tobytes()
            done += k
    wall = time.perf_counter() - t0
    # size-independent check: sampled files of the big run decode to the same icons rendered alone
    try:
        from PIL import Image
    except ImportError:
        Image = None
    same = close = 0
    for i, png in kept.items():
        single = eng.render(encode.encode_scene(synth.icon_scene(i), synth.icon_size()))["canvas"][: px * 4]
        if Image is None:
            continue
        im = np.asarray(Image.open(io.BytesIO(png))).reshape(-1)
        d = int(np.abs(im.astype(np.int16) - single.astype(np.int16)).max())
        same += int(d == 0)
        close += int(d <= 1)
    print(json.dumps({
        "config": f"c5 full size: {done} synthetic icons at 256 x 256, batches of {batch}, result = PNG files", "icons": done,
        "host_processes": procs, "wall_s": round(wall, 3), "render_png_s": round(t_render, 3),
        "worker_seconds_generating_scenes": round(t_gen, 2), "worker_seconds_encoding": round(t_enc, 2),
        "mpx_s_wall": round(done * px / wall / 1e6, 1), "mpx_s_render_calls": round(done * px / t_render / 1e6, 1),
        "png_bytes": png_bytes, "crc32_of_all_png_bytes": crc,
        "sampled_icons_equal_to_single_renders": f"{same}/{len(kept)} identical, {close}/{len(kept)} within 1 LSB"}))
    assert done == n and (Image is None or close == len(kept))


if __name__ == "__main__":
    main()
