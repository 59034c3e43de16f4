#!/usr/bin/env python
"""BASELINE.json's fifth configuration at its full size: 100 000 synthetic icon SVGs (seed = icon index) at
256 x 256, encoded on the host cores (worker processes, native scene encoder) and rendered through the public call
(Engine.render_png, host buffers in and out) in batches of 2048.  The workers build their Scene objects first -- the
stand-in for the reference's SVG parser, which is not part of the hot path -- and the clock runs from Scene objects
to PNG files.  Prints one JSON line: wall times, Mpx/s with and without the host encoding, a 64-bit sum of all result bytes, and a check that does not depend on the size of the run -- sampled icons of the big run are
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
SUB = 2048  # icons per encoding job = one render call: nothing is concatenated in the GPU process


def worker(w, n_workers, spans, ready, go, out_q):
    """Owns spans w, w + n_workers, ...: first generates their scenes (the stand-in for parsing, not part of the hot
    path), reports, waits for the start signal, then encodes span after span with the native walk."""
    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200 import native, synth

    mine = list(range(w, len(spans), n_workers))
    t0 = time.perf_counter()
    scenes = {k: [(synth.icon_scene(i), synth.icon_size(), False) for i in range(*spans[k])] for k in mine}
    ready.put((w, time.perf_counter() - t0))
    go.wait()
    for k in mine:
        t1 = time.perf_counter()
        prog = native.encode_batch(scenes.pop(k))
        out = prog.to_program() if isinstance(prog, native.NativeProgram) else prog
        out.t_encode = time.perf_counter() - t1
        out_q.put((k, out))


def main():
    import io
    import queue
    import threading

    import torch

    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200 import encode, synth
    from svgrasterize_b200.engine import Engine

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    procs = max(1, (os.cpu_count() or 2) - 2)
    n_render = 3  # render threads (one context + stream each): a call's copies and planning overlap another's kernels
    px = synth.icon_size()[0] * synth.icon_size()[1]
    rng = np.random.default_rng(0)
    sample = sorted(set(int(v) for v in rng.integers(0, n, 48)))
    spans = [(lo, min(lo + batch, n)) for lo in range(0, n, batch)]
    ctx = mp.get_context("spawn")
    ready, go = ctx.Queue(), ctx.Event()
    queues = [ctx.Queue(maxsize=2) for _ in range(procs)]  # bounded: a worker runs at most 2 batches ahead of the GPU
    workers = [ctx.Process(target=worker, args=(w, procs, spans, ready, go, queues[w]), daemon=True) for w in range(procs)]
    t_start = time.perf_counter()
    for p in workers:
        p.start()
    engines = [Engine(0) for _ in range(n_render)]
    # PNG files: far below the raw bytes
    outs = [torch.empty(batch * px * 2, dtype=torch.uint8, pin_memory=True).numpy() for _ in range(n_render)]
    t_gen = sum(ready.get()[1] for _ in workers)
    wall_generate = time.perf_counter() - t_start

    # a thread of this process takes the encoded programs off the workers' queues in order (unpickling holds the GIL,
    # the render calls do not) and deals them to the render threads; every render thread checksums and samples its own
    # results (numpy's sum releases the GIL too)
    arrived = [queue.Queue(maxsize=2) for _ in range(n_render)]

    def fetch():
        for k in range(len(spans)):  # batch k comes from worker k mod procs and goes to render thread k mod n_render
            kk, prog = queues[k % procs].get()
            assert kk == k
            arrived[k % n_render].put((k, prog))
        for q in arrived:
            q.put(None)

    stats = [dict(t_render=0.0, t_enc=0.0, done=0, png_bytes=0, sum64=0, kept={}, err=None) for _ in range(n_render)]

    def render(r):
        st, eng, out = stats[r], engines[r], outs[r]
        try:
            while True:
                item = arrived[r].get()
                if item is None:
                    return
                k, prog = item
                first = spans[k][0]
                t1 = time.perf_counter()
                res = eng.render_png(prog, out=out)  # the result of the job: PNG files in host memory
                st["t_render"] += time.perf_counter() - t1
                off = res["offsets"]
                nbytes, n8 = int(off[-1]), int(off[-1]) // 8
                st["sum64"] += int(out[: 8 * n8].view(np.uint64).sum()) + int(out[8 * n8: nbytes].sum())
                st["png_bytes"] += nbytes
                st["t_enc"] += prog.t_encode
                cnt = len(prog.canvases)
                for i in sample:
                    if first <= i < first + cnt:
                        st["kept"][i] = out[off[i - first]: off[i - first + 1]].tobytes()
                st["done"] += cnt
        except Exception as exc:  # noqa: BLE001
            st["err"] = exc

    threads = [threading.Thread(target=fetch, daemon=True)] + [threading.Thread(target=render, args=(r,)) for r in range(n_render)]
    t0 = time.perf_counter()
    go.set()
    for t in threads:
        t.start()
    for t in threads[1:]:
        t.join()
    wall = time.perf_counter() - t0
    for st in stats:
        if st["err"] is not None:
            raise st["err"]
    for p in workers:
        p.join(timeout=30)
    done = sum(st["done"] for st in stats)
    t_render = max(st["t_render"] for st in stats)
    t_enc = sum(st["t_enc"] for st in stats)
    png_bytes = sum(st["png_bytes"] for st in stats)
    crc = sum(st["sum64"] for st in stats) & 0xFFFFFFFFFFFFFFFF
    kept = {}
    for st in stats:
        kept.update(st["kept"])
    eng = engines[0]
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
        "host_processes": procs,
        "wall_s": round(wall, 3), "what_wall_covers": "Scene objects in the workers' memory -> native encode -> "
        "programs to the GPU process -> svgr_render_png -> PNG files in pinned host memory",
        "render_threads": n_render, "render_png_s_busiest_thread": round(t_render, 3),
        "wall_s_generating_the_synthetic_scenes_before": round(wall_generate, 3),
        "worker_seconds_generating_scenes": round(t_gen, 2), "worker_seconds_encoding": round(t_enc, 2),
        "mpx_s_wall": round(done * px / wall / 1e6, 1),
        "png_bytes": png_bytes, "sum64_of_all_png_bytes": crc,
        "sampled_icons_equal_to_single_renders": f"{same}/{len(kept)} identical, {close}/{len(kept)} within 1 LSB"}))
    assert done == n and (Image is None or close == len(kept))


if __name__ == "__main__":
    main()
