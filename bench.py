#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its workload: Mpixels/s rendered for a batch of synthetic
256 px icon SVGs (config 5: gradients, clip, mask, strokes), whole-job over N GPUs (weak scaling: every
rank renders its own `--icons` icons per step; whole SVGs are the shard unit, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--icons B] [--impl reference]

A step is one pass of the hot path (stroke outlines -> flatten -> bounds -> bin -> coverage ->
paint + compose -> quantise) over one batch.
  value  device timing, scene program resident in HBM, RGBA8 result left in HBM
  e2e    the same through the C-ABI call a user makes (svgr_render): host (pinned) program arrays in,
         host RGBA8 out, copies inside the timed region
  --impl reference   the CPU restatement of the reference's numpy path (oracle/, pinned bit-exactly to
         the unmodified reference by tests/test_oracle_pins.py) on all host cores, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpixels/sec rendered (icon batch)"
UNIT = "Mpx/s"
ICON_PX = 256


def workload_name(icons):
    return (f"c5: batch of {icons} synthetic icon SVGs per GPU per step at {ICON_PX}x{ICON_PX} px "
            "(12 masks, ~1040 edges, 3 strokes, 3 gradients, clip + luminance mask per icon; seed = icon index)")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference, one process per host core
# ---------------------------------------------------------------------------------------------
def _cpu_render(seed):
    import svgrasterize_b200  # noqa: F401
    from oracle import render as O
    from svgrasterize_b200 import synth

    img = O.render_canvas(synth.icon_scene(seed), synth.icon_size())
    return int(img.shape[0] * img.shape[1])


def cpu_throughput(n_icons, cores, seed0=0):
    """-> (Mpx/s, seconds) rendering n_icons with `cores` processes."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_render, range(cores))  # start-up + library load, untimed
        t0 = time.perf_counter()
        px = sum(pool.map(_cpu_render, range(seed0, seed0 + n_icons), chunksize=max(1, n_icons // (cores * 8))))
        dt = time.perf_counter() - t0
    return px / dt / 1e6, dt


def run_reference(opts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 64)
    n_icons = max(cores * 32, 256)
    vals = []
    for step in range(opts.warmup + opts.steps):
        v, dt = cpu_throughput(n_icons, cores, seed0=step * n_icons)
        if step >= opts.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": opts.gpus, "steps": opts.steps, "warmup": opts.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(opts.icons), "sample": f"{n_icons} icons per step on {cores} processes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_icons} icons of the workload per step, one process per core, oracle/ "
                                   "(C + numpy restatement pinned to the unmodified reference)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pin_program(prog):
    """Re-home the program's arrays in pinned host memory (what a serving process would keep them in)."""
    import torch

    keep = []
    for name in prog.ARRAYS:
        arr = np.ascontiguousarray(getattr(prog, name))
        if arr.nbytes == 0:
            continue
        t = torch.empty(arr.nbytes, dtype=torch.uint8, pin_memory=True)
        view = t.numpy().view(arr.dtype).reshape(arr.shape)
        view[...] = arr
        setattr(prog, name, view)
        keep.append(t)
    return keep


def build_batch(icons, seed0, engine):
    """-> (one program for the whole batch, its (scene, size, linear_rgb) jobs, seconds spent encoding).  The scenes
    come from the synthetic generator (the stand-in for the reference's SVG parser, untimed); the encoding is the
    native walk (svgrasterize_b200.native: csrc/_flatten.c + csrc/encode_flat.cpp)."""
    from svgrasterize_b200 import native, synth

    jobs = [(synth.icon_scene(seed0 + i), synth.icon_size(), False) for i in range(icons)]
    t0 = time.perf_counter()
    prog = native.encode_batch(jobs, engine=engine)
    dt = time.perf_counter() - t0
    if isinstance(prog, native.NativeProgram):
        prog = prog.to_program()  # plain arrays: they are re-homed in pinned memory below
    return prog, jobs, dt


def run_e2e(batches, out_host_np, device, steps, warmup, workers, chunks, png=False):
    """End to end through the public call (Engine.render -> svgr_render) with host buffers: step k renders batch
    k (mod the number of distinct batches), cut into `chunks` programs that `workers` host threads (one Engine =
    one context + stream each) render back to back, so that one chunk's device->host copy overlaps the next
    chunk's compute.  Returns wall seconds per step (torch.cuda.synchronize on both sides)."""
    import torch

    from svgrasterize_b200 import encode
    from svgrasterize_b200.engine import Engine

    from svgrasterize_b200 import native

    all_parts, pins = [], []
    for jobs in batches:
        n = len(jobs)
        bounds = [n * c // chunks for c in range(chunks + 1)]
        parts, offs = [], [0]
        for c in range(chunks):
            part = native.encode_batch(jobs[bounds[c]: bounds[c + 1]])
            if isinstance(part, native.NativeProgram):
                part = part.to_program()
            pins.append(pin_program(part))
            parts.append(part)
            offs.append(offs[-1] + part.canvas_bytes)
        all_parts.append((parts, offs))
    engines = [Engine(device) for _ in range(workers)]
    errors = []
    step_no = [0] * workers
    d2h = [0] * workers

    def work(w, reps):
        try:
            for _ in range(reps):
                parts, offs = all_parts[step_no[w] % len(all_parts)]
                step_no[w] += 1
                for c in range(w, chunks, workers):
                    if png:  # PNG files made on the device: only they cross PCIe
                        res = engines[w].render_png(parts[c], out=out_host_np[offs[c]: offs[c + 1]])
                        d2h[w] += int(res["png_bytes"])
                    else:
                        engines[w].render(parts[c], out=out_host_np[offs[c]: offs[c + 1]])
                        d2h[w] += parts[c].canvas_bytes
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    def run(reps):
        threads = [threading.Thread(target=work, args=(w, reps)) for w in range(workers)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(max(1, min(warmup, 2)))
    for w in range(workers):
        d2h[w] = 0
    dt = run(steps)
    for e in engines:
        e.close()
    if errors:
        raise errors[0]
    h2d = sum(p.h2d_bytes() for p in all_parts[0][0])
    del pins
    return dt / steps, h2d, sum(d2h) // steps


def run_from_scene(batches, out_host_np, device, steps, warmup, enc_threads=3, parts=4):
    """The same step starting from Scene objects in memory (what the reference's parser hands to Scene.render):
    native encode (svgrasterize_b200.native: flatten under the GIL, then the C++ walk, which runs without it) ->
    svgr_render_png -> PNG files in pinned host memory.  One process: a batch is cut into `parts` slices that
    `enc_threads` host threads encode (the C++ walk of one slice overlaps the flattening of the next) while the
    main thread renders the slices in order as they arrive.
    Returns (wall seconds per step, encoder thread-seconds per step)."""
    import queue
    from concurrent.futures import ThreadPoolExecutor

    import torch

    from svgrasterize_b200 import native
    from svgrasterize_b200.engine import Engine

    eng = Engine(device)
    t_enc = [0.0]
    lock = threading.Lock()

    def encode_part(jobs):
        t0 = time.perf_counter()
        prog = native.encode_batch(jobs)
        dt = time.perf_counter() - t0
        with lock:
            t_enc[0] += dt
        return prog

    def run(first, count):
        t_enc[0] = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=enc_threads) as pool:
            pending = queue.Queue()
            todo = []
            for k in range(first, first + count):
                jobs = batches[k % len(batches)]
                n = len(jobs)
                todo += [jobs[n * c // parts: n * (c + 1) // parts] for c in range(parts)]
            ahead = 2 * enc_threads  # slices in flight: bounds the memory of encoded-but-not-rendered programs
            it = iter(todo)
            for _ in range(ahead):
                j = next(it, None)
                if j is not None:
                    pending.put(pool.submit(encode_part, j))
            while not pending.empty():
                prog = pending.get().result()
                j = next(it, None)
                if j is not None:
                    pending.put(pool.submit(encode_part, j))
                eng.render_png(prog, out=out_host_np)
                if hasattr(prog, "close"):
                    prog.close()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(0, max(1, min(warmup, 2)))
    dt = run(0, steps)
    enc = t_enc[0]
    eng.close()
    return dt / steps, enc / steps


def time_other_configs(device, peak, with_cpu):
    """BASELINE.json's single-render configurations at their stated sizes (c1 and c3 at the sizes the reference's own
    goldens were recorded at), timed like the icon batch (resident
    program, CUDA events from svgr_render's stage timing, median of 5) with their own SURVEY 8(d) rooflines:
      c2  demo/material-design.svg -w 4096 (scene + reference bytes from tests/golden_big): coverage 4 B per mask
          pixel + 36 B per binned edge; compose 36 B per layer pixel + 20 B per canvas pixel
      c4  feGaussianBlur 4 -> feMorphology dilate 3 -> feColorMatrix saturate on 8192 x 8192: 64 B/px per separable
          stencil (two passes of 16 B in + 16 B out), 32 B/px colour matrix, 36 B/px fill, 20 B/px quantise
    The CPU figure beside them is the oracle (1 core, the reference is single-threaded) on c2 at 4096 and on c4 at
    2048 x 2048 scaled by 16 (the reference cannot hold 8192 x 8192 in float64)."""
    import torch

    from svgrasterize_b200 import encode, sceneio, synth
    from svgrasterize_b200.engine import Engine

    out = {}
    eng = Engine(device)

    def timed(prog, reps=5):
        buf = torch.empty(max(prog.canvas_bytes, 4), dtype=torch.uint8, device="cuda")
        eng.render(prog, out=buf)
        eng.render_resident(buf)
        rows = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            st = eng.render_resident(buf, timing=False, stream=torch.cuda.current_stream())
            e1.record()
            torch.cuda.synchronize()
            rows.append((e0.elapsed_time(e1), st))
        rows.sort(key=lambda r: r[0])
        ms = rows[len(rows) // 2][0]
        st = eng.render_resident(buf, timing=True, stream=torch.cuda.current_stream())
        return ms, st, buf

    def roof(st, ms_key, bytes_key):
        ms = st[ms_key]
        gbs = st[bytes_key] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"ms": ms, "algorithmic_bytes": int(st[bytes_key]), "achieved": gbs, "peak": peak, "unit": "GB/s",
                "frac": gbs / peak}

    # ---- c2
    path = os.path.join(ROOT, "tests", "golden_big", "demo_material_w4096.npz")
    if os.path.exists(path):
        z = np.load(path, allow_pickle=False)
        scene, size = sceneio.load_scene(z), tuple(float(v) for v in z["size"])
        t0 = time.perf_counter()
        prog = encode.encode_scene(scene, size, bool(z["linear_rgb"]), engine=eng)
        t_enc = time.perf_counter() - t0
        ms, st, buf = timed(prog)
        got = buf.cpu().numpy()[: prog.canvas_bytes].reshape(z["canvas_u8"].shape)
        out["c2"] = {"workload": "demo/material-design.svg -w 4096: 4096x4096 canvas, 1924 masks, one 935-layer group",
                     "ms_per_render": ms, "mpx_s": size[0] * size[1] / ms / 1e3, "masks": int(len(prog.paths)),
                     "edges": int(st["n_edges"]), "mask_px": int(st["mask_pixels"]), "host_encode_s": t_enc,
                     "max_lsb_vs_reference_bytes": int(np.abs(got.astype(np.int16) - z["canvas_u8"].astype(np.int16)).max()),
                     "stage_ms": {k[3:]: v for k, v in st.items() if k.startswith("ms_") and v > 0.0005},
                     "roofline": {"coverage_kernel": roof(st, "ms_coverage", "coverage_bytes"),
                                  "compose_kernel": roof(st, "ms_compose_busy", "compose_bytes_8d")}}
        del buf
    # ---- c1 / c3: the reference's own CPU-runnable cases (scene + reference bytes from tests/golden)
    from svgrasterize_b200 import native

    for key, fixture, what in (("c1", "demo_icons_w512", "demo/icons.svg -w 512: 991 masks, filters, strokes"),
                               ("c3", "demo_prompt", "demo/prompt.svg: 10 masks")):
        path = os.path.join(ROOT, "tests", "golden", fixture + ".npz")
        if not os.path.exists(path):
            continue
        z = np.load(path, allow_pickle=False)
        sc, sz = sceneio.load_scene(z), tuple(float(v) for v in z["size"])
        t0 = time.perf_counter()
        prog = native.encode_batch([(sc, sz, bool(z["linear_rgb"]))], engine=eng)
        t_enc = time.perf_counter() - t0
        ms, st, buf = timed(prog)
        got = eng.canvas(prog, buf.cpu().numpy())
        out[key] = {"workload": what, "canvas": f"{int(sz[0])}x{int(sz[1])}", "ms_per_render": ms,
                    "mpx_s": sz[0] * sz[1] / ms / 1e3, "masks": int(len(prog.paths)), "edges": int(st["n_edges"]),
                    "host_encode_s": t_enc, "native_encoder": isinstance(prog, native.NativeProgram),
                    "max_lsb_vs_reference_bytes": int(np.abs(got.astype(np.int16) - z["canvas_u8"].astype(np.int16)).max()),
                    "stage_ms": {k[3:]: v for k, v in st.items() if k.startswith("ms_") and v > 0.0005}}
        del buf
        if hasattr(prog, "close"):
            prog.close()
    # ---- c4
    n = 8192
    t0 = time.perf_counter()
    prog = encode.encode_scene(synth.filter_stack_scene(n), (n, n), False, engine=eng)
    t_enc = time.perf_counter() - t0
    ms, st, buf = timed(prog, reps=3)
    out["c4"] = {"workload": f"filter stack blur(4) -> dilate(3) -> saturate(0.5) on a gradient-filled circle, {n}x{n} canvas",
                 "ms_per_render": ms, "mpx_s": n * n / ms / 1e3, "host_encode_s": t_enc,
                 "layer_GiB": st["layer_floats"] * 4 / 2**30,
                 "stage_ms": {k[3:]: v for k, v in st.items() if k.startswith("ms_") and v > 0.0005},
                 "roofline": {"stencil_and_compose_kernels": roof(st, "ms_compose_busy", "compose_bytes_8d")}}
    del buf
    eng.close()
    if with_cpu:
        from oracle import render as O

        if "c2" in out:
            t0 = time.perf_counter()
            O.render_canvas(scene, size, bool(z["linear_rgb"]))
            dt = time.perf_counter() - t0
            out["c2"]["cpu_baseline"] = {"seconds": dt, "mpx_s": size[0] * size[1] / dt / 1e6, "cores": 1, "kind": "port",
                                         "sample": "the same render, oracle/ on one core (the reference is single-threaded)"}
        t0 = time.perf_counter()
        O.render_canvas(synth.filter_stack_scene(2048), (2048, 2048))
        dt = time.perf_counter() - t0
        out["c4"]["cpu_baseline"] = {"seconds_scaled": dt * 16, "mpx_s": 2048 * 2048 / dt / 1e6, "cores": 1, "kind": "port",
                                     "sample": f"2048x2048 in {dt:.1f} s on one core, x16 for 8192x8192 (float64 temporaries "
                                               "of the full size do not fit)"}
    return out


def run_gpu(opts):
    import torch

    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.current_stream()

    # ---- distinct seeds per step: step k of every leg renders batch k mod n_batches, batch b of rank r holds the
    # icons [(b * world + r) * icons, ... + icons).  One Engine (context) per batch keeps every program resident;
    # the number of distinct batches is bounded by their arenas (~7 GiB per 2048 icons).
    n_batches = max(1, min(opts.batches, opts.steps))
    engines, progs, batches, pins = [], [], [], []
    t_encode = 0.0
    for b in range(n_batches):
        e = Engine(local)
        prog_b, jobs_b, dt = build_batch(opts.icons, opts.seed0 + (b * world + rank) * opts.icons, e)
        t_encode += dt
        pins.append(pin_program(prog_b))
        engines.append(e), progs.append(prog_b), batches.append(jobs_b)
    t_encode /= n_batches
    eng, prog = engines[0], progs[0]
    n_px = opts.icons * ICON_PX * ICON_PX
    assert all(p.canvas_bytes == prog.canvas_bytes for p in progs)
    out_dev = torch.empty(prog.canvas_bytes, dtype=torch.uint8, device="cuda")
    out_host = torch.empty(prog.canvas_bytes, dtype=torch.uint8, pin_memory=True)
    out_host_np = out_host.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- first render of every batch: uploads its program, sizes every buffer; then W warm-up steps
    for e, p in zip(engines, progs):
        e.render(p, out=out_dev, stream=stream)
    for k in range(max(opts.warmup, n_batches)):
        engines[k % n_batches].render_resident(out_dev, stream=stream)

    # ---- value: resident program, device timing
    sampler = ClockSampler(local)
    acc = {}
    barrier()
    sampler.start()  # clocks are sampled during the timed region only
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(opts.steps):
        engines[k % n_batches].render_resident(out_dev, stream=stream)
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.stop_flag = True
    barrier()
    ms_local = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_local)
    ranks_ms = [ms_local / opts.steps]
    if dist is not None:  # diagnostic only: the spread over ranks behind the max
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = ms_local / opts.steps
        dist.all_reduce(t)
        ranks_ms = [round(float(v), 3) for v in t.tolist()]
    sampler.join()
    ms_step = ms_total / opts.steps
    value = world * n_px / (ms_step * 1e-3) / 1e6

    # ---- the same steps once more with per-stage CUDA events (these add ~2 % of synchronisation, which is why
    # they are not part of the timed region above): stage times and the roofline of the dominant kernel
    counts = {}
    for k in range(opts.steps):
        st = engines[k % n_batches].render_resident(out_dev, timing=True, stream=stream)
        for key, v in st.items():
            if key.startswith("ms_") or key.startswith("host_"):
                acc[key] = acc.get(key, 0.0) + v
            elif key.endswith("_bytes") or key.endswith("_bytes_8d") or key in ("mask_pixels", "n_edges", "n_kernels"):
                counts[key] = counts.get(key, 0) + v
    torch.cuda.synchronize()
    st = dict(st, **{key: v / opts.steps for key, v in counts.items()})  # per-step means over the distinct batches

    # ---- e2e: svgr_render with host buffers (H2D of the program, D2H of the RGBA8 result inside)
    barrier()
    # every host thread of the e2e leg needs a core to itself (it plans on the host between its launches)
    workers = max(1, min(opts.e2e_workers, (os.cpu_count() or 1) // max(world, 1) - 1))
    chunks = max(workers, opts.e2e_chunks if opts.e2e_chunks % workers == 0 else workers)
    for e in engines[1:]:
        e.close()  # their arenas are not needed any more
    sec_e2e, h2d_bytes, d2h_raw = run_e2e(batches, out_host_np, local, opts.steps, opts.warmup, workers, chunks)
    barrier()
    ms_e2e_raw = max_over_ranks(sec_e2e * 1e3)
    check = int(out_host_np[:: max(1, len(out_host_np) // 4096)].astype(np.int64).sum())
    # the same with the result delivered as PNG files encoded on the device (what the reference's main() writes)
    sec_png, _h2d, d2h_png = run_e2e(batches, out_host_np, local, opts.steps, opts.warmup, workers, chunks, png=True)
    barrier()
    ms_e2e_png = max_over_ranks(sec_png * 1e3)
    sec_scene, sec_scene_enc = run_from_scene(batches, out_host_np, local, opts.steps, opts.warmup)
    barrier()
    ms_from_scene = max_over_ranks(sec_scene * 1e3)
    use_png = ms_e2e_png < ms_e2e_raw
    ms_e2e = ms_e2e_png if use_png else ms_e2e_raw
    e2e_value = world * n_px / (ms_e2e * 1e-3) / 1e6

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (stage times are CUDA events on the launch stream)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    ms_cov = acc.get("ms_coverage", 0.0) / opts.steps
    ms_cmp = acc.get("ms_compose_busy", 0.0) / opts.steps  # device time of the compose launches themselves
    cov_gbs = st["coverage_bytes"] / (ms_cov * 1e-3) / 1e9 if ms_cov > 0 else 0.0
    # compose: SURVEY 8(d) counts one HBM pass per layer (36 B per layer pixel composited + 20 B per quantised
    # canvas pixel); the fold keeps the destination in registers, so the bytes it must move are fewer -- both
    # are reported, `achieved` is the 8(d) figure the contract asks for, `achieved_moved` the stricter one
    cmp_gbs = st["compose_bytes_8d"] / (ms_cmp * 1e-3) / 1e9 if ms_cmp > 0 else 0.0
    cmp_moved_gbs = st["compose_bytes"] / (ms_cmp * 1e-3) / 1e9 if ms_cmp > 0 else 0.0
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    def traffic_of(name):
        """Measured DRAM bytes per step of this kernel (ncu capture committed under profiles/), scaled to the
        batch size of this run when it differs from the captured one."""
        rec = traffic.get(name)
        if not isinstance(rec, dict) or not traffic.get("icons"):
            return None
        return rec["bytes_per_step"] * opts.icons / traffic["icons"]

    kernels = {
        "compose_kernel": {"bound": "hbm", "achieved": cmp_gbs, "peak": peak, "unit": "GB/s", "frac": cmp_gbs / peak,
                           "traffic": traffic_of("compose_kernel"), "ms_per_step": ms_cmp,
                           "launches_per_step": st["n_launches"] - 1,
                           "algorithmic_bytes_per_step": st["compose_bytes_8d"],
                           "achieved_moved": cmp_moved_gbs, "frac_moved": cmp_moved_gbs / peak,
                           "moved_bytes_per_step": st["compose_bytes"]},
        "coverage_kernel": {"bound": "hbm", "achieved": cov_gbs, "peak": peak, "unit": "GB/s", "frac": cov_gbs / peak,
                            "traffic": traffic_of("coverage_kernel"), "ms_per_step": ms_cov, "launches_per_step": 1,
                            "algorithmic_bytes_per_step": st["coverage_bytes"]},
    }
    dominant = "compose_kernel" if ms_cmp >= ms_cov else "coverage_kernel"
    roofline = dict(kernels[dominant], kernel=dominant, peak_source=peak_src,
                    note="achieved = algorithmic bytes per step (SURVEY 8(d): compose 36 B per layer pixel composited "
                         "+ 20 B per canvas pixel quantised; coverage 36 B per binned edge + 4 B per mask pixel) / "
                         "CUDA-event time of the kernel's launches in a step (first to last launch of every plan "
                         "chunk, on the launch stream); achieved_moved = the bytes the fused fold must actually move "
                         "(every source pixel read once, every output pixel written once) / the same time; traffic "
                         "= measured DRAM bytes per step (ncu, profiles/traffic.json)")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": opts.steps, "warmup": opts.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(opts.icons), "icons_per_gpu_per_step": opts.icons,
                   "distinct_batches": n_batches,
                   "seeds": f"step k renders batch k mod {n_batches}; batch b of rank r = icons "
                            f"[(b * {world} + r) * {opts.icons}, +{opts.icons}) -- {n_batches * world * opts.icons} distinct "
                            "icons per run, in all three legs (value, stage timing, e2e)",
                   "canvas_px_per_step_per_gpu": n_px, "mask_px_per_step_per_gpu": st["mask_pixels"],
                   "paths_per_step_per_gpu": int(len(prog.paths)), "edges_per_step_per_gpu": int(st["n_edges"]),
                   "l2": "inputs larger than L2 (coverage + layer arenas of "
                         f"{(st['cov_floats'] + st['layer_floats']) * 4 / 2**30:.1f} GiB per step)",
                   "parallelism": f"whole SVGs sharded over {world} GPU(s), no collective",
                   "arithmetic": "geometry f64, coverage/compose f32"},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_png if use_png else d2h_raw), "checksum": check,
                "result": ("PNG files encoded on the device (lossless: they decode to exactly the RGBA8 canvases, "
                           "tests/test_gpu_png.py)" if use_png else "raw RGBA8 canvases"),
                "how": ("Engine.render_png (svgr_render_png)" if use_png else "Engine.render (svgr_render)") +
                       f" on pinned host buffers; the batch goes through "
                       f"{chunks} calls on {workers} host threads (one context + stream each) "
                       "so that copies overlap compute; wall clock between device synchronisations",
                "raw_rgba8": {"value": world * n_px / (ms_e2e_raw * 1e-3) / 1e6, "ms_per_step": ms_e2e_raw,
                              "d2h_bytes_per_step": int(d2h_raw)},
                "png": {"value": world * n_px / (ms_e2e_png * 1e-3) / 1e6, "ms_per_step": ms_e2e_png,
                        "d2h_bytes_per_step": int(d2h_png)}},
        "e2e_from_scene": {"value": world * n_px / (ms_from_scene * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_from_scene,
                           "encode_ms_per_step": sec_scene_enc * 1e3,
                           "how": "Scene objects in memory -> native encoder (CPython flattener under the GIL + C++ "
                                  "walk without it) on 3 host threads, a batch in 4 slices -> svgr_render_png per slice "
                                  "-> PNG files in pinned host memory; encoding overlaps rendering (one process); "
                                  "encode_ms_per_step = encoder thread-time"},
        "gpu_launches": int(st["n_kernels"]) * opts.steps,
        "paths_per_s": world * len(prog.paths) / (ms_step * 1e-3),
        "stage_ms_per_step": {k: v / opts.steps for k, v in sorted(acc.items())},
        "ranks_ms_per_step": ranks_ms,
        "roofline": roofline,
        "roofline_other": {k: v for k, v in kernels.items() if k != dominant},
        "host_encode_s_per_batch": t_encode,
    }
    if world == 1 and not opts.no_configs:
        eng.close()
        line["configs"] = time_other_configs(local, peak, not opts.no_cpu)
    if world == 1 and not opts.no_cpu:
        cores = min(os.cpu_count() or 1, 32)
        n = max(cores * 48, 512)
        v, dt = cpu_throughput(n, cores, seed0=0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} icons of the same workload in {dt:.1f} s wall, one process per core, "
                                          "oracle/ (C + numpy restatement pinned to the unmodified reference)"}
    del pins
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_bands(opts):
    """--config c2 | c4: ONE big render cut into row bands over the N ranks (SURVEY.md 8(e), strong scaling): every
    rank keeps its band's program resident, a step = re-render the band + gather the RGBA8 bands on rank 0 (grouped
    ncclSend / ncclRecv, the only collective).  value = canvas Mpx/s, device events, max over ranks."""
    import torch

    import svgrasterize_b200  # noqa: F401
    from svgrasterize_b200 import parallel, sceneio, synth
    from svgrasterize_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    stream = torch.cuda.current_stream()
    eng = Engine(local)
    if opts.config == "c2":
        z = np.load(os.path.join(ROOT, "tests", "golden_big", "demo_material_w4096.npz"), allow_pickle=False)
        scene, size, lin = sceneio.load_scene(z), tuple(int(v) for v in z["size"]), bool(z["linear_rgb"])
        name = "c2: demo/material-design.svg -w 4096 (4096x4096, 1924 masks, one 935-layer group)"
    else:
        n = 8192
        scene, size, lin = synth.filter_stack_scene(n), (n, n), False
        name = f"c4: filter stack blur(4) -> dilate(3) -> saturate(0.5) on a gradient-filled circle, {n}x{n}"
    w, h = size
    prog, (a, b) = parallel.band_program(eng, scene, size, world, rank, lin)
    band = torch.empty(max(prog.canvas_bytes, 4), dtype=torch.uint8, device="cuda")
    eng.render(prog, out=band, stream=stream)
    view = band[: prog.canvas_bytes].view(b - a, w, 4)

    def step():
        st = eng.render_resident(band, stream=stream)
        return st, parallel.gather_bands(view, h, w, world, rank, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(opts.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(opts.steps):
        st, canvas = step()
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.stop_flag = True
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.join()
    if rank == 0:
        ms_step = ms / opts.steps
        check = None
        if opts.config == "c2":
            check = int(np.abs(canvas.cpu().numpy().astype(np.int16) - z["canvas_u8"].astype(np.int16)).max())
        print(json.dumps({
            "metric": "Mpixels/sec rendered (one canvas, row bands)", "value": w * h / (ms_step * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": opts.steps, "warmup": max(opts.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "bands": world, "rows_per_band": -(-h // world),
                       "collective": "grouped ncclSend / ncclRecv of RGBA8 bands to rank 0" if world > 1 else "none",
                       "l2": f"inputs larger than L2 (layer arena {st['layer_floats'] * 4 / 2**30:.2f} GiB on rank 0)"},
            "clocks": sampler.summary(), "gpu_launches": int(st["n_kernels"]) * opts.steps,
            "max_lsb_vs_reference_bytes": check}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--icons", type=int, default=2048, help="icons per GPU per step")
    ap.add_argument("--seed0", type=int, default=0, help="first icon seed (diagnostics: render another rank's batch)")
    ap.add_argument("--batches", type=int, default=8, help="distinct icon batches kept resident (steps cycle through them)")
    ap.add_argument("--no-configs", action="store_true", help="skip timing BASELINE.json's configs c2 / c4")
    ap.add_argument("--config", default="c5", choices=["c5", "c2", "c4"],
                    help="c5 (default): the icon batch, whole SVGs per rank; c2 / c4: one big render in row bands")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-workers", type=int, default=3, help="host threads (contexts) of the e2e leg")
    ap.add_argument("--e2e-chunks", type=int, default=3, help="svgr_render calls per step in the e2e leg")
    opts = ap.parse_args()
    if opts.impl == "reference":
        run_reference(opts)
    elif opts.config != "c5":
        run_bands(opts)
    else:
        run_gpu(opts)


if __name__ == "__main__":
    main()
