"""Multi-GPU partitioning: one process per GPU, torch.distributed for the plumbing.

The hot path has no exchange step, so there are exactly two shard shapes and one collective:

* batches      whole SVGs are independent: rank r renders scenes shard_range(n, world, r); nothing is
               communicated (bench.py runs this way, weak scaling);
* one huge render   the canvas is cut into row bands (the scanline prefix sum runs along columns, so bands
               are independent); filters get their halo by redundant compute, not by exchange; the
               finished RGBA8 bands are gathered to rank 0 -- over NCCL / NVLink when the tensors are on
               the GPU, over gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous, balanced [start, stop) of n independent items for `rank` of `world`."""
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def band_rows(height: int, world: int, rank: int):
    """Rows [a, b) of the canvas that `rank` renders; every band has ceil(height / world) rows except
    the tail, so that the gather can use equal-sized buffers."""
    step = -(-int(height) // int(world))
    a = min(height, rank * step)
    return a, min(height, a + step)


def gather_bands(band, height: int, width: int, world: int, rank: int, dist=None, dst: int = 0):
    """band: (rows, width, 4) uint8 tensor of this rank (CUDA with NCCL, CPU with gloo).  Returns the
    (height, width, 4) canvas on `dst`, None elsewhere.  The only collective of a row-band render (SURVEY.md
    8(e)): every other rank sends its band, `dst` receives each into its rows of the canvas -- one grouped
    send / receive (ncclGroupStart ... ncclSend / ncclRecv ... ncclGroupEnd under torch's batch_isend_irecv),
    nothing is broadcast and nobody but `dst` allocates the canvas."""
    import torch

    if world == 1:
        return band[:height]
    ops, out = [], None
    if rank == dst:
        out = torch.empty((height, width, 4), dtype=torch.uint8, device=band.device)
        a, b = band_rows(height, world, rank)
        out[a:b] = band
        for r in range(world):
            ra, rb = band_rows(height, world, r)
            if r != dst and rb > ra:
                ops.append(dist.P2POp(dist.irecv, out[ra:rb], r))
    elif band.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, band.contiguous(), dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def band_program(engine, scene, size, world: int, rank: int, linear_rgb: bool = False):
    """The scene program of this rank's row band (masks clipped to the band + the filters' reach, filter results
    placed where the whole-canvas render places them) and the band's rows [a, b)."""
    from .encode import Encoder

    h = int(size[1])
    a, b = band_rows(h, world, rank)
    probe = Encoder(engine)
    probe.add_scene(scene, size, linear_rgb)  # host-only pass to learn the filter reach
    halo = probe.filter_reach()
    enc = Encoder(engine)
    enc.add_scene_band(scene, size, (a, b), halo, linear_rgb)
    return enc.finish(), (a, b)


def render_band(engine, scene, size, world: int, rank: int, linear_rgb: bool = False, device_out: bool = False):
    """Render this rank's row band of `scene` (canvas size = (width, height)) -> (rows, width, 4) uint8: a numpy
    array, or with device_out a torch CUDA tensor that never left the GPU (what gather_bands sends)."""
    w = int(size[0])
    prog, (a, b) = band_program(engine, scene, size, world, rank, linear_rgb)
    if device_out:
        import torch

        out = torch.empty(max(prog.canvas_bytes, 4), dtype=torch.uint8, device=f"cuda:{engine.device}")
        engine.render(prog, out=out)
        return out[: prog.canvas_bytes].view(b - a, w, 4)
    res = engine.render(prog)
    return np.asarray(res["canvas"]).reshape(b - a, w, 4) if b > a else np.zeros((0, w, 4), np.uint8)


def render_distributed(engine, scene, size, linear_rgb: bool = False, dist=None):
    """Row-band render of one scene over all ranks of the default process group: every band is rendered into device
    memory and goes to rank 0 over NCCL from there; rank 0 returns the canvas (a CUDA tensor), the others None."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    band = render_band(engine, scene, size, world, rank, linear_rgb, device_out=True)
    return gather_bands(band, int(size[1]), int(size[0]), world, rank, dist)
