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
    (height, width, 4) canvas on `dst`, None elsewhere."""
    import torch

    step = -(-int(height) // int(world))
    if world == 1:
        return band[:height]
    buf = torch.zeros((step, width, 4), dtype=torch.uint8, device=band.device)
    buf[: band.shape[0]] = band
    out = torch.empty((world * step, width, 4), dtype=torch.uint8, device=band.device) if rank == dst else None
    if dist.get_backend() == "nccl":
        # NCCL has no gather primitive for unequal roots in older torch; all ranks contribute, dst keeps
        full = out if out is not None else torch.empty((world * step, width, 4), dtype=torch.uint8, device=band.device)
        dist.all_gather_into_tensor(full.view(-1), buf.view(-1))
        return full[:height] if rank == dst else None
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat(parts, dim=0)[:height]


def render_band(engine, scene, size, world: int, rank: int, linear_rgb: bool = False):
    """Render this rank's row band of `scene` (canvas size = (width, height)); returns (rows, width, 4) uint8."""
    from .encode import Encoder

    w, h = int(size[0]), int(size[1])
    a, b = band_rows(h, world, rank)
    probe = Encoder(engine)
    probe.add_scene(scene, size, linear_rgb)  # host-only pass to learn the filter reach
    halo = probe.filter_reach()
    enc = Encoder(engine)
    enc.add_scene_band(scene, size, (a, b), halo, linear_rgb)
    prog = enc.finish()
    res = engine.render(prog)
    return np.asarray(res["canvas"]).reshape(b - a, w, 4) if b > a else np.zeros((0, w, 4), np.uint8)


def render_distributed(engine, scene, size, linear_rgb: bool = False, dist=None):
    """Row-band render of one scene over all ranks of the default process group; the canvas lands on rank 0."""
    import torch

    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    band = torch.from_numpy(render_band(engine, scene, size, world, rank, linear_rgb))
    if dist is not None and dist.get_backend() == "nccl":
        band = band.cuda()
    out = gather_bands(band, int(size[1]), int(size[0]), world, rank, dist)
    return None if out is None else out.cpu().numpy()
