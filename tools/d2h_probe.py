"""Pinned device->host (and host->device) copy rate of this box for the bench's result sizes: the floor of the e2e
leg.  Alone: `python tools/d2h_probe.py`.  All GPUs of a box at once (the rate the ranks of a multi-GPU bench share):
`python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/d2h_probe.py` -- every rank copies
on its own GPU between two barriers, rank 0 prints the per-rank and aggregate rates as one JSON line."""
import json
import os

import torch

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

out = {}
for name, n in (("raw_rgba8_2048_icons", 2048 * 256 * 256 * 4), ("png_2048_icons", 107_000_000)):
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for direction in ("d2h", "h2d"):
        src, dst = (d, h) if direction == "d2h" else (h, d)
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            allms = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allms, t)
            allms = [float(x) for x in allms]
        else:
            allms = [ms]
        out[f"{name}_{direction}"] = {"bytes": n, "ms_per_rank": [round(x, 3) for x in allms],
                                     "GBps_per_rank": [round(n / x / 1e6, 1) for x in allms],
                                     "GBps_aggregate": round(world * n / max(allms) / 1e6, 1)}
if rank == 0:
    print(json.dumps({"n_gpus": world, **out}))
if world > 1:
    dist.destroy_process_group()
