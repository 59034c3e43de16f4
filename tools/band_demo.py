"""Row-band render of ONE big SVG over all ranks: every rank renders its band of the
canvas, the RGBA8 bands are gathered to rank 0 over NCCL, and rank 0 checks the result against its own
full-canvas render.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/band_demo.py [size]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import svgrasterize_b200 as B  # noqa: E402
from conftest import load_golden  # noqa: E402
from svgrasterize_b200 import parallel, synth  # noqa: E402
from svgrasterize_b200.engine import Engine  # noqa: E402
from svgrasterize_b200.scene import Transform  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Engine(local)
    scene, size, lin, _z = load_golden("demo_material_w1024")
    jobs = [("material-design", scene.transform(Transform().scale(n / 1024.0)), (n, n), lin),
            ("filter stack", synth.filter_stack_scene(n), (n, n), False)]
    for name, sc, sz, lin in jobs:
        parallel.render_distributed(eng, sc, sz, lin, dist)  # warm-up (buffers, NCCL)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = parallel.render_distributed(eng, sc, sz, lin, dist)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            t1 = time.perf_counter()
            enc = B.encode.Encoder(eng)
            enc.add_scene(sc, sz, lin)
            prog = enc.finish()
            full = eng.canvas(prog, eng.render(prog)["canvas"])
            dt_full = time.perf_counter() - t1
            out = out.cpu().numpy()
            diff = int(np.abs(out.astype(int) - full.astype(int)).max())
            print(f"{name} {sz[0]}x{sz[1]} on {world} GPUs: bands + NCCL gather {dt * 1e3:.1f} ms "
                  f"(one GPU, full canvas, incl. encode: {dt_full * 1e3:.1f} ms), max |bands - full| = {diff} LSB",
                  flush=True)
            assert out.shape == full.shape and diff <= 1
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
