"""Worker of tests/test_multi.py::test_multi_driver_nccl_two_processes: one process per GPU, the C++ driver exchanges over
NCCL; torch.distributed only hands the NCCL unique id around and gathers the result."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent))


def main():
    import torch
    import torch.distributed as dist
    from pibiti_b200 import host, lib
    from test_multi import stir

    title, steps, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    recut = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ids = [lib.MultiSystem.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    s = host.CSph(device=-1)
    s.select_scene(title)
    pos, vel = s.host_arrays()
    m = lib.MultiSystem(s.params, capacity_per_slab=s.n, rank=rank, world=world, unique_id=ids[0], device=local)
    m.set_state(pos, stir(vel))
    m.set_recut_interval(recut)
    for _ in range(steps):
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)
    p, v, d, _, written = m.get_state(density=True)
    parts = [None] * world
    dist.all_gather_object(parts, (p, v, d, written))
    if rank == 0:
        P, V, D = parts[0][0].copy(), parts[0][1].copy(), parts[0][2].copy()
        for q in parts[1:]:
            mask = ~np.isnan(q[2])
            P[mask], V[mask], D[mask] = q[0][mask], q[1][mask], q[2][mask]
        assert sum(q[3] for q in parts) == s.n and not np.isnan(D).any()
        np.savez(out, pos=P, vel=V, dens=D, recuts=m.recut_count())
    m.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
