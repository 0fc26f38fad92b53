"""TEST INFRASTRUCTURE: one rank of a multi-process slab run (launched by tests/test_slab.py through
torch.distributed.run with the gloo backend).  argv: <oracle|gpu> <scene title> <steps> <out.npz>"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    from pibiti_b200 import host, slab

    kind, title, steps, out_path = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")

    s = host.CSph(device=-1)
    s.select_scene(title)
    par = s.params
    pos, vel = s.host_arrays()
    vel[:, 2] = 1.5 * np.sin(np.arange(vel.shape[0], dtype=np.float32) * np.float32(0.37)).astype(np.float32)   # == test_slab.stir
    cuts, parts = slab.split_initial_state(par, pos, vel, world)
    caps = slab.SlabCaps.for_state(par, pos, cuts)
    if kind == "oracle":
        from oracle import oracle as orc
        from slab_oracle import OracleSlabBackend
        be = OracleSlabBackend(orc.load("port"), par, cuts[rank], cuts[rank + 1], rank > 0, rank < world - 1, caps)
    else:
        cap = int(parts[rank].shape[0] * 1.5) + 4 * caps.rows
        be = slab.GpuSlabBackend(par, cap, cuts[rank], cuts[rank + 1], rank > 0, rank < world - 1, 0, caps)
    be.set_owned(parts[rank])
    comm = slab.DistComm(rank, world)
    for _ in range(steps):
        s.UpdateEmitter()
        be.set_params(s.params)
        slab.slab_step([be], comm)
    mine = be.get_owned()
    mine = mine.cpu().numpy() if hasattr(mine, "cpu") else np.asarray(mine)
    gathered = [None] * world
    dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        rec = slab.gather_by_id(gathered, s.n)
        np.savez(out_path, rec=rec, cuts=np.array(cuts), owned=np.array([g.shape[0] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
