"""A/B/C of the exchange paths of the multi-GPU driver in ONE process: ncclSend/ncclRecv (ncclCommInitAll), direct peer copies
(cudaMemcpyPeerAsync over NVLink) and peer STORES (the packing kernels write the neighbours' inboxes themselves: no copy,
only the live records travel), same scene, same steps.  Prints one JSON line with ms/step and the phase
profile of every slab for both.  `python bench_exchange.py --gpus 2 --steps 50`."""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np


def run(mode: str, gpus: int, steps: int, warmup: int, title: str) -> dict:
    import torch
    from pibiti_b200 import host, lib
    os.environ["SPH_B200_MULTI_XCHG"] = mode
    s = host.CSph(device=-1)
    s.select_scene(title)
    pos, vel = s.host_arrays()
    n = s.n
    m = lib.MultiSystem(s.params, capacity_per_slab=int(n / gpus * 1.25) + 600000, devices=list(range(gpus)))
    m.set_state(pos, vel)
    del pos, vel
    m.enable_phase_timing(True)

    def one():
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)

    for _ in range(warmup):
        one()
    m.sync()
    streams = [torch.cuda.ExternalStream(m.stream(k), device=k) for k in range(gpus)]
    ev = []
    for k in range(gpus):
        with torch.cuda.device(k):
            ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            ev[k][0].record(streams[k])
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    for k in range(gpus):
        with torch.cuda.device(k):
            ev[k][1].record(streams[k])
    m.sync()
    wall = time.perf_counter() - t0
    ms = max(ev[k][0].elapsed_time(ev[k][1]) for k in range(gpus)) / steps
    out = {"exchange": mode, "ms_per_step": round(ms, 4), "particle_updates_per_s": n / (ms * 1e-3), "wall_ms_per_step": round(wall / steps * 1e3, 4),
           "bytes_sent_per_step": m.info()["bytes_sent"] // (steps + warmup), "phase_ms_last_step": [m.phase_ms(k) for k in range(gpus)]}
    m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None)
    a = ap.parse_args()
    title = a.workload or {2: "wave tank 16M", 4: "wave tank 32M", 8: "wave tank 64M"}[a.gpus]
    res = [run(mode, a.gpus, a.steps, a.warmup, title) for mode in ("nccl", "copy", "peer")]
    print(json.dumps({"workload": title, "gpus": a.gpus, "steps": a.steps, "shape": "one process, one host thread", "runs": res}))


if __name__ == "__main__":
    main()
