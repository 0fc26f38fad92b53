#!/usr/bin/env python
"""Neighbour-search microbenchmark (BASELINE config 4): uniform random particle boxes.

  python bench_sweep.py [--sizes 0.1,0.25,0.5,1,2,4,8,16,32] [--steps 5] [--out profiles/sweep.json]

For every (N, mean occupancy lambda, h/cell, maxParInCell) a scene is built THROUGH the Scene path (an XML
element handed to the host layer, so every derived constant is the reference's), filled with positions drawn
i.i.d. uniform in [worldMinD, worldMaxD] from numpy's PCG64 seeded 0x5EED5EED (not rand()), velocities 0,
gravity 0 and a 1e-7 s time step so that the configuration stays the drawn one.  Reported per point:
  build   particles/s through hash + histogram + scan + stable counting sort + reorder (stages 0-2)
  search  particles/s and candidate-visits/s through the density kernel (the 27-cell neighbour walk)
as device time from CUDA events (sph_get_timings), plus the mean neighbour count.  One JSON line per point.
tests/test_gpu_parity.py::test_uniform_random_boxes checks sorted pairs, cell table and neighbour counts of such
boxes bit-for-bit against the oracle.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 0x5EED5EED
CELL = 0.008


def scene_xml(n: int, lam: float, h_over_cell: float, max_par: int) -> str:
    """A cubic box whose grid holds about n/lam cells (n a multiple of 1024)."""
    cells_per_side = max(8, round((n / lam) ** (1.0 / 3.0)))
    side = cells_per_side * CELL
    return (f'<Scene name="uniform box" ParticlesK="{n // 1024}" World="{side:.6f} {side:.6f} {side:.6f}" '
            f'CellSize="{CELL}" particleH="{h_over_cell * CELL:.6f}" maxParInCell="{max_par}" '
            f'TimeStep="0.0000001" Gravity="0 0 0" InitType="1" />')


def uniform_positions(par, n: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    lo = np.asarray(par["worldMinD"][0], np.float64)
    hi = np.asarray(par["worldMaxD"][0], np.float64)
    pos = np.ones((n, 4), np.float32)
    pos[:, :3] = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    return pos


def build_system(n: int, lam: float, h_over_cell: float, max_par: int, device: int = 0):
    from pibiti_b200 import host, lib
    s = host.CSph(device=-1)
    idx = s.add_scene_xml(scene_xml(n, lam, h_over_cell, max_par))
    par = s.scene_params(idx)
    s.close()
    g = lib.SphSystem(par, device) if device >= 0 else None
    pos = uniform_positions(par, n)
    vel = np.zeros((n, 4), np.float32)
    if g is not None:
        g.set_array(lib.SPH_POS, pos)
        g.set_array(lib.SPH_VEL, vel)
    return g, par, pos, vel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="0.1,0.25,0.5,1,2,4,8,16,32", help="millions of particles")
    ap.add_argument("--lambdas", default="1,3,8,16")
    ap.add_argument("--ratios", default="1.0,1.25,1.5")
    ap.add_argument("--maxpar", default="16,64")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from pibiti_b200 import lib

    out_lines = []
    for size in (float(x) for x in args.sizes.split(",")):
        n = max(1024, int(size * 1e6) // 1024 * 1024)
        for lam in (float(x) for x in args.lambdas.split(",")):
            for ratio in (float(x) for x in args.ratios.split(",")):
                for max_par in (int(x) for x in args.maxpar.split(",")):
                    g, par, pos, vel = build_system(n, lam, ratio, max_par)
                    g.step(2)
                    g.enable_timings(True)
                    acc = {k: 0.0 for k in lib.STAGE_NAMES}
                    for _ in range(args.steps):
                        g.step(1)
                        t = g.timings()
                        for k in acc:
                            acc[k] += t[k] / args.steps
                    counts = g.dump(lib.DUMP_NEIGHBOR_COUNTS)
                    cs, ce = g.dump(lib.DUMP_CELL_START), g.dump(lib.DUMP_CELL_END)
                    occ = cs != 0xFFFFFFFF
                    cell_counts = (ce[occ] - cs[occ]).astype(np.int64)
                    build_ms = acc["integrate_hash"] + acc["sort"] + acc["reorder"]
                    # candidates visited per particle ~ 27 * (mean of min(count, maxPar) seen from a particle's cell)
                    visits = 27.0 * float(np.minimum(cell_counts, max_par).sum()) / max(int(par["numCells"][0]), 1) * 1.0
                    line = {"n": n, "lambda": lam, "h_over_cell": ratio, "maxParInCell": max_par,
                            "grid": [int(x) for x in par["gridSize"][0]], "mean_per_occupied_cell": round(float(cell_counts.mean()), 2),
                            "max_cell": int(cell_counts.max()), "mean_neighbours": round(float(counts.mean()), 2),
                            "build_ms": round(build_ms, 4), "build_particles_per_s": n / (build_ms * 1e-3),
                            "search_ms": round(acc["density"], 4), "search_particles_per_s": n / (acc["density"] * 1e-3),
                            "search_candidate_visits_per_s": n * visits / (acc["density"] * 1e-3),
                            "force_ms": round(acc["force"], 4), "seed": hex(SEED)}
                    print(json.dumps(line), flush=True)
                    out_lines.append(line)
                    g.close()
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in out_lines) + "\n")


if __name__ == "__main__":
    main()
