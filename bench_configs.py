#!/usr/bin/env python
"""Measurement lines for the BASELINE configurations that are not the default bench line.

  python bench_configs.py config1 [--steps 1000] [--out profiles/config1_r02.json]
      BASELINE config 1: the scene the reference loads first ("box small default", reference Scenes.xml:172) and the
      shipped dam break ("Stiff  Dam break", Scenes.xml:364-368), headless, CPU (the host-compiled reference, all host
      cores) against one B200, fixed seed (neither scene uses rand()), K steps -- wall time of both, and the SURVEY.md 8d
      drift protocol: free-running drift at 100 steps (share of particles within spacing/2) and at K steps (centre of
      mass, kinetic energy, density distribution, maximum height within 1 %), plus a re-synchronised run that
      resets the GPU state to the oracle's every 50 steps and applies the one-step bar each time.

  python bench_configs.py config2 [--steps 100] [--out profiles/config2_r02.json]
      BASELINE config 2: "dam break 1M" and the shipped "Extreme box 1 M": bench.py's line (throughput, per-stage
      roofline table, parity check against the oracle, CPU baseline) for each.

The oracle is used as the checker and as the timed CPU baseline only.
"""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

REL = 1e-5


def one_step_bar(g, o, lib) -> dict:
    """integers bit for bit, density and velocity within REL (velocity relative to |v|max)"""
    do, dg = o.dump(5), g.dump(lib.DUMP_DENSITY)
    vo, vg = o.get_array(1), g.get_array(lib.SPH_VEL)
    vmax = max(float(np.abs(vo[:, :3]).max()), 1e-3)
    return {
        "pairs": bool(np.array_equal(g.dump(lib.DUMP_SORTED_PAIRS), o.dump(0))),
        "cell_start": bool(np.array_equal(g.dump(lib.DUMP_CELL_START), o.dump(1))),
        "neighbour_counts": bool(np.array_equal(g.dump(lib.DUMP_NEIGHBOR_COUNTS), o.dump(6))),
        "positions": bool(np.array_equal(g.get_array(lib.SPH_POS), o.get_array(0))),
        "density_rel": float(np.max(np.abs(dg - do) / np.maximum(np.abs(do), 1e-30))),
        "velocity_rel_vmax": float(np.abs(vg - vo).max() / vmax),
    }


def bar_ok(b: dict) -> bool:
    return b["pairs"] and b["cell_start"] and b["neighbour_counts"] and b["positions"] and b["density_rel"] <= REL \
        and b["velocity_rel_vmax"] <= REL


def aggregate(pos, vel, dens, par) -> dict:
    """the aggregate quantities of SURVEY.md 8d: centre of mass, kinetic energy, density distribution, maximum height"""
    wmin = np.asarray(par["worldMin"][0], np.float64)
    p = pos[:, :3].astype(np.float64)
    return {"com": p.mean(0).tolist(), "ke": float((vel[:, :3].astype(np.float64) ** 2).sum()),
            "max_height": float(p[:, 1].max() - wmin[1]),
            "density_mean": float(dens.astype(np.float64).mean()),
            "density_quantiles": np.quantile(dens.astype(np.float64), [0.05, 0.25, 0.5, 0.75, 0.95]).tolist()}


def rel_diff(a, b, scale=None) -> float:
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    s = float(np.abs(b).max()) if scale is None else scale
    return float(np.abs(a - b).max() / max(s, 1e-30))


def drift_protocol(title: str, steps: int = 1000, resync_every: int = 50, device: int = 0, oracle=None) -> dict:
    from oracle import oracle as orc                  # checker + timed CPU baseline
    from pibiti_b200 import host, lib
    O = oracle or orc.load(None)

    def fresh():
        s = host.CSph(device=device)
        s.select_scene(title)
        par = s.params
        pos, vel = s.host_arrays()
        o = O.system(par)
        o.set_array(0, pos)
        o.set_array(1, vel)
        return s, s.solver(), o, par

    out = {"scene": title, "steps": steps, "oracle": O.kind, "cpu_threads": O.threads()}

    # ---- free-running: K steps on both sides, no exchange of state ----
    s, g, o, par = fresh()
    n = g.n
    spacing = float(s.scene_extra()[8])
    extent = float(np.abs(np.asarray(par["worldSize"][0])).max())
    out.update({"particles": n, "spacing": spacing, "time_step": float(par["timeStep"][0])})
    marks = sorted({min(100, steps), steps})
    free, done = {}, 0
    t_gpu = t_cpu = 0.0
    for m in marks:
        k = m - done
        g.sync()
        t0 = time.perf_counter()
        s.Update(k)
        g.sync()
        t_gpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        o.step(k)
        t_cpu += time.perf_counter() - t0
        done = m
        pg, po = g.get_array(lib.SPH_POS), o.get_array(0)
        vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
        dg, do = g.dump(lib.DUMP_DENSITY), o.dump(5)
        d = np.linalg.norm(pg[:, :3].astype(np.float64) - po[:, :3].astype(np.float64), axis=1)
        ag, ao = aggregate(pg, vg, dg, par), aggregate(po, vo, do, par)
        free[str(m)] = {
            "share_within_half_spacing": float((d <= 0.5 * spacing).mean()),
            "displacement_q50_q99_max_over_spacing": [float(np.quantile(d, q) / spacing) for q in (0.5, 0.99, 1.0)],
            "com_diff_over_world": rel_diff(ag["com"], ao["com"], extent),
            "ke_rel_diff": abs(ag["ke"] - ao["ke"]) / max(ao["ke"], 1e-30),
            "max_height_rel_diff": abs(ag["max_height"] - ao["max_height"]) / max(ao["max_height"], 1e-30),
            "density_mean_rel_diff": abs(ag["density_mean"] - ao["density_mean"]) / ao["density_mean"],
            "density_quantiles_rel_diff": rel_diff(ag["density_quantiles"], ao["density_quantiles"]),
            "finite": bool(np.isfinite(pg).all() and np.isfinite(vg).all()),
        }
    out["free_running"] = free
    out["wall_s"] = {"gpu": round(t_gpu, 4), "cpu": round(t_cpu, 3), "cpu_over_gpu": round(t_cpu / max(t_gpu, 1e-9), 1),
                     "gpu_particle_updates_per_s": n * steps / t_gpu, "cpu_particle_updates_per_s": n * steps / t_cpu,
                     "what": "host wall clock around the K steps, state resident on each side, includes launch overhead"}
    o.close()
    s.close()

    # ---- re-synchronised: every `resync_every` steps the GPU state is reset to the oracle's and one step is compared ----
    s, g, o, par = fresh()
    worst = {"density_rel": 0.0, "velocity_rel_vmax": 0.0}
    checks = failures = 0
    for k in range(0, steps, resync_every):
        g.set_array(lib.SPH_POS, o.get_array(0))
        g.set_array(lib.SPH_VEL, o.get_array(1))
        g.step(1)
        o.step(1)
        b = one_step_bar(g, o, lib)
        checks += 1
        failures += 0 if bar_ok(b) else 1
        worst = {q: max(worst[q], b[q]) for q in worst}
        rest = min(resync_every, steps - k) - 1
        if rest > 0:
            o.step(rest)
    out["resync"] = {"every": resync_every, "checks": checks, "failures": failures, "worst": worst, "tolerance": REL}
    o.close()
    s.close()
    return out


def config1(args):
    res = [drift_protocol(t, args.steps) for t in ("box small default", "Stiff  Dam break")]
    doc = {"config": "BASELINE config 1: default scene + shipped dam break, headless, CPU vs 1 B200, K steps", "results": res}
    text = json.dumps(doc, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text + "\n")


def config2(args):
    lines = []
    for title in ("dam break 1M", "Extreme box 1 M"):
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", title, "--steps", str(args.steps), "--warmup", "10"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise SystemExit(f"bench.py failed on '{title}'")
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
        print(line, flush=True)
        lines.append(json.loads(line))
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(x) for x in lines) + "\n")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["config1", "config2"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 1000 if a.what == "config1" else 100
    {"config1": config1, "config2": config2}[a.what](a)
