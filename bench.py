#!/usr/bin/env python
"""Benchmark of the SPH solver step (BASELINE.json metric: particle-updates/s per SPH step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full solver step (integrate+hash -> sort -> reorder -> density -> force) over every
particle of the workload.  N=1 runs BASELINE config 2, the 8M-particle drop-into-tank scene
("tank 8M drop" of scenes/Scenes.xml, one cSPH::Drop at step 0); under torchrun with N>1 ranks it runs
the z-slab-decomposed wave tank with 8M particles per GPU (weak scaling; 64M at N=8).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the dominant kernel (pair force) against the measured HBM peak, plus per-stage figures
  cpu_baseline  the CPU oracle timed on this box's host cores on a bounded sample (rank 0, N=1 only)
  e2e           the same metric through the cSPH-shaped host API with HOST buffers: every step copies
                positions+velocities host->device from pinned memory and reads them back

`--impl reference` times the reference's own implementation of the path on the host cores (the
host-compiled reference text in oracle/_ref when present, else the oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "particle-updates/sec per SPH step"
UNIT = "particle-updates/s"
# algorithmic bytes per particle per stage (SURVEY.md section 8d / BASELINE.md section 3)
STAGE_BYTES = {"integrate_hash": 72, "sort": 24, "reorder": 72, "density": 24, "force": 56}
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12            # non-tensor FP32, for context
PAIR_FLOP_PER_PARTICLE = 2700                                # SURVEY.md 8d estimate for density+force


def measured_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_for(n_gpus: int, name: str | None) -> str:
    if name:
        return name
    return "tank 8M drop" if n_gpus == 1 else {2: "wave tank 16M", 4: "wave tank 32M", 8: "wave tank 64M"}[n_gpus]


# --------------------------------------------------------------------------------------------------
def run_ours_single(args) -> dict:
    import torch
    from pibiti_b200 import host, lib

    torch.cuda.set_device(0)
    title = workload_for(1, args.workload)
    s = host.CSph(device=0)
    s.select_scene(title)
    if "drop" in title:
        s.Drop(False)                                   # deterministic centre position (SPH_Init.cpp:89-100)
    g = s.solver()
    n = g.n
    par = s.params
    stream = torch.cuda.ExternalStream(g.stream())
    wave = "wave" in title

    def one_step():
        if wave:                                        # host prologue advances the wave phase each step
            s.UpdateEmitter()
        s.Update()

    for _ in range(max(args.warmup, 3)):
        one_step()
    g.sync()

    # ---- device-resident timed region ----
    g.enable_timings(True)
    stage_ms = {k: 0.0 for k in lib.STAGE_NAMES}
    sampler = ClockSampler(0)
    sampler.start()
    l0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    stage_samples = []
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    g.sync()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    launches = g.launch_count() - l0
    clocks = sampler.stop()
    # per-stage device times: a second pass of the same K steps with a read-back after each step
    # (the read-back would perturb the timed region above, so it is kept out of it)
    for _ in range(args.steps):
        one_step()
        t = g.timings()
        for k in stage_ms:
            stage_ms[k] += t[k]
    for k in stage_ms:
        stage_ms[k] /= args.steps
    g.enable_timings(False)
    ms_per_step = ms_total / args.steps
    value = n * args.steps / (ms_total * 1e-3)

    # ---- end to end through the host API with HOST buffers ----
    hpos = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    hvel = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    hpos.numpy()[:] = s.getArray(False)
    hvel.numpy()[:] = s.getArray(True)
    L = lib.load()
    import ctypes as C
    ppos, pvel = C.c_void_p(hpos.data_ptr()), C.c_void_p(hvel.data_ptr())

    def e2e_step():
        # setArray(pos), setArray(vel) from pinned host memory; Update; getArray(pos), getArray(vel)
        assert L.sph_set_array(g.h, lib.SPH_POS, ppos, 0, n) == 0
        assert L.sph_set_array(g.h, lib.SPH_VEL, pvel, 0, n) == 0
        one_step()
        assert L.sph_get_array(g.h, lib.SPH_POS, ppos, 0, n) == 0
        assert L.sph_get_array(g.h, lib.SPH_VEL, pvel, 0, n) == 0

    e2e_steps = max(3, min(args.steps, 10))
    e2e_step()
    g.sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    g.sync()
    e2e_s = time.perf_counter() - t0
    e2e_value = n * e2e_steps / e2e_s
    assert np.isfinite(hvel.numpy()).all(), "non-finite velocities after the benchmark"

    hbm, hbm_src = measured_peaks()
    # the dominant kernel is whichever pair kernel took longer in this run (density since the force kernel
    # consumes neighbour lists)
    dom = "density" if stage_ms["density"] >= stage_ms["force"] else "force"
    dom_kernel = {"density": "k_density_l1 (density/pressure + neighbour-list build)",
                  "force": "k_force_l1 (pair force over neighbour lists)"}[dom]
    achieved = STAGE_BYTES[dom] * n / (stage_ms[dom] * 1e-3) / 1e9
    stages = {k: {"ms": round(stage_ms[k], 4), "algorithmic_GBps": round(STAGE_BYTES[k] * n / (stage_ms[k] * 1e-3) / 1e9, 1),
                  "hbm_frac": round(STAGE_BYTES[k] * n / (stage_ms[k] * 1e-3) / 1e9 / hbm, 4),
                  "dram_traffic_bytes_ncu": TRAFFIC_BYTES.get(k)} for k in stage_ms}
    pair_s = (stage_ms["density"] + stage_ms["force"]) * 1e-3
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": title, "particles": n, "grid": [int(x) for x in par["gridSize"][0]],
                   "scene_file": "scenes/Scenes.xml", "l2": "state (140 B/particle) is far larger than L2; no flush needed",
                   "timing": "CUDA events on the solver stream"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 32), "d2h_bytes_per_step": int(n * 32),
                "steps": e2e_steps, "api": "cSPH setArray(pos,vel) -> Update -> getArray(pos,vel), pinned host buffers"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": round(achieved, 1), "peak": hbm,
                     "unit": "GB/s", "frac": round(achieved / hbm, 4), "traffic": TRAFFIC_BYTES.get(dom),
                     "peak_source": hbm_src, "algorithmic_bytes_per_particle": STAGE_BYTES[dom],
                     "note": "density+force are FP32/shared-memory bound, not HBM bound (SURVEY.md D7); see fp32_frac",
                     "fp32_frac_density_force": round(PAIR_FLOP_PER_PARTICLE * n / pair_s / 1e12 / FP32_PEAK_TFLOPS, 4),
                     "binding_unit_ncu": NCU_BINDING.get(dom),
                     "stages": stages},
    }
    return out, s


# measured with `ncu --set full` on "tank 8M drop" (profiles/r01_*_l1.txt): dram__bytes_read.sum + dram__bytes_write.sum
# per launch.  Both pair kernels move ~3x their algorithmic bytes because of the neighbour lists (8.4M x ~27 x 4 B =
# 0.9 GB written by density, read by force) -- the price of a force kernel with a third of the instructions.
TRAFFIC_BYTES: dict = {"force": 1_363_952_864, "density": 1_573_461_464}
# what actually bounds the two pair kernels (same captures): percentages of the sustained peak of the unit
NCU_BINDING: dict = {
    "density": {"l1tex_throughput_pct": 90.2, "issue_active_pct": 77.4, "dram_throughput_pct": 25.2, "warps_active_pct": 94.1,
                "source": "profiles/r01_density_l1.txt"},
    "force": {"l1tex_throughput_pct": 93.8, "issue_active_pct": 64.0, "dram_throughput_pct": 27.5, "warps_active_pct": 47.0,
              "source": "profiles/r01_force_l1.txt"},
}


def cpu_baseline(title: str, threads: int | None = None, full_steps: int = 1) -> dict:
    """The CPU oracle on this box's host cores, on a bounded sample of the workload."""
    from oracle import oracle as orc                  # the one place bench.py executes oracle/
    from pibiti_b200 import host
    O = orc.load(None)
    O.set_threads(threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    s = host.CSph(device=-1)
    s.select_scene(title)
    if "drop" in title:
        s.Drop(False)
    par = s.params
    pos, vel = s.host_arrays()
    n = s.n
    o = O.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    t0 = time.perf_counter()
    o.step(full_steps)
    dt = time.perf_counter() - t0
    o.close()
    s.close()
    return {"value": n * full_steps / dt, "unit": UNIT, "cores": O.threads(), "kind": O.kind,
            "sample": f"{full_steps} step(s) of '{title}' ({n} particles), OpenMP over emulated thread blocks",
            "seconds": round(dt, 2)}


def run_reference(args) -> dict:
    """--impl reference: the reference's own CPU implementation of the path, all host threads."""
    n_gpus = args.gpus
    title = workload_for(n_gpus, args.workload)
    # bounded sample: the 1M-particle member of the same scene family, so K+W steps end within minutes
    sample_title = "Extreme box 1 M" if "tank 8M" in title else "wave tank 256k" if "wave" in title else title
    from oracle import oracle as orc
    from pibiti_b200 import host
    O = orc.load(None)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would serialise the baseline)
    O.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    s = host.CSph(device=-1)
    s.select_scene(sample_title)
    if "drop" in title:
        s.Drop(False)
    par = s.params
    pos, vel = s.host_arrays()
    n = s.n
    o = O.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    o.step(max(args.warmup, 1))
    t0 = time.perf_counter()
    o.step(args.steps)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{args.steps} steps of '{sample_title}' ({n} particles)"
    if sample_title != title:
        sample += f": the 1/8-scale member of the '{title}' scene family; particle-updates/s is intensive in N"
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": title, "sample_scene": sample_title, "particles": n},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.threads(), "kind": O.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--slab", action="store_true", help="N=1 through the slab path (orchestration overhead check)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return

    from pibiti_b200 import build
    build.build_cuda()                                  # no-op when the in-tree .so is current

    if world > 1 or args.gpus > 1 or args.slab:
        from pibiti_b200 import slab
        out = slab.bench_multi(args, METRIC, UNIT, STAGE_BYTES, measured_peaks(), ClockSampler)
        if rank == 0:
            print(json.dumps(out), flush=True)
        return

    out, s = run_ours_single(args)
    if not args.no_cpu_baseline:
        s.close()
        out["cpu_baseline"] = cpu_baseline(out["config"]["workload"])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
