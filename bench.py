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
CELL_TABLE_BYTES = 16                                        # per cell, sort stage: histogram R+zero, scan W, bucket R


def stage_bytes(stage: str, n: int, cells: int) -> int:
    """Algorithmic bytes one launch of the stage moves."""
    return STAGE_BYTES[stage] * n + (CELL_TABLE_BYTES * cells if stage == "sort" else 0)


def kernel_source_hash() -> str:
    """Identifies the kernel sources an ncu capture belongs to (profiles/ncu_pair_kernels.json carries the same hash)."""
    import hashlib
    h = hashlib.sha256()
    for name in ("sph_pair_kernels.cu", "sph_stream_kernels.cu", "sph_device.cuh"):
        h.update((ROOT / "pibiti_b200" / "csrc" / name).read_bytes())
    return h.hexdigest()[:16]


def ncu_evidence(stage: str, workload: str) -> dict:
    """dram bytes per launch and the busiest unit of the stage's kernel, from the committed ncu summary of the SAME
    kernel sources and workload (profiles/summarize.py writes it).  A capture of other sources is reported as stale
    and contributes no number."""
    p = ROOT / "profiles" / "ncu_pair_kernels.json"
    if not p.exists():
        return {"traffic": None, "stale": True, "why": "no profiles/ncu_pair_kernels.json"}
    try:
        d = json.loads(p.read_text())
    except Exception as e:                                   # noqa: BLE001
        return {"traffic": None, "stale": True, "why": f"unreadable: {e}"}
    if d.get("source_hash") != kernel_source_hash():
        return {"traffic": None, "stale": True, "why": "kernel sources changed since the capture", "capture_hash": d.get("source_hash")}
    k = d.get("kernels", {}).get(stage)
    if not k or d.get("workload") != workload:
        return {"traffic": None, "stale": True, "why": f"no capture of {stage} on '{workload}' (have '{d.get('workload')}')"}
    return {"traffic": k.get("dram_bytes"), "stale": False, "kernel": k.get("kernel"), "bound_unit": k.get("bound_unit"),
            "units_pct": k.get("units_pct"), "l2_hit_pct": k.get("l2_hit_pct"), "l1_hit_pct": k.get("l1_hit_pct"),
            "source": d.get("source")}
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
    e2e_serial_s = time.perf_counter() - t0
    assert np.isfinite(hvel.numpy()).all(), "non-finite velocities after the benchmark"

    # The same traffic with the duplex accessor: every step still uploads its inputs (32 B/particle) from pinned host memory
    # and every step's result (32 B/particle) is still read back inside the timed region, but the download of step k's
    # result and the upload of step k+1's inputs are ONE call (cSPH::exchangeArrays) whose two directions overlap on the
    # link; the last result is fetched by a plain getArray pair after the loop, inside the timing.
    opos = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)
    ovel = torch.empty((n, 4), dtype=torch.float32, pin_memory=True)

    def e2e_duplex(k):
        assert L.sph_exchange_arrays(g.h, C.c_void_p(opos.data_ptr()), C.c_void_p(ovel.data_ptr()), ppos, pvel) == 0
        one_step()

    e2e_duplex(0)
    g.sync()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_duplex(k)
    assert L.sph_get_array(g.h, lib.SPH_POS, C.c_void_p(opos.data_ptr()), 0, n) == 0
    assert L.sph_get_array(g.h, lib.SPH_VEL, C.c_void_p(ovel.data_ptr()), 0, n) == 0
    g.sync()
    e2e_s = time.perf_counter() - t0
    e2e_value = n * e2e_steps / e2e_s
    assert np.isfinite(ovel.numpy()).all(), "non-finite velocities after the benchmark"

    hbm, hbm_src = measured_peaks()
    cells = int(par["numCells"][0])
    # the dominant kernel is whichever pair kernel took longer in this run
    dom = "density" if stage_ms["density"] >= stage_ms["force"] else "force"
    ev = {k: ncu_evidence(k, title) for k in stage_ms}
    gbps = {k: stage_bytes(k, n, cells) / (stage_ms[k] * 1e-3) / 1e9 for k in stage_ms}
    stages = {k: {"ms": round(stage_ms[k], 4), "algorithmic_bytes": stage_bytes(k, n, cells), "algorithmic_GBps": round(gbps[k], 1),
                  "hbm_frac": round(gbps[k] / hbm, 4), "dram_traffic_bytes_ncu": ev[k]["traffic"]} for k in stage_ms}
    pair_s = (stage_ms["density"] + stage_ms["force"]) * 1e-3
    dom_ev = ev[dom]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": title, "particles": n, "grid": [int(x) for x in par["gridSize"][0]],
                   "scene_file": "scenes/Scenes.xml", "l2": "state (140 B/particle) is far larger than L2; no flush needed",
                   "timing": "CUDA events on the solver stream", "pair_kernels": g.pair_variant()},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 32), "d2h_bytes_per_step": int(n * 32),
                "steps": e2e_steps,
                "api": "sph_exchange_arrays / cSPH::exchangeArrays (out <- result of the previous step, in <- this step's inputs; downloads "
                       "overlap uploads) -> Update, pinned host buffers; the last result by getArray(pos,vel) inside the timed region",
                "serial_value": n * e2e_steps / e2e_serial_s,
                "serial_api": "cSPH setArray(pos,vel) -> Update -> getArray(pos,vel): the four copies one after the other"},
        "gpu_launches": int(launches),
        # achieved / peak / frac are ALGORITHMIC HBM bytes over the measured copy bandwidth, as the contract asks;
        # `bound` names the unit ncu shows busiest for this kernel (the pair kernels are not HBM-bound)
        "roofline": {"bound": dom_ev.get("bound_unit") or "unknown (no current ncu capture)", "roofline_against": "hbm",
                     "kernel": dom_ev.get("kernel") or dom, "achieved": round(gbps[dom], 1), "peak": hbm,
                     "unit": "GB/s", "frac": round(gbps[dom] / hbm, 4), "traffic": dom_ev["traffic"],
                     "traffic_stale": dom_ev["stale"], "traffic_note": dom_ev.get("why") or dom_ev.get("source"),
                     "peak_source": hbm_src, "algorithmic_bytes_per_particle": STAGE_BYTES[dom],
                     "fp32_frac_density_force": round(PAIR_FLOP_PER_PARTICLE * n / pair_s / 1e12 / FP32_PEAK_TFLOPS, 4),
                     "units_pct_ncu": dom_ev.get("units_pct"), "l2_hit_pct_ncu": dom_ev.get("l2_hit_pct"),
                     "stages": stages},
    }
    return out, s


# --------------------------------------------------------------------------------------------------
def _multi_parity_check(dist, rank, world, local, uid_fn) -> dict:
    """Before any timing: 'wave tank 256k', particles stirred along z so that they migrate, ten steps through the SAME
    driver and the SAME NCCL exchange the timed run uses; the gathered result must equal a single-GPU run on rank 0 bit
    for bit (positions, velocities, densities)."""
    from pibiti_b200 import host, lib
    title, steps = "wave tank 256k", 10
    s = host.CSph(device=-1)
    s.select_scene(title)
    pos, vel = s.host_arrays()
    vel = vel.copy()
    vel[:, 2] = 1.5 * np.sin(np.arange(vel.shape[0], dtype=np.float32) * np.float32(0.37)).astype(np.float32)
    m = lib.MultiSystem(s.params, capacity_per_slab=s.n, rank=rank, world=world, unique_id=uid_fn(), device=local)
    m.set_state(pos, vel)
    owned0 = m.info()["owned"][0]
    for _ in range(steps):
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)
    p, v, d, _, written = m.get_state(density=True)
    owned1 = m.info()["owned"][0]
    m.close()
    parts = [None] * world
    dist.all_gather_object(parts, (p, v, d, written, owned0 != owned1))
    res = {"config": title, "steps": steps, "slabs": world, "exchange": "ncclSend/ncclRecv (C++ driver)", "ok": False}
    if rank == 0:
        P, V, D = parts[0][0], parts[0][1], parts[0][2]
        for q in parts[1:]:
            mk = ~np.isnan(q[2])
            P[mk], V[mk], D[mk] = q[0][mk], q[1][mk], q[2][mk]
        ref = host.CSph(device=local)
        ref.select_scene(title)
        g = ref.solver()
        g.set_array(lib.SPH_VEL, vel)
        for _ in range(steps):
            ref.UpdateEmitter()
            ref.Update()
        res.update({"particles_conserved": bool(sum(q[3] for q in parts) == s.n and not np.isnan(D).any()),
                    "migration_happened": bool(any(q[4] for q in parts)),
                    "positions_bit_exact": bool(np.array_equal(P, g.get_array(lib.SPH_POS))),
                    "velocities_bit_exact": bool(np.array_equal(V, g.get_array(lib.SPH_VEL))),
                    "densities_bit_exact": bool(np.array_equal(D, g.get_array(lib.SPH_DENSITY)))})
        res["ok"] = all(res[k] for k in ("particles_conserved", "positions_bit_exact", "velocities_bit_exact", "densities_bit_exact"))
        ref.close()
    return res


_CK_PRIMES = np.array([0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x27D4EB2F165667C5, 0x85EBCA77C2B2AE63,
                       0xD6E8FEB86659FD93, 0xFF51AFD7ED558CCD, 0xC4CEB9FE1A85EC53, 0x94D049BB133111EB], dtype=np.uint64)


def state_checksum(ids: np.ndarray, pos: np.ndarray, vel: np.ndarray, dens: np.ndarray) -> int:
    """Order-independent 64-bit checksum of {(original index, position, velocity, density)}: bit-equal particle sets give
    equal sums whichever slab holds which particle (a checksum of per-particle checksums, mod 2^64)."""
    words = np.concatenate([np.ascontiguousarray(pos, np.float32).view(np.uint32).reshape(-1, 4),
                            np.ascontiguousarray(vel, np.float32).view(np.uint32).reshape(-1, 4),
                            np.ascontiguousarray(dens, np.float32).view(np.uint32).reshape(-1, 1)], axis=1).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = (words * _CK_PRIMES[None, :]).sum(axis=1, dtype=np.uint64)
        return int((h * (ids.astype(np.uint64) * np.uint64(2) + np.uint64(1))).sum(dtype=np.uint64))


def run_ours_multi(args) -> dict | None:
    """bench.py under torchrun (one process per GPU): weak scaling, 8M particles per GPU, the wave tank cut into z slabs and
    stepped by the C++ multi-GPU driver (sph_multi_*: NCCL send/recv from C++, no torch.distributed on the data path --
    torch.distributed only hands the NCCL unique id around and reduces the timings).  Also serves `--slab` at N=1."""
    import torch
    import torch.distributed as dist
    from pibiti_b200 import host, lib

    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0"), ("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29533")):
        os.environ.setdefault(k, v)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()

    def unique_id():
        ids = [lib.MultiSystem.unique_id() if rank == 0 and world > 1 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    parity = _multi_parity_check(dist, rank, world, local, unique_id) if world > 1 else None

    title = args.workload or {1: "wave tank 8M", 2: "wave tank 16M", 4: "wave tank 32M", 8: "wave tank 64M"}[world]
    s = host.CSph(device=-1)                     # scene + initial lattice on the host (every rank builds the same one)
    s.select_scene(title)
    par = s.params
    pos, vel = s.host_arrays()
    n = s.n
    m = lib.MultiSystem(par, capacity_per_slab=int(n / world * 1.25) + 600000, rank=rank, world=world, unique_id=unique_id(), device=local)
    m.set_state(pos, vel)
    del pos, vel
    info0 = m.info()
    view = m.slab(0)
    view.enable_timings(True)                    # event records around the pair kernels: no synchronisation
    m.enable_phase_timing(True)                  # a dozen more event records per step, no synchronisation
    stream = torch.cuda.ExternalStream(m.stream(0))

    replay = []                                  # the parameter block of every step, for the full-size check below

    def one_step():
        s.UpdateEmitter()                        # wave phase: identical host arithmetic on every rank
        par_k = s.params
        replay.append(par_k.tobytes())
        m.set_params(par_k)
        m.step(1)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        one_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                          # spawns nvidia-smi (tens of ms): before the barrier, or the other ranks'
    m.sync()                                     # first exchange would wait for it inside their timed region
    dist.barrier()
    torch.cuda.synchronize()
    dist.barrier()
    l0 = view.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
    e1.record(stream)
    host_enqueue_s = time.perf_counter() - t0
    m.sync()                                     # also surfaces message overflow / lost particles
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.barrier()
    wall = time.perf_counter() - t0
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = view.launch_count() - l0
    kernel_ms = {k: v for k, v in view.timings().items() if v >= 0}        # last timed step, this rank
    phases = [None] * world
    dist.all_gather_object(phases, m.phase_ms(0))
    clocks = sampler.stop() if rank == 0 else None
    info1 = m.info()
    owned_timed = info1["owned"][0]
    owned = torch.tensor([owned_timed], device="cuda", dtype=torch.int64)
    counts = [torch.zeros_like(owned) for _ in range(world)]
    dist.all_gather(counts, owned)
    per_gpu = [int(c.item()) for c in counts]
    ms_total = float(ms.item())

    # ---- full-size parity: the state after warm-up + timed steps, as a checksum over all slabs, against ONE GPU running the
    # whole system through the same parameter blocks (the B200 holds 64M particles: ~27 GB).  Bit-exact or not at all.
    host_rec = torch.empty((m.capacity, 12), dtype=torch.float32, pin_memory=True)
    cnt = m.fetch_owned(0, host_rec.data_ptr(), m.capacity)
    rec = host_rec.numpy()[:cnt]
    ck = state_checksum(np.ascontiguousarray(rec[:, 8]).view(np.uint32), rec[:, 0:4], rec[:, 4:8], np.ascontiguousarray(rec[:, 9]))
    limbs = torch.tensor([ck & 0xFFFFFFFF, ck >> 32, cnt], device="cuda", dtype=torch.int64)
    dist.all_reduce(limbs)
    ck_multi = (int(limbs[0].item()) + (int(limbs[1].item()) << 32)) & 0xFFFFFFFFFFFFFFFF
    full_check = {"config": title, "particles": n, "steps": len(replay), "what": "checksum of (id, pos, vel, rho) over all slabs vs one "
                  "GPU stepping the whole system through the same parameter blocks", "ok": None}
    if rank == 0 and world > 1 and not os.environ.get("SPH_BENCH_SKIP_FULL_CHECK"):
        try:
            free_b, _ = torch.cuda.mem_get_info()
            need = n * 520 + int(par["numCells"][0]) * 10 + (1 << 30)
            if free_b < need:
                full_check["skipped"] = f"single-GPU replay needs ~{need >> 30} GiB, {free_b >> 30} GiB free beside this rank's slab"
            else:
                s1 = host.CSph(device=-1)
                s1.select_scene(title)
                p1, v1 = s1.host_arrays()
                g1 = lib.SphSystem(s1.params, local)
                g1.set_array(lib.SPH_POS, p1)
                g1.set_array(lib.SPH_VEL, v1)
                del p1, v1
                for blk in replay:
                    g1.set_params(np.frombuffer(blk, dtype=lib.SIMPARAMS_DTYPE))
                    g1.step(1)
                ck_one = state_checksum(np.arange(n, dtype=np.uint32), g1.get_array(lib.SPH_POS), g1.get_array(lib.SPH_VEL),
                                        g1.get_array(lib.SPH_DENSITY))
                g1.close()
                s1.close()
                full_check.update({"ok": bool(ck_one == ck_multi and int(limbs[2].item()) == n), "checksum_slabs": f"{ck_multi:016x}",
                                   "checksum_one_gpu": f"{ck_one:016x}"})
        except Exception as e:                               # noqa: BLE001
            full_check["error"] = f"{type(e).__name__}: {e}"
    dist.barrier()

    # end to end with HOST buffers: every step uploads this rank's owned records from pinned memory and reads them back
    e2e_steps = max(3, min(args.steps, 5))
    # Two public accessor paths, both timed (every step uploads this rank's inputs and every step's result is read back
    # inside the timed region): (a) put_owned -> step -> fetch_owned, one direction at a time; (b) exchange_owned -> step:
    # the previous result comes out while this step's inputs go in (duplex), the last result by a plain fetch before the clock
    # stops.  Which one wins depends on how many GPUs share the host's PCIe uplinks: (b) at N <= 2, (a) at N = 8 on this box.
    out_rec = torch.empty((m.capacity, 12), dtype=torch.float32, pin_memory=True)
    in_cnt = cnt

    def leg_serial():
        nonlocal cnt
        h = d = 0
        for _ in range(e2e_steps):
            m.put_owned(0, host_rec.data_ptr(), in_cnt)
            h += in_cnt * 48
            one_step()
            cnt = m.fetch_owned(0, out_rec.data_ptr(), m.capacity)
            d += cnt * 48
        return h, d

    def leg_duplex():
        nonlocal cnt
        h = d = 0
        for _ in range(e2e_steps):
            got = m.exchange_owned(0, out_rec.data_ptr(), m.capacity, host_rec.data_ptr(), in_cnt)
            h += in_cnt * 48
            d += got * 48
            one_step()
        cnt = m.fetch_owned(0, out_rec.data_ptr(), m.capacity)
        d += cnt * 48
        return h, d

    legs = {}
    for name, leg in (("serial", leg_serial), ("duplex", leg_duplex)):
        m.put_owned(0, host_rec.data_ptr(), in_cnt)          # both legs start from the same state
        m.sync()
        dist.barrier()
        t0 = time.perf_counter()
        h2d, d2h = leg()
        m.sync()
        dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        legs[name] = (float(t.item()), h2d, d2h)
    best = min(legs, key=lambda k: legs[k][0])
    h2d, d2h = legs[best][1], legs[best][2]
    host_rec = out_rec
    e2e_s = torch.tensor([legs[best][0]], device="cuda")
    io = torch.tensor([h2d, d2h], device="cuda", dtype=torch.int64)
    dist.all_reduce(io)
    finite = bool(np.isfinite(host_rec.numpy()[:cnt, :8]).all())
    owned_end = torch.tensor([cnt], device="cuda", dtype=torch.int64)
    dist.all_reduce(owned_end)
    flags = torch.tensor([int(finite)], device="cuda", dtype=torch.int64)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)

    hbm, hbm_src = measured_peaks()
    total_steps = warm + args.steps + e2e_steps
    out = {
        "metric": METRIC, "value": n * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": title, "particles": n, "particles_per_gpu": per_gpu,
                   "slab_cuts_z_layers": info1["cuts"], "grid": [int(x) for x in par["gridSize"][0]], "scene_file": "scenes/Scenes.xml",
                   "parallelism": f"z-slab x{world}, 1-layer halo, ncclSend/ncclRecv from the C++ driver (sph_multi_*), one process per GPU",
                   "message_records": {"leavers": info1["cap_leavers"], "boundary": info1["cap_boundary"]},
                   "l2": "per-GPU state is far larger than L2; no flush needed", "pair_kernels": view.pair_variant(),
                   "timing": "CUDA events on the solver stream, max over ranks", "wall_s": round(wall, 3),
                   "host_enqueue_s": round(host_enqueue_s, 3)},
        "clocks": clocks,
        "e2e": {"value": n * e2e_steps / float(e2e_s.item()), "unit": UNIT,
                "h2d_bytes_per_step": int(io[0].item()) // e2e_steps, "d2h_bytes_per_step": int(io[1].item()) // e2e_steps,
                "steps": e2e_steps, "path": best,
                "api": {"serial": "per rank: sph_multi_put_owned (pinned host -> device), sph_multi_step, sph_multi_fetch_owned",
                        "duplex": "per rank: sph_multi_exchange_owned (previous result out, this step's inputs in, download overlapping "
                                  "upload; pinned host memory), sph_multi_step; the last result by sph_multi_fetch_owned inside the "
                                  "timed region"}[best],
                "both_paths": {k: n * e2e_steps / v[0] for k, v in legs.items()}},
        "gpu_launches": int(launches),
        "halo_bytes_per_step_rank0": info1["bytes_sent"] // max(total_steps, 1),
        "phase_ms_last_step_by_rank": phases,
        "after_run": {"particles_conserved": bool(int(owned_end.item()) == n), "all_finite": bool(flags.item() == 1)},
        "parity_check": parity,
        "parity_check_full_size": full_check if world > 1 else None,
        "roofline": None,
    }
    dom = max((k for k in ("density", "force") if k in kernel_ms), key=lambda k: kernel_ms[k], default=None)
    if dom is not None and kernel_ms[dom] > 0:
        achieved = STAGE_BYTES[dom] * owned_timed / (kernel_ms[dom] * 1e-3) / 1e9
        ev = ncu_evidence(dom, "tank 8M drop")
        out["roofline"] = {"bound": ev.get("bound_unit") or "unknown (no current ncu capture)", "roofline_against": "hbm",
                           "kernel": {"density": "k_density_rm", "force": "k_force_rm"}[dom],
                           "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                           "traffic": None, "peak_source": hbm_src, "algorithmic_bytes_per_particle": STAGE_BYTES[dom],
                           "note": "rank 0, last timed step, owned particles only (force: interior + boundary passes and the wait for "
                                   "the rho,p rows in between); the single-GPU line carries the ncu traffic",
                           "kernel_ms": {k: round(v, 4) for k, v in kernel_ms.items()}}
    m.close()
    return out if rank == 0 else None



def scaling_base(steps: int, warmup: int) -> dict:
    """The N=1 point of the weak-scaling curve on the SAME path the N>1 runs take: 'wave tank 8M' as one slab of the C++
    multi-GPU driver (no neighbours, so no exchange) -- so that 1 -> N compares one code path on one scene family."""
    import torch
    from pibiti_b200 import host, lib
    title = "wave tank 8M"
    s = host.CSph(device=-1)
    s.select_scene(title)
    pos, vel = s.host_arrays()
    n = s.n
    m = lib.MultiSystem(s.params, capacity_per_slab=int(n * 1.25) + 600000, devices=[0])
    m.set_state(pos, vel)
    del pos, vel
    stream = torch.cuda.ExternalStream(m.stream(0))

    def one_step():
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)

    for _ in range(max(warmup, 3)):
        one_step()
    m.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        one_step()
    e1.record(stream)
    m.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    m.close()
    return {"workload": title, "particles": n, "path": "sph_multi_* driver, one slab (the N=1 point of the slab-decomposed runs)",
            "steps": steps, "ms_per_step": round(ms, 4), "value": n / (ms * 1e-3), "unit": UNIT}


def _host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def cpu_baseline(title: str, threads: int | None = None, full_steps: int = 1) -> tuple[dict, dict]:
    """The CPU oracle on this box's host cores, on a bounded sample of the workload -- and, because the oracle's
    step is already paid for, the one-step parity bar of the workload the numbers are quoted on: oracle and GPU start
    from the same state, both step once, integers must agree bit for bit and rho / v within 1e-5."""
    from oracle import oracle as orc                  # the one place bench.py executes oracle/ (as the checker)
    from pibiti_b200 import host, lib
    O = orc.load(None)
    O.set_threads(threads or _host_threads())
    s = host.CSph(device=0)
    s.select_scene(title)
    if "drop" in title:
        s.Drop(False)
    if "wave" in title:
        s.UpdateEmitter()
    par = s.params
    pos, vel = s.host_arrays()
    n = s.n
    o = O.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    t0 = time.perf_counter()
    o.step(full_steps)
    dt = time.perf_counter() - t0
    base = {"value": n * full_steps / dt, "unit": UNIT, "cores": O.threads(), "kind": O.kind,
            "sample": f"{full_steps} step(s) of '{title}' ({n} particles), OpenMP over emulated thread blocks",
            "seconds": round(dt, 2)}

    check = {"config": title, "particles": n, "steps": full_steps, "oracle": O.kind, "ok": False}
    try:
        g = s.solver()
        s.Update(full_steps)
        REL = 1e-5
        res = {
            "sorted_pairs_bit_exact": bool(np.array_equal(g.dump(lib.DUMP_SORTED_PAIRS), o.dump(0))),
            "cell_start_bit_exact": bool(np.array_equal(g.dump(lib.DUMP_CELL_START), o.dump(1))),
            "neighbour_counts_bit_exact": bool(np.array_equal(g.dump(lib.DUMP_NEIGHBOR_COUNTS), o.dump(6))),
            "positions_bit_exact": bool(np.array_equal(g.get_array(lib.SPH_POS), o.get_array(0))),
        }
        dg, do = g.dump(lib.DUMP_DENSITY), o.dump(5)
        res["density_max_rel_err"] = float(np.max(np.abs(dg - do) / np.maximum(np.abs(do), 1e-30)))
        vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
        vmax = max(float(np.abs(vo[:, :3]).max()), 1e-3)
        res["velocity_max_err_over_vmax"] = float(np.abs(vg - vo).max() / vmax)
        res["tolerance"] = REL
        check.update(res)
        check["ok"] = bool(res["sorted_pairs_bit_exact"] and res["cell_start_bit_exact"] and res["neighbour_counts_bit_exact"]
                           and res["positions_bit_exact"] and res["density_max_rel_err"] <= REL
                           and res["velocity_max_err_over_vmax"] <= REL)
    except Exception as e:                                   # noqa: BLE001
        check["error"] = f"{type(e).__name__}: {e}"
    o.close()
    s.close()
    return base, check


def run_reference(args) -> dict:
    """--impl reference: the reference's own CPU implementation of the path (its kernel text host-compiled,
    oracle/_ref/libsphref.so), all host threads, on the SAME workload the GPU arm runs at this N, from the initial state
    the reference's own scene code builds.  Nothing of pibiti_b200/ is loaded into this process."""
    n_gpus = args.gpus
    title = workload_for(n_gpus, args.workload)
    from oracle import oracle as orc
    O = orc.load(None)
    O.set_threads(_host_threads())
    wave = "wave" in title
    if O.kind == "reference":
        from oracle.ref_scene import RefScene
        sc = RefScene(ROOT / "scenes", title)
        if "drop" in title:
            sc.drop(False)
        par = sc.params()
        pos, vel = sc.arrays()
        n = sc.n
        prologue = sc.update_emitter if wave else None
    else:                                               # no reference build on this box: the port, scene from the host layer
        from pibiti_b200 import host
        hs = host.CSph(device=-1)
        hs.select_scene(title)
        if "drop" in title:
            hs.Drop(False)
        par = hs.params
        pos, vel = hs.host_arrays()
        n = hs.n

        def prologue():
            hs.UpdateEmitter()
            return hs.params
        if not wave:
            prologue = None
    o = O.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    del pos, vel

    def run(k):
        for _ in range(k):
            if prologue is not None:
                o.set_params(prologue())
            o.step(1)

    # bounded in TIME, never in space: the workload is always the GPU arm's; when K+W steps of it would run for much longer
    # than REF_BUDGET_S on these host cores (16M particles and up), fewer steps are timed and the line says how many
    budget = float(os.environ.get("SPH_REF_BUDGET_S", "150"))
    t0 = time.perf_counter()
    run(1)
    first = time.perf_counter() - t0
    warm = max(1, min(max(args.warmup, 1), int(0.2 * budget / first)))
    run(warm - 1)
    timed = max(1, min(args.steps, int(0.8 * budget / first)))
    t0 = time.perf_counter()
    run(timed)
    dt = time.perf_counter() - t0
    value = n * timed / dt
    sample = f"{timed} of {args.steps} steps of '{title}' ({n} particles): the full workload of the GPU arm, {warm} warm-up step(s)"
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "steps_timed": timed, "warmup": warm, "ms_per_step": dt / timed * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": title, "particles": n, "scene_file": "scenes/Scenes.xml",
                       "initial_state": "reference scene code (oracle/_ref)" if O.kind == "reference" else "host layer (port)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.threads(), "kind": O.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "native_so_loaded": repo_libraries_mapped()}


def repo_libraries_mapped() -> list[str]:
    """Shared libraries of this repository mapped into this process (the reference arm must show oracle/ only)."""
    libs = set()
    try:
        for ln in open("/proc/self/maps"):
            path = ln.split(None, 5)[-1].strip() if ln.count(" ") >= 5 else ""
            if path.startswith(str(ROOT)) and ".so" in path:
                libs.add(os.path.relpath(path, ROOT))
    except OSError:
        pass
    return sorted(libs)


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
        out = run_ours_multi(args)
        if rank == 0:
            print(json.dumps(out), flush=True)
        return

    out, s = run_ours_single(args)
    if not args.no_cpu_baseline:
        s.close()
        out["cpu_baseline"], out["parity_check"] = cpu_baseline(out["config"]["workload"])
        if args.workload is None:
            try:
                out["scaling_base"] = scaling_base(min(args.steps, 30), 5)
            except Exception as e:                           # noqa: BLE001 -- an extra, never worth losing the line for
                out["scaling_base"] = {"error": f"{type(e).__name__}: {e}"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
