"""Slab decomposition (multi-GPU path): an R-rank run must equal the single-domain run.

CPU tests drive pibiti_b200.slab with a backend built on the oracle port (tests/slab_oracle.py): in one
process (LocalComm) and as two processes over gloo (DistComm) -- the same communicator class the NCCL
runs use.  GPU tests (`-m gpu`) do the same with the real CUDA backend, all ranks sharing cuda:0.
Equality is bit-exact: a rank sorts its cells by (hash, original index) and sees the same neighbours
in the same order as the single-domain run."""
from __future__ import annotations

import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from pibiti_b200 import host, lib, slab
from slab_oracle import OracleSlabBackend


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def stir(vel):
    """Deterministic initial z velocities (+-1.5 m/s) so that particles cross slab boundaries within a few steps."""
    v = vel.copy()
    v[:, 2] = 1.5 * np.sin(np.arange(v.shape[0], dtype=np.float32) * np.float32(0.37)).astype(np.float32)
    return v


def scene_state(title):
    s = host.CSph(device=-1)
    s.select_scene(title)
    pos, vel = s.host_arrays()
    return s, s.params, pos, stir(vel)


def single_oracle_run(oracle, title, steps):
    s, par, pos, vel = scene_state(title)
    o = oracle.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    for _ in range(steps):
        s.UpdateEmitter()
        o.set_params(s.params)
        o.step(1)
    idx = o.dump(0)[:, 1]
    dens = np.empty(s.n, np.float32)
    dens[idx] = o.dump(5)
    out = o.get_array(0), o.get_array(1), dens
    o.close()
    return out


def check_records(rec, ref):
    pos, vel, dens = ref
    assert np.array_equal(rec[:, 0:4], pos), "positions differ"
    assert np.array_equal(rec[:, 4:8], vel), "velocities differ"
    assert np.array_equal(rec[:, 9], dens), "densities differ"


def test_cut_layers_balances_and_keeps_minimum_thickness():
    zc = np.concatenate([np.full(1000, 3), np.full(1000, 4), np.arange(5, 45).repeat(50)])
    cuts = slab.cut_layers(zc, 64, 4)
    assert cuts[0] == 0 and cuts[-1] == 64 and all(b - a >= 2 for a, b in zip(cuts, cuts[1:]))
    counts = [int(((zc >= a) & (zc < b)).sum()) for a, b in zip(cuts, cuts[1:])]
    assert max(counts) <= 2 * min(counts) + 1000
    with pytest.raises(lib.SphError):
        slab.cut_layers(zc, 6, 4)


def test_records_roundtrip():
    pos = np.random.default_rng(0).random((10, 4), np.float32)
    ids = np.array([9, 8, 7, 6, 5, 4, 3, 2, 1, 0], np.uint32)
    rec = slab.make_records(pos, pos * 2, ids)
    assert np.array_equal(slab.record_ids(rec), ids)
    g = slab.gather_by_id([rec[:4], rec[4:]], 10)
    assert np.array_equal(g[:, 0:4], pos[::-1])


@pytest.mark.parametrize("title,ranks", [("mini waves", 2), ("mini waves", 3), ("mini box", 2), ("mini collider accel", 3)])
def test_oracle_slabs_equal_single_domain(oracle_port, title, ranks):
    steps = 8
    s, par, pos, vel = scene_state(title)
    cuts, parts = slab.split_initial_state(par, pos, vel, ranks)
    caps = slab.SlabCaps.for_state(par, pos, cuts)
    bes = [OracleSlabBackend(oracle_port, par, cuts[r], cuts[r + 1], r > 0, r < ranks - 1, caps) for r in range(ranks)]
    for b, p in zip(bes, parts):
        b.set_owned(p)
    comm = slab.LocalComm()
    moved = 0
    for _ in range(steps):
        s.UpdateEmitter()
        before = [set(slab.record_ids(b.rec).tolist()) for b in bes]
        for b in bes:
            b.set_params(s.params)
        slab.slab_step(bes, comm)
        moved += sum(len(set(slab.record_ids(b.rec).tolist()) - bf) for b, bf in zip(bes, before))
    rec = slab.gather_by_id([b.get_owned() for b in bes], s.n)
    check_records(rec, single_oracle_run(oracle_port, title, steps))
    if title == "mini waves":
        assert moved > 0, "the test should exercise migration between slabs"


def run_workers(kind, title, steps, tmp_path, nproc=2):
    out = tmp_path / "slab.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "tests" / "slab_worker.py"), kind, title, str(steps), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out)


def test_two_process_gloo_slabs_equal_single_domain(oracle_port, tmp_path):
    steps = 6
    got = run_workers("oracle", "mini waves", steps, tmp_path)
    assert got["owned"].sum() == 16384 and (got["owned"] > 0).all()
    check_records(got["rec"], single_oracle_run(oracle_port, "mini waves", steps))


# ---- GPU -----------------------------------------------------------------------------------------

def single_gpu_run(title, steps):
    s = host.CSph(device=0)
    s.select_scene(title)
    g = s.solver()
    g.set_array(lib.SPH_VEL, stir(s.host_arrays()[1]))
    for _ in range(steps):
        s.UpdateEmitter()
        s.Update()
    out = g.get_array(lib.SPH_POS), g.get_array(lib.SPH_VEL), g.get_array(lib.SPH_DENSITY)
    s.close()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["l1,128,1344,48", "tma,128,1344,48", "rm,128,32,48"])
@pytest.mark.parametrize("title,ranks", [("mini waves", 2), ("mini waves", 3), ("wave tank 256k", 4)])
def test_gpu_slabs_equal_single_gpu(title, ranks, variant, monkeypatch):
    monkeypatch.setenv("SPH_B200_PAIR_CFG", variant)
    steps = 10
    s, par, pos, vel = scene_state(title)
    cuts, parts = slab.split_initial_state(par, pos, vel, ranks)
    caps = slab.SlabCaps.for_state(par, pos, cuts)
    bes = [slab.GpuSlabBackend(par, int(p.shape[0] * 1.5) + 4 * caps.rows, cuts[r], cuts[r + 1], r > 0, r < ranks - 1, 0, caps)
           for r, p in enumerate(parts)]
    for b, p in zip(bes, parts):
        b.set_owned(p)
    comm = slab.LocalComm()
    for _ in range(steps):
        s.UpdateEmitter()
        for b in bes:
            b.set_params(s.params)
        slab.slab_step(bes, comm)
    rec = slab.gather_by_id([b.get_owned().cpu().numpy() for b in bes], s.n)
    check_records(rec, single_gpu_run(title, steps))


@pytest.mark.gpu
def test_gpu_slabs_split_integrate_and_pack_calls(monkeypatch):
    """The separate sph_slab_integrate / sph_slab_pack entry points (slab_step uses the fused one by default)."""
    monkeypatch.setenv("SPH_SLAB_SPLIT_PACK", "1")
    test_gpu_slabs_equal_single_gpu("mini waves", 3, "l1,128,1344,48", monkeypatch)


@pytest.mark.gpu
def test_gpu_two_process_gloo_slabs_equal_single_gpu(tmp_path):
    steps = 6
    got = run_workers("gpu", "mini waves", steps, tmp_path)
    check_records(got["rec"], single_gpu_run("mini waves", steps))


def test_cpp_cut_planner_matches_the_protocol_model():
    """The multi-GPU driver's cut planner (sph_multi_plan_cuts, C++) against slab.cut_layers (the Python model the gloo tests
    run on): same boundaries on random layer histograms, thin grids refused by both."""
    import ctypes as C
    L = lib.load()
    rng = np.random.default_rng(7)
    for trial in range(200):
        gz = int(rng.integers(4, 400))
        ranks = int(rng.integers(1, 9))
        n = int(rng.integers(1, 200000))
        zc = np.minimum((rng.random(n) ** rng.uniform(0.3, 3.0) * gz).astype(np.int64), gz - 1)
        hist = np.bincount(zc, minlength=gz).astype(np.int64)
        cuts = (C.c_int * (ranks + 1))()
        rc = L.sph_multi_plan_cuts(hist.ctypes.data_as(C.c_void_p), gz, ranks, cuts)
        try:
            want = slab.cut_layers(zc, gz, ranks)
        except lib.SphError:
            assert rc != 0, (gz, ranks)
            continue
        assert rc == 0 and list(cuts) == want, (gz, ranks, list(cuts), want)
