"""The C++ multi-GPU driver (sph_multi_*, pibiti_b200/csrc/sph_multi.cu): a run cut into R z slabs must equal the
single-GPU run bit for bit -- positions, velocities and densities -- with particles stirred along z so that they migrate
between slabs.  On a one-GPU box all slabs share cuda:0 (the driver then copies the neighbours' buffers directly instead
of calling NCCL); tests/multi_worker.py runs the one-process-per-GPU shape over NCCL where two GPUs are visible."""
from __future__ import annotations

import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from pibiti_b200 import host, lib

pytestmark = pytest.mark.gpu


def stir(vel):
    v = vel.copy()
    v[:, 2] = 1.5 * np.sin(np.arange(v.shape[0], dtype=np.float32) * np.float32(0.37)).astype(np.float32)
    return v


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


def multi_run(title, steps, slabs, monkeypatch=None, variant=None, check_every=0, recut_every=0):
    s = host.CSph(device=-1)                      # scene + initial state on the host only
    s.select_scene(title)
    par = s.params
    pos, vel = s.host_arrays()
    vel = stir(vel)
    n = s.n
    m = lib.MultiSystem(par, capacity_per_slab=int(n / slabs * 1.6) + 40000, devices=[0] * slabs)
    m.set_state(pos, vel)
    m.set_recut_interval(recut_every)
    owned0 = m.info()["owned"]
    assert sum(owned0) == n
    for k in range(steps):
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)
        if check_every and (k + 1) % check_every == 0:
            m.sync()
    p, v, d, _, written = m.get_state(density=True)
    info = m.info()
    info["recuts"] = m.recut_count()
    m.close()
    assert written == n and sum(info["owned"]) == n
    return (p, v, d), owned0, info


@pytest.mark.parametrize("title,slabs", [("mini waves", 2), ("mini waves", 3), ("wave tank 256k", 4), ("mini box", 1)])
def test_multi_driver_equals_single_gpu(title, slabs):
    steps = 10
    got, owned0, info = multi_run(title, steps, slabs)
    ref = single_gpu_run(title, steps)
    for a, b, what in zip(got, ref, ("positions", "velocities", "densities")):
        assert np.array_equal(a, b), f"{what} differ"
    if slabs > 1 and "waves" in title:
        assert info["owned"] != owned0, "the test should exercise migration between slabs"
        assert info["bytes_sent"] > 0


@pytest.mark.parametrize("variant", ["l1,128,1344,48", "rm,64,24,48"])
def test_multi_driver_other_pair_variants(monkeypatch, variant):
    monkeypatch.setenv("SPH_B200_PAIR_CFG", variant)
    got, _, _ = multi_run("mini waves", 6, 3)
    ref = single_gpu_run("mini waves", 6)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("slabs", [2, 4])
def test_multi_driver_peer_store_exchange(monkeypatch, slabs):
    """SPH_B200_MULTI_XCHG=peer: no exchange copy at all -- the packing kernels store leavers, boundary copies and rho,p rows
    straight into the neighbours' inboxes and the receivers wait for the senders' kernels.  Same results, with a re-cut."""
    monkeypatch.setenv("SPH_B200_MULTI_XCHG", "peer")
    got, owned0, info = multi_run("wave tank 256k" if slabs == 4 else "mini waves", 10, slabs, recut_every=4)
    ref = single_gpu_run("wave tank 256k" if slabs == 4 else "mini waves", 10)
    for a, b, what in zip(got, ref, ("positions", "velocities", "densities")):
        assert np.array_equal(a, b), f"{what} differ"
    assert info["owned"] != owned0 and info["bytes_sent"] == 0


def test_multi_driver_recut_keeps_the_results(monkeypatch):
    """Re-cutting the slabs every three steps changes who owns what, never the result."""
    got, _, info = multi_run("mini waves", 10, 3, recut_every=3)
    assert info["recuts"] == 3
    ref = single_gpu_run("mini waves", 10)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b)


def test_multi_driver_reports_message_overflow(monkeypatch):
    """Sections sized far below a layer's population: the device flags it and the next sync reports it."""
    monkeypatch.setenv("SPH_B200_SLAB_SAFETY", "0.01")
    s = host.CSph(device=-1)
    s.select_scene("wave tank 256k")
    pos, vel = s.host_arrays()
    m = lib.MultiSystem(s.params, capacity_per_slab=s.n, devices=[0, 0])
    m.set_state(pos, stir(vel))
    m.step(2)
    with pytest.raises(lib.SphError, match="overflow"):
        m.sync()
    m.close()


def test_multi_driver_rejects_unsupported_boundaries():
    s = host.CSph(device=-1)
    s.select_scene("mini pump square")
    m = lib.MultiSystem(s.params, capacity_per_slab=s.n, devices=[0, 0])
    pos, vel = s.host_arrays()
    with pytest.raises(lib.SphError, match="pump"):
        m.set_state(pos, vel)
    m.close()


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_multi_driver_one_process_two_gpus_nccl_and_copies(monkeypatch):
    """One process driving two DIFFERENT GPUs: ncclCommInitAll + grouped ncclSend/ncclRecv, and the default peer copies."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ref = single_gpu_run("wave tank 256k", 8)
    for mode in ("nccl", "copy", "peer"):
        monkeypatch.setenv("SPH_B200_MULTI_XCHG", mode)
        s = host.CSph(device=-1)
        s.select_scene("wave tank 256k")
        pos, vel = s.host_arrays()
        m = lib.MultiSystem(s.params, capacity_per_slab=s.n, devices=[0, 1])
        m.set_state(pos, stir(vel))
        for _ in range(8):
            s.UpdateEmitter()
            m.set_params(s.params)
            m.step(1)
        p, v, d, _, written = m.get_state(density=True)
        sent = m.info()["bytes_sent"]
        m.close()
        assert written == s.n and (sent > 0) == (mode != "peer")
        assert np.array_equal(p, ref[0]) and np.array_equal(v, ref[1]) and np.array_equal(d, ref[2]), mode


def test_multi_driver_nccl_two_processes(tmp_path):
    """One process per GPU over ncclSend/ncclRecv (the shape bench.py runs under torchrun); needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "multi.npz"
    steps = 8
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(ROOT / "tests" / "multi_worker.py"), "wave tank 256k", str(steps), str(out), "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    assert int(got["recuts"]) == 2, "the worker re-cuts every three steps (NCCL all-reduce / all-gather + send/recv)"
    ref = single_gpu_run("wave tank 256k", steps)
    assert np.array_equal(got["pos"], ref[0]) and np.array_equal(got["vel"], ref[1]) and np.array_equal(got["dens"], ref[2])


def test_csph_device_list_checkpoint_round_trip(tmp_path):
    """SaveState / LoadState through the host mirrors of a multi-GPU cSPH: a run resumed from the file (into a fresh
    one-device object) continues exactly like the uninterrupted multi-GPU run."""
    a = host.CSph(device=0, devices=[0, 0])
    a.select_scene("mini waves")
    for _ in range(4):
        a.UpdateEmitter()
        a.Update()
    a.SaveState(tmp_path / "multi.ckp")
    for _ in range(3):
        a.UpdateEmitter()
        a.Update()
    want = a.getArray(False).copy(), a.getArray(True).copy()
    a.close()
    for devices in (None, [0, 0, 0]):
        b = host.CSph(device=0, devices=devices)
        b.LoadState(tmp_path / "multi.ckp")
        for _ in range(3):
            b.UpdateEmitter()
            b.Update()
        assert np.array_equal(b.getArray(False), want[0]) and np.array_equal(b.getArray(True), want[1]), devices
        b.close()


def test_csph_with_a_device_list_equals_one_device():
    """cSPH(device list): Reset through the mirrors, a full-range setArray, steps, a PARTIAL setArray mid-run (what an
    emitter does) and getArray -- identical to the one-device cSPH."""
    def run(devices):
        s = host.CSph(device=0, devices=devices)
        s.select_scene("mini waves")
        _, vel = s.host_arrays()
        s.setArray(True, stir(vel))
        for k in range(8):
            s.UpdateEmitter()
            s.Update()
            if k == 3:
                v = s.getArray(True)[100:228].copy()
                v[:, 1] += 0.25
                s.setArray(True, v, start=100)
        out = s.getArray(False).copy(), s.getArray(True).copy()
        s.close()
        return out
    one, three = run(None), run([0, 0, 0])
    assert np.array_equal(one[0], three[0]) and np.array_equal(one[1], three[1])


def test_fetch_put_owned_round_trip_is_transparent():
    """The per-process host accessors of the one-process-per-GPU shape (what bench.py's end-to-end leg uses): fetching a slab's
    owned records to the host and putting them back between steps does not change the trajectory."""
    import torch
    s = host.CSph(device=-1)
    s.select_scene("mini waves")
    pos, vel = s.host_arrays()
    vel = stir(vel)
    m = lib.MultiSystem(s.params, capacity_per_slab=s.n, devices=[0, 0])
    m.set_state(pos, vel)
    buf = torch.empty((s.n, 12), dtype=torch.float32, pin_memory=True)
    for k in range(6):
        s.UpdateEmitter()
        m.set_params(s.params)
        m.step(1)
        if k % 2 == 1:
            for local in range(2):
                cnt = m.fetch_owned(local, buf.data_ptr(), s.n)
                assert 0 < cnt < s.n
                ids = buf.numpy()[:cnt, 8].copy().view(np.uint32)
                assert len(np.unique(ids)) == cnt
                if k == 1:
                    m.put_owned(local, buf.data_ptr(), cnt)
                else:                   # the duplex form: what comes out is what went in
                    out = torch.empty((s.n, 12), dtype=torch.float32, pin_memory=True)
                    got = m.exchange_owned(local, out.data_ptr(), s.n, buf.data_ptr(), cnt)
                    assert got == cnt and np.array_equal(out.numpy()[:cnt], buf.numpy()[:cnt])
    p, v, written = m.get_state()
    m.close()
    ref = single_gpu_run("mini waves", 6)
    assert written == s.n and np.array_equal(p, ref[0]) and np.array_equal(v, ref[1])


def test_multi_driver_with_obstacles_and_collider():
    """Height map / rotor obstacles and the sphere collider run behind the force passes of every slab."""
    for title in ("mini heightmap XZ", "mini collider accel"):
        got, _, _ = multi_run(title, 5, 2)
        ref = single_gpu_run(title, 5)
        for a, b, what in zip(got, ref, ("positions", "velocities", "densities")):
            assert np.array_equal(a, b), f"{title}: {what} differ"
