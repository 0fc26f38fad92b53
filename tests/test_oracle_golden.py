"""Pins the CPU oracle.

tests/golden/steps.npz was produced by the REFERENCE's own kernel text compiled for the host
(oracle/_ref, tests/golden/make_golden.py).  The port (oracle/sph_port.cpp) must reproduce every
digest bit-for-bit; where the reference build is present it is re-checked against its own vectors
and compared live with the port.
"""
from __future__ import annotations

import numpy as np
import pytest

from conftest import REFERENCE_XML, sha
from oracle import oracle as orc
from pibiti_b200 import host

ITEMS = {"pairs": 0, "cellStart": 1, "sortedPos": 2, "sortedVel": 3, "pressure": 4, "density": 5, "counts": 6,
         "color": 7, "dye": 8}


def initial_state(title):
    s = host.CSph(device=-1)
    s.select_scene(title)
    par = s.params
    pos, vel = s.host_arrays()
    s.close()
    return par, pos, vel


def run_against_golden(oracle, golden, title):
    key = title.replace(" ", "_")
    par, pos, vel = initial_state(title)
    gpar = golden[f"{key}/params"]
    for name in par.dtype.names:                # the product's scene layer produced the oracle's inputs
        assert par[name].tobytes() == gpar[name].tobytes(), name
    sysm = oracle.system(par)
    sysm.set_array(0, pos)
    sysm.set_array(1, vel)
    done = 0
    for steps in (1, 5):
        sysm.step(steps - done)
        done = steps
        for name, what in ITEMS.items():
            assert sha(sysm.dump(what)) == str(golden[f"{key}/{steps}/{name}_sha"]), (title, steps, name)
        assert sha(sysm.get_array(0)) == str(golden[f"{key}/{steps}/pos_sha"]), (title, steps, "pos")
        assert sha(sysm.get_array(1)) == str(golden[f"{key}/{steps}/vel_sha"]), (title, steps, "vel")
    sysm.close()


def titles(golden):
    return [str(t) for t in golden["titles"]]


def test_golden_lists_expected_scenes(golden_steps):
    t = titles(golden_steps)
    assert "box small default" in t and "Stiff  Dam break" in t and "mini dense cells" in t and len(t) >= 17
    # the truncation quirk is really exercised by one fixture
    assert int(golden_steps["mini_dense_cells/1/max_cell"]) > 8


@pytest.mark.parametrize("title", ["box small default", "Stiff  Dam break", "mini box", "mini dense cells", "mini random",
                                   "mini cylinder Y", "mini cylinder Z", "mini sphere", "mini wrap Z", "mini cycle Z",
                                   "mini waves", "mini collider accel", "mini heightmap XZ", "mini heightmap YZ holes",
                                   "mini rotor Z", "mini rotor Y", "mini propeller pair"])
def test_port_reproduces_reference_vectors(oracle_port, golden_steps, title):
    run_against_golden(oracle_port, golden_steps, title)


@pytest.mark.skipif(not orc.available("reference"), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("title", ["box small default", "mini dense cells", "mini waves"])
def test_reference_build_reproduces_its_own_vectors(golden_steps, title):
    run_against_golden(orc.load("reference"), golden_steps, title)


@pytest.mark.skipif(not (orc.available("reference") and REFERENCE_XML.exists()), reason="needs /root/reference")
@pytest.mark.parametrize("scene", [6, 28, 46, 55, 63, 76, 78, 116])
def test_port_equals_reference_on_reference_scenes(oracle_port, scene):
    """Boundary types, rotors, pump, height map, colour modes and dye on the reference's own scenes."""
    ref = orc.load("reference")
    s = host.CSph(REFERENCE_XML, device=-1)
    s.select_scene(scene)
    par = s.params
    pos, vel = s.host_arrays()
    s.close()
    a, b = ref.system(par), oracle_port.system(par)
    for m in (a, b):
        m.set_array(0, pos)
        m.set_array(1, vel)
    for rnd in range(2):
        a.step(2)
        b.step(2)
        for what in range(9):
            assert a.dump(what).tobytes() == b.dump(what).tobytes(), (scene, rnd, what)
        assert a.get_array(1).tobytes() == b.get_array(1).tobytes()
        par2 = par.copy()
        par2["dyeClear"] = 0
        par2["dyeType"] = 1 + scene % 2
        par2["clrType"] = (scene + rnd) % 6
        par2["rAngle"] += 0.2
        a.set_params(par2)
        b.set_params(par2)
    a.close()
    b.close()


def test_stage_functions_compose_to_a_step(oracle_port):
    """orc_integrate / calc_hash / sort_pairs used one by one equal the system object's first half."""
    par, pos, vel = initial_state("mini box")
    oracle_port.set_params(par)
    npos, nvel = oracle_port.integrate(pos, vel)
    pairs = oracle_port.sort_pairs(oracle_port.calc_hash(npos))
    sysm = oracle_port.system(par)
    sysm.set_array(0, pos)
    sysm.set_array(1, vel)
    sysm.step(1)
    assert np.array_equal(pairs, sysm.dump(0))
    assert np.array_equal(npos, sysm.get_array(0))
    assert np.array_equal(npos[pairs[:, 1]], sysm.dump(2))
    assert np.array_equal(nvel[pairs[:, 1]], sysm.dump(3))
    # stable: inside a cell the original indices ascend
    same = pairs[1:, 0] == pairs[:-1, 0]
    assert np.all(pairs[1:, 1][same] > pairs[:-1, 1][same])
    sysm.close()
