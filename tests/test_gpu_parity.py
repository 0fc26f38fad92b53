"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Needs a B200: `-m gpu`.

Bars:
  * bit-exact: cell hashes, sorted (hash, index) pairs, cell table, neighbour counts, and -- because
    the streaming kernels evaluate floats exactly as the CPU does -- post-integration positions
    and velocities;
  * density / pressure / new velocities: relative 1e-5 after one step (REL below), with the floors
    written next to each check (pressure cancels against rho0*k; velocity components against |v|max).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

from conftest import ROOT, sha
from pibiti_b200 import host, lib

pytestmark = pytest.mark.gpu

REL = 1e-5
TITLES = ["box small default", "Stiff  Dam break", "mini box", "mini dense cells", "mini random", "mini cylinder Y",
          "mini cylinder Z", "mini sphere", "mini wrap Z", "mini cycle Z", "mini waves", "mini collider accel",
          "mini heightmap XZ", "mini heightmap YZ holes", "mini rotor Z", "mini rotor Y", "mini propeller pair"]
# Scenes whose obstacles are sphere lattices placed with sinf/cosf (System.cu:254-372): the sphere centres differ in
# the last bits between the device's and the host's libm, and the contact spring multiplies that by `spring`.
LIBM_SCENES = {"mini heightmap XZ", "mini heightmap YZ holes", "mini rotor Z", "mini rotor Y", "mini propeller pair"}
REL_LIBM = 1e-4


def start(title, oracle, device=0):
    s = host.CSph(device=device)
    s.select_scene(title)
    par = s.params
    pos, vel = s.host_arrays()
    o = oracle.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    return s, s.solver(), o, par


def check_floats(g, o, par, rel=REL):
    rho0k = float(par["restDensity"][0]) * float(par["stiffness"][0])
    dg, do = g.dump(lib.DUMP_DENSITY), o.dump(5)
    assert np.all(np.abs(dg - do) <= rel * np.abs(do) + 1e-30), "density"
    pg, po = g.dump(lib.DUMP_PRESSURE), o.dump(4)
    assert np.all(np.abs(pg - po) <= rel * (np.abs(po) + rho0k)), "pressure"
    vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
    vmax = max(float(np.abs(vo[:, :3]).max()), 1e-3)
    assert np.all(np.abs(vg - vo) <= rel * vmax), f"velocity: {np.abs(vg - vo).max()} vs scale {vmax}"
    return vmax


def check_integers_exact(g, o):
    assert np.array_equal(g.dump(lib.DUMP_SORTED_PAIRS), o.dump(0)), "sorted (hash,index) pairs"
    cs = g.dump(lib.DUMP_CELL_START)
    assert np.array_equal(cs, o.dump(1)), "cellStart"
    assert np.array_equal(g.dump(lib.DUMP_NEIGHBOR_COUNTS), o.dump(6)), "neighbour counts"
    assert np.array_equal(g.dump(lib.DUMP_SORTED_POS), o.dump(2)), "sortedPos (bit-exact integrate)"
    assert np.array_equal(g.dump(lib.DUMP_SORTED_VEL), o.dump(3)), "sortedVel (bit-exact integrate)"
    assert np.array_equal(g.get_array(lib.SPH_POS), o.get_array(0)), "positions, original order"


# the variants of the density/force pair (see pibiti_b200/csrc/sph_device.cuh): "tma,..." stages the candidates in
# shared memory by TMA bulk copies, "l1,..." reads them through L1 (one particle per thread, index lists), "rm,..." is
# the same walk with {bit mask, first index} neighbour records instead of index lists.  Format: mode,threads,cap,kMax
# (rm: cap = records per particle, kMax unused)
PAIR_VARIANTS = {"l1": "l1,128,1344,48", "tma": "tma,128,1344,48", "rm": "rm,128,32,48"}


@pytest.mark.parametrize("variant", ["l1", "tma", "rm"])
@pytest.mark.parametrize("title", TITLES)
def test_one_step_parity(oracle_any, golden_steps, title, variant, monkeypatch):
    monkeypatch.setenv("SPH_B200_PAIR_CFG", PAIR_VARIANTS[variant])
    s, g, o, par = start(title, oracle_any)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    rel = REL_LIBM if title in LIBM_SCENES else REL
    check_floats(g, o, par, rel)
    # and against the committed vectors of the reference build
    key = title.replace(" ", "_")
    assert sha(g.dump(lib.DUMP_SORTED_PAIRS)) == str(golden_steps[f"{key}/1/pairs_sha"])
    assert sha(g.dump(lib.DUMP_NEIGHBOR_COUNTS)) == str(golden_steps[f"{key}/1/counts_sha"])
    assert sha(g.dump(lib.DUMP_CELL_START)) == str(golden_steps[f"{key}/1/cellStart_sha"])
    assert sha(g.get_array(lib.SPH_POS)) == str(golden_steps[f"{key}/1/pos_sha"])
    gd = golden_steps[f"{key}/1/density_sample"]
    assert np.all(np.abs(g.dump(lib.DUMP_DENSITY)[::64] - gd) <= REL * np.abs(gd) + 1e-30)
    gv = golden_steps[f"{key}/1/vel_sample"]
    assert np.all(np.abs(g.get_array(lib.SPH_VEL)[::64] - gv) <= rel * max(float(np.abs(gv).max()), 1e-3))
    o.close()


@pytest.mark.parametrize("title", ["mini rotor Z", "mini propeller pair", "mini heightmap XZ"])
def test_obstacle_scenes_follow_the_oracle_with_moving_rotors(oracle_any, title):
    """Rotor angle advances every step (UpdateEmitter); ten steps with resync."""
    s, g, o, par = start(title, oracle_any)
    for step in range(10):
        s.UpdateEmitter()
        par = s.params
        g.set_params(par)
        o.set_params(par)
        g.step(1)
        o.step(1)
        check_integers_exact(g, o)
        check_floats(g, o, par, REL_LIBM)
        g.set_array(lib.SPH_POS, o.get_array(0))
        g.set_array(lib.SPH_VEL, o.get_array(1))
    vmax = float(np.abs(o.get_array(1)[:, :3]).max())
    assert vmax > 0.05, "the obstacle should be pushing the fluid"
    o.close()


@pytest.mark.parametrize("clr_type", [0, 1, 2, 3, 4, 5, 6])
def test_colour_and_dye_outputs(oracle_any, clr_type):
    """The visual-only outputs of computeForceD (System.cu:406-546): every colour mode, hue mode, dye box/sphere
    with fading and the dyeClear countdown.  Tolerance 1e-4 absolute on O(1) colours."""
    s, g, o, par = start("mini collider accel", oracle_any)
    g.set_visual(True)
    par = par.copy()
    par["clrType"] = clr_type
    par["iHue"] = clr_type % 2
    par["dyeType"] = 1 + clr_type % 2
    par["dyePos"] = (0.06, -0.085, 0.06)                     # inside the fluid, away from the collider sphere
    par["dyeSize"] = (0.03, 0.03, 0.03)
    for step in range(4):
        par["dyeClear"] = max(0, 1 - step)                      # cleared on the first step, then dyed and fading
        g.set_params(par)
        o.set_params(par)
        g.step(1)
        o.step(1)
        if clr_type == 6:                                       # CLR_None: the reference returns before colour AND dye
            continue
        assert np.all(np.abs(g.get_array(lib.SPH_DYE) - o.dump(8)) <= 1e-6), ("dye", step)
        assert np.all(np.abs(g.get_array(lib.SPH_COLOR) - o.dump(7)) <= 1e-4), ("colour", step)
        g.set_array(lib.SPH_POS, o.get_array(0))
        g.set_array(lib.SPH_VEL, o.get_array(1))
    if clr_type != 6:
        assert float(g.get_array(lib.SPH_DYE).max()) > 0.9      # some particles sit in the dye volume
    o.close()


@pytest.mark.parametrize("variant", ["l1", "tma", "rm"])
@pytest.mark.parametrize("title", ["box small default", "Stiff  Dam break", "mini dense cells", "mini waves", "mini wrap Z"])
def test_trajectory_parity_with_resync(oracle_any, title, variant, monkeypatch):
    monkeypatch.setenv("SPH_B200_PAIR_CFG", PAIR_VARIANTS[variant])
    """Follow the oracle's trajectory for 12 steps; before every step the GPU state is reset to the
    oracle's, then both take one step and must meet the one-step bar (integers bit-exact)."""
    s, g, o, par = start(title, oracle_any)
    for step in range(12):
        if "waves" in title:                                    # host prologue advances the wave phase
            s.UpdateEmitter()
            par = s.params
            g.set_params(par)
            o.set_params(par)
        g.step(1)
        o.step(1)
        check_integers_exact(g, o)
        check_floats(g, o, par)
        g.set_array(lib.SPH_POS, o.get_array(0))
        g.set_array(lib.SPH_VEL, o.get_array(1))
    o.close()


@pytest.mark.parametrize("title", ["box small default", "Stiff  Dam break"])
def test_free_running_drift_is_bounded(oracle_any, title):
    """50 steps without resync.  SPH is chaotic, so the bar is statistical: 99 % of the
    particles within half a lattice spacing, centre of mass / kinetic energy / density mean within 1 %."""
    s, g, o, par = start(title, oracle_any)
    spacing = float(s.scene_extra()[8])
    g.step(50)
    o.step(50)
    pg, po = g.get_array(lib.SPH_POS), o.get_array(0)
    d = np.linalg.norm(pg[:, :3] - po[:, :3], axis=1)
    assert np.quantile(d, 0.99) <= 0.5 * spacing, np.quantile(d, 0.99)
    extent = float(np.abs(po[:, :3]).max())
    assert np.all(np.abs(pg[:, :3].mean(0) - po[:, :3].mean(0)) <= 0.01 * extent)
    vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
    keg, keo = float((vg[:, :3] ** 2).sum()), float((vo[:, :3] ** 2).sum())
    assert abs(keg - keo) <= 0.01 * keo
    assert abs(float(g.dump(lib.DUMP_DENSITY).mean()) - float(o.dump(5).mean())) <= 0.01 * float(o.dump(5).mean())
    o.close()


def test_deterministic_and_order_independent(oracle_any):
    """The sort uses atomics for bucketing but ranks by original index: two runs give identical bits,
    and so does a run whose particles were handed over in a shuffled slot order."""
    s, g, o, par = start("mini dense cells", oracle_any)
    pos, vel = s.host_arrays()
    g.step(3)
    a = [g.dump(k).copy() for k in (lib.DUMP_SORTED_PAIRS, lib.DUMP_DENSITY)] + [g.get_array(lib.SPH_VEL)]
    g2 = lib.SphSystem(par)
    g2.set_array(lib.SPH_POS, pos)
    g2.set_array(lib.SPH_VEL, vel)
    g2.step(1)                                                  # now slot order != original order
    g2.set_array(lib.SPH_POS, pos)                              # same particles, written through the permutation
    g2.set_array(lib.SPH_VEL, vel)
    g2.step(3)
    b = [g2.dump(k).copy() for k in (lib.DUMP_SORTED_PAIRS, lib.DUMP_DENSITY)] + [g2.get_array(lib.SPH_VEL)]
    for x, y in zip(a, b):
        assert x.tobytes() == y.tobytes()
    g2.close()
    o.close()


def test_graph_replay_equals_direct_launches(monkeypatch):
    """Steps with unchanged parameters are replayed as CUDA graphs; a parameter change falls back to direct
    launches and re-captures.  Both ways must give identical bits."""
    results = []
    for graphs in ("1", "0"):
        monkeypatch.setenv("SPH_B200_GRAPHS", graphs)
        s = host.CSph(device=0)
        s.select_scene("mini collider accel")
        g = s.solver()
        g.step(7)                                               # direct, direct, then graphs of both parities
        par = s.params
        par["gravity"] = (0.5, -9.81, 0.0)
        g.set_params(par)                                       # invalidates the graphs
        g.step(6)
        results.append((g.get_array(lib.SPH_POS), g.get_array(lib.SPH_VEL), g.dump(lib.DUMP_SORTED_PAIRS), g.launch_count()))
        s.close()
    for a, b in zip(*results):
        assert np.array_equal(a, b)


def test_set_get_array_ranges(oracle_any):
    s, g, o, par = start("mini box", oracle_any)
    n = g.n
    rng = np.random.default_rng(5)
    g.step(2)                                                   # state is in sorted order now
    o.step(2)
    g.set_array(lib.SPH_POS, o.get_array(0))                    # full write through the permutation
    g.set_array(lib.SPH_VEL, o.get_array(1))
    assert np.array_equal(g.get_array(lib.SPH_POS), o.get_array(0))
    patch = rng.uniform(-0.05, 0.05, (100, 4)).astype(np.float32)
    patch[:, 3] = 1
    g.set_array(lib.SPH_POS, patch, start=1234)
    o.set_array(0, patch, start=1234)
    vpatch = rng.uniform(-1, 1, (77, 4)).astype(np.float32)
    g.set_array(lib.SPH_VEL, vpatch, start=n - 77)
    o.set_array(1, vpatch, start=n - 77)
    assert np.array_equal(g.get_array(lib.SPH_POS), o.get_array(0))
    assert np.array_equal(g.get_array(lib.SPH_VEL), o.get_array(1))
    assert np.array_equal(g.get_array(lib.SPH_POS, 1200, 200), o.get_array(0, 1200, 200))
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    check_floats(g, o, par)
    # density / pressure in original order == oracle's sorted arrays un-permuted
    idx = o.dump(0)[:, 1]
    dens = np.empty(n, np.float32)
    dens[idx] = o.dump(5)
    assert np.all(np.abs(g.get_array(lib.SPH_DENSITY) - dens) <= REL * dens)
    with pytest.raises(lib.SphError):
        g.set_array(lib.SPH_POS, patch, start=n - 5)            # range check
    o.close()


def test_unstaged_fallback_path_matches(oracle_any, monkeypatch):
    """A staging buffer too small for any CTA forces the global-memory walk; results must not change."""
    monkeypatch.setenv("SPH_B200_PAIR_CFG", "tma,128,16,16")
    s, g, o, par = start("mini dense cells", oracle_any)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    check_floats(g, o, par)
    o.close()


@pytest.mark.parametrize("variant", ["l1", "tma", "rm"])
def test_neighbour_list_overflow_path_matches(oracle_any, monkeypatch, variant):
    """Lists shorter than the neighbour count make the force kernel take its filtering walk (staged)."""
    # rm: one record per particle is never enough -> every stream overflows and the force kernel walks
    monkeypatch.setenv("SPH_B200_PAIR_CFG", "rm,128,1,48" if variant == "rm" else variant + ",128,1536,8")
    s, g, o, par = start("mini box", oracle_any)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    check_floats(g, o, par)
    o.close()


@pytest.mark.parametrize("cfg", ["tma,64,1024,64", "tma,256,3072,48", "l1,64,16,32", "l1,256,16,64",
                                 "rm,64,24,48", "rm,32,32,48", "rm,128,12,48"])
def test_other_cta_shapes_match(oracle_any, monkeypatch, cfg):
    monkeypatch.setenv("SPH_B200_PAIR_CFG", cfg)
    s, g, o, par = start("mini dense cells" if cfg.startswith("rm") else "Stiff  Dam break", oracle_any)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    check_floats(g, o, par)
    o.close()


def test_host_layer_update_and_changed_flag(oracle_any):
    """cSPH::Update uploads scn.params only when the changed flag is set (SPH_Update.cpp:19-27)."""
    s, g, o, par = start("mini box", oracle_any)
    s.Update()
    o.step(1)
    assert np.array_equal(s.getArray(False), o.get_array(0))    # inverted flag: False = positions
    par2 = par.copy()
    par2["gravity"] = (0, -3.0, 0)
    s.set_params(par2)                                          # marks changed
    o.set_params(par2)
    s.Update(2)
    o.step(2)
    vmax = max(float(np.abs(o.get_array(1)).max()), 1e-3)
    assert np.all(np.abs(s.getArray(True) - o.get_array(1)) <= 1e-4 * vmax)
    o.close()


def test_one_million_particles_full_compare(oracle_any):
    """BASELINE config 1b (shipped "Extreme box 1 M") against the oracle, one step, everything."""
    s, g, o, par = start("Extreme box 1 M", oracle_any)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    check_floats(g, o, par)
    o.close()


def sampled_neighbor_counts(pairs, cell_start, cell_end, spos, par, sample):
    """Pure-numpy restatement of the neighbour walk for a few particles (small cases only)."""
    gx, gyx, ncell = int(par["gridSize"][0][0]), int(par["gridSize_yx"][0]), int(par["numCells"][0])
    mp, h2 = int(par["maxParInCell"][0]), np.float32(par["h2"][0])
    out = []
    for i in sample:
        key = int(pairs[i, 0])
        cnt = 0
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    h = key + dz * gyx + dy * gx + dx
                    if h < 0 or h >= ncell:
                        continue
                    a, e = int(cell_start[h]), int(cell_end[h])
                    if a == 0xFFFFFFFF:
                        continue
                    e = min(e, a + mp)
                    for j in range(a, e):
                        if j == i:
                            continue
                        d = spos[i, :3] - spos[j, :3]
                        r2 = np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])
                        cnt += bool(np.float32(r2) < h2)
        out.append(cnt)
    return np.array(out, np.uint32)


def test_eight_million_particles_properties():
    """BASELINE config 2 at full size, through size-independent properties: the sorted keys ascend, the
    index array is a permutation, the cell table brackets exactly the runs of equal keys, one Drop lands
    where cSPH::Drop put it, state stays finite, and sampled neighbour counts match a numpy walk."""
    s = host.CSph(device=0)
    s.select_scene("tank 8M drop")
    s.Drop(False)
    g = s.solver()
    par = s.params
    n = g.n
    hpos, _ = s.host_arrays()
    assert np.array_equal(g.get_array(lib.SPH_POS, 0, 4096), hpos[:4096])
    g.step(3)
    pairs = g.dump(lib.DUMP_SORTED_PAIRS)
    keys = pairs[:, 0]
    assert np.all(keys[1:] >= keys[:-1])
    seen = np.zeros(n, bool)
    seen[pairs[:, 1]] = True
    assert seen.all()
    same = keys[1:] == keys[:-1]
    assert np.all(pairs[1:, 1][same] > pairs[:-1, 1][same])     # stable by original index
    cs, ce = g.dump(lib.DUMP_CELL_START), g.dump(lib.DUMP_CELL_END)
    occ = np.unique(keys)
    first = np.searchsorted(keys, occ, "left")
    last = np.searchsorted(keys, occ, "right")
    assert np.array_equal(cs[occ], first.astype(np.uint32)) and np.array_equal(ce[occ], last.astype(np.uint32))
    empty = np.ones(len(cs), bool)
    empty[occ] = False
    assert np.all(cs[empty] == 0xFFFFFFFF)
    spos = g.dump(lib.DUMP_SORTED_POS)
    dens = g.dump(lib.DUMP_DENSITY)
    assert np.isfinite(spos).all() and np.isfinite(dens).all() and (dens >= 0).all()
    wmin, wmax = par["worldMin"][0], par["worldMax"][0]
    assert np.all(spos[:, :3] >= wmin) and np.all(spos[:, :3] <= wmax)
    sample = np.random.default_rng(1).integers(0, n, 48)
    counts = g.dump(lib.DUMP_NEIGHBOR_COUNTS)
    assert np.array_equal(counts[sample], sampled_neighbor_counts(pairs, cs, ce, spos, par, sample))
    vel = g.get_array(lib.SPH_VEL)
    assert np.isfinite(vel).all()


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """Save after 5 steps, run 5 more; a fresh system resumed from the file must arrive at the same bits."""
    s = host.CSph(device=0)
    s.select_scene("mini waves")
    for _ in range(5):
        s.UpdateEmitter()
        s.Update()
    s.SaveState(tmp_path / "ck.bin")
    for _ in range(5):
        s.UpdateEmitter()
        s.Update()
    a = (s.getArray(False), s.getArray(True))
    t = host.CSph(device=0)
    t.LoadState(tmp_path / "ck.bin")
    for _ in range(5):
        t.UpdateEmitter()
        t.Update()
    assert np.array_equal(t.getArray(False), a[0]) and np.array_equal(t.getArray(True), a[1])


@pytest.mark.parametrize("variant", ["l1", "rm"])
@pytest.mark.parametrize("lam,ratio,max_par", [(1, 1.0, 16), (3, 1.25, 16), (8, 1.5, 16), (16, 1.25, 16), (16, 1.0, 64)])
def test_uniform_random_boxes(oracle_any, lam, ratio, max_par, variant, monkeypatch):
    """BASELINE config 4: uniform random boxes (documented PCG64 seed) at several occupancies, h/cell ratios and
    maxParInCell: sorted pairs, cell table and neighbour counts bit-exact, density within 1e-5.  lambda = 16 with
    maxParInCell = 16 exercises the truncating walk on about half of the cells."""
    sys.path.insert(0, str(ROOT))
    import bench_sweep
    monkeypatch.setenv("SPH_B200_PAIR_CFG", PAIR_VARIANTS[variant])
    n = 32768
    g, par, pos, vel = bench_sweep.build_system(n, lam, ratio, max_par)
    o = oracle_any.system(par)
    o.set_array(0, pos)
    o.set_array(1, vel)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    dg, do = g.dump(lib.DUMP_DENSITY), o.dump(5)
    assert np.all(np.abs(dg - do) <= REL * np.abs(do) + 1e-30)
    # the pair force on up to ~130 neighbours per particle (no floor on the velocity scale: dt is 1e-7 s here)
    vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
    assert np.all(np.abs(vg - vo) <= REL * float(np.abs(vo[:, :3]).max()))
    if lam == 16 and max_par == 16:
        assert int(np.bincount(g.dump(lib.DUMP_SORTED_PAIRS)[:, 0]).max()) > 16
    g.close()
    o.close()


def _scene_fill(par, extra, n, seed):
    """A jittered lattice at the scene's rest spacing inside its init volume (bottom-up), topped up with uniform
    random points; velocities of a few cm/s.  Not cSPH::Reset -- only the scene's parameters are under test here."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lo, hi, sp = extra[0:3].astype(np.float64), extra[3:6].astype(np.float64), float(extra[8])
    wlo, whi = par["worldMinD"][0].astype(np.float64), par["worldMaxD"][0].astype(np.float64)
    lo, hi = np.maximum(lo, wlo), np.minimum(hi, whi)
    hi = np.maximum(hi, lo + sp)
    dims = np.maximum(((hi - lo) / sp).astype(np.int64), 1)
    m = int(min(n, dims.prod()))
    k = np.arange(m)
    ix, iz, iy = k % dims[0], (k // dims[0]) % dims[2], k // (dims[0] * dims[2])
    pts = lo + (np.stack([ix, iy, iz], 1) + 0.5 + rng.uniform(-0.1, 0.1, (m, 3))) * sp
    if m < n:
        pts = np.concatenate([pts, lo + (hi - lo) * rng.random((n - m, 3))])
    pos = np.ones((n, 4), np.float32)
    pos[:, :3] = pts.astype(np.float32)
    vel = np.zeros((n, 4), np.float32)
    vel[:, :3] = rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
    return pos, vel


def test_every_reference_scene_parameter_set(oracle_any, golden_ref_scenes):
    """the parameter blocks of all 119 scenes of the reference's Scenes.xml (as its own loader
    produced them, tests/golden/ref_scenes.npz) run on the GPU: one step against the oracle on every scene, then ten
    more free-running steps that must stay finite and inside the world.  Particle counts are cut to 16K."""
    g = golden_ref_scenes
    assert len(g["live"]) == 119
    n = 16384
    for i in range(len(g["live"])):
        par = g["live"][i:i + 1].copy()
        par["numParticles"] = n
        pos, vel = _scene_fill(par, g["extra"][i], n, 1000 + i)
        s = lib.SphSystem(par, 0)
        o = oracle_any.system(par)
        for t in (s, o):
            t.set_array(lib.SPH_POS, pos)
            t.set_array(lib.SPH_VEL, vel)
        s.step(1)
        o.step(1)
        try:
            check_integers_exact(s, o)
            libm = int(par["iHmap"][0]) != 0 or int(par["rotType"][0]) != 0
            check_floats(s, o, par, REL_LIBM if libm else REL)
        except AssertionError as e:
            raise AssertionError(f"reference scene {i}: {e}") from None
        s.step(10)
        p = s.get_array(lib.SPH_POS)[:, :3]
        assert np.isfinite(p).all() and np.isfinite(s.get_array(lib.SPH_VEL)).all(), f"reference scene {i}: not finite"
        assert (p >= par["worldMin"][0] - 1e-6).all() and (p <= par["worldMax"][0] + 1e-6).all(), f"reference scene {i}: left the world"
        s.close()
        o.close()


def test_gl_interop_fails_cleanly_without_a_gl_context():
    """sph_gl_register / sph_gl_update cannot be exercised without OpenGL; what can be checked
    is that, with no GL context, registration reports an error instead of crashing and the update is a no-op.  Runs in
    a child process so that a misbehaving GL stack cannot take the test session down."""
    import subprocess
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from pibiti_b200 import host, lib\n"
        "s = host.CSph(device=0); s.select_scene('mini box'); g = s.solver()\n"
        "rc = g.lib.sph_gl_register(g.h, lib.SPH_POS, 12345)\n"
        "assert rc != 0, rc\n"
        "assert g.lib.sph_gl_update(g.h) == 0\n"
        "assert g.lib.sph_gl_register(g.h, lib.SPH_POS, 0) == 0\n"
        "s.Update(2); print('ok', rc)\n" % str(ROOT))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("variant", ["l1", "tma", "rm"])
@pytest.mark.parametrize("n", [1, 37, 1000, 4097])
def test_ragged_particle_counts(oracle_port, n, variant, monkeypatch):
    """the reference's kernels do not bounds-check and need N to be a multiple of 512; these do, for any N
    (the port oracle loops over particles, so it takes any N too)."""
    monkeypatch.setenv("SPH_B200_PAIR_CFG", PAIR_VARIANTS[variant])
    h = host.CSph(device=-1)
    par = h.scene_params(h.scene_index("mini dense cells")).copy()
    ex = h.scene_extra(h.scene_index("mini dense cells"))
    par["numParticles"] = n
    pos, vel = _scene_fill(par, ex, n, 77 + n)
    h.close()
    g = lib.SphSystem(par, 0)
    o = oracle_port.system(par)
    for t in (g, o):
        t.set_array(lib.SPH_POS, pos)
        t.set_array(lib.SPH_VEL, vel)
    for _ in range(3):
        g.step(1)
        o.step(1)
        check_integers_exact(g, o)
        check_floats(g, o, par)
        g.set_array(lib.SPH_POS, o.get_array(0))          # resync: the next step starts from the oracle's floats
        g.set_array(lib.SPH_VEL, o.get_array(1))
    g.close()
    o.close()


# ---- round 2 additions ---------------------------------------------------------------------------------------

def test_particle_count_change_after_a_step_keeps_the_first_particles(oracle_any):
    """sph_set_params with a smaller numParticles after a step: slots are in sorted order by then, so the library first
    restores original order; the survivors must be exactly particles [0, m) -- what the reference, which simply runs its
    kernels over the first numParticles entries, would keep -- and every row of get_array is written.  Growing back
    re-exposes the old rows once, never a duplicate."""
    s, g, o, par = start("mini box", oracle_any)
    g.step(2)
    pos_full, vel_full = g.get_array(lib.SPH_POS), g.get_array(lib.SPH_VEL)
    n, m = g.n, g.n // 2 + 512                             # the reference kernels need a multiple of 512
    small = par.copy()
    small["numParticles"] = m
    g.set_params(small)
    g.n = m
    pg, vg = g.get_array(lib.SPH_POS), g.get_array(lib.SPH_VEL)
    assert np.array_equal(pg, pos_full[:m]) and np.array_equal(vg, vel_full[:m])
    o2 = oracle_any.system(small)
    o2.set_array(0, pos_full[:m])
    o2.set_array(1, vel_full[:m])
    g.step(1)
    o2.step(1)
    check_integers_exact(g, o2)
    check_floats(g, o2, small)
    g.set_params(par)                                       # grow back to the allocation
    g.n = n
    ids_pos = g.get_array(lib.SPH_POS)
    assert np.array_equal(ids_pos[:m], o2.get_array(0)) and np.array_equal(ids_pos[m:], pos_full[m:])
    o.close()
    o2.close()


@pytest.mark.parametrize("title", ["mini pump square", "mini pump S"])
def test_pump_boundary_and_exit_teleport(oracle_any, title):
    """Pump boundary (System.cu:100-158): cylinder frame, outlet box, inlet hole, and the exit -> inlet teleport.
    Twenty resynchronised steps on the scene as Reset fills it, then particles are put just inside the exit with an
    outward velocity so that the teleport branch runs: the same particles must jump, to the same place (the inlet
    position goes through sinf/cosf per particle: 1e-6 of the world size), everything else bit for bit."""
    s, g, o, par = start(title, oracle_any)
    for _ in range(20):
        s.UpdateEmitter()                                   # rotor angle
        p = s.params
        g.set_params(p)
        o.set_params(p)
        g.step(1)
        o.step(1)
        check_integers_exact(g, o)
        check_floats(g, o, p, REL_LIBM)                     # rotor spheres placed with sinf/cosf
        g.set_array(lib.SPH_POS, o.get_array(0))
        g.set_array(lib.SPH_VEL, o.get_array(1))
    p = s.params
    pos, vel = o.get_array(0), o.get_array(1)
    wmax, wmin = p["worldMax"][0], p["worldMin"][0]
    k = 96
    rng = np.random.Generator(np.random.PCG64(11))
    pos[:k, 0] = rng.uniform(-0.02, 0.02, k) if float(p["angOut"][0]) < 0.5 else rng.uniform(0.09, 0.14, k)
    # inside the exit strip but outside the soft zone of the +y wall, and below the inlet hole's frame: a soft-boundary
    # push would take vel.y back under rVexit before the teleport test (System.cu:140)
    pos[:k, 1] = wmax[1] - float(p["rDexit"][0]) * rng.uniform(0.82, 0.98, k)
    pos[:k, 2] = rng.uniform(wmin[2] * 0.8, float(p["hClose"][0]) - 0.035, k)
    vel[:k, :3] = 0
    vel[:k, 1] = float(p["rVexit"][0]) + rng.uniform(0.2, 1.0, k)
    for q in (g, o):
        q.set_array(0, pos)
        q.set_array(1, vel)
    g.step(1)
    o.step(1)
    pg, po = g.get_array(lib.SPH_POS), o.get_array(0)
    jumped_o = np.abs(po[:k, 1] - pos[:k, 1]) > 0.05
    jumped_g = np.abs(pg[:k, 1] - pos[:k, 1]) > 0.05
    assert jumped_o.sum() >= k // 2, "the fixture should reach the teleport branch"
    assert np.array_equal(jumped_o, jumped_g)
    extent = float(np.abs(p["worldSize"][0]).max())
    assert np.abs(pg[:, :3] - po[:, :3]).max() <= 1e-6 * extent
    rest = np.ones(g.n, bool)
    rest[:k] = False
    assert np.array_equal(pg[rest], po[rest]), "particles that did not teleport are bit-exact"
    vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
    assert np.all(np.abs(vg - vo) <= REL_LIBM * max(float(np.abs(vo[:, :3]).max()), 1e-3))
    o.close()


def test_scene_switch_reuses_the_device_buffers(oracle_any):
    """cSPH::InitScene keeps the solver handle when the next scene fits its buffers (the reference frees and reallocates
    everything, SPH_Scenes.cpp:9-13): same handle, and a step after the switch still meets the one-step bar."""
    s = host.CSph(device=0)
    s.select_scene("box small default")                     # 57K particles: the largest of the three
    h0 = s.solver().h.value
    s.Update(3)
    for title in ("mini waves", "Stiff  Dam break", "mini dense cells"):
        s.select_scene(title)
        assert s.solver().h.value == h0, "switching to a smaller scene must not reallocate"
        par = s.params
        pos, vel = s.host_arrays()
        g = s.solver()
        o = oracle_any.system(par)
        o.set_array(0, pos)
        o.set_array(1, vel)
        assert np.array_equal(g.get_array(lib.SPH_POS), pos)
        g.step(1)
        o.step(1)
        check_integers_exact(g, o)
        check_floats(g, o, par)
        o.close()
    s.select_scene("Extreme box 1 M")                       # does not fit: new buffers
    assert s.n == 1024 * 1024 and s.solver().h.value is not None
    s.Update(1)
    assert np.isfinite(s.getArray(True)).all()


def test_checkpoint_resumes_into_an_object_on_another_scene(tmp_path):
    """The targets the per-step prologue drags the collider and the dye source towards (App::colliderPos / dyePos) travel
    with the checkpoint: resuming inside an object that sits on a scene with another collider position continues
    bit-identically."""
    s = host.CSph(device=0)
    s.select_scene("mini collider accel")
    s.set_targets(np.array([0.03, -0.05, 0.02, 0], np.float32), np.array([0.01, -0.06, 0.0], np.float32), None)
    for _ in range(4):
        s.UpdateEmitter()
        s.Update()
    s.SaveState(tmp_path / "ck.bin")
    for _ in range(6):
        s.UpdateEmitter()
        s.Update()
    a = (s.getArray(False), s.getArray(True), s.params.tobytes())
    t = host.CSph(device=0)
    t.select_scene("mini rotor Z")                          # different collider, rotor on
    t.Update(2)
    t.LoadState(tmp_path / "ck.bin")
    for _ in range(6):
        t.UpdateEmitter()
        t.Update()
    assert np.array_equal(t.getArray(False), a[0]) and np.array_equal(t.getArray(True), a[1]) and t.params.tobytes() == a[2]
    bad = tmp_path / "bad.bin"
    raw = bytearray((tmp_path / "ck.bin").read_bytes())
    raw[16:20] = (12345).to_bytes(4, "little")              # sizeof(Scene) guard
    bad.write_bytes(bytes(raw))
    with pytest.raises(Exception):
        t.LoadState(bad)


@pytest.mark.parametrize("variant", ["rm", "l1"])
def test_pile_up_cells_take_the_big_cell_rank_path(oracle_any, monkeypatch, variant):
    """Cells with more than 64 entries are ranked by a whole CTA each (k_rank_big_cells) instead of the per-entry counting
    loop: a pile-up of 700 and one of 90 particles inside single cells must still give the reference's stable order,
    the truncated neighbour walk (maxParInCell) and densities / velocities within the bar."""
    monkeypatch.setenv("SPH_B200_PAIR_CFG", PAIR_VARIANTS[variant])
    s, g, o, par = start("mini box", oracle_any)
    pos, vel = s.host_arrays()
    rng = np.random.Generator(np.random.PCG64(5))
    cs = np.asarray(par["cellSize"][0], np.float32)
    wmin = np.asarray(par["worldMin"][0], np.float32)
    for first, count, cell in ((1000, 700, (9, 7, 9)), (4000, 90, (12, 7, 9))):
        lo = wmin + cs * np.asarray(cell, np.float32)
        pos[first:first + count, :3] = (lo + cs * rng.uniform(0.05, 0.95, (count, 3))).astype(np.float32)
    for q in (g, o):
        q.set_array(0, pos)
        q.set_array(1, vel)
    g.step(1)
    o.step(1)
    check_integers_exact(g, o)
    assert int(np.bincount(g.dump(lib.DUMP_SORTED_PAIRS)[:, 0]).max()) > 64
    dg, do = g.dump(lib.DUMP_DENSITY), o.dump(5)
    assert np.all(np.abs(dg - do) <= REL * np.abs(do) + 1e-30)
    vg, vo = g.get_array(lib.SPH_VEL), o.get_array(1)
    assert np.all(np.abs(vg - vo) <= REL * max(float(np.abs(vo[:, :3]).max()), 1e-3))
    o.close()


def test_exchange_arrays_equals_get_then_set(oracle_any):
    """cSPH::exchangeArrays (downloads overlapping uploads) == getArray(pos), getArray(vel), setArray(pos), setArray(vel):
    same data out, same trajectory afterwards -- on a stepped (sorted) state, with pinned and with pageable buffers."""
    import torch
    def run(use_exchange, pinned):
        s = host.CSph(device=0)
        s.select_scene("mini waves")
        s.Update(3)
        n = s.n
        rng = np.random.Generator(np.random.PCG64(9))
        new_pos = s.getArray(False).copy()
        new_pos[:, :3] += rng.uniform(-1e-4, 1e-4, (n, 3)).astype(np.float32)
        new_vel = (s.getArray(True) * np.float32(0.5)).astype(np.float32)
        mk = (lambda a: torch.from_numpy(a.copy()).pin_memory()) if pinned else (lambda a: torch.from_numpy(a.copy()))
        ip, iv = mk(new_pos), mk(new_vel)
        op, ov = mk(np.zeros((n, 4), np.float32)), mk(np.zeros((n, 4), np.float32))
        if use_exchange:
            s.exchangeArrays(op.data_ptr(), ov.data_ptr(), ip.data_ptr(), iv.data_ptr())
            outs = op.numpy().copy(), ov.numpy().copy()
        else:
            outs = s.getArray(False).copy(), s.getArray(True).copy()
            s.setArray(False, new_pos)
            s.setArray(True, new_vel)
        assert np.array_equal(s.getArray(False), new_pos) and np.array_equal(s.getArray(True), new_vel)
        s.Update(2)
        res = outs + (s.getArray(False).copy(), s.getArray(True).copy())
        s.close()
        return res
    ref = run(False, False)
    for pinned in (True, False):
        got = run(True, pinned)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
