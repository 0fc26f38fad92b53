"""BASELINE config 1 / SURVEY.md 8d drift protocol on the two scenes the reference ships for it: 100 and 1000
free-running steps, and a run re-synchronised every 50 steps.  Needs a B200: `-m gpu`."""
from __future__ import annotations

import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("title", ["box small default", "Stiff  Dam break"])
def test_drift_protocol(oracle_any, title):
    sys.path.insert(0, str(ROOT))
    import bench_configs
    r = bench_configs.drift_protocol(title, steps=1000, resync_every=50, oracle=oracle_any)
    # re-synchronised: the one-step bar (integers bit-exact, rho and v within 1e-5) holds at every one of the 20 checks
    assert r["resync"]["checks"] == 20 and r["resync"]["failures"] == 0, r["resync"]
    f100, f1000 = r["free_running"]["100"], r["free_running"]["1000"]
    assert f100["finite"] and f1000["finite"]
    # 100 steps: at least 99 % of the particles within half a lattice spacing of the CPU trajectory
    assert f100["share_within_half_spacing"] >= 0.99, f100
    # 1000 steps: SPH is chaotic, so aggregates: centre of mass, kinetic energy, density distribution, maximum height
    assert f1000["com_diff_over_world"] <= 0.01, f1000
    assert f1000["density_mean_rel_diff"] <= 0.01 and f1000["density_quantiles_rel_diff"] <= 0.01, f1000
    assert f1000["max_height_rel_diff"] <= 0.01, f1000
    assert f1000["ke_rel_diff"] <= 0.01 or f1000["share_within_half_spacing"] >= 0.99, f1000
