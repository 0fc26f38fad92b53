"""The reference arm of bench.py runs on the host alone, so its JSON line can be checked without a GPU."""
from __future__ import annotations

import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(env_extra=None, *args):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "mini box",
                        "--steps", "2", "--warmup", "1", *args], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    (line,) = _run()
    d = json.loads(line)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "particle-updates/sec per SPH step" and d["unit"] == "particle-updates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["workload"] == "mini box"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm times the reference's own code from the reference's own initial state: the product library is not even mapped
    if cb["kind"] == "reference":
        assert d["native_so_loaded"] == ["oracle/_ref/libsphref.so"], d["native_so_loaded"]
        assert d["config"]["initial_state"].startswith("reference scene code")


def test_reference_arm_under_torchrun_only_rank_zero_prints():
    assert len(_run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")) == 1
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2") == []


def test_state_checksum_is_order_independent_and_bit_sensitive():
    """bench.py's full-size multi-GPU parity compares an order-independent checksum of (id, pos, vel, rho): equal for any
    distribution of the same particles over slabs, different as soon as one bit of one particle differs."""
    import numpy as np
    sys.path.insert(0, str(ROOT))
    import bench
    rng = np.random.default_rng(3)
    n = 5000
    ids = np.arange(n, dtype=np.uint32)
    pos, vel = rng.random((n, 4), np.float32), rng.random((n, 4), np.float32)
    dens = rng.random(n, np.float32)
    whole = bench.state_checksum(ids, pos, vel, dens)
    perm = rng.permutation(n)
    parts = np.array_split(perm, 3)
    split = sum(bench.state_checksum(ids[p], pos[p], vel[p], dens[p]) for p in parts) & 0xFFFFFFFFFFFFFFFF
    assert split == whole
    v2 = vel.copy()
    v2.view(np.uint32)[1234, 2] ^= 1                      # one ulp of one component of one particle
    assert bench.state_checksum(ids, pos, v2, dens) != whole
    swapped = ids.copy()
    swapped[[10, 11]] = swapped[[11, 10]]                 # two particles exchanging their states
    assert bench.state_checksum(swapped, pos, vel, dens) != whole
