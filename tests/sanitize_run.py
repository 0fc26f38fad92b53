#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck), run on the GPU box as
  compute-sanitizer --tool memcheck python tests/sanitize_run.py
Covers the places where a stray access would hide: the predicated inline-PTX list store of the density kernel,
the pair-kernel variants (row-mask records incl. record overflow, index lists, TMA-staged), list overflow (walk fallback), the truncating walk ("mini dense cells"), the host
accessors' permutation kernels and a 3-slab step with device-count-bounded kernels (all ranks in this process)."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def single(cfg: str, title: str, steps: int = 3):
    from pibiti_b200 import host, lib
    os.environ["SPH_B200_PAIR_CFG"] = cfg
    s = host.CSph(device=0)
    s.select_scene(title)
    g = s.solver()
    for _ in range(steps):
        s.UpdateEmitter()
        s.Update()
    g.dump(lib.DUMP_NEIGHBOR_COUNTS)
    v = s.getArray(True)
    assert np.isfinite(v).all()
    s.setArray(True, v[:100], 17)
    s.Update()
    s.close()
    print("ok", cfg, title, flush=True)


def slabs():
    import test_slab                      # the 3-slab LocalComm fixture of the test-suite
    test_slab.run_gpu_slabs_for_sanitizer()


def multi():
    """The C++ multi-GPU driver: three slabs on this GPU (peer-copy exchange), stirred particles, a re-cut in between --
    the kernels that find their ranges in device words and are launched over upper bounds."""
    import test_multi
    got, _, info = test_multi.multi_run("mini waves", 5, 3, recut_every=2)
    assert np.isfinite(got[1]).all() and info["recuts"] == 2
    print("ok multi", info, flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["single", "slab", "multi"]
    if "single" in what:
        for cfg in os.environ.get("SPH_SANITIZE_CFGS", "rm,128,32,48;rm,128,2,48;l1,128,1344,48;tma,128,1344,48;l1,128,1344,8").split(";"):
            for title in ("mini dense cells", "mini waves"):
                single(cfg, title)
    if "slab" in what:
        slabs()
    if "multi" in what:
        multi()
    print("sanitize_run done")
