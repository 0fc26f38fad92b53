#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own code (oracle/_ref/libsphref.so).

Run here (where /root/reference exists):  python tests/golden/make_golden.py

Writes
  ref_scenes.npz    every <Scene> of the reference's own Scenes.xml through the reference's scene
                    loader: SimParams blocks, the non-SimParams scene fields, and SHA-256 digests of
                    the particle arrays cSPH::Reset produces (plus Drop).
  repo_scenes.npz   the same for this repository's scenes/Scenes.xml -- i.e. what the reference's
                    loader derives from OUR scene file.
  steps.npz         for a set of scenes of scenes/Scenes.xml: the reference kernels (host-compiled)
                    stepped from the Reset state; digests of every integer/float output after 1 and
                    5 steps, plus strided float samples for tolerance comparisons.

Uninitialised bytes of the reference's SimParams (ff2 and the struct padding; Scene.cpp never
sets them) are zeroed before storing so that the fixtures are reproducible.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc                      # noqa: E402
from pibiti_b200.lib import SIMPARAMS_DTYPE           # noqa: E402

REFERENCE_DIR = Path("/root/reference")
OUT = Path(__file__).resolve().parent
SAMPLE_STRIDE = 64
MAX_RESET_PARTICLES = 20_000_000
STEP_SCENES = ["box small default", "Stiff  Dam break", "mini box", "mini dense cells", "mini random",
               "mini cylinder Y", "mini cylinder Z", "mini sphere", "mini wrap Z", "mini cycle Z", "mini waves",
               "mini collider accel", "mini heightmap XZ", "mini heightmap YZ holes", "mini rotor Z", "mini rotor Y",
               "mini propeller pair", "mini pump square", "mini pump S"]


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def clean_params(block: np.ndarray) -> np.ndarray:
    """Copy only the named fields (drops padding) and zero ff2."""
    out = np.zeros(1, SIMPARAMS_DTYPE)
    for name in SIMPARAMS_DTYPE.names:
        out[name] = block[name]
    out["ff2"] = 0
    return out


def scene_titles(xml: Path) -> list[str]:
    import re
    text = re.sub(r"<!--.*?-->", "", xml.read_text(errors="replace"), flags=re.S)
    return re.findall(r"<Scene\s+name=\"([^\"]*)\"", text)


def scene_table(L, xml_dir: Path) -> dict:
    n = L.refh_load(str(xml_dir).encode())
    start_scene = int(L.refh_cur_scene())                 # the scene LoadScenes starts on
    params = np.zeros(n, SIMPARAMS_DTYPE)
    live = np.zeros(n, SIMPARAMS_DTYPE)
    extra = np.zeros((n, 64), np.float32)
    reset_sha, drop_sha, head = [], [], np.zeros((n, 8, 4), np.float32)
    counts = np.zeros(n, np.int64)
    for i in range(n):
        blk = np.zeros(1, SIMPARAMS_DTYPE)
        L.refh_scene_params(i, vp(blk))
        params[i] = clean_params(blk)[0]
        L.refh_scene_extra(i, vp(extra[i]))
    for i in range(n):
        counts[i] = int(params[i]["numParticles"])
        if counts[i] > MAX_RESET_PARTICLES:               # 32M/64M tanks: params only (host memory)
            live[i] = params[i]
            reset_sha.append("skipped")
            drop_sha.append("skipped")
            continue
        npar = L.refh_select_scene(i)                     # srand(1); UpdScene -> InitScene -> Reset
        assert npar == counts[i]
        blk = np.zeros(1, SIMPARAMS_DTYPE)
        L.refh_live_params(vp(blk))
        live[i] = clean_params(blk)[0]
        pos = np.zeros((npar, 4), np.float32)
        vel = np.zeros((npar, 4), np.float32)
        L.refh_get_host(vp(pos), vp(vel))
        reset_sha.append(sha(pos) + sha(vel))
        head[i] = pos[:8]
        # one fixed and one random drop (xyz only: the reference leaves w undefined, SPH_Init.cpp:109)
        L.refh_srand(7)
        L.refh_drop(0)
        e = L.refh_drop(1)
        L.refh_get_host(vp(pos), vp(vel))
        drop_sha.append(sha(pos[:, :3]) + sha(vel) + f"{e:08x}")
    return dict(cur_scene=np.int64(start_scene), params=params, live=live, extra=extra,
                num_particles=counts, reset_sha=np.array(reset_sha), drop_sha=np.array(drop_sha), reset_head=head)


def step_vectors(L, O, xml_dir: Path) -> dict:
    titles = scene_titles(xml_dir / "Scenes.xml")
    L.refh_load(str(xml_dir).encode())
    out = {"titles": np.array(STEP_SCENES)}
    for title in STEP_SCENES:
        idx = titles.index(title)
        npar = L.refh_select_scene(idx)
        par = np.zeros(1, SIMPARAMS_DTYPE)
        L.refh_live_params(vp(par))
        par = clean_params(par)
        pos = np.zeros((npar, 4), np.float32)
        vel = np.zeros((npar, 4), np.float32)
        L.refh_get_host(vp(pos), vp(vel))
        sysm = O.system(par)
        sysm.set_array(0, pos)
        sysm.set_array(1, vel)
        key = title.replace(" ", "_")
        out[f"{key}/params"] = par
        done = 0
        for steps in (1, 5):
            sysm.step(steps - done)
            done = steps
            items = {"pairs": sysm.dump(0), "cellStart": sysm.dump(1), "sortedPos": sysm.dump(2), "sortedVel": sysm.dump(3),
                     "pressure": sysm.dump(4), "density": sysm.dump(5), "counts": sysm.dump(6), "color": sysm.dump(7),
                     "dye": sysm.dump(8), "pos": sysm.get_array(0), "vel": sysm.get_array(1)}
            for name, arr in items.items():
                out[f"{key}/{steps}/{name}_sha"] = np.array(sha(arr))
                if arr.dtype == np.float32:
                    out[f"{key}/{steps}/{name}_sample"] = arr[::SAMPLE_STRIDE].copy()
            out[f"{key}/{steps}/counts_hist"] = np.bincount(items["counts"], minlength=64).astype(np.int64)
            out[f"{key}/{steps}/max_cell"] = np.int64(np.bincount(items["pairs"][:, 0]).max())
        sysm.close()
        print("  steps:", title, npar)
    return out


def main():
    if not orc.available("reference"):
        sys.exit("oracle/_ref/libsphref.so missing: run `python -m pibiti_b200.build` where /root/reference exists")
    O = orc.load("reference")
    L = O.L
    L.refh_load.argtypes = [C.c_char_p]
    print("reference Scenes.xml ...")
    np.savez_compressed(OUT / "ref_scenes.npz", **scene_table(L, REFERENCE_DIR))
    print("repo scenes/Scenes.xml ...")
    np.savez_compressed(OUT / "repo_scenes.npz", **scene_table(L, ROOT / "scenes"))
    print("step vectors ...")
    np.savez_compressed(OUT / "steps.npz", **step_vectors(L, O, ROOT / "scenes"))
    for f in ("ref_scenes.npz", "repo_scenes.npz", "steps.npz"):
        print(f, (OUT / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
