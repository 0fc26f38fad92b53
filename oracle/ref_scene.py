"""TEST INFRASTRUCTURE ONLY -- scenes, initial particle state and the per-step host prologue from the
REFERENCE's own host code (Scene.cpp, Scene_Load.cpp, SPH_Init.cpp, SPH_Scenes.cpp, App/Update.cpp), as compiled
into oracle/_ref/libsphref.so by oracle/build_ref.sh.

bench.py's `--impl reference` arm uses this so that nothing of the product (pibiti_b200/) is loaded into the
process that times the reference.  Needs the prebuilt library (it travels to the GPU box); there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_LIB = _HERE / "_ref" / "libsphref.so"

# include/sph_params.h offsets the arm needs (560-byte SimParams block; checked against the library's sizeof)
_OFF_NUM_PARTICLES = 4


def scene_titles(xml: Path) -> list[str]:
    """Titles in file order.  The reference does not read `name` on Linux (Scene_Load.cpp:37-39)."""
    text = re.sub(r"<!--.*?-->", "", Path(xml).read_text(errors="replace"), flags=re.S)
    return re.findall(r"<Scene\s+name=\"([^\"]*)\"", text)


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """The reference's cSPH object (host part) on one scene of a Scenes.xml."""

    def __init__(self, xml_dir: Path, title: str):
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} not built (oracle/build_ref.sh needs the reference tree)")
        self.L = L = C.CDLL(str(REF_LIB))
        if not hasattr(L, "refh_update_emitter"):
            raise RuntimeError("libsphref.so predates refh_update_emitter: rebuild it where the reference tree exists")
        assert L.orc_sizeof_params() == 560
        titles = scene_titles(Path(xml_dir) / "Scenes.xml")
        if L.refh_load(str(xml_dir).encode()) != len(titles):
            raise RuntimeError("scene count of the reference loader and of the title scan differ")
        self.n = int(L.refh_select_scene(titles.index(title)))          # srand(1); UpdScene -> InitScene -> Reset
        self.title = title

    def params(self) -> np.ndarray:
        """Live SimParams as a 560-byte block (uint8)."""
        blk = np.zeros(560, np.uint8)
        self.L.refh_live_params(_vp(blk))
        return blk

    def arrays(self):
        pos, vel = np.zeros((self.n, 4), np.float32), np.zeros((self.n, 4), np.float32)
        self.L.refh_get_host(_vp(pos), _vp(vel))
        return pos, vel

    def drop(self, random: bool = False) -> int:
        return int(self.L.refh_drop(int(random)))

    def update_emitter(self) -> np.ndarray:
        """App::UpdateEmitter (wave / rotor phase, lags, emitters, rain); returns the new parameter block."""
        self.L.refh_update_emitter()
        return self.params()
