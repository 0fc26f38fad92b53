"""ctypes binding of the C++ host layer (include/sph_host_c.h): the cSPH-shaped system object with
its Scenes.xml loader, Reset / Drop initialisers and the per-step host prologue.

`CSph` mirrors the reference's `class cSPH` (source/SPH/SPH.h:9-50): same method names and meaning
(`Update`, `Reset`, `Drop`, `NextScene`, `PrevScene`, `getArray(bool)` with its inverted flag,
`setArray`).  All logic lives in the C++ library; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import lib as _lib
from .lib import SIMPARAMS_DTYPE, SphError

DEFAULT_SCENES_XML = Path(__file__).resolve().parent.parent / "scenes" / "Scenes.xml"

HOST_ABI_SYMBOLS = (
    "sphh_create", "sphh_destroy", "sphh_last_error", "sphh_num_scenes", "sphh_cur_scene", "sphh_scene_title",
    "sphh_scene_params", "sphh_scene_extra", "sphh_live_extra", "sphh_live_params", "sphh_set_live_params",
    "sphh_select_scene", "sphh_add_scene_xml", "sphh_next_scene", "sphh_prev_scene", "sphh_host_arrays",
    "sphh_reset", "sphh_drop", "sphh_emit_id", "sphh_srand", "sphh_update_emitter", "sphh_update",
    "sphh_mark_changed", "sphh_get_array", "sphh_set_array", "sphh_solver", "sphh_load_options",
    "sphh_save_state", "sphh_load_state", "sphh_set_targets", "sphh_set_emitter", "sphh_cnt_rain", "sphh_changed_flag",
    "sphh_register_gl", "sphh_pos_buffer", "sphh_timer_fps", "sphh_create_multi", "sphh_multi_solver",
    "sphh_exchange_arrays",
)

EXTRA_FIELDS = ("initMin", "initMax", "initType", "initLast", "spacing", "fCellSize", "dropR", "rain", "rVel", "r2Vel",
                "camPos", "camRot", "bChapter")

_bound = False


def _bind():
    global _bound
    L = _lib.load()
    if _bound:
        return L
    vp, ci, cs = C.c_void_p, C.c_int, C.c_char_p
    L.sphh_create.restype = vp;            L.sphh_create.argtypes = [cs, ci]
    L.sphh_create_multi.restype = vp;      L.sphh_create_multi.argtypes = [cs, C.POINTER(ci), ci]
    L.sphh_multi_solver.restype = vp;      L.sphh_multi_solver.argtypes = [vp]
    L.sphh_exchange_arrays.argtypes = [vp, vp, vp, vp, vp]
    L.sphh_destroy.argtypes = [vp]
    L.sphh_last_error.restype = cs;        L.sphh_last_error.argtypes = [vp]
    L.sphh_num_scenes.argtypes = [vp];     L.sphh_cur_scene.argtypes = [vp]
    L.sphh_scene_title.restype = cs;       L.sphh_scene_title.argtypes = [vp, ci]
    L.sphh_scene_params.argtypes = [vp, ci, vp]
    L.sphh_scene_extra.argtypes = [vp, ci, vp]
    L.sphh_live_extra.argtypes = [vp, vp]
    L.sphh_live_params.argtypes = [vp, vp]
    L.sphh_set_live_params.argtypes = [vp, vp]
    L.sphh_select_scene.argtypes = [vp, ci]
    L.sphh_add_scene_xml.argtypes = [vp, cs]
    L.sphh_next_scene.argtypes = [vp, ci]; L.sphh_prev_scene.argtypes = [vp, ci]
    L.sphh_host_arrays.argtypes = [vp, vp, vp]
    L.sphh_reset.argtypes = [vp, ci]
    L.sphh_drop.argtypes = [vp, ci]
    L.sphh_emit_id.argtypes = [vp]
    L.sphh_srand.argtypes = [C.c_uint]
    L.sphh_update_emitter.argtypes = [vp]
    L.sphh_update.argtypes = [vp, ci]
    L.sphh_mark_changed.argtypes = [vp]
    L.sphh_get_array.argtypes = [vp, ci, vp]
    L.sphh_set_array.argtypes = [vp, ci, vp, ci, ci]
    L.sphh_solver.restype = vp;            L.sphh_solver.argtypes = [vp]
    L.sphh_load_options.argtypes = [cs, vp]
    L.sphh_save_state.argtypes = [vp, cs]
    L.sphh_load_state.argtypes = [vp, cs]
    L.sphh_set_targets.argtypes = [vp, vp, vp, vp]
    L.sphh_set_emitter.argtypes = [vp, ci, vp, vp, C.c_float, ci, ci]
    L.sphh_cnt_rain.argtypes = [vp]
    L.sphh_changed_flag.argtypes = [vp, ci]
    L.sphh_register_gl.argtypes = [vp, C.c_uint, C.c_uint]
    L.sphh_pos_buffer.argtypes = [vp];     L.sphh_pos_buffer.restype = C.c_uint
    L.sphh_timer_fps.argtypes = [vp];      L.sphh_timer_fps.restype = C.c_double
    _bound = True
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class _SolverView(_lib.SphSystem):
    """SphSystem methods over a handle owned by a cSPH object (never destroys it)."""

    def __init__(self, handle, params):  # noqa: super().__init__ deliberately not called
        self.lib = _lib.load()
        self.h = C.c_void_p(handle)
        self.params = params
        self.n = int(params["numParticles"][0])
        self.num_cells = int(params["numCells"][0])

    def close(self):
        self.h = C.c_void_p()


class CSph:
    """The reference's `cSPH`, headless.  device < 0: scene layer only (no GPU is touched)."""

    def __init__(self, scenes_xml: str | Path = DEFAULT_SCENES_XML, device: int = 0, seed: int | None = 1, devices=None):
        """devices: a list of GPUs -> the system is cut into z slabs, one per device (cSPH's device-list constructor)."""
        self.L = _bind()
        if seed is not None:
            self.L.sphh_srand(seed)          # the reference never seeds rand(): glibc's default seed is 1
        self.devices = list(devices) if devices is not None and len(devices) > 1 else None
        if self.devices:
            arr = (C.c_int * len(self.devices))(*self.devices)
            self.h = C.c_void_p(self.L.sphh_create_multi(str(scenes_xml).encode(), arr, len(self.devices)))
            device = self.devices[0]
        else:
            self.h = C.c_void_p(self.L.sphh_create(str(scenes_xml).encode(), device))
        if not self.h:
            raise SphError("sphh_create failed")
        self.device = device
        self._check_solver()

    def _check_solver(self):
        if self.devices:
            if not self.L.sphh_multi_solver(self.h):
                raise SphError("no multi-GPU solver: " + self.last_error())
        elif self.device >= 0 and not self.L.sphh_solver(self.h):
            raise SphError("no solver: " + self.last_error())

    def close(self):
        if getattr(self, "h", None):
            self.L.sphh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return (self.L.sphh_last_error(self.h) or b"").decode()

    # ---- scenes ---------------------------------------------------------------------------------
    @property
    def num_scenes(self) -> int:
        return self.L.sphh_num_scenes(self.h)

    @property
    def curScene(self) -> int:
        return self.L.sphh_cur_scene(self.h)

    def scene_title(self, idx: int) -> str:
        return self.L.sphh_scene_title(self.h, idx).decode(errors="replace")

    def scene_index(self, title: str) -> int:
        for i in range(self.num_scenes):
            if self.scene_title(i) == title:
                return i
        raise KeyError(title)

    def scene_params(self, idx: int) -> np.ndarray:
        a = np.zeros(1, SIMPARAMS_DTYPE)
        self.L.sphh_scene_params(self.h, idx, _p(a))
        return a

    def scene_extra(self, idx: int | None = None) -> np.ndarray:
        a = np.zeros(64, np.float32)
        if idx is None:
            self.L.sphh_live_extra(self.h, _p(a))
        else:
            self.L.sphh_scene_extra(self.h, idx, _p(a))
        return a

    @property
    def params(self) -> np.ndarray:
        """Copy of scn.params (the live scene)."""
        a = np.zeros(1, SIMPARAMS_DTYPE)
        self.L.sphh_live_params(self.h, _p(a))
        return a

    def set_params(self, params: np.ndarray):
        """scn.params = params, and mark them changed (slider edit + ParamBase::Changed())."""
        a = np.ascontiguousarray(params).copy()
        self.L.sphh_set_live_params(self.h, _p(a))

    def select_scene(self, idx_or_title, seed: int | None = 1) -> int:
        idx = self.scene_index(idx_or_title) if isinstance(idx_or_title, str) else idx_or_title
        if seed is not None:
            self.L.sphh_srand(seed)
        n = self.L.sphh_select_scene(self.h, idx)
        if n < 0:
            raise SphError(f"no scene {idx_or_title}")
        self._check_solver()
        return n

    def add_scene_xml(self, element_text: str) -> int:
        idx = self.L.sphh_add_scene_xml(self.h, element_text.encode())
        if idx < 0:
            raise SphError("cannot parse scene element")
        return idx

    def NextScene(self, chapter: bool = False):
        self.L.sphh_next_scene(self.h, int(chapter))
        self._check_solver()

    def PrevScene(self, chapter: bool = False):
        self.L.sphh_prev_scene(self.h, int(chapter))
        self._check_solver()

    # ---- particles ------------------------------------------------------------------------------
    @property
    def n(self) -> int:
        return int(self.params["numParticles"][0])

    def host_arrays(self):
        n = self.n
        pos, vel = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        self.L.sphh_host_arrays(self.h, _p(pos), _p(vel))
        return pos, vel

    def Reset(self, type_: int = 0):
        self.L.sphh_reset(self.h, type_)

    def Drop(self, bRandom: bool = False) -> int:
        return self.L.sphh_drop(self.h, int(bRandom))

    @property
    def emitId(self) -> int:
        return self.L.sphh_emit_id(self.h)

    def srand(self, seed: int):
        self.L.sphh_srand(seed)

    def UpdateEmitter(self):
        self.L.sphh_update_emitter(self.h)

    def set_targets(self, collider=None, dye=None, acc=None):
        a = [None if v is None else np.ascontiguousarray(v, np.float32) for v in (collider, dye, acc)]
        self.L.sphh_set_targets(self.h, *[None if v is None else _p(v) for v in a])

    def set_emitter(self, e: int, pos_lag, rot_lag, vel: float, size: int, size2: int = 0):
        p, r = np.ascontiguousarray(pos_lag, np.float32), np.ascontiguousarray(rot_lag, np.float32)
        self.L.sphh_set_emitter(self.h, e, _p(p), _p(r), vel, size, size2)

    @property
    def cntRain(self) -> int:
        return self.L.sphh_cnt_rain(self.h)

    def changed_flag(self, clear: bool = False) -> bool:
        return bool(self.L.sphh_changed_flag(self.h, int(clear)))

    def Update(self, nsteps: int = 1):
        rc = self.L.sphh_update(self.h, nsteps)
        if rc != 0:
            raise SphError(f"cSPH::Update failed ({rc}): {self.last_error()}")

    def getArray(self, pos: bool) -> np.ndarray:
        """NB the reference's inverted flag: getArray(False) -> positions, getArray(True) -> velocities."""
        out = np.empty((self.n, 4), np.float32)
        if self.L.sphh_get_array(self.h, int(pos), _p(out)) != 0:
            raise SphError("getArray: not initialised")
        return out

    def setArray(self, pos: bool, data: np.ndarray, start: int = 0):
        data = np.ascontiguousarray(data, np.float32).reshape(-1, 4)
        self.L.sphh_set_array(self.h, int(pos), _p(data), start, data.shape[0])

    def exchangeArrays(self, out_pos, out_vel, in_pos, in_vel):
        """cSPH::exchangeArrays on raw host pointers (ints) or numpy arrays: current state out, new state in, one call."""
        def ptr(a):
            return C.c_void_p(a) if isinstance(a, int) else _p(a)
        rc = self.L.sphh_exchange_arrays(self.h, ptr(out_pos), ptr(out_vel), ptr(in_pos), ptr(in_vel))
        if rc != 0:
            raise SphError(f"cSPH::exchangeArrays failed ({rc}): {self.last_error()}")

    def SaveState(self, path):
        if self.L.sphh_save_state(self.h, str(path).encode()) != 0:
            raise SphError("SaveState: " + self.last_error())

    def LoadState(self, path):
        if self.L.sphh_load_state(self.h, str(path).encode()) != 0:
            raise SphError("LoadState: " + self.last_error())
        self._check_solver()

    def solver(self) -> _SolverView:
        h = self.L.sphh_solver(self.h)
        if not h:
            raise SphError("no solver (constructed with device < 0?): " + self.last_error())
        return _SolverView(h, self.params)


def load_options(scenes_xml: str | Path = DEFAULT_SCENES_XML) -> dict:
    L = _bind()
    o = np.zeros(7, np.int32)
    L.sphh_load_options(str(scenes_xml).encode(), _p(o))
    keys = ("Windowed", "WSizeX", "WSizeY", "VSyncOff", "timAvgCnt", "barsScale", "showInfo")
    return dict(zip(keys, o.tolist()))
