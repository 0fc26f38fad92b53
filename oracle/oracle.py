"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the CPU oracles (see oracle/oracle_api.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  Nothing under pibiti_b200/ does.

  kind="reference": oracle/_ref/libsphref.so, the reference's own kernel text host-compiled
                    (built where /root/reference exists; the prebuilt .so travels to the GPU box)
  kind="port":      oracle/libsphport.so, this repo's plain C++ restatement (always buildable)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
REF_LIB = _HERE / "_ref" / "libsphref.so"
PORT_LIB = _HERE / "libsphport.so"

_cache: dict[str, "Oracle"] = {}


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, path: Path):
        self.path = path
        L = self.L = C.CDLL(str(path))
        L.orc_kind.restype = C.c_char_p
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p]
        for f in ("orc_destroy",):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_sys_set_params.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sys_set_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_sys_get_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_sys_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_sys_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.kind = L.orc_kind().decode()
        assert L.orc_sizeof_params() == 560

    # -- stage-level -----------------------------------------------------------------------------
    def set_params(self, par: np.ndarray):
        self.L.orc_set_params(_p(par))

    def set_threads(self, n: int):
        self.L.orc_set_threads(int(n))

    def threads(self) -> int:
        return int(self.L.orc_get_threads())

    def integrate(self, pos, vel):
        n = pos.shape[0]
        npos, nvel = np.empty_like(pos), np.empty_like(vel)
        self.L.orc_integrate(_p(pos), _p(vel), _p(npos), _p(nvel), n)
        return npos, nvel

    def calc_hash(self, pos):
        pairs = np.empty((pos.shape[0], 2), np.uint32)
        self.L.orc_calc_hash(_p(pos), _p(pairs), pos.shape[0])
        return pairs

    def sort_pairs(self, pairs):
        out = np.ascontiguousarray(pairs.copy())
        self.L.orc_sort_pairs(_p(out), out.shape[0])
        return out

    def reorder(self, pairs, old_pos, old_vel, num_cells):
        n = pairs.shape[0]
        cell_start = np.empty(num_cells, np.uint32)
        spos, svel = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        self.L.orc_reorder(_p(pairs), _p(cell_start), _p(old_pos), _p(old_vel), _p(spos), _p(svel), n, num_cells)
        return cell_start, spos, svel

    def density(self, spos, pairs, cell_start):
        n = spos.shape[0]
        pres, dens = np.empty(n, np.float32), np.empty(n, np.float32)
        self.L.orc_density(_p(spos), _p(pairs), _p(cell_start), _p(pres), _p(dens), n, cell_start.shape[0])
        return pres, dens

    def force(self, spos, svel, pres, dens, pairs, cell_start):
        n = spos.shape[0]
        new_vel, clr, dye = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros(n, np.float32)
        self.L.orc_force(_p(spos), _p(svel), _p(pres), _p(dens), _p(pairs), _p(cell_start), _p(new_vel), _p(clr), _p(dye),
                         n, cell_start.shape[0])
        return new_vel

    def system(self, par: np.ndarray) -> "OracleSystem":
        return OracleSystem(self, par)


class OracleSystem:
    """cSPH::Update on host arrays (oracle/oracle_system.inc)."""

    def __init__(self, orc: Oracle, par: np.ndarray):
        self.o = orc
        self.par = np.ascontiguousarray(par).copy()
        if self.par.dtype == np.uint8:          # raw 560-byte SimParams block (include/sph_params.h offsets 4 and 56)
            self.n = int(self.par[4:8].view(np.uint32)[0])
            self.num_cells = int(self.par[56:60].view(np.uint32)[0])
        else:
            self.n = int(self.par["numParticles"][0])
            self.num_cells = int(self.par["numCells"][0])
        self.h = C.c_void_p(orc.L.orc_create(_p(self.par)))

    def close(self):
        if self.h:
            self.o.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, par):
        self.par = np.ascontiguousarray(par).copy()
        self.o.L.orc_sys_set_params(self.h, _p(self.par))

    def set_array(self, which, data, start=0):
        data = np.ascontiguousarray(data, np.float32).reshape(-1, 4)
        self.o.L.orc_sys_set_array(self.h, which, _p(data), start, data.shape[0])

    def get_array(self, which, start=0, count=None):
        count = self.n - start if count is None else count
        out = np.empty((count, 4), np.float32)
        self.o.L.orc_sys_get_array(self.h, which, _p(out), start, count)
        return out

    def step(self, nsteps=1):
        self.o.L.orc_sys_step(self.h, nsteps)

    def dump(self, what):
        n, c = self.n, self.num_cells
        spec = {0: ((n, 2), np.uint32), 1: ((c,), np.uint32), 2: ((n, 4), np.float32), 3: ((n, 4), np.float32),
                4: ((n,), np.float32), 5: ((n,), np.float32), 6: ((n,), np.uint32), 7: ((n, 4), np.float32),
                8: ((n,), np.float32)}[what]
        out = np.empty(spec[0], spec[1])
        self.o.L.orc_sys_dump(self.h, what, _p(out))
        return out


def available(kind: str) -> bool:
    return (REF_LIB if kind == "reference" else PORT_LIB).exists()


def load(kind: str | None = None) -> Oracle:
    """kind None: the reference build if present, else the port."""
    if kind is None:
        kind = "reference" if REF_LIB.exists() else "port"
    if kind not in _cache:
        path = REF_LIB if kind == "reference" else PORT_LIB
        if not path.exists():
            raise FileNotFoundError(f"oracle '{kind}' not built: {path} (python -m pibiti_b200.build)")
        _cache[kind] = Oracle(path)
    return _cache[kind]
