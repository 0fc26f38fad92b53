"""ctypes binding of the C ABI in include/sph_b200.h (libsph_b200.so).

There is no fallback of any kind: if the CUDA library is missing or no sm_100 GPU is usable,
loading / sph_create raise.  SimParams is handled as the 560-byte block it is (numpy structured
dtype generated from the field table below, which mirrors include/sph_params.h).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsph_b200.so"

_f3 = ("f4", (3,))
_u3 = ("u4", (3,))
_i3 = ("i4", (3,))
_accel = np.dtype([("pos", *_f3), ("size", *_f3), ("acc", *_f3), ("type", "u4")])

# field order == include/sph_params.h == reference source/CUDA/Params.cuh:53-114
SIMPARAMS_DTYPE = np.dtype(
    {
        "names": [],
        "formats": [],
        "offsets": [],
        "itemsize": 560,
    }
)


def _build_dtype() -> np.dtype:
    fields = [
        ("timeStep", "f4"), ("numParticles", "u4"), ("maxParInCell", "u4"),
        ("gravity", *_f3), ("globalDamping", "f4"),
        ("gridSize", *_u3), ("cellSize", *_f3), ("gridSize_yx", "u4"), ("numCells", "u4"),
        ("worldMin", *_f3), ("worldMax", *_f3), ("worldSize", *_f3),
        ("worldMinD", *_f3), ("worldMaxD", *_f3), ("worldSizeD", *_f3),
        ("particleR", "f4"), ("h", "f4"), ("h2", "f4"),
        ("SpikyKern", "f4"), ("LapKern", "f4"), ("Poly6Kern", "f4"),
        ("particleMass", "f4"), ("restDensity", "f4"), ("stiffness", "f4"), ("viscosity", "f4"),
        ("minDens", "f4"), ("minDist", "f4"),
        ("distBndHard", "f4"), ("distBndSoft", "f4"), ("bndDamp", "f4"), ("bndStiff", "f4"), ("bndDampC", "f4"),
        ("bndType", "u4"), ("bndEffZ", "u4"),
        ("collPos", "f4", (4,)), ("collR", "f4"), ("spring", "f4"), ("damping", "f4"), ("shear", "f4"),
        ("clrType", "u4"), ("iHue", "i4"), ("brightness", "f4"), ("contrast", "f4"),
        ("dyeType", "i4"), ("dyeClear", "i4"), ("dyeFade", "f4"), ("dyePos", *_f3), ("dyeSize", *_f3),
        ("acc", _accel, (4,)), ("iHmap", "i4"),
        ("angOut", "f4"), ("hClose", "f4"), ("radIn", "f4"), ("rVexit", "f4"), ("rDexit", "f4"),
        ("s1", "f4"), ("s2", "f4"), ("s3", "f4"), ("s4", "f4"), ("s5", "f4"), ("s6", "f4"),
        ("rAngle", "f4"), ("rTwist", "f4"), ("rotType", "i4"), ("rotBlades", "i4"), ("rotSize", *_i3),
        ("rotR", "f4"), ("rotSpc", "f4"), ("r2Dist", "f4"), ("r2Angle", "f4"), ("r2twist", "f4"), ("ff2", "f4"),
    ]
    # natural packing, except that float4 collPos is 16-byte aligned (offset 208) and the block is
    # padded to 560 bytes
    names, formats, offsets = [], [], []
    off = 0
    for f in fields:
        name, fmt = f[0], f[1:]
        dt = np.dtype(fmt[0] if len(fmt) == 1 else (fmt[0], fmt[1]))
        if name == "collPos":
            off = (off + 15) // 16 * 16
        names.append(name); formats.append(dt); offsets.append(off)
        off += dt.itemsize
    assert off <= 560, off
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": 560})


SIMPARAMS_DTYPE = _build_dtype()
assert SIMPARAMS_DTYPE.fields["collPos"][1] == 208 and SIMPARAMS_DTYPE.fields["acc"][1] == 292 \
    and SIMPARAMS_DTYPE.fields["rAngle"][1] == 500, "SimParams layout drifted from include/sph_params.h"

SPH_POS, SPH_VEL, SPH_DENSITY, SPH_PRESSURE, SPH_COLOR, SPH_DYE = range(6)
(DUMP_SORTED_PAIRS, DUMP_CELL_START, DUMP_SORTED_POS, DUMP_SORTED_VEL, DUMP_PRESSURE, DUMP_DENSITY,
 DUMP_NEIGHBOR_COUNTS, DUMP_CELL_END) = range(8)
STAGE_NAMES = ("integrate_hash", "sort", "reorder", "density", "force")

# every symbol include/sph_b200.h declares (tests check the library exports all of them)
ABI_SYMBOLS = (
    "sph_create", "sph_destroy", "sph_set_params", "sph_get_params", "sph_reset_state", "sph_set_visual", "sph_set_dye", "sph_step", "sph_sync",
    "sph_set_array", "sph_get_array", "sph_exchange_arrays", "sph_set_array_device", "sph_get_array_device", "sph_device_buffers",
    "sph_debug_dump", "sph_get_timings", "sph_kernel_launch_count", "sph_cuda_stream", "sph_last_error",
    "sph_version", "sph_pair_variant", "sph_gl_register", "sph_gl_update",
    "sph_slab_configure", "sph_slab_set_owned", "sph_slab_get_owned", "sph_slab_integrate", "sph_slab_pack", "sph_slab_integrate_pack",
    "sph_slab_unpack", "sph_slab_sort", "sph_slab_density", "sph_slab_pack_dp", "sph_slab_ghost_counts",
    "sph_slab_unpack_dp", "sph_slab_force", "sph_slab_force_part", "sph_slab_stats",
    "sph_multi_unique_id", "sph_multi_create", "sph_multi_create_rank", "sph_multi_destroy", "sph_multi_last_error",
    "sph_multi_set_params", "sph_multi_set_state", "sph_multi_step", "sph_multi_sync", "sph_multi_get_state",
    "sph_multi_local_slabs", "sph_multi_handle", "sph_multi_stream", "sph_multi_info", "sph_multi_fetch_owned", "sph_multi_put_owned",
    "sph_multi_phase_ms", "sph_multi_exchange_owned", "sph_multi_recut", "sph_multi_set_recut_interval", "sph_multi_recut_count",
    "sph_multi_plan_cuts",
)


class SphError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libsph_b200.so.  Raises if it has not been built -- there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SphError(f"{LIB_PATH} is missing: build it with `python -m pibiti_b200.build` "
                       "(CUDA extension for sm_100a; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp, ci = C.c_void_p, C.c_int
    lib.sph_create.argtypes = [vp, ci, C.POINTER(vp)]
    lib.sph_destroy.argtypes = [vp]
    lib.sph_set_params.argtypes = [vp, vp]
    lib.sph_get_params.argtypes = [vp, vp]
    lib.sph_set_visual.argtypes = [vp, ci]
    lib.sph_reset_state.argtypes = [vp]
    lib.sph_set_dye.argtypes = [vp, vp, ci, ci]
    lib.sph_step.argtypes = [vp, ci]
    lib.sph_sync.argtypes = [vp]
    lib.sph_set_array.argtypes = [vp, ci, vp, ci, ci]
    lib.sph_get_array.argtypes = [vp, ci, vp, ci, ci]
    lib.sph_exchange_arrays.argtypes = [vp, vp, vp, vp, vp]
    lib.sph_set_array_device.argtypes = [vp, ci, vp, ci, ci]
    lib.sph_get_array_device.argtypes = [vp, ci, vp, ci, ci]
    lib.sph_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.sph_debug_dump.argtypes = [vp, ci, vp, C.c_size_t]
    lib.sph_gl_register.argtypes = [vp, ci, C.c_uint]
    lib.sph_gl_update.argtypes = [vp]
    lib.sph_get_timings.argtypes = [vp, vp, ci]
    lib.sph_kernel_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.sph_cuda_stream.argtypes = [vp]
    lib.sph_cuda_stream.restype = vp
    lib.sph_last_error.argtypes = [vp]
    lib.sph_last_error.restype = C.c_char_p
    lib.sph_version.restype = C.c_char_p
    lib.sph_pair_variant.argtypes = [vp]
    lib.sph_pair_variant.restype = C.c_char_p
    ip = C.POINTER(ci)
    lib.sph_slab_configure.argtypes = [vp, ci, ci, ci, ci]
    lib.sph_slab_set_owned.argtypes = [vp, vp, ci]
    lib.sph_slab_get_owned.argtypes = [vp, vp, ci, ip]
    lib.sph_slab_integrate.argtypes = [vp]
    lib.sph_slab_pack.argtypes = [vp, vp, vp, ci, ci]
    lib.sph_slab_unpack.argtypes = [vp, vp, vp, vp, vp, ci, ci]
    lib.sph_slab_sort.argtypes = [vp, ip]
    lib.sph_slab_density.argtypes = [vp]
    lib.sph_slab_pack_dp.argtypes = [vp, vp, vp, ci, ip]
    lib.sph_slab_ghost_counts.argtypes = [vp, ip]
    lib.sph_slab_unpack_dp.argtypes = [vp, vp, ci, vp, ci]
    lib.sph_slab_force.argtypes = [vp]
    lib.sph_slab_force_part.argtypes = [vp, ci]
    lib.sph_slab_integrate_pack.argtypes = [vp, vp, vp, ci, ci]
    lib.sph_slab_stats.argtypes = [vp, ip]
    lib.sph_multi_unique_id.argtypes = [vp]
    lib.sph_multi_create.argtypes = [vp, ci, ip, ci, C.POINTER(vp)]
    lib.sph_multi_create_rank.argtypes = [vp, ci, ci, vp, ci, ci, C.POINTER(vp)]
    lib.sph_multi_destroy.argtypes = [vp]
    lib.sph_multi_last_error.argtypes = [vp]
    lib.sph_multi_last_error.restype = C.c_char_p
    lib.sph_multi_set_params.argtypes = [vp, vp]
    lib.sph_multi_set_state.argtypes = [vp, vp, vp, ci, ip]
    lib.sph_multi_step.argtypes = [vp, ci]
    lib.sph_multi_sync.argtypes = [vp]
    lib.sph_multi_get_state.argtypes = [vp, vp, vp, vp, vp, ci, ip]
    lib.sph_multi_fetch_owned.argtypes = [vp, ci, vp, ci, ip]
    lib.sph_multi_put_owned.argtypes = [vp, ci, vp, ci]
    lib.sph_multi_exchange_owned.argtypes = [vp, ci, vp, ci, ip, vp, ci]
    lib.sph_multi_phase_ms.argtypes = [vp, ci, ci, vp]
    lib.sph_multi_recut.argtypes = [vp]
    lib.sph_multi_plan_cuts.argtypes = [vp, ci, ci, ip]
    lib.sph_multi_set_recut_interval.argtypes = [vp, ci]
    lib.sph_multi_recut_count.argtypes = [vp]
    lib.sph_multi_local_slabs.argtypes = [vp]
    lib.sph_multi_handle.argtypes = [vp, ci]
    lib.sph_multi_handle.restype = vp
    lib.sph_multi_stream.argtypes = [vp, ci]
    lib.sph_multi_stream.restype = vp
    lib.sph_multi_info.argtypes = [vp, ip, ip, ip, C.POINTER(C.c_ulonglong)]
    _lib = lib
    return lib


def params_array(src=None) -> np.ndarray:
    """A one-element structured array holding a SimParams block (optionally copied from bytes)."""
    a = np.zeros(1, SIMPARAMS_DTYPE)
    if src is not None:
        a.view(np.uint8)[:] = np.frombuffer(bytes(src), np.uint8, 560)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class SphSystem:
    """Thin object wrapper over one sph_t handle."""

    def __init__(self, params: np.ndarray, device: int = 0):
        self.lib = load()
        self.params = params_array(params.tobytes())
        self.h = C.c_void_p()
        rc = self.lib.sph_create(_ptr(self.params), device, C.byref(self.h))
        if rc != 0:
            raise SphError(f"sph_create failed ({rc}): {self.lib.sph_last_error(None).decode()}")
        self.n = int(self.params["numParticles"][0])
        self.num_cells = int(self.params["numCells"][0])

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SphError(f"{what} failed ({rc}): {self.lib.sph_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.sph_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params: np.ndarray):
        self.params = params_array(params.tobytes())
        self._check(self.lib.sph_set_params(self.h, _ptr(self.params)), "sph_set_params")

    def set_visual(self, on: bool = True):
        self._check(self.lib.sph_set_visual(self.h, int(on)), "sph_set_visual")

    def step(self, nsteps: int = 1):
        self._check(self.lib.sph_step(self.h, nsteps), "sph_step")

    def sync(self):
        self._check(self.lib.sph_sync(self.h), "sph_sync")

    def set_array(self, which: int, data: np.ndarray, start: int = 0):
        data = np.ascontiguousarray(data, np.float32).reshape(-1, 4)
        self._check(self.lib.sph_set_array(self.h, which, _ptr(data), start, data.shape[0]), "sph_set_array")

    def get_array(self, which: int, start: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - start if count is None else count
        shape = (count, 4) if which in (SPH_POS, SPH_VEL, SPH_COLOR) else (count,)
        out = np.empty(shape, np.float32)
        self._check(self.lib.sph_get_array(self.h, which, _ptr(out), start, count), "sph_get_array")
        return out

    def dump(self, what: int) -> np.ndarray:
        n, c = self.n, self.num_cells
        spec = {
            DUMP_SORTED_PAIRS: ((n, 2), np.uint32), DUMP_CELL_START: ((c,), np.uint32), DUMP_CELL_END: ((c,), np.uint32),
            DUMP_SORTED_POS: ((n, 4), np.float32), DUMP_SORTED_VEL: ((n, 4), np.float32),
            DUMP_PRESSURE: ((n,), np.float32), DUMP_DENSITY: ((n,), np.float32),
            DUMP_NEIGHBOR_COUNTS: ((n,), np.uint32),
        }[what]
        out = np.empty(spec[0], spec[1])
        self._check(self.lib.sph_debug_dump(self.h, what, _ptr(out), out.nbytes), "sph_debug_dump")
        return out

    def enable_timings(self, on: bool = True):
        self._check(self.lib.sph_get_timings(self.h, None, 1 if on else 0), "sph_get_timings")

    def timings(self) -> dict:
        ms = np.zeros(5, np.float32)
        self._check(self.lib.sph_get_timings(self.h, _ptr(ms), 1), "sph_get_timings")
        return dict(zip(STAGE_NAMES, ms.tolist()))

    def launch_count(self) -> int:
        v = C.c_longlong(0)
        self._check(self.lib.sph_kernel_launch_count(self.h, C.byref(v)), "sph_kernel_launch_count")
        return int(v.value)

    def pair_variant(self) -> str:
        return (self.lib.sph_pair_variant(self.h) or b"").decode()

    def stream(self) -> int:
        return int(self.lib.sph_cuda_stream(self.h) or 0)

    def device_buffers(self):
        p, v, i, cs = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.sph_device_buffers(self.h, C.byref(p), C.byref(v), C.byref(i), C.byref(cs)), "sph_device_buffers")
        return p.value, v.value, i.value, cs.value


class _SlabView(SphSystem):
    """One slab of a MultiSystem as an SphSystem (timings, launch counts, dumps); the handle stays owned by the driver."""

    def __init__(self, lib, handle, params):  # noqa: super().__init__ deliberately not called
        self.lib, self.h = lib, C.c_void_p(handle)
        self.params = params_array(params.tobytes())
        self.n = int(self.params["numParticles"][0])
        self.num_cells = int(self.params["numCells"][0])

    def close(self):
        self.h = C.c_void_p()


class MultiSystem:
    """ctypes wrapper over sph_multi_* (include/sph_b200.h): the whole system on several GPUs, z-slab decomposed, driven by
    the C++ multi-GPU driver (NCCL send/recv or peer copies between the slabs).  Either `devices` (one process drives them
    all) or `rank`/`world`/`unique_id`/`device` (one process per GPU)."""

    def __init__(self, params: np.ndarray, capacity_per_slab: int, devices=None, rank=None, world=None, unique_id: bytes | None = None,
                 device: int = 0):
        self.lib = load()
        self.params = params_array(params.tobytes())
        self.h = C.c_void_p()
        self.capacity = int(capacity_per_slab)
        if devices is not None:
            self.world = len(devices)
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.sph_multi_create(_ptr(self.params), len(devices), arr, self.capacity, C.byref(self.h))
        else:
            self.world = int(world)
            buf = (C.c_ubyte * 128).from_buffer_copy(unique_id) if unique_id is not None else None
            rc = self.lib.sph_multi_create_rank(_ptr(self.params), int(rank), int(world), buf, int(device), self.capacity, C.byref(self.h))
        if rc != 0:
            raise SphError(f"sph_multi_create failed ({rc}): {self.lib.sph_multi_last_error(None).decode()}")
        self.n = 0

    @staticmethod
    def unique_id() -> bytes:
        lib = load()
        buf = (C.c_ubyte * 128)()
        if lib.sph_multi_unique_id(buf) != 0:
            raise SphError("sph_multi_unique_id failed: " + lib.sph_multi_last_error(None).decode())
        return bytes(buf)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SphError(f"{what} failed ({rc}): {self.lib.sph_multi_last_error(self.h).decode()}")

    def close(self):
        if self.h:
            self.lib.sph_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params: np.ndarray):
        self.params = params_array(params.tobytes())
        self._check(self.lib.sph_multi_set_params(self.h, _ptr(self.params)), "sph_multi_set_params")

    def set_state(self, pos: np.ndarray, vel: np.ndarray, cuts=None):
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 4)
        vel = np.ascontiguousarray(vel, np.float32).reshape(-1, 4)
        assert pos.shape == vel.shape
        c = (C.c_int * (self.world + 1))(*cuts) if cuts is not None else None
        self.n = pos.shape[0]
        self._check(self.lib.sph_multi_set_state(self.h, _ptr(pos), _ptr(vel), self.n, c), "sph_multi_set_state")

    def step(self, nsteps: int = 1):
        self._check(self.lib.sph_multi_step(self.h, nsteps), "sph_multi_step")

    def sync(self):
        self._check(self.lib.sph_multi_sync(self.h), "sph_multi_sync")

    def get_state(self, density: bool = False):
        """(pos, vel[, dens, pres], written): rows of particles this process does not own stay NaN."""
        n = self.n
        pos, vel = np.full((n, 4), np.nan, np.float32), np.full((n, 4), np.nan, np.float32)
        dens = np.full(n, np.nan, np.float32) if density else None
        pres = np.full(n, np.nan, np.float32) if density else None
        w = C.c_int(0)
        self._check(self.lib.sph_multi_get_state(self.h, _ptr(pos), _ptr(vel), _ptr(dens) if density else None,
                                                 _ptr(pres) if density else None, n, C.byref(w)), "sph_multi_get_state")
        return (pos, vel, dens, pres, w.value) if density else (pos, vel, w.value)

    def fetch_owned(self, local: int, host_ptr: int, capacity_records: int) -> int:
        """Owned records of a local slab into host memory at `host_ptr` (e.g. a pinned torch tensor); returns the count."""
        c = C.c_int(0)
        self._check(self.lib.sph_multi_fetch_owned(self.h, local, C.c_void_p(host_ptr), capacity_records, C.byref(c)), "sph_multi_fetch_owned")
        return c.value

    def exchange_owned(self, local: int, out_ptr: int, out_capacity: int, in_ptr: int, in_count: int) -> int:
        """fetch_owned + put_owned as one call whose download overlaps its upload; returns the number of records written out."""
        c = C.c_int(0)
        self._check(self.lib.sph_multi_exchange_owned(self.h, local, C.c_void_p(out_ptr), out_capacity, C.byref(c),
                                                      C.c_void_p(in_ptr), in_count), "sph_multi_exchange_owned")
        return c.value

    def put_owned(self, local: int, host_ptr: int, count: int):
        self._check(self.lib.sph_multi_put_owned(self.h, local, C.c_void_p(host_ptr), count), "sph_multi_put_owned")

    PHASES = ("edge_integrate_pack", "interior_integrate_hist", "wait_particle_exchange", "unpack_arrivals", "scan_bucket_gather",
              "density", "pack_rho_p", "interior_force", "wait_rho_p_exchange", "unpack_rho_p", "boundary_force")

    def enable_phase_timing(self, on: bool = True):
        self._check(self.lib.sph_multi_phase_ms(self.h, 0, int(on), None), "sph_multi_phase_ms")

    def phase_ms(self, local: int = 0) -> dict:
        out = np.zeros(11, np.float32)
        self._check(self.lib.sph_multi_phase_ms(self.h, local, 1, _ptr(out)), "sph_multi_phase_ms")
        return dict(zip(self.PHASES, [round(float(x), 4) for x in out]))

    def recut(self):
        self._check(self.lib.sph_multi_recut(self.h), "sph_multi_recut")

    def set_recut_interval(self, steps: int):
        self._check(self.lib.sph_multi_set_recut_interval(self.h, steps), "sph_multi_set_recut_interval")

    def recut_count(self) -> int:
        return int(self.lib.sph_multi_recut_count(self.h))

    def local_slabs(self) -> int:
        return int(self.lib.sph_multi_local_slabs(self.h))

    def slab(self, local: int = 0) -> _SlabView:
        par = self.params.copy()
        par["numParticles"] = self.capacity
        return _SlabView(self.lib, self.lib.sph_multi_handle(self.h, local), par)

    def stream(self, local: int = 0) -> int:
        return int(self.lib.sph_multi_stream(self.h, local) or 0)

    def info(self, owned: bool = True) -> dict:
        cuts = (C.c_int * (self.world + 1))()
        own = (C.c_int * max(self.local_slabs(), 1))()
        caps = (C.c_int * 2)()
        sent = C.c_ulonglong(0)
        self._check(self.lib.sph_multi_info(self.h, cuts, own if owned else None, caps, C.byref(sent)), "sph_multi_info")
        return {"cuts": list(cuts), "owned": list(own)[: self.local_slabs()] if owned else None, "cap_leavers": caps[0],
                "cap_boundary": caps[1], "bytes_sent": int(sent.value)}
