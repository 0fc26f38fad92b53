"""In-tree build of the native code (no JIT cache: the .so files travel with the repo snapshot).

  pibiti_b200/libsph_b200.so   CUDA kernels (sm_100a) + C ABI (include/sph_b200.h) + C++ host layer
  oracle/libsphport.so         CPU restatement of the reference algorithm   (test infrastructure)
  oracle/_ref/libsphref.so     the reference's own text, host-compiled      (test infrastructure,
                               only where /root/reference exists)

`python -m pibiti_b200.build` builds everything that is out of date.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "pibiti_b200" / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = ROOT / "pibiti_b200" / "libsph_b200.so"
PORT_LIB = ROOT / "oracle" / "libsphport.so"
REF_LIB = ROOT / "oracle" / "_ref" / "libsphref.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-I", str(CSRC)]

# translation unit -> extra flags
CUDA_UNITS = {
    "sph_stream_kernels.cu": ["-fmad=false"],     # bit-exact float evaluation vs the CPU oracle
    "sph_pair_kernels.cu": [],
    "sph_extras_kernels.cu": [],
    "sph_capi.cu": [],
    "sph_multi.cu": [],           # multi-GPU driver; NCCL is dlopen'ed at run time
}
HOST_UNITS = sorted(p.name for p in (CSRC / "host").glob("*.cpp")) if (CSRC / "host").is_dir() else []


def _nccl_include() -> list[str]:
    """nccl.h for sph_multi.cu (types only; the library is dlopen'ed at run time): the system header, else the copy that ships
    with the nvidia-nccl wheel PyTorch depends on."""
    if os.path.exists("/usr/include/nccl.h"):
        return []
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            inc = Path(loc) / "include"
            if (inc / "nccl.h").exists():
                return ["-I", str(inc)]
    except Exception:
        pass
    return []


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return exe


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    if r.stderr.strip() and os.environ.get("SPH_BUILD_VERBOSE"):
        sys.stderr.write(r.stderr)


def _newer(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.exists() and d.stat().st_mtime > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA translation unit for sm_100a and link libsph_b200.so."""
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h")) + list((CSRC / "host").glob("*.h"))
    objs = []
    nvcc = _nvcc()
    for unit, extra in CUDA_UNITS.items():
        src, obj = CSRC / unit, OBJ / (unit + ".o")
        if force or _newer(obj, [src] + headers):
            cmd = [nvcc] + ARCH + NVCC_COMMON + _nccl_include() + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(src), "-o", str(obj)]
            _run(cmd)
        objs.append(obj)
    for unit in HOST_UNITS:
        src, obj = CSRC / "host" / unit, OBJ / (unit + ".o")
        if force or _newer(obj, [src] + headers):
            _run(["g++", "-O2", "-std=c++17", "-fPIC", "-I", str(ROOT / "include"), "-I", str(CSRC), "-I", str(CSRC / "host"),
                  "-I", "/usr/local/cuda/include", "-c", str(src), "-o", str(obj)])
        objs.append(obj)
    if force or _newer(LIB, objs):
        _run([nvcc] + ARCH + ["-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    return LIB


def build_oracle_port(force: bool = False) -> Path | None:
    src = ROOT / "oracle" / "sph_port.cpp"
    if not src.exists():
        return None
    deps = [src, ROOT / "oracle" / "oracle_api.h", ROOT / "oracle" / "oracle_system.inc", ROOT / "include" / "sph_params.h"]
    if force or _newer(PORT_LIB, deps):
        _run(["g++", "-O2", "-fopenmp", "-fPIC", "-std=c++14", "-ffp-contract=off", "-shared",
              "-I", str(ROOT / "oracle"), "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
              str(src), "-o", str(PORT_LIB)])
    return PORT_LIB


def build_oracle_ref(force: bool = False) -> Path | None:
    """Host-compile the reference's own text.  Only possible where /root/reference exists."""
    ref = Path(os.environ.get("SPH_REFERENCE", "/root/reference"))
    if not (ref / "source" / "CUDA" / "System.cu").exists():
        return REF_LIB if REF_LIB.exists() else None
    deps = [ROOT / "oracle" / f for f in ("ref_driver.cpp", "ref_host.cpp", "oracle_system.inc", "oracle_api.h",
                                          "build_ref.sh", "shim/ref_shim.h")]
    if force or _newer(REF_LIB, deps):
        _run(["bash", str(ROOT / "oracle" / "build_ref.sh")])
    return REF_LIB


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_oracle_port(force)
    build_oracle_ref(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB, PORT_LIB if PORT_LIB.exists() else "(no port yet)", REF_LIB if REF_LIB.exists() else "(no _ref)")
