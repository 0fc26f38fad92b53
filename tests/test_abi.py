"""The drop-in boundary: the shared library loads, exports every symbol the headers declare, the
SimParams block has the reference layout, and nothing computes without a GPU."""
from __future__ import annotations

import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from pibiti_b200 import host, lib


def declared_functions(header: Path) -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return sorted(set(re.findall(r"\b(sphh?_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    for header, listed in (("sph_b200.h", lib.ABI_SYMBOLS), ("sph_host_c.h", host.HOST_ABI_SYMBOLS)):
        declared = declared_functions(ROOT / "include" / header)
        assert declared, header
        assert sorted(listed) == declared, f"{header}: python symbol table out of date"
        for name in declared:
            assert hasattr(L, name), f"{name} declared in {header} but not exported"


def test_simparams_layout_matches_header(tmp_path):
    """Compile a probe against include/sph_params.h and compare every offset with the numpy dtype."""
    names = list(lib.SIMPARAMS_DTYPE.names)
    src = tmp_path / "probe.cpp"
    body = "\n".join(f'  printf("{n} %zu\\n", offsetof(SimParams, {n}));' for n in names)
    src.write_text('#include <cstdio>\n#include <cstddef>\n#include "sph_params.h"\nint main(){\n'
                   'printf("sizeof %zu\\n", sizeof(SimParams));\n' + body + "\nreturn 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["g++", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == 560 == lib.SIMPARAMS_DTYPE.itemsize
    for n in names:
        assert int(out[n]) == lib.SIMPARAMS_DTYPE.fields[n][1], n


def test_c_header_is_plain_c(tmp_path):
    """include/*.h (the C ABI) must compile as C with nothing but the CUDA vector types."""
    src = tmp_path / "probe.c"
    src.write_text('#include "sph_b200.h"\n#include "sph_host_c.h"\nint main(void){ struct SimParams p; (void)p; return SPH_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
                    "-c", str(src), "-o", str(tmp_path / "probe.o")], check=True)


@pytest.mark.skipif(has_gpu(), reason="checks the behaviour without a GPU")
def test_no_gpu_means_loud_failure_not_fallback(golden_repo_scenes):
    L = lib.load()
    par = golden_repo_scenes["live"][10:11].copy()              # "mini box"
    h = C.c_void_p()
    rc = L.sph_create(par.ctypes.data_as(C.c_void_p), 0, C.byref(h))
    assert rc == 2 and not h                                    # SPH_ERR_CUDA
    msg = L.sph_last_error(None).decode()
    assert "no CPU path" in msg or "CUDA" in msg
    with pytest.raises(lib.SphError):
        lib.SphSystem(par)
    # the C++ host layer refuses to step, too
    s = host.CSph(device=-1)
    with pytest.raises(lib.SphError):
        s.Update()
    # ... and so does the multi-GPU driver (both shapes), with the reason in its own error channel
    with pytest.raises(lib.SphError, match="CUDA|no CPU path"):
        lib.MultiSystem(par, capacity_per_slab=8192, devices=[0, 0])
    with pytest.raises(lib.SphError, match="CUDA|no CPU path"):
        lib.MultiSystem(par, capacity_per_slab=8192, rank=0, world=1, device=0)


def test_bad_params_rejected():
    L = lib.load()
    par = lib.params_array()
    par["numParticles"] = 1024
    par["gridSize"] = (2, 8, 8)
    par["gridSize_yx"] = 16
    par["numCells"] = 128
    h = C.c_void_p()
    assert L.sph_create(par.ctypes.data_as(C.c_void_p), 0, C.byref(h)) == 3      # SPH_ERR_PARAMS (checked before the device)
    assert b">= 4 cells" in L.sph_last_error(None)


def test_product_never_touches_the_oracle():
    """Nothing under pibiti_b200/ may import, link or execute anything under oracle/."""
    forbidden = ("import oracle", "from oracle", "oracle/", "oracle_api", "libsphport", "libsphref", "orc_")
    for path in (ROOT / "pibiti_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") and path.name != "build.py":
            text = path.read_text(errors="replace")
            for pat in forbidden:
                assert pat not in text, f"{path} mentions {pat!r}"
    out = subprocess.run(["ldd", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sphport" not in out and "sphref" not in out


APP_STYLE_CLIENT = r'''
// Written the way the reference's App layer uses `App::psys` (grep 'psys->' source/App): direct access to scn.params,
// scenes, curScene, hPos/hVel, and the cSPH methods -- compiled against include/sph_host.h only.
#include <cstdio>
#include <cmath>
#include "sph_host.h"
int main(int argc, char** argv)
{
    SphOptions opt = cSPH::LoadOptions(argv[1]);
    cSPH* psys = new cSPH(argv[1], -1);                       // scene layer only: no GPU in this test
    SimParams* p = &psys->scn.params;                         // Sliders.cpp binds slider pointers like this
    printf("scenes %d cur %d n %u title %s windowed %d\n", (int)psys->scenes.size(), psys->curScene, p->numParticles,
           psys->scn.title, (int)opt.bWindowed);
    psys->NextScene(false);
    psys->PrevScene(false);
    psys->Reset(0);
    psys->Drop(false);
    psys->UpdateEmitter();
    p->viscosity *= 2.f;  psys->app.bChangedAny = true;       // ParamBase::Changed()
    float sum = 0.f;
    for (unsigned i = 0; i < p->numParticles; i++) sum += psys->hPos[i].w;
    printf("w-sum %.1f emitId %d\n", sum, psys->app.emitId);
    int rc = psys->Update();                                  // must fail loudly: there is no CPU path
    printf("update rc %d: %s\n", rc, psys->lastError());
    delete psys;
    return rc != 0 && std::isfinite(sum) ? 0 : 1;
}
'''


def test_app_style_cpp_client_compiles_and_links_against_the_host_header(tmp_path):
    src, exe = tmp_path / "client.cpp", tmp_path / "client"
    src.write_text(APP_STYLE_CLIENT)
    libdir = lib.LIB_PATH.parent
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", str(src),
                    "-o", str(exe), f"-L{libdir}", "-lsph_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe), str(host.DEFAULT_SCENES_XML)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "scenes" in r.stdout and "update rc" in r.stdout and "update rc 0" not in r.stdout


MULTI_GPU_CLIENT = r'''
// What an App that holds a cSPH on SEVERAL GPUs writes (INTEGRATION.md section 5): the constructor is the only change.
#include "sph_host.h"
#include <cstdio>
int main(int argc, char** argv)
{
    const int devices[2] = {0, 1};
    cSPH* psys = new cSPH(argc > 1 ? argv[1] : "Scenes.xml", devices, 2);
    psys->curScene = 0;
    psys->UpdScene();
    std::printf("scenes %zu, multi solver %s\n", psys->scenes.size(), psys->multiSolver() ? "up" : "absent");
    psys->UpdateEmitter();
    int rc = psys->Update();                         // no GPUs on this box: must fail, loudly
    std::printf("update rc %d: %s\n", rc, psys->lastError());
    float4* pos = psys->getArray(false);
    int bad = pos == nullptr;
    // the C entry points underneath
    sph_multi_t* m = nullptr;
    int rc2 = sph_multi_create(&psys->scn.params, 2, devices, 65536, &m);
    std::printf("sph_multi_create rc %d: %s\n", rc2, sph_multi_last_error(nullptr));
    if (m) sph_multi_destroy(m);
    delete psys;
    return (rc != 0 && rc2 != 0 && !bad) ? 0 : 1;
}
'''


@pytest.mark.skipif(has_gpu(), reason="asserts the no-GPU behaviour")
def test_multi_gpu_cpp_client_compiles_links_and_fails_loudly_without_gpus(tmp_path):
    src, exe = tmp_path / "client_multi.cpp", tmp_path / "client_multi"
    src.write_text(MULTI_GPU_CLIENT)
    libdir = lib.LIB_PATH.parent
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include", str(src),
                    "-o", str(exe), f"-L{libdir}", "-lsph_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe), str(host.DEFAULT_SCENES_XML)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "multi solver absent" in r.stdout and "sph_multi_create rc 2" in r.stdout
