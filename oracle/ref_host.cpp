// TEST INFRASTRUCTURE ONLY -- second translation unit of oracle/_ref/libsphref.so.
//
// Compiles the REFERENCE's own host-side scene code so that golden vectors for the scene /
// parameter / initialiser layer come from the reference itself:
//   source/SPH/Scene.cpp, Scene_Load.cpp   (Scene::InitDefault/_FromXML/_UpdatePar/_UpdateGrid)
//   source/SPH/SPH_Init.cpp                (cSPH::cSPH, Reset, Drop)
//   source/SPH/SPH_Scenes.cpp              (LoadScenes, InitScene, Next/PrevScene)
// Those files cannot be compiled as they stand: their first include, pch/header.h, pulls in
// GL/glew.h + GL/glut.h (absent here) and App/App.h pulls in the renderer.  build_ref.sh
// therefore strips only their #include lines into oracle/_ref/gen/*.inc at build time, and this
// file provides the few things those includes would have provided: the GL-free part of
// pch/header.h (extracted by line range at build time, too), a data-only stand-in for class App,
// and host-memory versions of the cSPH methods that would touch CUDA/GL.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cassert>
#include <vector>
#include <unistd.h>
#include <cuda_runtime.h>
#include <helper_math.h>            // reference source/external/helper_math.h (header.h:10)
#include <tinyxml.h>                // reference source/external/tinyxml/tinyxml.h
typedef unsigned int uint;
using namespace std;

#include "header_ops.inc"           // pch/header.h:32-46   float4/float3 mixed operators
#include "header_util.inc"          // pch/header.h:123-156 length3, frand, random, toVec3, toFloat, toInt

#include "SPH/SPH.h"                // the reference class declaration itself (-I $REF/source)

// ---- data-only stand-ins for what App/App.h and Graphics/param.h would declare ----------------
struct ParamBase { static bool bChangedAny; static void Changed() { bChangedAny = true; } };
bool ParamBase::bChangedAny = false;

struct App {
    static float3 camPosLag, camRotLag, dyePos;         // App.cpp:12  (zero-initialised statics)
    static float4 colliderPos;                          // App.cpp:13
    static int emitId, cntRain;                         // App.cpp:14
    static bool bWindowed, bVsyncOff, bShowInfo;
    static int WSizeX, WSizeY, timAvgCnt;
    static float barsScale;
    static void updHue() {}
};
float3 App::camPosLag, App::camRotLag, App::dyePos;
float4 App::colliderPos;
int App::emitId = 0, App::cntRain = 0;
bool App::bWindowed = true, App::bVsyncOff = false, App::bShowInfo = true;
int App::WSizeX = 0, App::WSizeY = 0, App::timAvgCnt = 0;
float App::barsScale = 30.f;

Timer::Timer() {}
bool Timer::update(bool) { return false; }

// ---- host-memory versions of the CUDA/GL-touching members (SPH_Mem.cpp, SPH_Util.cpp) ---------
void cSPH::_InitMem()
{
    if (bInitialized) return;  bInitialized = true;
    uint npar = scn.params.numParticles;
    hPos = new float4[npar];  memset(hPos, 0, npar * sizeof(float4));      // SPH_Mem.cpp:20-21
    hVel = new float4[npar];  memset(hVel, 0, npar * sizeof(float4));
}
void cSPH::_FreeMem()
{
    if (!bInitialized) return;  bInitialized = false;
    delete[] hPos;  hPos = 0;  delete[] hVel;  hVel = 0;
}
// Reset/Drop fill hPos/hVel and then upload them; the upload is the identity here.
void cSPH::setArray(bool, const float4*, int, int) {}
float4* cSPH::getArray(bool pos) { return pos ? hVel : hPos; }           // inverted flag: SPH_Util.cpp:44-56
void cSPH::Update() {}

// ---- the reference text ------------------------------------------------------------------------
#include "Scene.inc"
#include "Scene_Load.inc"
#include "SPH_Init.inc"
#include "SPH_Scenes.inc"

// ---- C interface for tests/golden/make_golden.py ----------------------------------------------
static cSPH* g_sys = 0;

extern "C" int refh_load(const char* dirWithScenesXml)
{
    // LoadScenes reads "Scenes.xml" from the current directory (SPH_Scenes.cpp:56)
    char cwd[4096];  if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(dirWithScenesXml) != 0) return -1;
    srand(1);                                   // the reference never seeds: glibc default seed 1
    App::emitId = 0;  App::cntRain = 0;
    delete g_sys;  g_sys = new cSPH();
    if (chdir(cwd) != 0) return -1;
    return (int)g_sys->scenes.size();
}
extern "C" int refh_cur_scene() { return g_sys->curScene; }

extern "C" void refh_scene_params(int idx, void* out) { memcpy(out, &g_sys->scenes[idx].params, sizeof(SimParams)); }

// 64 floats of the non-SimParams scene state, fixed order (see tests/golden/make_golden.py)
extern "C" void refh_scene_extra(int idx, float* o)
{
    const Scene& s = g_sys->scenes[idx];  int k = 0;
    o[k++] = s.initMin.x; o[k++] = s.initMin.y; o[k++] = s.initMin.z;
    o[k++] = s.initMax.x; o[k++] = s.initMax.y; o[k++] = s.initMax.z;
    o[k++] = (float)s.initType; o[k++] = (float)s.initLast; o[k++] = s.spacing; o[k++] = s.fCellSize;
    o[k++] = s.dropR; o[k++] = (float)s.rain; o[k++] = s.rVel; o[k++] = s.r2Vel;
    o[k++] = s.camPos.x; o[k++] = s.camPos.y; o[k++] = s.camPos.z; o[k++] = s.camRot.x; o[k++] = s.camRot.y;
    o[k++] = s.bChapter ? 1.f : 0.f;
    for (int e = 0; e < NumEmit; e++) {
        const Emitter& m = s.emit[e];
        o[k++] = m.pos.x; o[k++] = m.pos.y; o[k++] = m.pos.z; o[k++] = m.rot.x; o[k++] = m.rot.y;
        o[k++] = m.vel; o[k++] = (float)m.size; o[k++] = (float)m.size2;
    }
    while (k < 64) o[k++] = 0.f;
}

// scn is the live copy: params after InitScene (dyePos overwritten, SPH_Scenes.cpp:23)
extern "C" void refh_live_params(void* out) { memcpy(out, &g_sys->scn.params, sizeof(SimParams)); }

extern "C" int refh_select_scene(int idx)
{
    srand(1);
    g_sys->curScene = idx;  g_sys->UpdScene();          // scn = scenes[cur]; InitScene() -> Reset
    return (int)g_sys->scn.params.numParticles;
}
extern "C" void refh_next_scene(int chapter) { g_sys->NextScene(chapter != 0); }
extern "C" void refh_prev_scene(int chapter) { g_sys->PrevScene(chapter != 0); }

extern "C" void refh_get_host(float* pos, float* vel)
{
    size_t n = g_sys->scn.params.numParticles;
    if (pos) memcpy(pos, g_sys->hPos, n * sizeof(float4));
    if (vel) memcpy(vel, g_sys->hVel, n * sizeof(float4));
}
extern "C" void refh_reset(int type) { g_sys->Reset(type); }
extern "C" int  refh_drop(int bRandom) { g_sys->Drop(bRandom != 0);  return App::emitId; }
extern "C" void refh_srand(unsigned s) { srand(s); }
extern "C" int  refh_emit_id() { return App::emitId; }
