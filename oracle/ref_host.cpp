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

typedef float GLfloat;
struct App {
    static float3 camPosLag, camRotLag, dyePos;         // App.cpp:12  (zero-initialised statics)
    static float4 colliderPos;                          // App.cpp:13
    static int emitId, cntRain;                         // App.cpp:14
    static float modelView[16], inertia;                // App.cpp:15
    static bool bWindowed, bVsyncOff, bShowInfo;
    static int WSizeX, WSizeY, timAvgCnt;
    static float barsScale, fSimTime;
    static cSPH* psys;                                  // App.h:44
    static void updHue() {}
    static void UpdateEmitter();                        // App/Update.cpp:9-97 (text compiled below)
    static void mulTr(float* v, float* r, GLfloat* m);  // App/Input.cpp:588-593 (text compiled below)
};
float3 App::camPosLag, App::camRotLag, App::dyePos;
float4 App::colliderPos;
int App::emitId = 0, App::cntRain = 0;
float App::modelView[16], App::inertia = 0.06;
bool App::bWindowed = true, App::bVsyncOff = false, App::bShowInfo = true;
int App::WSizeX = 0, App::WSizeY = 0, App::timAvgCnt = 0;
float App::barsScale = 30.f, App::fSimTime = 0.f;
cSPH* App::psys = 0;

// The three fixed-function GL calls UpdateEmitter uses to build the emitter orientation (Update.cpp:73-75),
// restated from the OpenGL specification (column-major 4x4, glRotatef post-multiplies the current matrix;
// sine and cosine of angle*pi/180 evaluated in double and rounded to float, as Mesa's _math_matrix_rotate does).
#define GL_MODELVIEW_MATRIX 0x0BA6
static float g_glm[16];
static void glLoadIdentity() { for (int i = 0; i < 16; i++) g_glm[i] = (i % 5 == 0) ? 1.f : 0.f; }
static void glRotatef(float angle, float x, float y, float z)
{
    const float s = (float)sin((double)angle * (M_PI / 180.0)), c = (float)cos((double)angle * (M_PI / 180.0));
    float r[16];
    for (int i = 0; i < 16; i++) r[i] = (i % 5 == 0) ? 1.f : 0.f;
    // axis-aligned axes only (the emitter uses (1,0,0) and (0,1,0)); columns of the rotation matrix
    if (x != 0.f && y == 0.f && z == 0.f) { r[5] = c;  r[10] = c;  if (x < 0.f) { r[6] = -s; r[9] = s; } else { r[6] = s;  r[9] = -s; } }
    else if (x == 0.f && y != 0.f && z == 0.f) { r[0] = c;  r[10] = c;  if (y < 0.f) { r[8] = -s; r[2] = s; } else { r[8] = s;  r[2] = -s; } }
    else if (x == 0.f && y == 0.f && z != 0.f) { r[0] = c;  r[5] = c;  if (z < 0.f) { r[1] = -s; r[4] = s; } else { r[1] = s;  r[4] = -s; } }
    float o[16];
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) {
            float a = 0.f;
            for (int k = 0; k < 4; k++) a += g_glm[k * 4 + row] * r[col * 4 + k];
            o[col * 4 + row] = a;
        }
    memcpy(g_glm, o, sizeof o);
}
static void glGetFloatv(int, float* m) { memcpy(m, g_glm, sizeof g_glm); }

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
// Reset/Drop fill hPos/hVel and then upload them: the upload is the identity here.  The emitters pass their
// own buffers (App/Update.cpp:86-87): those land in the host arrays (clipped to the array, where the
// reference would write past the end of its GL buffer).  Inverted flag as in SPH_Util.cpp:59-71.
static long g_setArrayCalls = 0;
void cSPH::setArray(bool pos, const float4* data, int start, int count)
{
    float4* dst = pos ? hVel : hPos;
    const int n = (int)scn.params.numParticles;
    if (start < 0 || start >= n) return;
    if (count > n - start) count = n - start;
    if (data != dst + start) { memcpy(dst + start, data, (size_t)count * sizeof(float4));  g_setArrayCalls++; }
}
float4* cSPH::getArray(bool pos) { return pos ? hVel : hPos; }           // inverted flag: SPH_Util.cpp:44-56
void cSPH::Update() {}

// ---- the reference text ------------------------------------------------------------------------
#include "Scene.inc"
#include "Scene_Load.inc"
#include "SPH_Init.inc"
#include "SPH_Scenes.inc"

#include "App_Update.inc"           // App::UpdateEmitter, the per-step host prologue (App/Update.cpp:9-97)
#include "App_mulTr.inc"            // App::mulTr (App/Input.cpp:588-593)

// ---- C interface for tests/golden/make_golden.py ----------------------------------------------
static cSPH* g_sys = 0;

extern "C" int refh_load(const char* dirWithScenesXml)
{
    // LoadScenes reads "Scenes.xml" from the current directory (SPH_Scenes.cpp:56)
    char cwd[4096];  if (!getcwd(cwd, sizeof cwd)) return -1;
    if (chdir(dirWithScenesXml) != 0) return -1;
    srand(1);                                   // the reference never seeds: glibc default seed 1
    App::emitId = 0;  App::cntRain = 0;  App::fSimTime = 0.f;
    App::dyePos = make_float3(0.f, 0.f, 0.f);   // process-start values of the statics (App.cpp:12-15)
    delete g_sys;  g_sys = new cSPH();
    App::psys = g_sys;
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

// ---- App::UpdateEmitter -------------------------------------------------------
// One call of the reference's per-step prologue on the live scene.  The App statics it reads are set the way
// InitScene leaves them (SPH_Scenes.cpp:15-26); refh_set_targets moves the collider / dye / accelerator targets
// like the mouse handlers do (App/Input.cpp).
extern "C" void refh_update_emitter() { App::UpdateEmitter();  App::fSimTime += g_sys->scn.params.timeStep; }
extern "C" void refh_set_targets(const float* collider4, const float* dye3, const float* acc3)
{
    if (collider4) App::colliderPos = make_float4(collider4[0], collider4[1], collider4[2], collider4[3]);
    if (dye3) App::dyePos = make_float3(dye3[0], dye3[1], dye3[2]);
    if (acc3) g_sys->scn.accPos[g_sys->scn.ca] = make_float3(acc3[0], acc3[1], acc3[2]);
}
// emitter lag state is advanced by the camera code in the reference (App/Input.cpp:555-560); the test drives it
extern "C" void refh_set_emitter(int e, const float* posLag3, const float* rotLag2, float vel, int size, int size2)
{
    Emitter& em = g_sys->scn.emit[e];
    em.posLag = make_float3(posLag3[0], posLag3[1], posLag3[2]);
    em.rotLag = make_float3(rotLag2[0], rotLag2[1], 0.f);
    em.vel = vel;  em.size = size;  em.size2 = size2;
}
extern "C" int  refh_cnt_rain() { return App::cntRain; }
extern "C" int  refh_changed_flag(int clear) { int v = ParamBase::bChangedAny ? 1 : 0;  if (clear) ParamBase::bChangedAny = false;  return v; }
extern "C" long refh_set_array_calls() { return g_setArrayCalls; }
