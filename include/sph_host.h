// sph_host.h -- the C++ host layer that sits on the C ABI (sph_b200.h): scene parameters,
// Scenes.xml loading, particle initialisers and the cSPH-shaped system object.
//
// It mirrors the interface the reference's App / Graphics layers program against:
//   class Scene    reference source/SPH/Scene.h:24-62   (Scene.cpp, Scene_Load.cpp)
//   class Emitter  reference source/SPH/Scene.h:7-19
//   class cSPH     reference source/SPH/SPH.h:9-50      (SPH_Init/Mem/Update/Util/Scenes.cpp)
// Same member names and meaning, so code written against `psys->scn.params.*`, `psys->Update()`,
// `psys->setArray(...)`, `psys->NextScene()` compiles against this header.  What differs, because
// this layer is headless (no GL, no GLUT):
//   * positions live in solver-owned device memory.  The renderer keeps ownership of its GL buffers:
//     registerGLBuffers(posVbo, colorVbo) hands their ids over once, Update() then refreshes them on the
//     device after every step, and getPosBuffer() returns the position buffer's GL id exactly as the
//     reference does (SPH.h:29).  Headless callers read getPosDevice() / getArray() instead;
//   * the App:: statics the reference's SPH layer reaches into (App::emitId, App::dyePos,
//     App::colliderPos, camera lag, ParamBase::bChangedAny) are members of cSPH::app;
//   * errors are reported through lastError()/return codes instead of exit(1).
#ifndef SPH_HOST_H
#define SPH_HOST_H

#include <string>
#include <vector>
#include "sph_params.h"
#include "sph_b200.h"

static const int NumEmit = 4;       // Scene.h:21

class Emitter {                     // Scene.h:7-19
public:
    float3 pos, rot, posLag, rotLag;
    float vel;
    int size, size2;
    Emitter();
    void IncSize() { if (size < 10) size++; }
    void DecSize() { if (size > 0) size--; }
};

namespace sphxml { struct Element; }

class Scene {                       // Scene.h:24-62
public:
    char title[40];
    bool bChapter;

    SimParams params;
    float fCellSize;

    float3 initMin, initMax;        // init volume
    int initType, initLast;         // 0 volume / 1 random;  last axis 0 x, 1 y, 2 z
    float spacing;

    float3 camPos, camRot;
    float4 collidPos;

    int ce;                         // current emitter
    Emitter emit[NumEmit];
    float dropR;
    int rain;

    int ca;                         // current accelerator
    float3 accPos[SPH_NUM_ACC];

    float rVel, r2Vel;              // rotor / wave angular velocity

    Scene();                                        // default scene only
    explicit Scene(const sphxml::Element* s);       // from a <Scene> element

    void InitDefault();
    void _FromXML(const sphxml::Element* s);
    void Update();                                  // _UpdatePar + _UpdateGrid
    void _UpdatePar();
    void _UpdateGrid();
};

// pch/timer.h:23-36: wall-clock timer the App layer reads through psys->tim (tim.FR, tim.dt;
// App/RenderText.cpp:19,69,81).  The reference's body is QueryPerformanceCounter inside #if _WIN32 and
// therefore dead on Linux; this one runs on std::chrono with the same fields.
class Timer {
    double st, st1;  int iFR;
public:
    double t, dt, FR, iv, iv1;      // time, delta time, frame rate, interval, frame-rate interval
    Timer();
    bool update(bool updFR = false);
};

// program options of the <Options> element (the reference stores them in App:: statics,
// SPH_Scenes.cpp:99-109)
struct SphOptions {
    bool bWindowed = true, bVsyncOff = false, bShowInfo = true;
    int WSizeX = 0, WSizeY = 0, timAvgCnt = 0;
    float barsScale = 30.f;
};

// the App:: state the reference's SPH layer reads and writes
struct SphAppState {
    float3 camPosLag, camRotLag, dyePos;    // App.cpp:12 (zero-initialised)
    float4 colliderPos;                     // App.cpp:13
    int emitId = 0, cntRain = 0;            // App.cpp:14
    float inertia = 0.06f;                  // App.cpp:15
    float fSimTime = 0.f;
    bool bChangedAny = false;               // ParamBase::bChangedAny (Graphics/param.h:33-34)
    SphAppState();
};

class cSPH {                        // SPH.h:9-50
public:
    // device >= 0: allocate the solver on that GPU.  device < 0: scene/initialiser layer only
    // (no GPU touched; Update() fails) -- used to load and inspect Scenes.xml.
    explicit cSPH(const char* scenesXmlPath = "Scenes.xml", int device = 0);
    // Several GPUs of one box: the system is cut into z slabs, one per device, and stepped by the multi-GPU driver
    // (sph_multi_*, sph_b200.h) -- same Update / getArray / setArray / scene interface, results bit-identical to one GPU.
    // Not available in this mode: GL interop, the dye / colour outputs, scenes with a Z wrap/cycle or pump boundary.
    cSPH(const char* scenesXmlPath, const int* devices, int ndev);
    ~cSPH();

    void _InitMem(), _FreeMem();
    bool bInitialized;

    int Update();                               // one solver step (SPH_Update.cpp:12-81); 0 on success
    int Update(int nsteps);
    void Reset(int type);                       // fill the init volume (SPH_Init.cpp:23-79)
    void Drop(bool bRandom);                    // drop a sphere of particles (SPH_Init.cpp:84-118)
    void UpdateEmitter();                       // per-step host prologue (App/Update.cpp:9-97)
    float3 DropPos;

    static SphOptions LoadOptions(const char* scenesXmlPath = "Scenes.xml");
    void LoadScenes();
    void InitScene();
    void NextScene(bool chapter = false);
    void PrevScene(bool chapter = false);
    void UpdScene();
    std::vector<Scene> scenes;
    Scene scn;
    int curScene;

    float4* getArray(bool pos);                 // NB inverted flag as in the reference: false = positions, true = velocities
    void setArray(bool pos, const float4* data, int start, int count);
    // getArray(false), getArray(true) into the caller's buffers, then setArray of both from the caller's new state, as one
    // call whose downloads overlap its uploads (sph_exchange_arrays; pinned buffers make the overlap real).  0 on success.
    int exchangeArrays(float4* outPos, float4* outVel, const float4* inPos, const float4* inVel);
    uint getPosBuffer() const { return posVbo[curPosRead]; }    // GL id of the position buffer (SPH.h:29); 0 when headless
    const float4* getPosDevice() const;         // device pointer of the live positions (sorted order), for headless callers
    // Hands the renderer's GL buffers to the solver (the reference creates them itself in _InitMem,
    // SPH_Mem.cpp:26-29).  Update() refreshes them after every step.  0 on success; without a current GL context
    // the registration fails and the ids stay 0.
    int registerGLBuffers(uint positionsVbo, uint colorsVbo);

    // Checkpoint / resume (the reference has none): parameters, positions and velocities in
    // original particle order, and the ring / rain / time counters.  0 on success.
    int SaveState(const char* path);
    int LoadState(const char* path);

    sph_t* solver() const { return sys; }
    sph_multi_t* multiSolver() const { return msys; }         // non-null when constructed with a device list
    const char* lastError() const { return err.c_str(); }

    float4 *hPos, *hVel;                        // host mirrors, original particle order
    int* hCounters;                             // debug counters (SPH.h:43; always zero here, as in the reference's build)
    uint posVbo[2], colorVbo;                   // GL ids handed over by registerGLBuffers (both posVbo entries are the same buffer)
    uint curPosRead, curPosWrite;               // kept for source compatibility: positions do not ping-pong between VBOs here
    Timer tim;                                  // updated once per Update(), SPH_Update.cpp:16
    SphAppState app;

private:
    std::string xmlPath;
    int device;
    sph_t* sys;
    std::vector<int> devices;                   // multi-GPU mode: the device list (empty otherwise)
    sph_multi_t* msys;
    bool multiDirty;                            // multi-GPU mode: the host mirrors hold changes the slabs have not seen yet
    int multiSync();                            // ... bring the mirrors up to date with the slabs (before a partial write)
    int multiFlush();                           // ... push the mirrors to the slabs (before a step)
    std::string err;
    size_t memParticles, memCells;              // what the device buffers of `sys` were allocated for
};

#endif  // SPH_HOST_H
