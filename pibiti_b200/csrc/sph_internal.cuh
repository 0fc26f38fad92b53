// sph_internal.cuh -- the handle behind sph_t and the helpers shared by the translation units that implement the C ABI
// (sph_capi.cu: single-GPU entry points and the slab phases; sph_multi.cu: the multi-GPU driver).
#pragma once
#include "sph_b200.h"
#include "sph_device.cuh"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <string>

struct sph_system {
    int device = 0;
    cudaStream_t stream = nullptr;
    SimParams par;
    int nAlloc = 0, cellsAlloc = 0;

    float4 *pos[2] = {nullptr, nullptr}, *vel = nullptr, *velS = nullptr, *posP = nullptr, *velD = nullptr, *io = nullptr;
    uint32_t *idx[2] = {nullptr, nullptr}, *keyU = nullptr, *rankU = nullptr, *keyS = nullptr, *counts = nullptr;
    uint2* pairT = nullptr;
    void* nlist = nullptr;
    float4* clr = nullptr;  float* dye = nullptr;  bool visual = false;   // colour / dye outputs (sph_set_visual)
    uint16_t* ncount = nullptr;
    uint32_t *cellCount = nullptr, *cellStart = nullptr, *tileSums = nullptr, *maxCount = nullptr, *ctaRows = nullptr;
    int cur = 0;                    // live pos/idx buffer
    bool stepped = false;           // sorted scratch (keyS, posP, velD, cellStart) is valid
    bool wantCounts = false;

    SphPairConfig cfg;
    long long launches = 0;

    // CUDA-graph replay of a step (see replay_step_graph)
    struct StepGraph { cudaGraphExec_t exec = nullptr; unsigned long long version = 0; int kernels = 0; };
    StepGraph graph[2];
    unsigned long long stateVersion = 1;    // bumped by anything that changes what a step launches
    int stepsSinceChange = 0;
    bool useGraphs = true;

    // slab mode (sph_slab_*): owned z layers [zLo,zHi), one ghost layer towards each existing neighbour
    struct Slab {
        bool on = false;
        int zLo = 0, zHi = 0, hasLower = 0, hasUpper = 0, lowLayers = 0, highLayers = 0;
        long long keyOffset = 0;
        int numCellsLocal = 0;
        int first = 0, count = 0;           // live owned particles: slots [first, first+count) of pos[cur]/vel/idx[cur]
        int work = 0;                       // end of the work set (owned + appended arrivals + ghosts)
        int g0 = 0, g1 = 0, g2 = 0;         // after sort: ghosts below [0,g0), owned [g0,g1), ghosts above [g1,g2)
        int bLoEnd = 0, bHiStart = 0;       // first owned layer [g0,bLoEnd), last owned layer [bHiStart,g1)
        bool sorted = false;
        bool unpacked = false;             // sph_slab_unpack ran: the work-set size lives on the device
        int workBound = 0;                  // launch bound for kernels over the work set
        SimParams parLocal;                 // par with numCells = numCellsLocal, for the neighbour walk
    } slab;
    cudaGraphicsResource* glRes[2] = {nullptr, nullptr};    // registered GL buffers: positions, colours
    uint32_t* counters = nullptr;           // device: 4 append counters
    uint32_t* keyMax = nullptr;             // device: slab scan bound (sph_device.cuh kKeyMaxSlots)
    uint32_t* hostInts = nullptr;           // pinned: read-back of counters / cell-table entries
    // sph_exchange_arrays: three more staging buffers and a second stream, allocated on first use
    float4* xio[3] = {nullptr, nullptr, nullptr};
    cudaStream_t ioStream = nullptr;
    cudaEvent_t evIn = nullptr;

    bool timing = false;
    cudaEvent_t ev[SPH_STAGE_COUNT + 1] = {};
    cudaEvent_t evForce[2] = {};            // slab mode: the force phase does not start where the density phase ends
    float stageMs[SPH_STAGE_COUNT] = {};

    std::string err;
};

extern std::string g_sphCreateError;

inline int sph_fail(sph_system* s, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;  va_start(ap, fmt);  vsnprintf(buf, sizeof buf, fmt, ap);  va_end(ap);
    if (s) s->err = buf; else g_sphCreateError = buf;
    return code;
}
#define fail sph_fail

#define CU_TRY(s, call)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (call);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail((s), SPH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline SphLaunch sph_launcher(sph_system* s) { SphLaunch L;  L.stream = s->stream;  L.launches = &s->launches;  return L; }
