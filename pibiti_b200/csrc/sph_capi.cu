// sph_capi.cu -- the C ABI of include/sph_b200.h: buffer ownership, stage sequencing, accessors.
//
// Takes the place of the reference's host wrappers (source/CUDA/System.cu:613-756) and of the
// sequencing in cSPH::Update (source/SPH/SPH_Update.cpp:12-81): one stream, no host
// synchronisation inside a step, no texture binds, no GL map/unmap.
#include "sph_internal.cuh"
#include <algorithm>
#include <string>
#include <vector>

std::string g_sphCreateError;
#define g_createError g_sphCreateError

static const char* check_params(const SimParams* p)
{
    if (p->numParticles == 0) return "numParticles is 0";
    if (p->gridSize.x < 4 || p->gridSize.y < 4 || p->gridSize.z < 4)
        return "every grid dimension must be >= 4 cells (SURVEY.md Q3: unclamped neighbour hashes must miss)";
    if ((unsigned long long)p->gridSize.x * p->gridSize.y != p->gridSize_yx) return "gridSize_yx != gridSize.y*gridSize.x";
    if ((unsigned long long)p->gridSize_yx * p->gridSize.z != p->numCells) return "numCells != gridSize.x*y*z";
    if (p->numCells > 0x3fffff00u) return "numCells too large";      // the neighbour walk does its hash arithmetic in int32
    return nullptr;
}

static int check_range(sph_system* s, int start, int count);
static SphLaunch launcher(sph_system* s) { return sph_launcher(s); }

template <class T> static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T)); }

extern "C" const char* sph_version(void) { return "pibiti_b200 0.1 (sm_100a)"; }

// which density/force kernel pair this handle launches, e.g. "l1,threads=128,kMax=48"
extern "C" const char* sph_pair_variant(sph_t* s)
{
    if (!s) return "";
    static thread_local char buf[96];
    snprintf(buf, sizeof buf, "%s,threads=%d,cap=%d,kMax=%d", sph_pair_mode_name(s->cfg.mode), s->cfg.threads, s->cfg.cap, s->cfg.kMax);
    return buf;
}

extern "C" const char* sph_last_error(sph_t* s) { return s ? s->err.c_str() : g_createError.c_str(); }

extern "C" int sph_destroy(sph_t* s)
{
    if (!s) return SPH_ERR_ARG;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    void* bufs[] = {s->pos[0], s->pos[1], s->vel, s->velS, s->posP, s->velD, s->io, s->idx[0], s->idx[1], s->keyU,
                    s->rankU, s->keyS, s->counts, s->pairT, s->nlist, s->ncount, s->cellCount, s->cellStart, s->tileSums, s->maxCount, s->ctaRows, s->counters, s->keyMax, s->clr, s->dye};
    for (void* b : bufs) if (b) cudaFree(b);
    for (float4* b : s->xio) if (b) cudaFree(b);
    if (s->ioStream) cudaStreamDestroy(s->ioStream);
    if (s->evIn) cudaEventDestroy(s->evIn);
    for (auto& g : s->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (auto& r : s->glRes) if (r) cudaGraphicsUnregisterResource(r);
    if (s->hostInts) cudaFreeHost(s->hostInts);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    for (auto& e : s->evForce) if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return SPH_OK;
}

extern "C" int sph_create(const struct SimParams* params, int device, sph_t** out)
{
    if (!params || !out) return fail(nullptr, SPH_ERR_ARG, "sph_create: null argument");
    *out = nullptr;
    if (const char* why = check_params(params)) return fail(nullptr, SPH_ERR_PARAMS, "sph_create: %s", why);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SPH_ERR_CUDA, "sph_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail(nullptr, SPH_ERR_ARG, "sph_create: device %d out of range", device);
    cudaDeviceProp prop;
    CU_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, SPH_ERR_CUDA, "sph_create: device %d is sm_%d%d; this library is built for sm_100a only",
                    device, prop.major, prop.minor);
    CU_TRY(nullptr, cudaSetDevice(device));

    sph_system* s = new sph_system();
    s->device = device;
    s->par = *params;
    const size_t n = params->numParticles, C = params->numCells;
    s->nAlloc = (int)n;  s->cellsAlloc = (int)C;
    const size_t tiles = (C + SPH_SCAN_TILE - 1) / SPH_SCAN_TILE;

#define ALLOC(ptr, count)                                                                              \
    do {                                                                                               \
        cudaError_t _e = dalloc(&(ptr), (count));                                                      \
        if (_e != cudaSuccess) {                                                                       \
            int rc = fail(nullptr, SPH_ERR_CUDA, "sph_create: cudaMalloc(%s, %zu) failed: %s", #ptr,   \
                          (size_t)(count), cudaGetErrorString(_e));                                    \
            sph_destroy(s);                                                                            \
            return rc;                                                                                 \
        }                                                                                              \
    } while (0)

    sph_pair_default_config(&s->cfg);
    if (const char* env = getenv("SPH_B200_PAIR_CFG")) {     // "tma|l1|rm,threads,cap,kMax" -- tuning / test aid
        char mode[8] = {0};  int t = 0, c = 0, k = 0;
        if (sscanf(env, "%7[a-z0-9],%d,%d,%d", mode, &t, &c, &k) == 4 && t >= 32 && t <= 256 && t % 32 == 0 && c > 0 && c <= 3500 && k > 0 && k <= 1024 && k % 4 == 0) {
            const bool tma = strcmp(mode, "tma") == 0, rm = strcmp(mode, "rm") == 0;
            // staged variant: cap candidates (32 B each in the force kernel) plus the list block in shared memory;
            // rm variant: cap = records per particle (kMax unused)
            const size_t smemNeed = tma ? (size_t)c * 32 + (size_t)k * t * 2 + 1024 : 0;
            if (smemNeed <= prop.sharedMemPerBlockOptin && !(rm && (c > 64 || t > 128))) {
                s->cfg.mode = tma ? SPH_PAIR_TMA : rm ? SPH_PAIR_RM : SPH_PAIR_L1;
                s->cfg.threads = t;  s->cfg.cap = c;  s->cfg.kMax = k;
            }
        }
    }
    ALLOC(s->pos[0], n);  ALLOC(s->pos[1], n);  ALLOC(s->vel, n);  ALLOC(s->velS, n);
    ALLOC(s->posP, n);    ALLOC(s->velD, n);    ALLOC(s->io, n);
    ALLOC(s->idx[0], n);  ALLOC(s->idx[1], n);  ALLOC(s->keyU, n);  ALLOC(s->rankU, n);  ALLOC(s->keyS, n);
    ALLOC(s->counts, n);  ALLOC(s->pairT, n);
    { unsigned char* lb = nullptr;  ALLOC(lb, sph_pair_list_bytes(s->cfg, (int)n));  s->nlist = lb; }  ALLOC(s->ncount, n);
    ALLOC(s->ctaRows, sph_pair_blocks(s->cfg, (int)n));
    ALLOC(s->counters, 16);  ALLOC(s->keyMax, kKeyMaxSlots + 8);
    ALLOC(s->cellCount, C + 16);  ALLOC(s->cellStart, C + 16);  ALLOC(s->tileSums, tiles + 1);  ALLOC(s->maxCount, kMaxCountWords);
#undef ALLOC

    // failures past this point release what was allocated (the handle, its buffers, stream and events)
#define CREATE_TRY(call)                                                                               \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            int rc = fail(nullptr, SPH_ERR_CUDA, "sph_create: %s failed: %s", #call, cudaGetErrorString(_e)); \
            cudaGetLastError();                                                                        \
            sph_destroy(s);                                                                            \
            return rc;                                                                                 \
        }                                                                                              \
    } while (0)
    CREATE_TRY(cudaMallocHost((void**)&s->hostInts, 64));
    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& ev : s->ev) CREATE_TRY(cudaEventCreate(&ev));
    for (auto& ev : s->evForce) CREATE_TRY(cudaEventCreate(&ev));

    CREATE_TRY(sph_pair_prepare(s->cfg));
    if (const char* env = getenv("SPH_B200_GRAPHS")) s->useGraphs = atoi(env) != 0;

    CREATE_TRY(cudaMemsetAsync(s->pos[0], 0, n * sizeof(float4), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->vel, 0, n * sizeof(float4), s->stream));
    // posP / velD rows of ghost slots are read by the filtering walk before the first rho,p exchange fills them
    CREATE_TRY(cudaMemsetAsync(s->posP, 0, n * sizeof(float4), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->velD, 0, n * sizeof(float4), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->cellCount, 0, (C + 16) * sizeof(uint32_t), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->maxCount, 0, kMaxCountWords * sizeof(uint32_t), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->counters, 0, 16 * sizeof(uint32_t), s->stream));
    CREATE_TRY(cudaMemsetAsync(s->keyMax, 0, (kKeyMaxSlots + 8) * sizeof(uint32_t), s->stream));
    sph_launch_iota(launcher(s), s->idx[0], (int)n);
    CREATE_TRY(cudaStreamSynchronize(s->stream));
#undef CREATE_TRY
    *out = s;
    return SPH_OK;
}

// boundary effects a z-slab decomposition cannot honour: they move a particle further than one cell layer in a step
static const char* slab_unsupported(const SimParams* p)
{
    if (p->bndEffZ == BND_EFF_WRAP || p->bndEffZ == BND_EFF_CYCLE)
        return "slab mode does not support the Z wrap/cycle teleport (bndEffZ=1,2): it would make the first and last slab neighbours";
    if (p->bndType == BND_PUMP_Y)
        return "slab mode does not support the pump boundary (bndType=5): its exit->inlet teleport moves particles across slabs";
    return nullptr;
}

extern "C" int sph_set_params(sph_t* s, const struct SimParams* p)
{
    if (!s || !p) return SPH_ERR_ARG;
    if (const char* why = check_params(p)) return fail(s, SPH_ERR_PARAMS, "sph_set_params: %s", why);
    if ((int)p->numParticles > s->nAlloc || (int)p->numCells > s->cellsAlloc)
        return fail(s, SPH_ERR_PARAMS, "sph_set_params: numParticles/numCells exceed the allocation of sph_create");
    if (s->slab.on) {
        if (p->numParticles != s->par.numParticles)
            return fail(s, SPH_ERR_PARAMS, "sph_set_params: numParticles is the slab capacity and cannot change in slab mode");
        if (const char* why = slab_unsupported(p)) return fail(s, SPH_ERR_PARAMS, "sph_set_params: %s", why);
    }
    if (p->numParticles != s->par.numParticles && s->stepped) {
        // Slots are in sorted order after a step and idx[] is a permutation of the OLD particle range: bring the state
        // back to original order first, so that the new range [0, numParticles) keeps exactly the particles the
        // reference would keep (it simply runs its kernels over the first numParticles entries).
        CU_TRY(s, cudaSetDevice(s->device));
        const int nOld = (int)s->par.numParticles, cur = s->cur;
        SphLaunch L = launcher(s);
        sph_launch_unpermute4(L, s->pos[cur], s->idx[cur], s->pos[cur ^ 1], 0, nOld, nOld);
        sph_launch_unpermute4(L, s->vel, s->idx[cur], s->velS, 0, nOld, nOld);
        CU_TRY(s, cudaMemcpyAsync(s->vel, s->velS, (size_t)nOld * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream));
        s->cur = cur ^ 1;
        sph_launch_iota(L, s->idx[s->cur], s->nAlloc);
        CU_TRY(s, cudaGetLastError());
    }
    if (p->numCells != s->par.numCells || p->numParticles != s->par.numParticles) s->stepped = false;
    if (memcmp(&s->par, p, sizeof(SimParams)) != 0) { s->stateVersion++;  s->stepsSinceChange = 0; }
    s->par = *p;
    if (s->slab.on) {
        s->slab.parLocal = *p;
        s->slab.parLocal.numCells = (uint)s->slab.numCellsLocal;
    }
    return SPH_OK;
}

// Back to the state sph_create leaves behind (slot order == original order, nothing stepped), without touching
// the allocation: what a scene switch needs when the new scene fits the buffers (cSPH::InitScene).
extern "C" int sph_reset_state(sph_t* s)
{
    if (!s) return SPH_ERR_ARG;
    if (s->slab.on) return fail(s, SPH_ERR_STATE, "sph_reset_state: handle is in slab mode");
    CU_TRY(s, cudaSetDevice(s->device));
    const size_t n = (size_t)s->nAlloc;
    s->cur = 0;  s->stepped = false;  s->stateVersion++;  s->stepsSinceChange = 0;
    CU_TRY(s, cudaMemsetAsync(s->pos[0], 0, n * sizeof(float4), s->stream));
    CU_TRY(s, cudaMemsetAsync(s->vel, 0, n * sizeof(float4), s->stream));
    if (s->clr) CU_TRY(s, cudaMemsetAsync(s->clr, 0, n * sizeof(float4), s->stream));
    if (s->dye) CU_TRY(s, cudaMemsetAsync(s->dye, 0, n * sizeof(float), s->stream));
    sph_launch_iota(launcher(s), s->idx[0], (int)n);
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

// dye concentrations (dDyeColor, original particle order) from host memory: checkpoint restore
extern "C" int sph_set_dye(sph_t* s, const float* dye, int start, int count)
{
    if (!s || !dye) return SPH_ERR_ARG;
    if (int rc = check_range(s, start, count)) return rc;
    if (!s->dye) return fail(s, SPH_ERR_STATE, "sph_set_dye: enable the visual outputs first (sph_set_visual)");
    CU_TRY(s, cudaSetDevice(s->device));
    CU_TRY(s, cudaMemcpyAsync(s->dye + start, dye, (size_t)count * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

extern "C" int sph_get_params(sph_t* s, struct SimParams* out)
{
    if (!s || !out) return SPH_ERR_ARG;
    *out = s->par;
    return SPH_OK;
}

// the kernel sequence of one step, reading slot buffers `in` and writing `in^1`
static void enqueue_step(sph_system* s, int in, bool tm)
{
    const int n = (int)s->par.numParticles, C = (int)s->par.numCells, outb = in ^ 1;
    SphLaunch L = launcher(s);
    if (tm) cudaEventRecord(s->ev[0], s->stream);
    sph_launch_integrate_hash(L, s->par, s->pos[in], s->vel, s->keyU, s->rankU, s->cellCount, 0, n);
    if (tm) cudaEventRecord(s->ev[1], s->stream);
    sph_launch_scan(L, s->cellCount, s->cellStart, s->tileSums, s->maxCount, C, C);
    sph_launch_bucket(L, s->keyU, s->rankU, s->idx[in], s->cellStart, s->pairT, n);
    if (tm) cudaEventRecord(s->ev[2], s->stream);
    sph_launch_rank_gather(L, s->pairT, s->keyU, s->cellStart, s->pos[in], s->vel,
                           s->pos[outb], s->velS, s->idx[outb], s->keyS, n, nullptr, s->maxCount, C);
    if (tm) cudaEventRecord(s->ev[3], s->stream);
    sph_launch_density(L, s->cfg, s->par, s->pos[outb], s->velS, s->keyS, s->cellStart, s->maxCount,
                       s->posP, s->velD, s->wantCounts ? s->counts : nullptr, s->nlist, s->ncount, s->ctaRows, 0, n);
    if (tm) cudaEventRecord(s->ev[4], s->stream);
    sph_launch_force(L, s->cfg, s->par, s->posP, s->velD, s->velS, s->keyS, s->cellStart, s->maxCount,
                     s->nlist, s->ncount, s->ctaRows, s->vel, 0, n);
    if (sph_needs_obstacles(s->par)) sph_launch_obstacles(L, s->par, s->posP, s->velD, s->vel, 0, n);
    if (s->visual)
        sph_launch_color_dye(L, s->par, s->pos[outb], s->velS, s->velD, s->vel, s->keyS, s->cellStart, s->idx[outb],
                             s->clr, s->dye, 0, n);
    if (tm) cudaEventRecord(s->ev[5], s->stream);
}

// A step whose parameters did not change since the previous one is replayed as a CUDA graph (one per ping-pong
// parity): a launch-bound small scene (the default 57K-particle scene is ~60 us of kernels) then pays one graph
// launch instead of eight kernel launches.  Parameters are kernel arguments, so a change invalidates the graphs.
static bool replay_step_graph(sph_system* s, int in)
{
    sph_system::StepGraph& g = s->graph[in];
    if (g.exec && g.version != s->stateVersion) {
        cudaGraphExecDestroy(g.exec);  g.exec = nullptr;
    }
    if (!g.exec) {
        const long long before = s->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
        enqueue_step(s, in, false);
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        g.kernels = (int)(s->launches - before);
        s->launches = before;
        if (e != cudaSuccess || !graph) { cudaGetLastError(); return false; }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError();  g.exec = nullptr;  return false; }
        g.version = s->stateVersion;
    }
    if (cudaGraphLaunch(g.exec, s->stream) != cudaSuccess) { cudaGetLastError(); return false; }
    s->launches += g.kernels;
    return true;
}

extern "C" int sph_step(sph_t* s, int nsteps)
{
    if (!s || nsteps < 0) return SPH_ERR_ARG;
    if (s->slab.on) return fail(s, SPH_ERR_STATE, "sph_step: handle is in slab mode; drive it with the sph_slab_* phases");
    CU_TRY(s, cudaSetDevice(s->device));
    for (int it = 0; it < nsteps; it++) {
        const bool tm = s->timing && it == nsteps - 1;
        // graphs only once the parameters have been stable for two steps (a scene whose host prologue edits them
        // every step -- wave phase, rotor angle -- would re-capture each time)
        const bool useGraph = s->useGraphs && !s->timing && s->stepsSinceChange >= 2;
        if (!(useGraph && replay_step_graph(s, s->cur))) enqueue_step(s, s->cur, tm);
        s->cur ^= 1;
        s->stepped = true;
        if (s->stepsSinceChange < 1000) s->stepsSinceChange++;
    }
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_sync(sph_t* s)
{
    if (!s) return SPH_ERR_ARG;
    CU_TRY(s, cudaSetDevice(s->device));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

static int check_range(sph_system* s, int start, int count)
{
    if (start < 0 || count < 0 || (long long)start + count > (long long)s->par.numParticles)
        return fail(s, SPH_ERR_ARG, "particle range [%d,%d) outside [0,%u)", start, start + count, s->par.numParticles);
    return SPH_OK;
}

extern "C" int sph_set_array_device(sph_t* s, int which, const float* d_xyzw, int start, int count)
{
    if (!s || !d_xyzw) return SPH_ERR_ARG;
    if (int rc = check_range(s, start, count)) return rc;
    if (which != SPH_POS && which != SPH_VEL) return fail(s, SPH_ERR_ARG, "sph_set_array: only SPH_POS / SPH_VEL can be written");
    CU_TRY(s, cudaSetDevice(s->device));
    float4* dst = which == SPH_POS ? s->pos[s->cur] : s->vel;
    sph_launch_permute4(launcher(s), dst, s->idx[s->cur], (const float4*)d_xyzw, start, count, (int)s->par.numParticles);
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_set_array(sph_t* s, int which, const float* xyzw, int start, int count)
{
    if (!s || !xyzw) return SPH_ERR_ARG;
    if (int rc = check_range(s, start, count)) return rc;
    CU_TRY(s, cudaSetDevice(s->device));
    CU_TRY(s, cudaMemcpyAsync(s->io, xyzw, (size_t)count * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    int rc = sph_set_array_device(s, which, (const float*)s->io, start, count);
    // the staging buffer is reused by the next call: finish before returning (the reference's
    // setArray is a blocking glBufferSubData / cudaMemcpy as well)
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return rc;
}

extern "C" int sph_get_array_device(sph_t* s, int which, float* d_out, int start, int count)
{
    if (!s || !d_out) return SPH_ERR_ARG;
    if (int rc = check_range(s, start, count)) return rc;
    CU_TRY(s, cudaSetDevice(s->device));
    const int n = (int)s->par.numParticles;
    SphLaunch L = launcher(s);
    switch (which) {
    case SPH_POS: sph_launch_unpermute4(L, s->pos[s->cur], s->idx[s->cur], (float4*)d_out, start, count, n); break;
    case SPH_VEL: sph_launch_unpermute4(L, s->vel, s->idx[s->cur], (float4*)d_out, start, count, n); break;
    case SPH_DENSITY:
        if (!s->stepped) return fail(s, SPH_ERR_STATE, "density is only defined after a step");
        sph_launch_unpermute_w(L, s->velD, s->idx[s->cur], d_out, start, count, n); break;
    case SPH_PRESSURE:
        if (!s->stepped) return fail(s, SPH_ERR_STATE, "pressure is only defined after a step");
        sph_launch_unpermute_w(L, s->posP, s->idx[s->cur], d_out, start, count, n); break;
    case SPH_COLOR:
    case SPH_DYE: {
        if (!s->visual || !s->stepped) return fail(s, SPH_ERR_STATE, "colour / dye need sph_set_visual(s, 1) and a step");
        const void* src = which == SPH_COLOR ? (const void*)(s->clr + start) : (const void*)(s->dye + start);
        CU_TRY(s, cudaMemcpyAsync(d_out, src, (size_t)count * (which == SPH_COLOR ? 16 : 4), cudaMemcpyDeviceToDevice, s->stream));
        break;
    }
    default: return fail(s, SPH_ERR_ARG, "sph_get_array: array %d not available", which);
    }
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

// cuda_gl_interop.h needs the GL headers, which this image does not have; the entry point lives in libcudart
extern "C" cudaError_t CUDARTAPI cudaGraphicsGLRegisterBuffer(cudaGraphicsResource** resource, unsigned int buffer, unsigned int flags);

extern "C" int sph_gl_register(sph_t* s, int which, unsigned int glBuffer)
{
    if (!s || (which != SPH_POS && which != SPH_COLOR)) return SPH_ERR_ARG;
    CU_TRY(s, cudaSetDevice(s->device));
    cudaGraphicsResource*& res = s->glRes[which == SPH_POS ? 0 : 1];
    if (res) { CU_TRY(s, cudaStreamSynchronize(s->stream));  CU_TRY(s, cudaGraphicsUnregisterResource(res));  res = nullptr; }
    if (glBuffer == 0) return SPH_OK;
    cudaError_t e = cudaGraphicsGLRegisterBuffer(&res, glBuffer, cudaGraphicsRegisterFlagsWriteDiscard);
    if (e != cudaSuccess) {
        res = nullptr;
        cudaGetLastError();
        return fail(s, SPH_ERR_CUDA, "sph_gl_register: cudaGraphicsGLRegisterBuffer(%u) failed: %s (is a GL context current?)",
                    glBuffer, cudaGetErrorString(e));
    }
    return SPH_OK;
}

extern "C" int sph_gl_update(sph_t* s)
{
    if (!s) return SPH_ERR_ARG;
    CU_TRY(s, cudaSetDevice(s->device));
    const int n = (int)s->par.numParticles;
    for (int k = 0; k < 2; k++) {
        if (!s->glRes[k]) continue;
        if (k == 1 && !(s->visual && s->stepped)) continue;     // colours exist only after a visual step
        void* p = nullptr;  size_t bytes = 0;
        CU_TRY(s, cudaGraphicsMapResources(1, &s->glRes[k], s->stream));
        cudaError_t e = cudaGraphicsResourceGetMappedPointer(&p, &bytes, s->glRes[k]);
        int rc = SPH_OK;
        if (e != cudaSuccess) rc = fail(s, SPH_ERR_CUDA, "sph_gl_update: %s", cudaGetErrorString(e));
        else if (bytes < (size_t)n * sizeof(float4)) rc = fail(s, SPH_ERR_ARG, "sph_gl_update: GL buffer holds %zu bytes, %zu needed", bytes, (size_t)n * sizeof(float4));
        else rc = sph_get_array_device(s, k == 0 ? SPH_POS : SPH_COLOR, (float*)p, 0, n);
        CU_TRY(s, cudaGraphicsUnmapResources(1, &s->glRes[k], s->stream));
        if (rc != SPH_OK) return rc;
    }
    return SPH_OK;
}

extern "C" int sph_set_visual(sph_t* s, int enable)
{
    if (!s) return SPH_ERR_ARG;
    CU_TRY(s, cudaSetDevice(s->device));
    if (enable && !s->clr) {
        CU_TRY(s, cudaMalloc((void**)&s->clr, (size_t)s->nAlloc * sizeof(float4)));
        CU_TRY(s, cudaMalloc((void**)&s->dye, (size_t)s->nAlloc * sizeof(float)));
        CU_TRY(s, cudaMemsetAsync(s->clr, 0, (size_t)s->nAlloc * sizeof(float4), s->stream));
        CU_TRY(s, cudaMemsetAsync(s->dye, 0, (size_t)s->nAlloc * sizeof(float), s->stream));
    }
    s->visual = enable != 0;
    s->stateVersion++;
    return SPH_OK;
}

extern "C" int sph_get_array(sph_t* s, int which, float* out, int start, int count)
{
    if (!s || !out) return SPH_ERR_ARG;
    int rc = sph_get_array_device(s, which, (float*)s->io, start, count);
    if (rc) return rc;
    size_t elem = (which == SPH_POS || which == SPH_VEL || which == SPH_COLOR) ? sizeof(float4) : sizeof(float);
    CU_TRY(s, cudaMemcpyAsync(out, s->io, (size_t)count * elem, cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

// The whole state out to the host and a whole new state in, in one call: the device->host copies of the CURRENT positions and
// velocities run while the host->device copies of the NEW ones do (PCIe is full duplex; a get followed by a set would
// serialise them).  Semantics = sph_get_array(POS), sph_get_array(VEL), then sph_set_array(POS), sph_set_array(VEL), all over
// the full range, original particle order; returns when all four copies are done.  Pinned host memory on both sides is what
// makes the overlap real.
extern "C" int sph_exchange_arrays(sph_t* s, float* outPos, float* outVel, const float* inPos, const float* inVel)
{
    if (!s || !outPos || !outVel || !inPos || !inVel) return SPH_ERR_ARG;
    if (s->slab.on) return fail(s, SPH_ERR_STATE, "sph_exchange_arrays: handle is in slab mode");
    CU_TRY(s, cudaSetDevice(s->device));
    const int n = (int)s->par.numParticles;
    const size_t bytes = (size_t)n * sizeof(float4);
    if (!s->ioStream) {
        for (float4*& b : s->xio) CU_TRY(s, cudaMalloc((void**)&b, (size_t)s->nAlloc * sizeof(float4)));
        CU_TRY(s, cudaStreamCreateWithFlags(&s->ioStream, cudaStreamNonBlocking));
        CU_TRY(s, cudaEventCreateWithFlags(&s->evIn, cudaEventDisableTiming));
    }
    SphLaunch L = launcher(s);
    float4 *outP = s->io, *outV = s->xio[0], *inP = s->xio[1], *inV = s->xio[2];
    // in: host -> staging on the second stream, at once
    CU_TRY(s, cudaMemcpyAsync(inP, inPos, bytes, cudaMemcpyHostToDevice, s->ioStream));
    CU_TRY(s, cudaMemcpyAsync(inV, inVel, bytes, cudaMemcpyHostToDevice, s->ioStream));
    CU_TRY(s, cudaEventRecord(s->evIn, s->ioStream));
    // out: current state into original order, then to the host on the solver stream
    sph_launch_unpermute4(L, s->pos[s->cur], s->idx[s->cur], outP, 0, n, n);
    sph_launch_unpermute4(L, s->vel, s->idx[s->cur], outV, 0, n, n);
    CU_TRY(s, cudaMemcpyAsync(outPos, outP, bytes, cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(outVel, outV, bytes, cudaMemcpyDeviceToHost, s->stream));
    // the new state replaces the old one once both have been read / have arrived
    CU_TRY(s, cudaStreamWaitEvent(s->stream, s->evIn, 0));
    sph_launch_permute4(L, s->pos[s->cur], s->idx[s->cur], inP, 0, n, n);
    sph_launch_permute4(L, s->vel, s->idx[s->cur], inV, 0, n, n);
    CU_TRY(s, cudaGetLastError());
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

extern "C" int sph_device_buffers(sph_t* s, const float** d_pos, const float** d_vel, const uint32_t** d_index,
                                  const uint32_t** d_cellStart)
{
    if (!s) return SPH_ERR_ARG;
    if (d_pos) *d_pos = (const float*)s->pos[s->cur];
    if (d_vel) *d_vel = (const float*)s->vel;
    if (d_index) *d_index = s->idx[s->cur];
    if (d_cellStart) *d_cellStart = s->stepped ? s->cellStart : nullptr;
    return SPH_OK;
}

extern "C" int sph_debug_dump(sph_t* s, int what, void* out, size_t outBytes)
{
    if (!s || !out) return SPH_ERR_ARG;
    if (!s->stepped) return fail(s, SPH_ERR_STATE, "sph_debug_dump: no step has run yet");
    CU_TRY(s, cudaSetDevice(s->device));
    const size_t n = s->par.numParticles, C = s->par.numCells;
    SphLaunch L = launcher(s);
    const void* src = nullptr;  size_t bytes = 0;
    std::vector<float> tmp;
    switch (what) {
    case SPH_DUMP_SORTED_PAIRS:
        sph_launch_pack_pairs(L, s->keyS, s->idx[s->cur], s->pairT, (int)n);
        src = s->pairT;  bytes = n * 8;  break;
    case SPH_DUMP_CELL_START:
    case SPH_DUMP_CELL_END:
        // cellCount is all zeros between steps: borrow it as the output buffer and re-zero it afterwards
        sph_launch_cell_table_dump(L, s->cellStart, what == SPH_DUMP_CELL_START ? s->cellCount : nullptr,
                                   what == SPH_DUMP_CELL_END ? s->cellCount : nullptr, (int)C);
        src = s->cellCount;  bytes = C * 4;  break;
    case SPH_DUMP_SORTED_POS: src = s->pos[s->cur];  bytes = n * 16;  break;
    case SPH_DUMP_SORTED_VEL: src = s->velS;  bytes = n * 16;  break;
    case SPH_DUMP_PRESSURE:
    case SPH_DUMP_DENSITY: {
        if (outBytes < n * 4) return fail(s, SPH_ERR_ARG, "sph_debug_dump: output buffer too small");
        tmp.resize(4 * n);
        CU_TRY(s, cudaMemcpyAsync(tmp.data(), what == SPH_DUMP_PRESSURE ? s->posP : s->velD, n * 16,
                                  cudaMemcpyDeviceToHost, s->stream));
        CU_TRY(s, cudaStreamSynchronize(s->stream));
        float* o = (float*)out;
        for (size_t i = 0; i < n; i++) o[i] = tmp[4 * i + 3];
        return SPH_OK;
    }
    case SPH_DUMP_NEIGHBOR_COUNTS:
        // recompute density on the sorted state with counting enabled (same kernel, COUNT=true)
        sph_launch_density(L, s->cfg, s->par, s->pos[s->cur], s->velS, s->keyS, s->cellStart, s->maxCount,
                           s->posP, s->velD, s->counts, s->nlist, s->ncount, s->ctaRows, 0, (int)n);
        src = s->counts;  bytes = n * 4;  break;
    default: return fail(s, SPH_ERR_ARG, "sph_debug_dump: unknown item %d", what);
    }
    if (outBytes < bytes) return fail(s, SPH_ERR_ARG, "sph_debug_dump: output buffer too small (%zu < %zu)", outBytes, bytes);
    CU_TRY(s, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, s->stream));
    if (what == SPH_DUMP_CELL_START || what == SPH_DUMP_CELL_END)
        CU_TRY(s, cudaMemsetAsync(s->cellCount, 0, C * 4, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_get_timings(sph_t* s, float* msPerStage, int enable)
{
    if (!s) return SPH_ERR_ARG;
    if (msPerStage && s->timing && s->stepped) {
        CU_TRY(s, cudaSetDevice(s->device));
        CU_TRY(s, cudaStreamSynchronize(s->stream));
        for (int k = 0; k < SPH_STAGE_COUNT; k++) {
            float ms = 0.f;
            cudaEvent_t a = s->ev[k], b = s->ev[k + 1];
            if (s->slab.on) {                       // slab mode records the two pair kernels only
                if (k == SPH_STAGE_FORCE) { a = s->evForce[0];  b = s->evForce[1]; }
                else if (k != SPH_STAGE_DENSITY) { msPerStage[k] = -1.f;  continue; }
            }
            if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { ms = -1.f; cudaGetLastError(); }
            msPerStage[k] = ms;
        }
    } else if (msPerStage) {
        for (int k = 0; k < SPH_STAGE_COUNT; k++) msPerStage[k] = -1.f;
    }
    s->timing = enable != 0;
    return SPH_OK;
}

extern "C" int sph_kernel_launch_count(sph_t* s, long long* launches)
{
    if (!s || !launches) return SPH_ERR_ARG;
    *launches = s->launches;
    return SPH_OK;
}

extern "C" void* sph_cuda_stream(sph_t* s) { return s ? (void*)s->stream : nullptr; }

// ================================================================================================
// Slab decomposition
// ================================================================================================

#define SLAB_CHECK(s)                                                                    \
    if (!(s)) return SPH_ERR_ARG;                                                        \
    if (!(s)->slab.on) return fail((s), SPH_ERR_STATE, "handle is not in slab mode (sph_slab_configure)"); \
    CU_TRY((s), cudaSetDevice((s)->device));

extern "C" int sph_slab_configure(sph_t* s, int zLo, int zHi, int hasLower, int hasUpper)
{
    if (!s) return SPH_ERR_ARG;
    const int gz = (int)s->par.gridSize.z;
    if (zLo < 0 || zHi > gz || zHi - zLo < 1) return fail(s, SPH_ERR_ARG, "sph_slab_configure: bad layer range [%d,%d) of %d", zLo, zHi, gz);
    if ((hasLower && zLo == 0) || (hasUpper && zHi == gz)) return fail(s, SPH_ERR_ARG, "sph_slab_configure: neighbour beyond the grid");
    if (const char* why = slab_unsupported(&s->par)) return fail(s, SPH_ERR_PARAMS, "sph_slab_configure: %s", why);
    sph_system::Slab& b = s->slab;
    b.on = true;  b.zLo = zLo;  b.zHi = zHi;  b.hasLower = hasLower ? 1 : 0;  b.hasUpper = hasUpper ? 1 : 0;
    b.lowLayers = b.hasLower;  b.highLayers = b.hasUpper;
    b.keyOffset = (long long)(zLo - b.lowLayers) * s->par.gridSize_yx;
    b.numCellsLocal = (int)s->par.gridSize_yx * (zHi - zLo + b.lowLayers + b.highLayers);
    // the tables hold cellsAlloc+16 entries: room for the dummy cell behind the local cells
    if (b.numCellsLocal > s->cellsAlloc) return fail(s, SPH_ERR_PARAMS, "sph_slab_configure: local cell table exceeds the allocation");
    b.first = b.count = b.work = 0;  b.sorted = false;
    b.parLocal = s->par;
    b.parLocal.numCells = (uint)b.numCellsLocal;
    s->stepped = false;
    return SPH_OK;
}

extern "C" int sph_slab_set_owned(sph_t* s, const float* d_records, int count)
{
    SLAB_CHECK(s);
    if (count < 0 || count > s->nAlloc) return fail(s, SPH_ERR_ARG, "sph_slab_set_owned: %d particles exceed capacity %d", count, s->nAlloc);
    sph_launch_slab_append(launcher(s), d_records, count, s->pos[s->cur], s->vel, s->idx[s->cur], 0);
    s->slab.first = 0;  s->slab.count = count;  s->slab.work = count;  s->slab.sorted = false;  s->slab.unpacked = false;
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_slab_get_owned(sph_t* s, float* d_records, int capacity, int* count)
{
    SLAB_CHECK(s);
    if (!count) return SPH_ERR_ARG;
    *count = s->slab.count;
    if (s->slab.count > capacity) return fail(s, SPH_ERR_ARG, "sph_slab_get_owned: buffer holds %d records, need %d", capacity, s->slab.count);
    sph_launch_slab_export(launcher(s), s->pos[s->cur], s->vel, s->idx[s->cur], s->stepped ? s->posP : nullptr,
                           s->stepped ? s->velD : nullptr, s->slab.first, s->slab.count, d_records);
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

extern "C" int sph_slab_integrate(sph_t* s)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    SphLaunch L = launcher(s);
    // everything outside the live owned range is retired (ghosts of the previous step)
    sph_launch_fill_u32(L, s->idx[s->cur], SPH_DEAD_INDEX, 0, b.first);
    sph_launch_fill_u32(L, s->idx[s->cur], SPH_DEAD_INDEX, b.first + b.count, b.work - (b.first + b.count));
    b.work = b.first + b.count;
    sph_launch_integrate_hash(L, s->par, s->pos[s->cur], s->vel, nullptr, nullptr, nullptr, b.first, b.count);
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

// device words used by the slab phases: counters[4] = work-set size, counters[5] = overflow flag
static const int kDevWork = 4, kDevOverflow = 5;

extern "C" int sph_slab_pack(sph_t* s, float* d_msgDown, float* d_msgUp, int capL, int capB)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (!d_msgDown || !d_msgUp || capL < 1 || capB < 1) return SPH_ERR_ARG;
    SphLaunch L = launcher(s);
    // header rows (48 bytes) hold the two append counters of each message
    CU_TRY(s, cudaMemsetAsync(d_msgDown, 0, SPH_SLAB_RECORD_BYTES, s->stream));
    CU_TRY(s, cudaMemsetAsync(d_msgUp, 0, SPH_SLAB_RECORD_BYTES, s->stream));
    uint32_t* hd = reinterpret_cast<uint32_t*>(d_msgDown);
    uint32_t* hu = reinterpret_cast<uint32_t*>(d_msgUp);
    float* leavDown = d_msgDown + SPH_SLAB_RECORD_FLOATS;
    float* leavUp = d_msgUp + SPH_SLAB_RECORD_FLOATS;
    float* bndDown = d_msgDown + (size_t)SPH_SLAB_RECORD_FLOATS * (1 + capL);
    float* bndUp = d_msgUp + (size_t)SPH_SLAB_RECORD_FLOATS * (1 + capL);
    sph_launch_slab_take_leavers(L, s->par, s->pos[s->cur], s->vel, s->idx[s->cur], b.first, b.count,
                                 b.zLo, b.zHi, b.hasLower, b.hasUpper, leavDown, capL, leavUp, capL, hd + 0, hu + 0);
    sph_launch_slab_boundary(L, s->par, s->pos[s->cur], s->vel, s->idx[s->cur], b.work, b.zLo, b.zHi,
                             b.hasLower, b.hasUpper, bndDown, capB, bndUp, capB, hd + 1, hu + 1);
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

// sph_slab_integrate followed by sph_slab_pack, as one kernel: the state is read and written once
extern "C" int sph_slab_integrate_pack(sph_t* s, float* d_msgDown, float* d_msgUp, int capL, int capB)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (!d_msgDown || !d_msgUp || capL < 1 || capB < 1) return SPH_ERR_ARG;
    CU_TRY(s, cudaMemsetAsync(d_msgDown, 0, SPH_SLAB_RECORD_BYTES, s->stream));
    CU_TRY(s, cudaMemsetAsync(d_msgUp, 0, SPH_SLAB_RECORD_BYTES, s->stream));
    const int work = b.work;
    b.work = b.first + b.count;                 // the slots behind the owned range are retired by the kernel
    sph_launch_slab_integrate_pack(launcher(s), s->par, s->pos[s->cur], s->vel, s->idx[s->cur], b.first, b.count, work,
                                   b.zLo, b.zHi, b.hasLower, b.hasUpper,
                                   d_msgDown + SPH_SLAB_RECORD_FLOATS, d_msgUp + SPH_SLAB_RECORD_FLOATS, capL,
                                   d_msgDown + (size_t)SPH_SLAB_RECORD_FLOATS * (1 + capL),
                                   d_msgUp + (size_t)SPH_SLAB_RECORD_FLOATS * (1 + capL), capB,
                                   reinterpret_cast<uint32_t*>(d_msgDown), reinterpret_cast<uint32_t*>(d_msgUp));
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_slab_unpack(sph_t* s, const float* d_inBelow, const float* d_inAbove,
                               const float* d_ownDown, const float* d_ownUp, int capL, int capB)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    CU_TRY(s, cudaMemsetAsync(s->counters + kDevOverflow, 0, sizeof(uint32_t), s->stream));
    sph_launch_slab_unpack(launcher(s), b.hasLower ? d_inBelow : nullptr, b.hasUpper ? d_inAbove : nullptr,
                           b.hasLower ? d_ownDown : nullptr, b.hasUpper ? d_ownUp : nullptr, capL, capB,
                           s->pos[s->cur], s->vel, s->idx[s->cur], b.work, s->nAlloc, s->counters + kDevWork);
    // upper bound of the work set until the sort reads the real size back
    long long bound = (long long)b.work + 4LL * capL + 2LL * capB;
    b.workBound = (int)(bound < s->nAlloc ? bound : s->nAlloc);
    b.unpacked = true;
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_slab_sort(sph_t* s, int* counts3)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    SphLaunch L = launcher(s);
    const int in = s->cur, outb = s->cur ^ 1, CL = b.numCellsLocal;
    const int yx = (int)s->par.gridSize_yx, nz = b.zHi - b.zLo;
    if (!b.unpacked) {      // no exchange happened (single slab or first use): the work set is what the host knows
        uint32_t w = (uint32_t)b.work;
        CU_TRY(s, cudaMemcpyAsync(s->counters + kDevWork, &w, sizeof w, cudaMemcpyHostToDevice, s->stream));
        CU_TRY(s, cudaMemsetAsync(s->counters + kDevOverflow, 0, sizeof(uint32_t), s->stream));
        CU_TRY(s, cudaStreamSynchronize(s->stream));
        b.workBound = b.work;
    }
    const int W = b.workBound;
    const uint32_t* nDev = s->counters + kDevWork;
    // the scan covers the occupied part of the table only: largest live key + two layers (all neighbour lookups)
    sph_launch_slab_hash_hist(L, s->par, s->pos[in], s->idx[in], s->keyU, s->rankU, s->cellCount, W, nDev, b.keyOffset, CL,
                              s->keyMax, 2u * (uint32_t)yx);
    sph_launch_scan(L, s->cellCount, s->cellStart, s->tileSums, s->maxCount, CL + 1, CL, s->keyMax + kKeyMaxSlots);
    if (W > 0) {
        sph_launch_bucket(L, s->keyU, s->rankU, s->idx[in], s->cellStart, s->pairT, W, nDev);
        sph_launch_rank_gather(L, s->pairT, s->keyU, s->cellStart, s->pos[in], s->vel, s->pos[outb], s->velS, s->idx[outb], s->keyS, W, nDev, s->maxCount, CL);
    }
    // one read-back: five cell-table entries bracket the ghost / owned / boundary-layer ranges, plus size and overflow
    const int cells[5] = {b.lowLayers * yx, (b.lowLayers + nz) * yx, CL, (b.lowLayers + 1) * yx, (b.lowLayers + nz - 1) * yx};
    for (int k = 0; k < 5; k++)
        CU_TRY(s, cudaMemcpyAsync(s->hostInts + k, s->cellStart + cells[k], sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->hostInts + 5, s->counters + kDevWork, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->hostInts + 7, s->keyMax + kKeyMaxSlots, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemcpyAsync(s->hostInts + 9, s->counters + kDevWork + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaMemsetAsync(s->counters + kDevWork + 2, 0, sizeof(uint32_t), s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    if (s->hostInts[9])
        return fail(s, SPH_ERR_STATE, "slab decomposition lost a particle: it moved further than one cell layer along z in a step "
                    "(beyond the halo), or a boundary teleported it");
    // table entries at or above the scan bound were not written this step: no live key is that large, so they equal
    // the live total, which is the start of the dummy cell (always scanned)
    {
        const int lastTileStart = CL / SPH_SCAN_TILE * SPH_SCAN_TILE;
        for (int k = 0; k < 5; k++)
            if (cells[k] >= (int)s->hostInts[7] && cells[k] < lastTileStart) s->hostInts[k] = s->hostInts[2];
    }
    if (s->hostInts[6])
        return fail(s, SPH_ERR_ARG, "slab exchange overflow: a message section was too small (raise SlabCaps) or the work set "
                    "(%u of capacity %d) does not fit", s->hostInts[5], s->nAlloc);
    b.g0 = (int)s->hostInts[0];  b.g1 = (int)s->hostInts[1];  b.g2 = (int)s->hostInts[2];
    b.bLoEnd = (int)s->hostInts[3];  b.bHiStart = (int)s->hostInts[4];
    s->cur = outb;
    b.first = b.g0;  b.count = b.g1 - b.g0;  b.work = (int)s->hostInts[5];  b.sorted = true;  b.unpacked = false;
    if (counts3) { counts3[0] = b.g0;  counts3[1] = b.g1 - b.g0;  counts3[2] = b.g2 - b.g1; }
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_slab_density(sph_t* s)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (!b.sorted) return fail(s, SPH_ERR_STATE, "sph_slab_density: call sph_slab_sort first");
    if (s->timing) cudaEventRecord(s->ev[3], s->stream);
    sph_launch_density(launcher(s), s->cfg, b.parLocal, s->pos[s->cur], s->velS, s->keyS, s->cellStart, s->maxCount,
                       s->posP, s->velD, s->wantCounts ? s->counts : nullptr, s->nlist, s->ncount, s->ctaRows, b.first, b.count);
    if (s->timing) cudaEventRecord(s->ev[4], s->stream);
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

// rows of the first / last owned layer: nDown x (x,y,z,p) then nDown x (vx,vy,vz,rho); asynchronous
extern "C" int sph_slab_pack_dp(sph_t* s, float* d_down, float* d_up, int capRows, int* counts2)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (!b.sorted) return fail(s, SPH_ERR_STATE, "sph_slab_pack_dp: call sph_slab_sort first");
    const int nDown = b.hasLower ? b.bLoEnd - b.g0 : 0, nUp = b.hasUpper ? b.g1 - b.bHiStart : 0;
    if (nDown > capRows || nUp > capRows) return fail(s, SPH_ERR_ARG, "sph_slab_pack_dp: %d/%d rows exceed the buffers (%d)", nDown, nUp, capRows);
    if (nDown > 0) {
        CU_TRY(s, cudaMemcpyAsync(d_down, s->posP + b.g0, (size_t)nDown * 16, cudaMemcpyDeviceToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(d_down + 4 * (size_t)nDown, s->velD + b.g0, (size_t)nDown * 16, cudaMemcpyDeviceToDevice, s->stream));
    }
    if (nUp > 0) {
        CU_TRY(s, cudaMemcpyAsync(d_up, s->posP + b.bHiStart, (size_t)nUp * 16, cudaMemcpyDeviceToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(d_up + 4 * (size_t)nUp, s->velD + b.bHiStart, (size_t)nUp * 16, cudaMemcpyDeviceToDevice, s->stream));
    }
    counts2[0] = nDown;  counts2[1] = nUp;
    return SPH_OK;
}

// diagnostics: {largest real cell, work-set size, ghosts below, owned, ghosts above, retired slots}
extern "C" int sph_slab_stats(sph_t* s, int* out6)
{
    SLAB_CHECK(s);
    CU_TRY(s, cudaMemcpyAsync(s->hostInts + 8, s->maxCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CU_TRY(s, cudaStreamSynchronize(s->stream));
    const sph_system::Slab& b = s->slab;
    out6[0] = (int)s->hostInts[8];  out6[1] = b.work;  out6[2] = b.g0;  out6[3] = b.g1 - b.g0;  out6[4] = b.g2 - b.g1;
    out6[5] = b.work - b.g2;
    return SPH_OK;
}

extern "C" int sph_slab_ghost_counts(sph_t* s, int* counts2)
{
    SLAB_CHECK(s);
    counts2[0] = s->slab.g0;  counts2[1] = s->slab.g2 - s->slab.g1;
    return SPH_OK;
}

extern "C" int sph_slab_unpack_dp(sph_t* s, const float* d_below, int nBelow, const float* d_above, int nAbove)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (nBelow != b.g0 || nAbove != b.g2 - b.g1)
        return fail(s, SPH_ERR_ARG, "sph_slab_unpack_dp: got %d/%d rows for %d/%d ghosts (ghost sets out of step between ranks)",
                    nBelow, nAbove, b.g0, b.g2 - b.g1);
    if (nBelow > 0) {
        CU_TRY(s, cudaMemcpyAsync(s->posP, d_below, (size_t)nBelow * 16, cudaMemcpyDeviceToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(s->velD, d_below + 4 * (size_t)nBelow, (size_t)nBelow * 16, cudaMemcpyDeviceToDevice, s->stream));
    }
    if (nAbove > 0) {
        CU_TRY(s, cudaMemcpyAsync(s->posP + b.g1, d_above, (size_t)nAbove * 16, cudaMemcpyDeviceToDevice, s->stream));
        CU_TRY(s, cudaMemcpyAsync(s->velD + b.g1, d_above + 4 * (size_t)nAbove, (size_t)nAbove * 16, cudaMemcpyDeviceToDevice, s->stream));
    }
    return SPH_OK;
}

// part 0: all owned particles.  part 1: the CTAs whose particles all lie strictly between the first and the last owned
// layer -- they have no ghost neighbours, so they do not need the rho,p rows of sph_slab_unpack_dp and can run while
// that exchange is in flight.  part 2: the remaining CTAs.  (TMA variant: part 1 is empty, part 2 is everything.)
extern "C" int sph_slab_force_part(sph_t* s, int part)
{
    SLAB_CHECK(s);
    sph_system::Slab& b = s->slab;
    if (!b.sorted) return fail(s, SPH_ERR_STATE, "sph_slab_force: call sph_slab_sort first");
    if (part < 0 || part > 2) return SPH_ERR_ARG;
    SphLaunch L = launcher(s);
    const int T = sph_pair_particles_per_cta(s->cfg), blocks = (int)sph_pair_blocks(s->cfg, b.count);
    int cLo = 0, cHi = blocks;                  // interior CTAs [cLo, cHi)
    if (s->cfg.mode == SPH_PAIR_TMA) cHi = 0;
    else {
        if (b.hasLower) cLo = (b.bLoEnd - b.g0 + T - 1) / T;
        if (b.hasUpper) cHi = (b.bHiStart - b.g0) / T;
        if (cLo > blocks) cLo = blocks;
        if (cHi < cLo) cHi = cLo;
    }
    auto run = [&](int c0, int c1) {            // force (+ obstacles) on CTAs [c0, c1)
        if (c1 <= c0) return;
        sph_launch_force(L, s->cfg, b.parLocal, s->posP, s->velD, s->velS, s->keyS, s->cellStart, s->maxCount,
                         s->nlist, s->ncount, s->ctaRows, s->vel, b.first, b.count, c0, c1 - c0);
        if (sph_needs_obstacles(s->par)) {
            const int p0 = b.first + c0 * T, p1 = std::min(b.first + b.count, b.first + c1 * T);
            sph_launch_obstacles(L, b.parLocal, s->posP, s->velD, s->vel, p0, p1 - p0);
        }
    };
    if (s->timing && part != 2) cudaEventRecord(s->evForce[0], s->stream);
    if (part == 0) run(0, blocks);
    else if (part == 1) run(cLo, cHi);
    else { run(0, cLo);  run(cHi, blocks); }
    if (part != 1) {
        if (s->timing) cudaEventRecord(s->evForce[1], s->stream);
        s->stepped = true;
    }
    CU_TRY(s, cudaGetLastError());
    return SPH_OK;
}

extern "C" int sph_slab_force(sph_t* s) { return sph_slab_force_part(s, 0); }
