// sph_multi.cu -- the multi-GPU driver behind sph_multi_* (include/sph_b200.h): one SPH system cut into z slabs, one slab
// per GPU, stepped from ONE host thread with no host synchronisation inside a step.
//
// The reference is single-GPU (one cSPH, source/SPH/SPH.h:9-50); this is the layer that lets a cSPH hold more than one
// device.  Two shapes share the code:
//   * one process drives ndev GPUs            (sph_multi_create:      ncclCommInitAll, a C++ App holding a cSPH)
//   * one process per GPU, `world` processes  (sph_multi_create_rank: ncclCommInitRank from a shared unique id; how
//                                              bench.py runs under torchrun)
// Every slab is an sph_t handle in slab mode whose bookkeeping (work-set size, sorted ranges, boundary layers) lives in
// device words (SlabDevWord, sph_device.cuh); kernels are launched over upper bounds and find their own ranges.
//
// One step, per slab (S = the handle's stream, X = the slab's exchange stream):
//   S  A  integrate the first two / last two owned layers, pack leavers + boundary-layer copies      k_slab_boundary_integrate_pack
//   X     exchange 1: particle messages to / from the z neighbours (ncclSend / ncclRecv, grouped)    -- overlaps B
//   S  B  integrate the layers in between, retire last step's ghosts, hash + count every slot        k_slab_interior_hist
//   S     append arrivals and ghosts, hash + count them                                              k_slab_unpack_hist
//   S     bounded scan, bucket, stable rank + gather (same in-cell order as the single-GPU sort), ranges to device words
//   S     density of the owned particles, rho,p rows of the two boundary layers                      k_density_*, k_slab_pack_dp
//   X     exchange 2: rho,p rows                                                                     -- overlaps the interior force
//   S     force of the CTAs without ghost neighbours; then the received rows, then the remaining CTAs
// With one process holding both ends the exchange can also be direct peer copies (SPH_B200_MULTI_XCHG=copy) or no copy at
// all (=peer): the packing kernels store leavers, boundary copies and rho,p rows straight into the neighbours' inboxes over
// NVLink and the receivers wait for the SENDERS' kernels -- only the live records travel.
// Exchanges are nearest-neighbour only; there is no collective on the data path.  NCCL is loaded at run time
// (libnccl.so.2), so the single-GPU entry points do not depend on it.
#include "sph_internal.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <vector>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

// dlopen by soname first: in a process that already carries an NCCL (PyTorch's bundled copy) this returns that one
bool load_nccl(std::string& why)
{
    if (g_nccl.lib) return true;
    const char* names[] = {getenv("SPH_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    void* h = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { why = std::string("libnccl.so.2 not found (set SPH_B200_NCCL_LIB): ") + (dlerror() ? dlerror() : "");  return false; }
    NcclApi a;
    a.lib = h;
#define SYM(field, name)                                                              \
    *(void**)(&a.field) = dlsym(h, name);                                             \
    if (!a.field) { why = std::string("NCCL symbol missing: ") + name;  return false; }
    SYM(GetUniqueId, "ncclGetUniqueId")  SYM(CommInitRank, "ncclCommInitRank")  SYM(CommInitAll, "ncclCommInitAll")
    SYM(CommDestroy, "ncclCommDestroy")  SYM(GroupStart, "ncclGroupStart")      SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")                SYM(Recv, "ncclRecv")                  SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")   SYM(AllReduce, "ncclAllReduce")        SYM(AllGather, "ncclAllGather")
#undef SYM
    g_nccl = a;
    return true;
}

constexpr int kRecFloats = SPH_SLAB_RECORD_FLOATS;     // 12 floats = 48 bytes per particle record

struct MultiRank {
    sph_system* s = nullptr;
    int device = 0, rank = 0;                   // CUDA device, global slab index
    int zLo = 0, zHi = 0, hasLower = 0, hasUpper = 0;
    int capacity = 0;
    cudaStream_t xs = nullptr;                  // exchange stream
    ncclComm_t comm = nullptr;
    float *msgDown = nullptr, *msgUp = nullptr, *inBelow = nullptr, *inAbove = nullptr;      // particle messages
    float4 *dpDown = nullptr, *dpUp = nullptr, *dpBelow = nullptr, *dpAbove = nullptr;       // rho,p rows
    cudaEvent_t evA = nullptr, evX1 = nullptr, evDp = nullptr, evX2 = nullptr;
    cudaEvent_t evPhase[12] = {};               // phase profile of the last step (sph_multi_phase_ms)
    float* xstage = nullptr;                    // sph_multi_exchange_owned: staging of the incoming records (lazy)
    cudaEvent_t evXin = nullptr;
    uint32_t* hostSt = nullptr;                 // pinned copy of the device words
    int owned = 0;                              // as of the last sph_multi_sync / set_state
};

}  // namespace

struct sph_multi {
    std::vector<MultiRank> ranks;               // the slabs this process drives
    int world = 1;
    std::vector<int> cuts;                      // world + 1 z-layer boundaries
    SimParams par;                              // global parameters (numParticles = particles of the whole system)
    int capL = 0, capB = 0;                     // message sections: leavers, boundary-layer copies (records)
    bool haveState = false;
    bool phaseTiming = false;                   // record the phase events (a handful of event records per step)
    bool copyExchange = false;                  // one process: neighbours' buffers are copied directly (cudaMemcpyPeerAsync)
                                                // instead of ncclSend/ncclRecv -- also what lets several slabs share one GPU
    bool peerStores = false;                    // one process: no copy at all -- the packing kernels store straight into the
                                                // neighbours' inboxes over NVLink (only the live records travel)
    long long steps = 0;
    int n = 0;                                  // particles of the whole system (sph_multi_set_state)
    int recutEvery = 0, recuts = 0;             // re-cut the slabs every so many steps (0: never); how often it happened
    unsigned long long bytesSent = 0;
    std::string err;
};

namespace {

std::string g_multiCreateError;

int mfail(sph_multi* m, int code, const char* fmt, ...)
{
    char buf[640];
    va_list ap;  va_start(ap, fmt);  vsnprintf(buf, sizeof buf, fmt, ap);  va_end(ap);
    if (m) m->err = buf; else g_multiCreateError = buf;
    return code;
}

#define MCU(m, call)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (call);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return mfail((m), SPH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
#define MNCCL(m, call)                                                                                 \
    do {                                                                                               \
        ncclResult_t _r = (call);                                                                      \
        if (_r != ncclSuccess)                                                                         \
            return mfail((m), SPH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

size_t msg_rows(const sph_multi* m) { return 1 + (size_t)m->capL + m->capB; }
size_t msg_bytes(const sph_multi* m) { return msg_rows(m) * kRecFloats * sizeof(float); }
size_t dp_bytes(const sph_multi* m) { return (1 + 2 * (size_t)m->capB) * sizeof(float4); }

// z cell layer of a particle, the same float arithmetic as z_cell() in sph_stream_kernels.cu
inline int host_z_cell(const SimParams& p, float z) { return (int)floorf((z - p.worldMin.z) / p.cellSize.z); }

void free_messages(MultiRank& r)
{
    cudaSetDevice(r.device);
    void* bufs[] = {r.msgDown, r.msgUp, r.inBelow, r.inAbove, r.dpDown, r.dpUp, r.dpBelow, r.dpAbove};
    for (void* b : bufs) if (b) cudaFree(b);
    r.msgDown = r.msgUp = r.inBelow = r.inAbove = nullptr;
    r.dpDown = r.dpUp = r.dpBelow = r.dpAbove = nullptr;
}

int alloc_messages(sph_multi* m, MultiRank& r)
{
    free_messages(r);
    MCU(m, cudaSetDevice(r.device));
    float** pm[] = {&r.msgDown, &r.msgUp, &r.inBelow, &r.inAbove};
    for (float** p : pm) { MCU(m, cudaMalloc((void**)p, msg_bytes(m)));  MCU(m, cudaMemset(*p, 0, msg_bytes(m))); }
    float4** pd[] = {&r.dpDown, &r.dpUp, &r.dpBelow, &r.dpAbove};
    for (float4** p : pd) { MCU(m, cudaMalloc((void**)p, dp_bytes(m)));  MCU(m, cudaMemset(*p, 0, dp_bytes(m))); }
    return SPH_OK;
}

// layer boundaries giving every slab about the same particle count (at least minLayers layers each)
bool cut_layers(const std::vector<long long>& hist, int ranks, int minLayers, std::vector<int>& cuts)
{
    const int gz = (int)hist.size();
    std::vector<long long> cum(gz);
    long long run = 0;
    for (int k = 0; k < gz; k++) { run += hist[k];  cum[k] = run; }
    cuts.assign(1, 0);
    for (int r = 1; r < ranks; r++) {
        const double target = (double)run * r / ranks;
        int c = (int)(std::lower_bound(cum.begin(), cum.end(), (long long)std::ceil(target)) - cum.begin()) + 1;
        c = std::max(c, cuts.back() + minLayers);
        c = std::min(c, gz - minLayers * (ranks - r));
        cuts.push_back(c);
    }
    cuts.push_back(gz);
    for (int r = 0; r < ranks; r++) if (cuts[r + 1] - cuts[r] < minLayers) return false;
    return true;
}

// the sort half of a step: bounded scan, bucket, stable rank + gather, sorted ranges to the device words
void enqueue_sort(MultiRank& r)
{
    sph_system* s = r.s;
    sph_system::Slab& b = s->slab;
    SphLaunch L = sph_launcher(s);
    const int in = s->cur, outb = s->cur ^ 1, CL = b.numCellsLocal, yx = (int)s->par.gridSize_yx, nz = b.zHi - b.zLo;
    uint32_t* st = s->counters;
    sph_launch_slab_scan_bound(L, s->keyMax, 2u * (uint32_t)yx, CL);
    sph_launch_scan(L, s->cellCount, s->cellStart, s->tileSums, s->maxCount, CL + 1, CL, s->keyMax + kKeyMaxSlots);
    sph_launch_bucket(L, s->keyU, s->rankU, s->idx[in], s->cellStart, s->pairT, r.capacity, st + SD_WORK);
    sph_launch_rank_gather(L, s->pairT, s->keyU, s->cellStart, s->pos[in], s->vel, s->pos[outb], s->velS, s->idx[outb], s->keyS,
                           r.capacity, st + SD_WORK, s->maxCount, CL);
    const int lo = b.lowLayers;
    const int cells[7] = {lo * yx, (lo + nz) * yx, CL, (lo + 1) * yx, (lo + nz - 1) * yx,
                          (lo + std::min(2, nz)) * yx, (lo + std::max(nz - 2, 0)) * yx};
    sph_launch_slab_bounds(L, s->cellStart, s->keyMax + kKeyMaxSlots, st, cells, b.hasLower, b.hasUpper);
    s->cur = outb;
}

int check_flags(sph_multi* m, MultiRank& r)
{
    const uint32_t* h = r.hostSt;
    r.owned = (int)(h[SD_END] - h[SD_FIRST]);
    if (h[SD_OVERFLOW])
        return mfail(m, SPH_ERR_STATE, "slab %d: exchange overflow -- a message section (%d leavers / %d boundary records) or the "
                     "work set (%u of capacity %d) was too small", r.rank, m->capL, m->capB, h[SD_WORK], r.capacity);
    if (h[SD_LOST])
        return mfail(m, SPH_ERR_STATE, "slab %d: a particle moved more than one cell layer along z in a step; the slab "
                     "decomposition cannot follow it (time step too large for the cell size?)", r.rank);
    if (h[SD_DPERR])
        return mfail(m, SPH_ERR_STATE, "slab %d: ghost sets out of step between neighbouring slabs", r.rank);
    return SPH_OK;
}

int sync_rank(sph_multi* m, MultiRank& r)
{
    MCU(m, cudaSetDevice(r.device));
    MCU(m, cudaMemcpyAsync(r.hostSt, r.s->counters, SD_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, r.s->stream));
    MCU(m, cudaStreamSynchronize(r.s->stream));
    MCU(m, cudaStreamSynchronize(r.xs));
    return check_flags(m, r);
}

// `count` particle records staged on the device become the slab's owned set: one sort without integration establishes
// the sorted ranges the first step's edge-layer pass needs
int load_slab(sph_multi* m, MultiRank& r, const float* d_records, int count)
{
    sph_system* s = r.s;
    if (sph_slab_set_owned(s, d_records, count) != SPH_OK) return mfail(m, SPH_ERR_CUDA, "slab %d: %s", r.rank, sph_last_error(s));
    uint32_t st[SD_WORDS] = {};
    st[SD_WORK] = st[SD_WORK0] = st[SD_END] = st[SD_G2] = (uint32_t)count;
    st[SD_BLO] = st[SD_BLO2] = 0;  st[SD_BHI] = st[SD_BHI2] = (uint32_t)count;
    memcpy(r.hostSt, st, sizeof st);
    MCU(m, cudaMemcpyAsync(s->counters, r.hostSt, sizeof st, cudaMemcpyHostToDevice, s->stream));
    sph_system::Slab& b = s->slab;
    const int yx = (int)s->par.gridSize_yx, nz = b.zHi - b.zLo;
    sph_launch_slab_interior_hist(sph_launcher(s), s->par, s->pos[s->cur], s->vel, s->idx[s->cur], s->keyU, s->rankU, s->cellCount,
                                  s->counters, r.capacity, b.keyOffset, b.numCellsLocal, (uint32_t)(b.lowLayers * yx),
                                  (uint32_t)((b.lowLayers + nz) * yx), s->keyMax, 0);
    enqueue_sort(r);
    // the gather put the sorted velocities into velS (where a step's force kernel reads them); the live array is vel
    MCU(m, cudaMemcpyAsync(s->vel, s->velS, (size_t)r.capacity * sizeof(float4), cudaMemcpyDeviceToDevice, s->stream));
    MCU(m, cudaGetLastError());
    MCU(m, cudaStreamSynchronize(s->stream));
    s->stepped = false;
    r.owned = count;
    return SPH_OK;
}

// message sections from the fullest layer: every slab agrees on them (fixed-size messages, counts in the header row).
// Returns whether the buffers must be (re)allocated.
bool message_geometry(sph_multi* m, long long fullest)
{
    const char* se = getenv("SPH_B200_SLAB_SAFETY");
    const double safety = se ? atof(se) : 1.5;
    // (a safety below 1 is a test aid: no slack, so that the overflow report can be exercised)
    const int capB = (int)std::min<long long>(std::max<long long>((long long)(fullest * safety), 64) + (safety >= 1.0 ? 8192 : 0), 0x3fffffff);
    const int capL = safety >= 1.0 ? std::max(capB / 4, 4096) : std::max(capB / 4, 16);
    const bool resize = capB > m->capB || capL > m->capL;       // buffers only ever grow
    if (resize) { m->capB = std::max(capB, m->capB);  m->capL = std::max(capL, m->capL); }
    return resize;
}

// slab r takes the layers the current cuts give it; message buffers follow the current geometry
int configure_rank(sph_multi* m, MultiRank& r, bool resize)
{
    MCU(m, cudaSetDevice(r.device));
    r.zLo = m->cuts[r.rank];  r.zHi = m->cuts[r.rank + 1];
    r.hasLower = r.rank > 0;  r.hasUpper = r.rank < m->world - 1;
    if (sph_slab_configure(r.s, r.zLo, r.zHi, r.hasLower, r.hasUpper) != SPH_OK)
        return mfail(m, SPH_ERR_PARAMS, "slab %d: %s", r.rank, sph_last_error(r.s));
    if (resize || !r.msgDown) if (int rc = alloc_messages(m, r)) return rc;
    return SPH_OK;
}

int create_common(sph_multi* m, const SimParams* p, int capacity)
{
    m->par = *p;
    for (MultiRank& r : m->ranks) {
        SimParams local = *p;
        local.numParticles = (uint)capacity;
        r.capacity = capacity;
        sph_t* h = nullptr;
        if (sph_create(&local, r.device, &h) != SPH_OK)
            return mfail(m, SPH_ERR_CUDA, "sph_multi: slab %d on device %d: %s", r.rank, r.device, sph_last_error(nullptr));
        r.s = h;
        if (h->cfg.mode == SPH_PAIR_TMA)
            return mfail(m, SPH_ERR_PARAMS, "sph_multi: the TMA-staged pair kernels have no device-resident ranges (use rm or l1)");
        MCU(m, cudaSetDevice(r.device));
        // highest priority: the NCCL send/recv kernels must get SM slots while the interior kernels fill the GPU, or the
        // transfer they are supposed to hide starts late (measured at N=8: 0.1 ms of waiting per step without this)
        int prLeast = 0, prGreatest = 0;
        MCU(m, cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
        MCU(m, cudaStreamCreateWithPriority(&r.xs, cudaStreamNonBlocking, prGreatest));
        cudaEvent_t* evs[] = {&r.evA, &r.evX1, &r.evDp, &r.evX2};
        for (cudaEvent_t* e : evs) MCU(m, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (cudaEvent_t& e : r.evPhase) MCU(m, cudaEventCreate(&e));
        MCU(m, cudaMallocHost((void**)&r.hostSt, SD_WORDS * sizeof(uint32_t)));
        memset(r.hostSt, 0, SD_WORDS * sizeof(uint32_t));
    }
    return SPH_OK;
}

// One nearest-neighbour exchange of every local slab on its exchange stream, after `ready` (recorded on the slab's solver
// stream): out buffers go to the neighbours' in buffers.  NCCL: grouped ncclSend / ncclRecv.  Copy mode (one process):
// each slab pulls from its neighbours once THEIR buffers are ready.
template <class T>
int exchange(sph_multi* m, size_t bytes, cudaEvent_t MultiRank::*ready, T* MultiRank::*pOutDown, T* MultiRank::*pOutUp,
             T* MultiRank::*pInBelow, T* MultiRank::*pInAbove)
{
    auto outDown = [&](MultiRank& r) { return r.*pOutDown; };
    auto outUp = [&](MultiRank& r) { return r.*pOutUp; };
    auto inBelow = [&](MultiRank& r) { return r.*pInBelow; };
    auto inAbove = [&](MultiRank& r) { return r.*pInAbove; };
    if (m->copyExchange) {
        for (MultiRank& r : m->ranks) {
            MCU(m, cudaSetDevice(r.device));
            if (r.hasLower) {
                MultiRank& lo = m->ranks[r.rank - 1];
                MCU(m, cudaStreamWaitEvent(r.xs, lo.*ready, 0));
                MCU(m, cudaMemcpyPeerAsync(inBelow(r), r.device, outUp(lo), lo.device, bytes, r.xs));
                m->bytesSent += bytes;
            }
            if (r.hasUpper) {
                MultiRank& hi = m->ranks[r.rank + 1];
                MCU(m, cudaStreamWaitEvent(r.xs, hi.*ready, 0));
                MCU(m, cudaMemcpyPeerAsync(inAbove(r), r.device, outDown(hi), hi.device, bytes, r.xs));
                m->bytesSent += bytes;
            }
        }
        return SPH_OK;
    }
    MNCCL(m, g_nccl.GroupStart());
    for (MultiRank& r : m->ranks) {
        MCU(m, cudaSetDevice(r.device));
        MCU(m, cudaStreamWaitEvent(r.xs, r.*ready, 0));
        if (r.hasLower) {
            MNCCL(m, g_nccl.Send(outDown(r), bytes, ncclChar, r.rank - 1, r.comm, r.xs));
            MNCCL(m, g_nccl.Recv(inBelow(r), bytes, ncclChar, r.rank - 1, r.comm, r.xs));
            m->bytesSent += bytes;
        }
        if (r.hasUpper) {
            MNCCL(m, g_nccl.Send(outUp(r), bytes, ncclChar, r.rank + 1, r.comm, r.xs));
            MNCCL(m, g_nccl.Recv(inAbove(r), bytes, ncclChar, r.rank + 1, r.comm, r.xs));
            m->bytesSent += bytes;
        }
    }
    MNCCL(m, g_nccl.GroupEnd());
    return SPH_OK;
}

}  // namespace

extern "C" const char* sph_multi_last_error(sph_multi_t* m) { return m ? m->err.c_str() : g_multiCreateError.c_str(); }

extern "C" int sph_multi_unique_id(unsigned char* id128)
{
    std::string why;
    if (!id128) return SPH_ERR_ARG;
    if (!load_nccl(why)) return mfail(nullptr, SPH_ERR_CUDA, "sph_multi_unique_id: %s", why.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    MNCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return SPH_OK;
}

extern "C" int sph_multi_destroy(sph_multi_t* m)
{
    if (!m) return SPH_ERR_ARG;
    for (MultiRank& r : m->ranks) {
        cudaSetDevice(r.device);
        if (r.s && r.s->stream) cudaStreamSynchronize(r.s->stream);
        if (r.xs) cudaStreamSynchronize(r.xs);
        if (r.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(r.comm);
        free_messages(r);
        cudaEvent_t evs[] = {r.evA, r.evX1, r.evDp, r.evX2};
        for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : r.evPhase) if (e) cudaEventDestroy(e);
        if (r.xstage) cudaFree(r.xstage);
        if (r.evXin) cudaEventDestroy(r.evXin);
        if (r.xs) cudaStreamDestroy(r.xs);
        if (r.hostSt) cudaFreeHost(r.hostSt);
        if (r.s) sph_destroy(r.s);
    }
    delete m;
    return SPH_OK;
}

// One process, ndev GPUs.  capacityPerSlab = particle slots of every slab (owned + ghosts + arrivals of a step).
extern "C" int sph_multi_create(const struct SimParams* params, int ndev, const int* devices, int capacityPerSlab, sph_multi_t** out)
{
    if (!params || !out || ndev < 1 || !devices || capacityPerSlab < 1) return mfail(nullptr, SPH_ERR_ARG, "sph_multi_create: bad argument");
    std::string why;
    sph_multi* m = new sph_multi;
    m->world = ndev;
    m->ranks.resize(ndev);
    for (int k = 0; k < ndev; k++) { m->ranks[k].device = devices[k];  m->ranks[k].rank = k; }
    bool repeated = false;
    for (int a = 0; a < ndev; a++) for (int b = a + 1; b < ndev; b++) repeated |= devices[a] == devices[b];
    // one process holds both ends of every exchange: direct peer copies are the default (measured against the alternatives,
    // profiles/exchange_r02_*: copies 1.701 / NCCL 1.710 / peer stores 1.724 ms per step at two GPUs; 1.73 / 1.80 at eight)
    const char* xe = getenv("SPH_B200_MULTI_XCHG");          // "copy" (default) | "nccl" | "peer"
    m->peerStores = xe && strcmp(xe, "peer") == 0;
    m->copyExchange = repeated || m->peerStores || !(xe && strcmp(xe, "nccl") == 0);     // peer stores need the same peer access
    if (ndev > 1 && !m->copyExchange && !load_nccl(why)) { delete m;  return mfail(nullptr, SPH_ERR_CUDA, "sph_multi_create: %s", why.c_str()); }
    int rc = create_common(m, params, capacityPerSlab);
    if (rc == SPH_OK && ndev > 1 && m->copyExchange) {
        // direct NVLink copies between neighbouring slabs (without peer access the copies would stage through the host)
        for (int k = 0; k + 1 < ndev; k++) {
            const int a = devices[k], b = devices[k + 1];
            if (a == b) continue;
            int ab = 0, ba = 0;
            if (cudaDeviceCanAccessPeer(&ab, a, b) == cudaSuccess && ab) { cudaSetDevice(a);  cudaDeviceEnablePeerAccess(b, 0); }
            if (cudaDeviceCanAccessPeer(&ba, b, a) == cudaSuccess && ba) { cudaSetDevice(b);  cudaDeviceEnablePeerAccess(a, 0); }
            cudaGetLastError();                 // "already enabled" is fine
            // copies work without peer access (staged through the host); kernel-side peer stores do not
            if (!(ab && ba)) m->peerStores = false;
        }
    }
    if (rc == SPH_OK && ndev > 1 && !m->copyExchange) {
        std::vector<ncclComm_t> comms(ndev);
        ncclResult_t r = g_nccl.CommInitAll(comms.data(), ndev, devices);
        if (r != ncclSuccess) rc = mfail(m, SPH_ERR_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
        else for (int k = 0; k < ndev; k++) m->ranks[k].comm = comms[k];
    }
    if (rc != SPH_OK) { g_multiCreateError = m->err;  sph_multi_destroy(m);  return rc; }
    *out = m;
    return SPH_OK;
}

// One process per GPU: slab `rank` of `world`, communicator from a unique id every process received (sph_multi_unique_id
// on one of them, distributed by whatever launched the job).
extern "C" int sph_multi_create_rank(const struct SimParams* params, int rank, int world, const unsigned char* id128, int device,
                                     int capacityPerSlab, sph_multi_t** out)
{
    if (!params || !out || world < 1 || rank < 0 || rank >= world || capacityPerSlab < 1 || (world > 1 && !id128))
        return mfail(nullptr, SPH_ERR_ARG, "sph_multi_create_rank: bad argument");
    std::string why;
    if (world > 1 && !load_nccl(why)) return mfail(nullptr, SPH_ERR_CUDA, "sph_multi_create_rank: %s", why.c_str());
    sph_multi* m = new sph_multi;
    m->world = world;
    m->ranks.resize(1);
    m->ranks[0].device = device;  m->ranks[0].rank = rank;
    int rc = create_common(m, params, capacityPerSlab);
    if (rc == SPH_OK && world > 1) {
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        cudaSetDevice(device);
        ncclResult_t r = g_nccl.CommInitRank(&m->ranks[0].comm, world, id, rank);
        if (r != ncclSuccess) rc = mfail(m, SPH_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    }
    if (rc != SPH_OK) { g_multiCreateError = m->err;  sph_multi_destroy(m);  return rc; }
    *out = m;
    return SPH_OK;
}

extern "C" int sph_multi_set_params(sph_multi_t* m, const struct SimParams* p)
{
    if (!m || !p) return SPH_ERR_ARG;
    m->par = *p;
    for (MultiRank& r : m->ranks) {
        SimParams local = *p;
        local.numParticles = (uint)r.capacity;
        if (sph_set_params(r.s, &local) != SPH_OK) return mfail(m, SPH_ERR_PARAMS, "slab %d: %s", r.rank, sph_last_error(r.s));
    }
    return SPH_OK;
}

// The whole system from HOST arrays in original particle order (float4 rows; every process passes the same arrays).
// cuts = world+1 z-layer boundaries, or NULL to balance the particle counts.  Each slab takes the particles of its layers,
// sorts them once and is ready to step.
extern "C" int sph_multi_set_state(sph_multi_t* m, const float* pos, const float* vel, int n, const int* cuts)
{
    if (!m || !pos || !vel || n < 1) return SPH_ERR_ARG;
    m->n = n;
    const SimParams& P = m->par;
    const int gz = (int)P.gridSize.z, W = m->world;
    std::vector<int> zc(n);
    std::vector<long long> hist(gz, 0);
    for (int i = 0; i < n; i++) {
        int z = host_z_cell(P, pos[4 * (size_t)i + 2]);
        z = std::min(std::max(z, 0), gz - 1);
        zc[i] = z;
        hist[z]++;
    }
    if (cuts) m->cuts.assign(cuts, cuts + W + 1);
    else if (!cut_layers(hist, W, 2, m->cuts)) return mfail(m, SPH_ERR_ARG, "sph_multi_set_state: %d z layers are too few for %d slabs", gz, W);
    if (m->cuts.front() != 0 || m->cuts.back() != gz) return mfail(m, SPH_ERR_ARG, "sph_multi_set_state: cuts must run from 0 to gridSize.z");
    for (int r = 0; r < W; r++) if (m->cuts[r + 1] - m->cuts[r] < 1) return mfail(m, SPH_ERR_ARG, "sph_multi_set_state: empty slab %d", r);

    const long long fullest = *std::max_element(hist.begin(), hist.end());
    const bool resize = message_geometry(m, fullest);
    const int capB = m->capB, capL = m->capL;
    (void)capL;

    std::vector<float> rec;
    for (MultiRank& r : m->ranks) {
        sph_system* s = r.s;
        MCU(m, cudaSetDevice(r.device));
        if (int rc = configure_rank(m, r, resize)) return rc;
        long long mine = 0;
        for (int z = r.zLo; z < r.zHi; z++) mine += hist[z];
        // a step's work set: last step's ghosts below + owned + arrivals + new ghosts on both sides (+ own leavers)
        if (mine + 4LL * fullest + 1024 > r.capacity)
            return mfail(m, SPH_ERR_ARG, "sph_multi_set_state: slab %d owns %lld particles; with ghosts and arrivals (layers of up to "
                         "%lld) that exceeds its capacity of %d slots", r.rank, mine, fullest, r.capacity);
        rec.resize((size_t)std::max<long long>(mine, 1) * kRecFloats);
        size_t k = 0;
        for (int i = 0; i < n; i++) {
            if (zc[i] < r.zLo || zc[i] >= r.zHi) continue;
            float* o = rec.data() + k * kRecFloats;
            memcpy(o, pos + 4 * (size_t)i, 16);
            memcpy(o + 4, vel + 4 * (size_t)i, 16);
            const uint32_t id = (uint32_t)i;
            memcpy(o + 8, &id, 4);
            o[9] = o[10] = o[11] = 0.f;
            k++;
        }
        // staged through the list buffer (192 bytes per slot, rebuilt by every step)
        float* stage = reinterpret_cast<float*>(s->nlist);
        MCU(m, cudaMemcpyAsync(stage, rec.data(), (size_t)mine * kRecFloats * sizeof(float), cudaMemcpyHostToDevice, s->stream));
        if (int rc = load_slab(m, r, stage, (int)mine)) return rc;
    }
    m->haveState = true;
    return SPH_OK;
}

extern "C" int sph_multi_step(sph_multi_t* m, int nsteps)
{
    if (!m || nsteps < 0) return SPH_ERR_ARG;
    if (!m->haveState) return mfail(m, SPH_ERR_STATE, "sph_multi_step: call sph_multi_set_state first");
    const int capL = m->capL, capB = m->capB;
    const bool multi = m->world > 1;
    for (int step = 0; step < nsteps; step++) {
        if (m->recutEvery > 0 && m->steps > 0 && m->steps % m->recutEvery == 0)
            if (int rc = sph_multi_recut(m)) return rc;
        // ---- A: edge layers + particle messages; B: everything else, overlapping exchange 1
        for (MultiRank& r : m->ranks) {
            sph_system* s = r.s;
            sph_system::Slab& b = s->slab;
            MCU(m, cudaSetDevice(r.device));
            SphLaunch L = sph_launcher(s);
            const int yx = (int)s->par.gridSize_yx, nz = b.zHi - b.zLo;
            auto mark = [&](int k) { if (m->phaseTiming) cudaEventRecord(r.evPhase[k], s->stream); };
            mark(0);
            if (m->copyExchange && !m->peerStores && multi) {
                // the neighbours pulled this slab's out buffers on THEIR exchange streams: wait for last step's pulls
                if (r.hasLower) { MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank - 1].evX1, 0));  MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank - 1].evX2, 0)); }
                if (r.hasUpper) { MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank + 1].evX1, 0));  MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank + 1].evX2, 0)); }
            }
            MCU(m, cudaMemsetAsync(r.msgDown, 0, SPH_SLAB_RECORD_BYTES, s->stream));
            MCU(m, cudaMemsetAsync(r.msgUp, 0, SPH_SLAB_RECORD_BYTES, s->stream));
            if (multi) {
                // thin slabs: the two edge regions are the whole slab
                const long long boundA = nz < 4 ? (long long)r.capacity : std::min<long long>(r.capacity, 4LL * capB);
                sph_launch_slab_boundary_integrate_pack(L, s->par, s->pos[s->cur], s->vel, s->idx[s->cur], s->counters, (int)boundA,
                                                        b.zLo, b.zHi, b.hasLower, b.hasUpper,
                                                        r.msgDown + kRecFloats, r.msgUp + kRecFloats, capL,
                                                        r.msgDown + (size_t)kRecFloats * (1 + capL), r.msgUp + (size_t)kRecFloats * (1 + capL), capB,
                                                        reinterpret_cast<uint32_t*>(r.msgDown), reinterpret_cast<uint32_t*>(r.msgUp),
                                                        m->peerStores && r.hasLower ? m->ranks[r.rank - 1].inAbove : nullptr,
                                                        m->peerStores && r.hasUpper ? m->ranks[r.rank + 1].inBelow : nullptr);
                MCU(m, cudaEventRecord(r.evA, s->stream));
            }
            mark(1);
            sph_launch_slab_interior_hist(L, s->par, s->pos[s->cur], s->vel, s->idx[s->cur], s->keyU, s->rankU, s->cellCount,
                                          s->counters, r.capacity, b.keyOffset, b.numCellsLocal, (uint32_t)(b.lowLayers * yx),
                                          (uint32_t)((b.lowLayers + nz) * yx), s->keyMax, 1);
            mark(2);
        }
        if (multi && !m->peerStores)
            if (int rc = exchange(m, msg_bytes(m), &MultiRank::evA, &MultiRank::msgDown, &MultiRank::msgUp, &MultiRank::inBelow, &MultiRank::inAbove)) return rc;
        // ---- arrivals + ghosts, sort, density, rho,p rows
        for (MultiRank& r : m->ranks) {
            sph_system* s = r.s;
            sph_system::Slab& b = s->slab;
            MCU(m, cudaSetDevice(r.device));
            SphLaunch L = sph_launcher(s);
            const int yx = (int)s->par.gridSize_yx, nz = b.zHi - b.zLo;
            auto mark = [&](int k) { if (m->phaseTiming) cudaEventRecord(r.evPhase[k], s->stream); };
            if (multi) {
                if (m->peerStores) {            // the neighbours' packing kernels wrote this slab's inboxes: wait for THEM
                    if (r.hasLower) MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank - 1].evA, 0));
                    if (r.hasUpper) MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank + 1].evA, 0));
                } else {
                    MCU(m, cudaEventRecord(r.evX1, r.xs));
                    MCU(m, cudaStreamWaitEvent(s->stream, r.evX1, 0));
                }
                mark(3);
                sph_launch_slab_unpack_hist(L, s->par, r.hasLower ? r.inBelow : nullptr, r.hasUpper ? r.inAbove : nullptr,
                                            r.hasLower ? r.msgDown : nullptr, r.hasUpper ? r.msgUp : nullptr, capL, capB,
                                            s->pos[s->cur], s->vel, s->idx[s->cur], r.capacity, s->keyU, s->rankU, s->cellCount,
                                            s->counters, b.keyOffset, b.numCellsLocal, (uint32_t)(b.lowLayers * yx),
                                            (uint32_t)((b.lowLayers + nz) * yx), s->keyMax);
            } else mark(3);
            mark(4);
            enqueue_sort(r);
            mark(5);
            const uint32_t* dev = s->counters + SD_FIRST;
            if (s->timing) cudaEventRecord(s->ev[3], s->stream);
            sph_launch_density(L, s->cfg, b.parLocal, s->pos[s->cur], s->velS, s->keyS, s->cellStart, s->maxCount, s->posP, s->velD,
                               s->wantCounts ? s->counts : nullptr, s->nlist, s->ncount, s->ctaRows, 0, r.capacity, dev);
            if (s->timing) cudaEventRecord(s->ev[4], s->stream);
            mark(6);
            if (multi) {
                // peer stores: the rows go straight into the neighbours' row inboxes
                float4* down = m->peerStores ? (r.hasLower ? m->ranks[r.rank - 1].dpAbove : r.dpDown) : r.dpDown;
                float4* up = m->peerStores ? (r.hasUpper ? m->ranks[r.rank + 1].dpBelow : r.dpUp) : r.dpUp;
                sph_launch_slab_pack_dp(L, s->posP, s->velD, s->counters, down, up, capB);
                MCU(m, cudaEventRecord(r.evDp, s->stream));
            }
            mark(7);
        }
        if (multi && !m->peerStores)
            if (int rc = exchange(m, dp_bytes(m), &MultiRank::evDp, &MultiRank::dpDown, &MultiRank::dpUp, &MultiRank::dpBelow, &MultiRank::dpAbove)) return rc;
        // ---- force: CTAs without ghost neighbours while the rows travel, then the rest
        for (MultiRank& r : m->ranks) {
            sph_system* s = r.s;
            sph_system::Slab& b = s->slab;
            MCU(m, cudaSetDevice(r.device));
            SphLaunch L = sph_launcher(s);
            const uint32_t* dev = s->counters + SD_FIRST;
            auto mark = [&](int k) { if (m->phaseTiming) cudaEventRecord(r.evPhase[k], s->stream); };
            if (s->timing) cudaEventRecord(s->evForce[0], s->stream);
            auto force = [&](int part) {
                sph_launch_force(L, s->cfg, b.parLocal, s->posP, s->velD, s->velS, s->keyS, s->cellStart, s->maxCount, s->nlist, s->ncount,
                                 s->ctaRows, s->vel, 0, r.capacity, 0, -1, dev, part, 2 * (capB / s->cfg.threads + 2));
            };
            if (multi) {
                if (!m->peerStores) MCU(m, cudaEventRecord(r.evX2, r.xs));
                force(1);
                mark(8);
                if (m->peerStores) {
                    if (r.hasLower) MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank - 1].evDp, 0));
                    if (r.hasUpper) MCU(m, cudaStreamWaitEvent(s->stream, m->ranks[r.rank + 1].evDp, 0));
                } else MCU(m, cudaStreamWaitEvent(s->stream, r.evX2, 0));
                mark(9);
                sph_launch_slab_unpack_dp(L, r.hasLower ? r.dpBelow : nullptr, r.hasUpper ? r.dpAbove : nullptr, s->counters, s->posP, s->velD, capB);
                mark(10);
                force(2);
            } else { force(0);  mark(8);  mark(9);  mark(10); }
            if (sph_needs_obstacles(s->par)) sph_launch_obstacles(L, b.parLocal, s->posP, s->velD, s->vel, 0, r.capacity, dev);
            mark(11);
            if (s->timing) cudaEventRecord(s->evForce[1], s->stream);
            s->stepped = true;
            MCU(m, cudaGetLastError());
        }
        m->steps++;
    }
    return SPH_OK;
}

// waits for everything enqueued and reports what the device flagged (message overflow, lost particles, ghost mismatch)
extern "C" int sph_multi_sync(sph_multi_t* m)
{
    if (!m) return SPH_ERR_ARG;
    for (MultiRank& r : m->ranks) if (int rc = sync_rank(m, r)) return rc;
    return SPH_OK;
}

extern "C" int sph_multi_local_slabs(sph_multi_t* m) { return m ? (int)m->ranks.size() : 0; }
extern "C" sph_t* sph_multi_handle(sph_multi_t* m, int local) { return m && local >= 0 && local < (int)m->ranks.size() ? m->ranks[local].s : nullptr; }
extern "C" void* sph_multi_stream(sph_multi_t* m, int local) { sph_t* s = sph_multi_handle(m, local);  return s ? (void*)s->stream : nullptr; }

// {z cuts[world+1]}, {owned particles of every LOCAL slab}, message geometry and traffic; any pointer may be NULL
extern "C" int sph_multi_info(sph_multi_t* m, int* cuts, int* ownedLocal, int* capLB2, unsigned long long* bytesSent)
{
    if (!m) return SPH_ERR_ARG;
    if (ownedLocal) { if (int rc = sph_multi_sync(m)) return rc; }
    if (cuts) for (size_t k = 0; k < m->cuts.size(); k++) cuts[k] = m->cuts[k];
    if (ownedLocal) for (size_t k = 0; k < m->ranks.size(); k++) ownedLocal[k] = m->ranks[k].owned;
    if (capLB2) { capLB2[0] = m->capL;  capLB2[1] = m->capB; }
    if (bytesSent) *bytesSent = m->bytesSent;
    return SPH_OK;
}

// Owned particles of the local slabs into HOST arrays addressed by original particle index (rows of particles owned by
// other processes are left untouched).  pos / vel: float4 rows; dens / pres: one float per particle; any may be NULL.
extern "C" int sph_multi_get_state(sph_multi_t* m, float* pos, float* vel, float* dens, float* pres, int n, int* written)
{
    if (!m) return SPH_ERR_ARG;
    if (!m->haveState) return mfail(m, SPH_ERR_STATE, "sph_multi_get_state: no state");
    int total = 0;
    std::vector<float> rec;
    for (MultiRank& r : m->ranks) {
        if (int rc = sync_rank(m, r)) return rc;
        sph_system* s = r.s;
        const int first = (int)r.hostSt[SD_FIRST], count = (int)(r.hostSt[SD_END] - r.hostSt[SD_FIRST]);
        if (count <= 0) continue;
        float* stage = reinterpret_cast<float*>(s->nlist);
        sph_launch_slab_export(sph_launcher(s), s->pos[s->cur], s->vel, s->idx[s->cur], s->stepped ? s->posP : nullptr,
                               s->stepped ? s->velD : nullptr, first, count, stage);
        rec.resize((size_t)count * kRecFloats);
        MCU(m, cudaMemcpyAsync(rec.data(), stage, rec.size() * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        MCU(m, cudaStreamSynchronize(s->stream));
        for (int k = 0; k < count; k++) {
            const float* o = rec.data() + (size_t)k * kRecFloats;
            uint32_t id;
            memcpy(&id, o + 8, 4);
            if (id >= (uint32_t)n) return mfail(m, SPH_ERR_STATE, "slab %d: record %d carries particle id %u >= %d", r.rank, k, id, n);
            if (pos) memcpy(pos + 4 * (size_t)id, o, 16);
            if (vel) memcpy(vel + 4 * (size_t)id, o + 4, 16);
            if (dens) dens[id] = o[9];
            if (pres) pres[id] = o[10];
        }
        total += count;
    }
    if (written) *written = total;
    return SPH_OK;
}

// The owned particles of one local slab as 48-byte records {pos xyzw, vel xyzw, (original index, rho, p, 0)} to / from HOST
// memory (pinned memory makes the copies asynchronous until the final wait).  put: the records must lie in the slab's layers
// (what a fetch returned, possibly modified); the slab is re-sorted and ready to step.
extern "C" int sph_multi_fetch_owned(sph_multi_t* m, int local, float* hostRecords, int capacityRecords, int* count)
{
    if (!m || local < 0 || local >= (int)m->ranks.size() || !hostRecords || !count) return SPH_ERR_ARG;
    MultiRank& r = m->ranks[local];
    if (int rc = sync_rank(m, r)) return rc;
    sph_system* s = r.s;
    const int first = (int)r.hostSt[SD_FIRST], n = (int)(r.hostSt[SD_END] - r.hostSt[SD_FIRST]);
    *count = n;
    if (n > capacityRecords) return mfail(m, SPH_ERR_ARG, "sph_multi_fetch_owned: %d records, room for %d", n, capacityRecords);
    if (n <= 0) return SPH_OK;
    float* stage = reinterpret_cast<float*>(s->nlist);
    sph_launch_slab_export(sph_launcher(s), s->pos[s->cur], s->vel, s->idx[s->cur], s->stepped ? s->posP : nullptr,
                           s->stepped ? s->velD : nullptr, first, n, stage);
    MCU(m, cudaMemcpyAsync(hostRecords, stage, (size_t)n * kRecFloats * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    MCU(m, cudaStreamSynchronize(s->stream));
    return SPH_OK;
}

extern "C" int sph_multi_put_owned(sph_multi_t* m, int local, const float* hostRecords, int count)
{
    if (!m || local < 0 || local >= (int)m->ranks.size() || !hostRecords || count < 0) return SPH_ERR_ARG;
    if (!m->haveState) return mfail(m, SPH_ERR_STATE, "sph_multi_put_owned: call sph_multi_set_state first (it fixes the cuts)");
    MultiRank& r = m->ranks[local];
    if (count > r.capacity) return mfail(m, SPH_ERR_ARG, "sph_multi_put_owned: %d records exceed the slab capacity %d", count, r.capacity);
    sph_system* s = r.s;
    MCU(m, cudaSetDevice(r.device));
    MCU(m, cudaStreamSynchronize(r.xs));
    float* stage = reinterpret_cast<float*>(s->nlist);
    MCU(m, cudaMemcpyAsync(stage, hostRecords, (size_t)count * kRecFloats * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    return load_slab(m, r, stage, count);
}

// Phase profile of the LAST step of one local slab, milliseconds on its solver stream (enable first; costs a dozen event
// records per step).  out11 = {edge integrate + pack, interior integrate + histogram, wait for the particle exchange,
// unpack + histogram of arrivals, scan + bucket + gather, density, pack rho/p rows, interior force, wait for the rho/p
// exchange, unpack rho/p rows, boundary force}.
extern "C" int sph_multi_phase_ms(sph_multi_t* m, int local, int enable, float* out11)
{
    if (!m || local < 0 || local >= (int)m->ranks.size()) return SPH_ERR_ARG;
    MultiRank& r = m->ranks[local];
    if (out11) {
        if (int rc = sync_rank(m, r)) return rc;
        for (int k = 0; k < 11; k++) {
            float ms = -1.f;
            if (!m->phaseTiming || m->steps == 0 || cudaEventElapsedTime(&ms, r.evPhase[k], r.evPhase[k + 1]) != cudaSuccess) { ms = -1.f;  cudaGetLastError(); }
            out11[k] = ms;
        }
    }
    m->phaseTiming = enable != 0;
    return SPH_OK;
}

// ---- re-cutting the slabs: the cuts balance the particle counts at set_state time; a flow that
// piles the fluid up at one end (a wave running down the tank) unbalances them.  The slabs are synchronised, the layer
// histogram of the CURRENT state gives new cuts, particles move to their new owners, every slab is re-sorted.  Results
// do not depend on where the cuts are, so a run with re-cuts equals a run without, bit for bit.
static int recut_one_process(sph_multi* m)
{
    std::vector<float> pos((size_t)m->n * 4), vel((size_t)m->n * 4);
    int written = 0;
    if (int rc = sph_multi_get_state(m, pos.data(), vel.data(), nullptr, nullptr, m->n, &written)) return rc;
    if (written != m->n) return mfail(m, SPH_ERR_STATE, "sph_multi_recut: %d of %d particles found", written, m->n);
    return sph_multi_set_state(m, pos.data(), vel.data(), m->n, nullptr);
}

static int recut_one_process_per_slab(sph_multi* m)
{
    MultiRank& r = m->ranks[0];
    sph_system* s = r.s;
    const SimParams& P = m->par;
    const int W = m->world, gz = (int)P.gridSize.z, me = r.rank;
    if (int rc = sync_rank(m, r)) return rc;
    const int first = (int)r.hostSt[SD_FIRST], count = (int)(r.hostSt[SD_END] - r.hostSt[SD_FIRST]);
    float* stage = reinterpret_cast<float*>(s->nlist);
    std::vector<float> rec((size_t)std::max(count, 1) * kRecFloats);
    if (count > 0) {
        sph_launch_slab_export(sph_launcher(s), s->pos[s->cur], s->vel, s->idx[s->cur], nullptr, nullptr, first, count, stage);
        MCU(m, cudaMemcpyAsync(rec.data(), stage, (size_t)count * kRecFloats * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
        MCU(m, cudaStreamSynchronize(s->stream));
    }
    // global layer histogram (the only collectives of the driver: two per re-cut, none on the step path)
    std::vector<long long> hist(gz, 0);
    std::vector<int> zc(std::max(count, 1));
    for (int k = 0; k < count; k++) {
        int z = host_z_cell(P, rec[(size_t)k * kRecFloats + 2]);
        zc[k] = z = std::min(std::max(z, 0), gz - 1);
        hist[z]++;
    }
    struct DeviceTemp {                         // freed on every way out, the early error returns included
        void* p = nullptr;
        ~DeviceTemp() { if (p) cudaFree(p); }
    } scratchMem, newMem;
    const size_t scratchWords = (size_t)std::max(gz, W * W) + W;
    MCU(m, cudaMalloc(&scratchMem.p, scratchWords * sizeof(long long)));
    long long* dScratch = static_cast<long long*>(scratchMem.p);
    auto done = [&](int rc) { return rc; };
    MCU(m, cudaMemcpyAsync(dScratch, hist.data(), (size_t)gz * sizeof(long long), cudaMemcpyHostToDevice, r.xs));
    MNCCL(m, g_nccl.AllReduce(dScratch, dScratch, (size_t)gz, ncclInt64, ncclSum, r.comm, r.xs));
    MCU(m, cudaMemcpyAsync(hist.data(), dScratch, (size_t)gz * sizeof(long long), cudaMemcpyDeviceToHost, r.xs));
    MCU(m, cudaStreamSynchronize(r.xs));
    if (!cut_layers(hist, W, 2, m->cuts)) return done(mfail(m, SPH_ERR_STATE, "sph_multi_recut: %d z layers are too few for %d slabs", gz, W));
    const bool resize = message_geometry(m, *std::max_element(hist.begin(), hist.end()));

    // bin the records by new owner; counts to everybody
    std::vector<long long> sendCnt(W, 0), off(W + 1, 0);
    std::vector<int> dest(std::max(count, 1));
    for (int k = 0; k < count; k++) {
        int d = (int)(std::upper_bound(m->cuts.begin(), m->cuts.end(), zc[k]) - m->cuts.begin()) - 1;
        dest[k] = d = std::min(std::max(d, 0), W - 1);
        sendCnt[d]++;
    }
    for (int d = 0; d < W; d++) off[d + 1] = off[d] + sendCnt[d];
    std::vector<float> binned((size_t)std::max(count, 1) * kRecFloats);
    {
        std::vector<long long> at(off.begin(), off.end() - 1);
        for (int k = 0; k < count; k++)
            memcpy(binned.data() + (size_t)at[dest[k]]++ * kRecFloats, rec.data() + (size_t)k * kRecFloats, kRecFloats * sizeof(float));
    }
    std::vector<long long> all((size_t)W * W, 0);
    long long* dSend = dScratch + (size_t)W * W;
    MCU(m, cudaMemcpyAsync(dSend, sendCnt.data(), (size_t)W * sizeof(long long), cudaMemcpyHostToDevice, r.xs));
    MNCCL(m, g_nccl.AllGather(dSend, dScratch, (size_t)W, ncclInt64, r.comm, r.xs));
    MCU(m, cudaMemcpyAsync(all.data(), dScratch, (size_t)W * W * sizeof(long long), cudaMemcpyDeviceToHost, r.xs));
    MCU(m, cudaStreamSynchronize(r.xs));
    long long incoming = 0;
    std::vector<long long> roff(W + 1, 0);
    for (int p = 0; p < W; p++) { const long long c = p == me ? 0 : all[(size_t)p * W + me];  roff[p + 1] = roff[p] + c;  incoming += c; }
    const long long kept = sendCnt[me], total = kept + incoming;
    if (total > r.capacity) return done(mfail(m, SPH_ERR_STATE, "sph_multi_recut: slab %d would own %lld particles, capacity %d", me, total, r.capacity));

    // particles travel device to device: everything leaves from the staging area, the new owned set is assembled in a
    // temporary buffer (kept block first, then one block per sender)
    MCU(m, cudaMalloc(&newMem.p, (size_t)std::max<long long>(total, 1) * kRecFloats * sizeof(float)));
    float* dNew = static_cast<float*>(newMem.p);
    auto done2 = [&](int rc) { return done(rc); };
    if (count > 0) MCU(m, cudaMemcpyAsync(stage, binned.data(), (size_t)count * kRecFloats * sizeof(float), cudaMemcpyHostToDevice, r.xs));
    if (kept > 0) MCU(m, cudaMemcpyAsync(dNew, stage + (size_t)off[me] * kRecFloats, (size_t)kept * kRecFloats * sizeof(float), cudaMemcpyDeviceToDevice, r.xs));
    MNCCL(m, g_nccl.GroupStart());
    for (int p = 0; p < W; p++) {
        if (p == me) continue;
        if (sendCnt[p] > 0) MNCCL(m, g_nccl.Send(stage + (size_t)off[p] * kRecFloats, (size_t)sendCnt[p] * kRecFloats * sizeof(float), ncclChar, p, r.comm, r.xs));
        const long long c = all[(size_t)p * W + me];
        if (c > 0) MNCCL(m, g_nccl.Recv(dNew + (size_t)(kept + roff[p]) * kRecFloats, (size_t)c * kRecFloats * sizeof(float), ncclChar, p, r.comm, r.xs));
    }
    MNCCL(m, g_nccl.GroupEnd());
    MCU(m, cudaStreamSynchronize(r.xs));
    if (int rc = configure_rank(m, r, resize)) return done2(rc);
    return done2(load_slab(m, r, dNew, (int)total));
}

extern "C" int sph_multi_recut(sph_multi_t* m)
{
    if (!m) return SPH_ERR_ARG;
    if (!m->haveState) return mfail(m, SPH_ERR_STATE, "sph_multi_recut: no state");
    int rc = SPH_OK;
    if (m->world > 1) rc = (int)m->ranks.size() == m->world ? recut_one_process(m) : recut_one_process_per_slab(m);
    if (rc == SPH_OK) m->recuts++;
    return rc;
}

/* every `steps` steps sph_multi_step re-cuts the slabs before stepping on (0: never) */
extern "C" int sph_multi_set_recut_interval(sph_multi_t* m, int steps)
{
    if (!m || steps < 0) return SPH_ERR_ARG;
    m->recutEvery = steps;
    return SPH_OK;
}

extern "C" int sph_multi_recut_count(sph_multi_t* m) { return m ? m->recuts : 0; }

// fetch_owned and put_owned as ONE blocking call: the download of the slab's current owned records overlaps the upload of the
// new ones (the link is full duplex; a fetch followed by a put would serialise them).  The new records must lie in the slab's
// layers, as for sph_multi_put_owned.
extern "C" int sph_multi_exchange_owned(sph_multi_t* m, int local, float* outRecords, int outCapacity, int* outCount,
                                        const float* inRecords, int inCount)
{
    if (!m || local < 0 || local >= (int)m->ranks.size() || !outRecords || !outCount || !inRecords || inCount < 0) return SPH_ERR_ARG;
    if (!m->haveState) return mfail(m, SPH_ERR_STATE, "sph_multi_exchange_owned: call sph_multi_set_state first (it fixes the cuts)");
    MultiRank& r = m->ranks[local];
    if (inCount > r.capacity) return mfail(m, SPH_ERR_ARG, "sph_multi_exchange_owned: %d records exceed the slab capacity %d", inCount, r.capacity);
    if (int rc = sync_rank(m, r)) return rc;
    sph_system* s = r.s;
    const int first = (int)r.hostSt[SD_FIRST], n = (int)(r.hostSt[SD_END] - r.hostSt[SD_FIRST]);
    *outCount = n;
    if (n > outCapacity) return mfail(m, SPH_ERR_ARG, "sph_multi_exchange_owned: %d records, room for %d", n, outCapacity);
    if (!r.xstage) {
        MCU(m, cudaMalloc((void**)&r.xstage, (size_t)r.capacity * kRecFloats * sizeof(float)));
        MCU(m, cudaEventCreateWithFlags(&r.evXin, cudaEventDisableTiming));
    }
    // in: host -> staging on the exchange stream, at once
    MCU(m, cudaMemcpyAsync(r.xstage, inRecords, (size_t)inCount * kRecFloats * sizeof(float), cudaMemcpyHostToDevice, r.xs));
    MCU(m, cudaEventRecord(r.evXin, r.xs));
    // out: current owned records -> staging -> host on the solver stream
    if (n > 0) {
        float* stage = reinterpret_cast<float*>(s->nlist);
        sph_launch_slab_export(sph_launcher(s), s->pos[s->cur], s->vel, s->idx[s->cur], s->stepped ? s->posP : nullptr,
                               s->stepped ? s->velD : nullptr, first, n, stage);
        MCU(m, cudaMemcpyAsync(outRecords, stage, (size_t)n * kRecFloats * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    }
    MCU(m, cudaStreamWaitEvent(s->stream, r.evXin, 0));
    return load_slab(m, r, r.xstage, inCount);          // ends with a stream synchronise: both directions are done
}

// The cut planner on its own (no GPU needed): layer boundaries that give every slab about the same particle count, at least
// two layers each, from a z-layer histogram.  What sph_multi_set_state and sph_multi_recut use.
extern "C" int sph_multi_plan_cuts(const long long* layerHistogram, int gridZ, int world, int* cuts)
{
    if (!layerHistogram || !cuts || gridZ < 1 || world < 1) return SPH_ERR_ARG;
    std::vector<long long> hist(layerHistogram, layerHistogram + gridZ);
    std::vector<int> c;
    if (!cut_layers(hist, world, 2, c)) return mfail(nullptr, SPH_ERR_ARG, "sph_multi_plan_cuts: %d z layers are too few for %d slabs", gridZ, world);
    for (int k = 0; k <= world; k++) cuts[k] = c[k];
    return SPH_OK;
}
