/* sph_b200.h -- C ABI of the B200-native SPH solver step (libsph_b200.so).
 *
 * This is the drop-in boundary.  It replaces the reference's host->device interface
 *   source/CUDA/System.cuh:1-37   (threadSync, allocateArray, copyTo/FromDevice, setParameters,
 *                                  integrate, calcHash, reorder, collide)
 *   source/CUDA/radixsort.cuh:30-32 (RadixSort)
 * which `class cSPH` (source/SPH/SPH.h:9-50) drives from cSPH::Update (source/SPH/SPH_Update.cpp:12-81).
 * The reference interface is void-returning, exits the process on a CUDA error and passes positions
 * as OpenGL buffer ids; this one uses an opaque handle, int status codes, plain pointers and sizes,
 * and owns its device buffers.  The cSPH-shaped C++ class in sph_host.h sits on top of it, and
 * INTEGRATION.md shows the binding the reference's SPH layer would add.
 *
 * Threading: one handle = one device + one CUDA stream; calls on a handle must come from one thread
 * at a time (the reference is driven from the single GLUT thread).  Calls are asynchronous on the
 * handle's stream unless stated otherwise; every function that hands data to the host synchronises.
 *
 * There is no CPU fallback: every entry point fails with SPH_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>
#include <stdint.h>
#include "sph_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sph_system sph_t;

enum sph_status {
    SPH_OK = 0,
    SPH_ERR_ARG = 1,        /* bad argument (null handle, range outside [0,numParticles), ...)   */
    SPH_ERR_CUDA = 2,       /* a CUDA call failed; sph_last_error() has the text                 */
    SPH_ERR_PARAMS = 3,     /* SimParams rejected (grid < 4 cells in a dimension, numCells mismatch) */
    SPH_ERR_STATE = 4       /* call not valid in the current state                               */
};

/* which particle array */
enum sph_array {
    SPH_POS = 0,            /* float4 xyzw, ORIGINAL particle order (cSPH::getArray(false))      */
    SPH_VEL = 1,            /* float4 xyzw, ORIGINAL particle order (cSPH::getArray(true))       */
    SPH_DENSITY = 2,        /* float, original order                                             */
    SPH_PRESSURE = 3,       /* float, original order                                             */
    SPH_COLOR = 4,          /* float4, original order (the reference's colorVbo contents)        */
    SPH_DYE = 5             /* float, original order (dDyeColor)                                 */
};

/* scratch of the last step, for parity tests (sph_debug_dump) */
enum sph_dump {
    SPH_DUMP_SORTED_PAIRS = 0,  /* uint32[2n]: (cellHash, originalIndex) in sorted order == reference dParHash[0] after RadixSort */
    SPH_DUMP_CELL_START   = 1,  /* uint32[numCells]: first sorted index of each cell, 0xffffffff if empty == reference dCellStart */
    SPH_DUMP_SORTED_POS   = 2,  /* float4[n]  == dSortedPos */
    SPH_DUMP_SORTED_VEL   = 3,  /* float4[n]  == dSortedVel */
    SPH_DUMP_PRESSURE     = 4,  /* float[n], sorted order == dPressure */
    SPH_DUMP_DENSITY      = 5,  /* float[n], sorted order == dDensity  */
    SPH_DUMP_NEIGHBOR_COUNTS = 6, /* uint32[n], sorted order: visited j!=i with r2<h2 in the density walk */
    SPH_DUMP_CELL_END     = 7   /* uint32[numCells]: one past the last sorted index of each cell (start==end if empty) */
};

/* per-stage device time of the last sph_step call, milliseconds (sph_get_timings) */
enum sph_stage {
    SPH_STAGE_INTEGRATE_HASH = 0,   /* boundary + integrate + cell hash + cell histogram  (integrate, calcHash) */
    SPH_STAGE_SORT = 1,             /* cell-table scan + deterministic stable counting sort (RadixSort)         */
    SPH_STAGE_REORDER = 2,          /* gather into sorted SoA float4 buffers               (reorder)            */
    SPH_STAGE_DENSITY = 3,          /* computeDensityD                                     (collide, 1st half)  */
    SPH_STAGE_FORCE = 4,            /* computeForceD                                       (collide, 2nd half)  */
    SPH_STAGE_COUNT = 5
};

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Allocates every buffer for params->numParticles / params->numCells on `device` (replaces
 * cSPH::_InitMem + setParameters, SPH_Mem.cpp:11-54).  Positions/velocities start as zeros. */
int sph_create(const struct SimParams* params, int device, sph_t** out);
int sph_destroy(sph_t* s);                                   /* cSPH::_FreeMem, SPH_Mem.cpp:59-82 */

/* ---- parameters ---------------------------------------------------------------------------- */
/* Replaces setParameters (System.cu:634-637).  Takes effect at the next step.  numParticles,
 * numCells and gridSize must not exceed what sph_create allocated. */
int sph_set_params(sph_t* s, const struct SimParams* params);
int sph_get_params(sph_t* s, struct SimParams* out);
/* Returns the handle to the state sph_create left it in (all particles at zero, slot order == original order)
 * while keeping every allocation: cSPH::InitScene calls it when the next scene fits the buffers, where the
 * reference frees and reallocates everything (SPH_Scenes.cpp:9-13, SPH_Mem.cpp:11-82). */
int sph_reset_state(sph_t* s);

/* Colour (System.cu:406-515) and dye (:519-545) are visual-only outputs of the reference's force kernel.  They
 * are off by default (no cost in the step); when enabled every step also fills SPH_COLOR / SPH_DYE. */
int sph_set_visual(sph_t* s, int enable);
/* dye concentrations (dDyeColor, System.cu:519-545), original particle order, from host memory (checkpoint restore) */
int sph_set_dye(sph_t* s, const float* dye, int start, int count);

/* ---- stepping ------------------------------------------------------------------------------ */
/* nsteps x { integrate -> hash -> sort -> reorder -> density -> force }, the stage order of
 * cSPH::Update (SPH_Update.cpp:39-80).  Asynchronous. */
int sph_step(sph_t* s, int nsteps);
int sph_sync(sph_t* s);                                      /* threadSync, System.cu:613 */

/* ---- particle arrays, original particle order ---------------------------------------------- */
/* cSPH::setArray (SPH_Util.cpp:59-71): overwrite particles [start, start+count) from host memory. */
int sph_set_array(sph_t* s, int which /*SPH_POS|SPH_VEL*/, const float* xyzw, int start, int count);
/* cSPH::getArray (SPH_Util.cpp:44-56) generalised to a range and to the scalar arrays.  Blocking. */
int sph_get_array(sph_t* s, int which, float* out, int start, int count);
/* Whole state out and whole new state in with ONE call: semantics of sph_get_array(SPH_POS), sph_get_array(SPH_VEL) followed
 * by sph_set_array(SPH_POS), sph_set_array(SPH_VEL) over the full range, but the device->host copies of the current state
 * overlap the host->device copies of the new one (PCIe is full duplex).  Blocking; use pinned host memory.  No counterpart
 * in the reference (cSPH::getArray / setArray are one blocking copy each, SPH_Util.cpp:44-71). */
int sph_exchange_arrays(sph_t* s, float* outPos, float* outVel, const float* inPos, const float* inVel);
/* Device-resident variants (no host copy): src/dst are device pointers on the handle's device. */
int sph_set_array_device(sph_t* s, int which, const float* d_xyzw, int start, int count);
int sph_get_array_device(sph_t* s, int which, float* d_out, int start, int count);

/* Replaces getPosBuffer() (SPH.h:29): device pointers of the live buffers in SORTED order plus the
 * original-index array, for a renderer or a halo exchange that does not care about particle order.
 * Valid until the next sph_step / sph_set_array. */
int sph_device_buffers(sph_t* s, const float** d_pos, const float** d_vel, const uint32_t** d_index,
                       const uint32_t** d_cellStart /* numCells+1 entries */);

/* ---- OpenGL interop ------------------------------------------------
 * The reference keeps positions and colours in GL vertex buffers that its renderer draws (posVbo / colorVbo,
 * SPH_Mem.cpp:20-37; ParticleRenderer::setVertexBuffer / setColorBuffer, render_particles.cpp:54-81).  Here the
 * renderer keeps ownership of its buffers: register them once, then sph_gl_update after a step writes the arrays
 * (original particle order, float4 per particle) into them on the device -- map, un-permute, unmap, no host copy.
 * `which` is SPH_POS or SPH_COLOR (colour needs sph_set_visual(s, 1)); glBuffer = 0 unregisters.  The calling thread
 * needs a current GL context on the handle's device; without one the call fails with SPH_ERR_CUDA.
 * NOT exercised by the tests: the build and GPU boxes have no GL. */
int sph_gl_register(sph_t* s, int which, unsigned int glBuffer);
int sph_gl_update(sph_t* s);

/* ---- introspection ------------------------------------------------------------------------- */
int sph_debug_dump(sph_t* s, int what, void* out, size_t outBytes);      /* blocking */
int sph_get_timings(sph_t* s, float* msPerStage /*[SPH_STAGE_COUNT]*/, int enable);
/* device time of each of the last stepped stage kernels is only recorded when enabled */
int sph_kernel_launch_count(sph_t* s, long long* launches);             /* kernels launched so far */
void* sph_cuda_stream(sph_t* s);                                         /* cudaStream_t of the handle */
const char* sph_last_error(sph_t* s);                                    /* s may be NULL: last create error */
const char* sph_version(void);
const char* sph_pair_variant(sph_t* s);                                   /* density/force kernel variant in use */

/* ---- slab decomposition: one handle per GPU owns the z-cell layers [zLo, zHi) ------------------------
 * New in this library (the reference is single-GPU, source/App/App.cpp:142).  The grid is cut into
 * contiguous z ranges; because the reference only ever looks +-1 cell one ghost layer per
 * side reproduces its results exactly, and because the cell hash is z-major every layer is a contiguous
 * run of the sorted arrays.  The handle is created with numParticles = local CAPACITY (owned + ghosts);
 * all pointers below are DEVICE pointers on the handle's device; the caller moves the buffers between
 * ranks (pibiti_b200/slab.py does it with torch.distributed send/recv over NCCL).  One step is
 *   integrate -> pack -> [exchange 1: fixed-size messages] -> unpack -> sort -> density
 *             -> pack_dp -> [exchange 2: sizes known on both sides] -> unpack_dp -> force
 * Only sph_slab_sort synchronises the stream (one read-back per step); everything else is asynchronous.
 *
 * Particle records are SPH_SLAB_RECORD_FLOATS floats: pos xyzw, vel xyzw, (originalIndex as uint32, rho, p, 0).
 * A message is (1 + capL + capB) records: row 0 is a header of uint32 words {nLeavers, nBoundary, 0, ...};
 * rows [1, 1+capL) are particles that left towards the receiver (they become OWNED there); rows
 * [1+capL, 1+capL+capB) are copies of the sender's boundary layer (they become GHOSTS there).  The sender's
 * own leavers are also its ghosts on that side, so sph_slab_unpack takes the outgoing messages back as well.
 * A density/pressure message is n x (x,y,z,pressure) followed by n x (vx,vy,vz,density). */
#define SPH_SLAB_RECORD_FLOATS 12
int sph_slab_configure(sph_t* s, int zLo, int zHi, int hasLower, int hasUpper);
int sph_slab_set_owned(sph_t* s, const float* d_records, int count);            /* replaces all owned particles */
int sph_slab_get_owned(sph_t* s, float* d_records, int capacity, int* count);    /* blocking */
int sph_slab_integrate(sph_t* s);
int sph_slab_pack(sph_t* s, float* d_msgDown, float* d_msgUp, int capL, int capB);
/* sph_slab_integrate + sph_slab_pack in one pass over the state (same results, one kernel instead of four) */
int sph_slab_integrate_pack(sph_t* s, float* d_msgDown, float* d_msgUp, int capL, int capB);
int sph_slab_unpack(sph_t* s, const float* d_inBelow, const float* d_inAbove,
                    const float* d_ownDown, const float* d_ownUp, int capL, int capB);
int sph_slab_sort(sph_t* s, int* counts3 /* ghosts below, owned, ghosts above */);    /* blocking */
int sph_slab_density(sph_t* s);
int sph_slab_pack_dp(sph_t* s, float* d_down, float* d_up, int capRows, int* counts2 /* rows down, rows up */);
int sph_slab_ghost_counts(sph_t* s, int* counts2 /* rows expected from below, from above */);
int sph_slab_unpack_dp(sph_t* s, const float* d_below, int nBelow, const float* d_above, int nAbove);
int sph_slab_force(sph_t* s);
/* The same in two parts, so that the rho,p exchange can overlap the bulk of the force pass: part 1 = particles with no
 * ghost neighbours (needs only sph_slab_density), part 2 = the rest (after sph_slab_unpack_dp); part 0 = both. */
int sph_slab_force_part(sph_t* s, int part);
/* diagnostics after a step: {largest real cell, work-set size, ghosts below, owned, ghosts above, retired slots} */
int sph_slab_stats(sph_t* s, int* out6);

/* ---- multi-GPU driver: the whole system on several GPUs of one box, z-slab decomposed -----------------------------
 * New in this library (the reference is single-GPU: one cSPH, source/SPH/SPH.h:9-50, one device).  A sph_multi_t owns
 * one slab handle per GPU it drives and steps them from one host thread with no host synchronisation inside a step;
 * neighbouring slabs exchange particle messages and (density, pressure) rows with ncclSend / ncclRecv over NVLink on a
 * second stream per slab (no collective on the data path).  Results are bit-identical to the single-GPU step.
 * Two shapes:  sph_multi_create       one process drives `ndev` GPUs (what a C++ App holding a cSPH uses);
 *              sph_multi_create_rank  one process per GPU, `world` processes (bench.py under torchrun); the NCCL unique id
 *                                     comes from sph_multi_unique_id on one process and is handed to the others by the
 *                                     launcher.
 * capacityPerSlab = particle slots of a slab: its owned particles plus ghosts and arrivals of a step.
 * Not supported (rejected): the Z wrap/cycle teleport and the pump boundary, which move particles across slabs. */
typedef struct sph_multi sph_multi_t;
int sph_multi_unique_id(unsigned char* id128);                                   /* 128 bytes (ncclUniqueId) */
int sph_multi_create(const struct SimParams* params, int ndev, const int* devices, int capacityPerSlab, sph_multi_t** out);
int sph_multi_create_rank(const struct SimParams* params, int rank, int world, const unsigned char* id128, int device,
                          int capacityPerSlab, sph_multi_t** out);
int sph_multi_destroy(sph_multi_t* m);
const char* sph_multi_last_error(sph_multi_t* m);
int sph_multi_set_params(sph_multi_t* m, const struct SimParams* params);       /* setParameters, every slab */
/* The whole system from HOST float4 arrays in ORIGINAL particle order (cSPH::setArray for both arrays at once; every
 * process passes the same arrays).  cuts: world+1 z-layer boundaries, or NULL to balance the particle counts. */
int sph_multi_set_state(sph_multi_t* m, const float* pos, const float* vel, int n, const int* cuts);
/* cSPH::Update: nsteps solver steps, asynchronous.  Device-side problems (message overflow, a particle that moved more
 * than one cell layer in a step) surface at the next sph_multi_sync / sph_multi_get_state. */
int sph_multi_step(sph_multi_t* m, int nsteps);
int sph_multi_sync(sph_multi_t* m);
/* cSPH::getArray: owned particles of this process's slabs into HOST arrays indexed by original particle id (float4 rows
 * for pos / vel, one float for dens / pres; NULL to skip).  Rows owned by other processes are left untouched. */
int sph_multi_get_state(sph_multi_t* m, float* pos, float* vel, float* dens, float* pres, int n, int* written);
/* The owned particles of one LOCAL slab as 48-byte records {pos xyzw, vel xyzw, (original index, rho, p, 0)} to / from HOST
 * memory: the per-process accessor of the one-process-per-GPU shape (no scan of the whole system).  put expects records
 * that lie in the slab's layers (what fetch returned, possibly modified). */
int sph_multi_fetch_owned(sph_multi_t* m, int local, float* hostRecords, int capacityRecords, int* count);
int sph_multi_put_owned(sph_multi_t* m, int local, const float* hostRecords, int count);
/* fetch + put as one blocking call whose download overlaps its upload (full-duplex link; use pinned memory) */
int sph_multi_exchange_owned(sph_multi_t* m, int local, float* outRecords, int outCapacity, int* outCount,
                             const float* inRecords, int inCount);
/* phase profile of the last step of a local slab, ms on its solver stream: {edge integrate + pack, interior integrate +
 * histogram, wait for the particle exchange, unpack arrivals, scan + bucket + gather, density, pack rho/p rows, interior
 * force, wait for the rho/p exchange, unpack rho/p rows, boundary force}; enable it first */
int sph_multi_phase_ms(sph_multi_t* m, int local, int enable, float* out11);
/* Re-cut the slabs from the layer histogram of the CURRENT state (the cuts of set_state balance the particle counts at
 * t = 0 only): the slabs synchronise, particles move to their new owners, every slab re-sorts.  Results do not depend on
 * the cuts.  sph_multi_set_recut_interval(m, M) makes sph_multi_step do it every M steps (0 = never). */
int sph_multi_recut(sph_multi_t* m);
/* the planner alone (host only): cuts[world+1] from a z-layer histogram, every slab at least two layers */
int sph_multi_plan_cuts(const long long* layerHistogram, int gridZ, int world, int* cuts);
int sph_multi_set_recut_interval(sph_multi_t* m, int steps);
int sph_multi_recut_count(sph_multi_t* m);
int sph_multi_local_slabs(sph_multi_t* m);
sph_t* sph_multi_handle(sph_multi_t* m, int local);                             /* timings, dumps, launch counts of one slab */
void* sph_multi_stream(sph_multi_t* m, int local);                              /* cudaStream_t the slab's kernels run on */
/* z cuts [world+1], owned particles of the LOCAL slabs, {leaver, boundary} record capacities, bytes sent so far */
int sph_multi_info(sph_multi_t* m, int* cuts, int* ownedLocal, int* capLB2, unsigned long long* bytesSent);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
