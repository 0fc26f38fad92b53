// sph_device.cuh -- declarations shared by the kernel translation units and the C ABI.
//
// Data layout in HBM (n = numParticles, C = numCells), all in "slot order" = the sorted order
// of the last completed step:
//   pos[2]   float4[n]  ping-pong: the live positions are gathered into the other buffer by reorder
//   vel      float4[n]  live velocities (integrate in place; force writes the new ones here)
//   velS     float4[n]  post-integration velocities in sorted order (reference dSortedVel)
//   idx[2]   u32[n]     original particle index of each slot (ping-pong with pos)
//   keyU     u32[n]     cell hash of each slot after integrate (unsorted)
//   rankU    u32[n]     arrival rank inside its cell (from the histogram atomic)
//   pairT    uint2[n]   (source slot, original index) bucketed by cell, arbitrary order inside a cell
//   keyS     u32[n]     cell hash in sorted order
//   posP     float4[n]  (x, y, z, pressure) sorted  -- written by density, read by force
//   velD     float4[n]  (vx, vy, vz, density) sorted -- written by density, read by force
//   nlist    [ceil(n/T)][kMax][T]  neighbour lists, CTA-blocked: u32 sorted indices (L1 variant) or u16 smem slots (TMA variant)
//   ncount   u16[n]     list length per particle (0xFFFF: no list)
//   ctaRows  u32[ceil(n/T)]  rows of each CTA's list block that hold data
//   cellCount u32[C]    histogram, zeroed again by the scan
//   cellStart u32[C+1]  exclusive scan: cell c owns sorted slots [cellStart[c], cellStart[c+1])
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "sph_params.h"

#define SPH_SCAN_TILE 4096          // cells per scan block (256 threads x 4 chunks x 4 cells)
// slab mode: per-block maxima of the live keys land in kKeyMaxSlots words; word [kKeyMaxSlots] is the scan bound
static const int kKeyMaxSlots = 64;

struct SphLaunch {
    cudaStream_t stream;
    long long* launches;            // counter of kernels launched (host side)
};

// ---- sph_stream_kernels.cu (compiled with -fmad=false: bit-exact vs the CPU oracle) ----------
void sph_launch_integrate_hash(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel,
                               uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount /*null: integrate only*/,
                               int first, int count);
// scan of cells [0,numCells); the largest-cell statistic only looks at cells [0,maxCells).  boundCells (device,
// optional): tiles entirely at or above *boundCells, other than the last one, are known to be empty and are skipped
void sph_launch_scan(const SphLaunch& L, uint32_t* cellCount, uint32_t* cellStart, uint32_t* blockSums,
                     uint32_t* maxCount, int numCells, int maxCells, const uint32_t* boundCells = nullptr);
void sph_launch_bucket(const SphLaunch& L, const uint32_t* keyU, const uint32_t* rankU, const uint32_t* idxIn,
                       const uint32_t* cellStart, uint2* pairT, int n, const uint32_t* nDev = nullptr);
void sph_launch_rank_gather(const SphLaunch& L, const uint2* pairT, const uint32_t* keyU, const uint32_t* cellStart,
                            const float4* posIn, const float4* velIn,
                            float4* posOut, float4* velOut, uint32_t* idxOut, uint32_t* keyS, int n,
                            const uint32_t* nDev = nullptr,      // nDev: element count read on the device (n = launch bound)
                            const uint32_t* big = nullptr, int realCells = 0);   // big = the maxCount block of sph_launch_scan:
                                                                 // cells it listed as big are ranked by a CTA each
// layout of the `maxCount` block the scan fills: [0] largest real cell, [kBigCount] cells with more than kBigCell entries,
// [kBigList ..] their indices (kBigCap at most; beyond that the list is ignored and every cell takes the per-entry path)
static const int kBigCell = 64, kBigCount = 1, kBigList = 2, kBigCap = 4094, kMaxCountWords = kBigList + kBigCap;
void sph_launch_iota(const SphLaunch& L, uint32_t* idx, int n);
// original-order accessors: out[idx[j] - start] = src[j]  /  dst[j] = in[idx[j] - start]
void sph_launch_unpermute4(const SphLaunch& L, const float4* src, const uint32_t* idx, float4* out, int start, int count, int n);
void sph_launch_unpermute_w(const SphLaunch& L, const float4* src, const uint32_t* idx, float* out, int start, int count, int n);
void sph_launch_permute4(const SphLaunch& L, float4* dst, const uint32_t* idx, const float4* in, int start, int count, int n);
void sph_launch_cell_table_dump(const SphLaunch& L, const uint32_t* cellStart, uint32_t* outStart, uint32_t* outEnd, int numCells);
void sph_launch_pack_pairs(const SphLaunch& L, const uint32_t* keyS, const uint32_t* idx, uint2* out, int n);

// slab mode (records are 48 bytes: float4 pos, float4 vel, uint4 (originalIndex,0,0,0))
#define SPH_SLAB_RECORD_BYTES 48
#define SPH_DEAD_INDEX 0xFFFFFFFFu
void sph_launch_slab_take_leavers(const SphLaunch& L, const SimParams& par, const float4* pos, const float4* vel, uint32_t* idx,
                                  int first, int count, int zLo, int zHi, int hasLower, int hasUpper,
                                  void* down, int capDown, void* up, int capUp, uint32_t* ctrDown, uint32_t* ctrUp);
void sph_launch_slab_boundary(const SphLaunch& L, const SimParams& par, const float4* pos, const float4* vel, const uint32_t* idx,
                              int n, int zLo, int zHi, int hasLower, int hasUpper,
                              void* down, int capDown, void* up, int capUp, uint32_t* ctrDown, uint32_t* ctrUp);
void sph_launch_slab_append(const SphLaunch& L, const void* recs, int count, float4* pos, float4* vel, uint32_t* idx, int at);
void sph_launch_slab_export(const SphLaunch& L, const float4* pos, const float4* vel, const uint32_t* idx,
                            const float4* posP, const float4* velD, int first, int count, void* recs);
// sph_slab_integrate + sph_slab_pack in one pass over the work set (headers zeroed by the caller)
void sph_launch_slab_integrate_pack(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                    int first, int count, int work, int zLo, int zHi, int hasLower, int hasUpper,
                                    void* leavDown, void* leavUp, int capL, void* bndDown, void* bndUp, int capB,
                                    uint32_t* headDown, uint32_t* headUp);
void sph_launch_fill_u32(const SphLaunch& L, uint32_t* p, uint32_t v, int first, int count);
void sph_launch_slab_hash_hist(const SphLaunch& L, const SimParams& par, const float4* pos, const uint32_t* idx,
                               uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, int nMax, const uint32_t* nDev,
                               long long keyOffset, int numCellsLocal, uint32_t* keyMaxSlots, uint32_t guardCells);
void sph_launch_slab_unpack(const SphLaunch& L, const void* inBelow, const void* inAbove, const void* ownDown, const void* ownUp,
                            int capL, int capB, float4* pos, float4* vel, uint32_t* idx, int work0, int capacity, uint32_t* dev);

// slab mode with device-resident bookkeeping (multi-GPU driver): words of the handle's `counters` array
enum SlabDevWord {
    SD_WORK = 4,        // work-set size (owned + appended arrivals + ghosts + retired slots)
    SD_OVERFLOW = 5,    // a message section or the work set was too small
    SD_LOST = 6,        // an owned particle left the owned layers without being handed over (moved > 1 layer in a step)
    SD_WORK0 = 7,       // work-set size before the arrivals of this step were appended
    SD_FIRST = 8,       // sorted ranges after the sort: owned = [SD_FIRST, SD_END) ...
    SD_END = 9,
    SD_BLO = 10,        // ... first owned layer = [SD_FIRST, SD_BLO), last owned layer = [SD_BHI, SD_END)
    SD_BHI = 11,
    SD_G2 = 12,         // ghosts above = [SD_END, SD_G2); ghosts below = [0, SD_FIRST)
    SD_DPERR = 13,      // rho,p rows received != ghosts expected
    SD_BLO2 = 14,       // the first TWO owned layers = [SD_FIRST, SD_BLO2), the last two = [SD_BHI2, SD_END): every particle
    SD_BHI2 = 15,       // that can leave the slab or land in a boundary layer within one step (it moves at most one layer)
    SD_WORDS = 16
};
void sph_launch_slab_boundary_integrate_pack(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                             uint32_t* st, int bound /* threads: >= particles of the two-layer edge regions */,
                                             int zLo, int zHi, int hasLower, int hasUpper,
                                             void* leavDown, void* leavUp, int capL, void* bndDown, void* bndUp, int capB,
                                             uint32_t* headDown, uint32_t* headUp,
                                             void* peerDown = nullptr, void* peerUp = nullptr /* peer-store exchange: the
                                             neighbours' inboxes (whole messages, header row first), written over NVLink */);
void sph_launch_slab_interior_hist(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                   uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, uint32_t* st, int bound,
                                   long long keyOffset, int numCellsLocal, uint32_t ownedLo, uint32_t ownedHi,
                                   uint32_t* keyMaxSlots, int split);
void sph_launch_slab_unpack_hist(const SphLaunch& L, const SimParams& par, const void* inBelow, const void* inAbove,
                                 const void* ownDown, const void* ownUp, int capL, int capB, float4* pos, float4* vel, uint32_t* idx,
                                 int capacity, uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, uint32_t* st,
                                 long long keyOffset, int numCellsLocal, uint32_t ownedLo, uint32_t ownedHi, uint32_t* keyMaxSlots);
void sph_launch_slab_scan_bound(const SphLaunch& L, uint32_t* keyMaxSlots, uint32_t guardCells, int numCellsLocal);
void sph_launch_slab_bounds(const SphLaunch& L, const uint32_t* cellStart, const uint32_t* scanBound, uint32_t* st,
                            const int cells[7], int hasLower, int hasUpper);
void sph_launch_slab_pack_dp(const SphLaunch& L, const float4* posP, const float4* velD, uint32_t* st, float4* dpDown, float4* dpUp, int capRows);
void sph_launch_slab_unpack_dp(const SphLaunch& L, const float4* dpBelow, const float4* dpAbove, uint32_t* st, float4* posP, float4* velD, int capRows);

// ---- sph_pair_kernels.cu ----------------------------------------------------------------------
// Two variants of the density/force pair, same results:
//   SPH_PAIR_TMA  candidates of a CTA staged in shared memory by TMA bulk copies; neighbour lists hold
//                 shared-memory slot numbers (uint16).  Both kernels must tile and stage identically.
//   SPH_PAIR_L1   candidates read through L1 from the sorted arrays; lists hold global sorted indices (uint32).
//   SPH_PAIR_RM   like L1, but the hits travel as {32-candidate bit mask, first sorted index} records (cap = records per
//                 particle): no per-neighbour stores in the density kernel, no cap on the number of neighbours.
enum SphPairMode { SPH_PAIR_TMA = 0, SPH_PAIR_L1 = 1, SPH_PAIR_RM = 2 };
struct SphPairConfig { int mode; int threads; int cap; int kMax; };   // variant, CTA size, staged-candidate capacity, list length
void sph_pair_default_config(SphPairConfig* cfg);
const char* sph_pair_mode_name(int mode);
size_t sph_pair_blocks(const SphPairConfig& cfg, int n);           // CTAs of the pair kernels
int sph_pair_particles_per_cta(const SphPairConfig& cfg);
size_t sph_pair_list_bytes(const SphPairConfig& cfg, int n);       // size of the neighbour-list buffer
cudaError_t sph_pair_prepare(const SphPairConfig& cfg);
void sph_launch_density(const SphLaunch& L, const SphPairConfig& cfg, const SimParams& par,
                        const float4* posS, const float4* velS, const uint32_t* keyS, const uint32_t* cellStart,
                        const uint32_t* maxCount, float4* posP, float4* velD, uint32_t* neighborCounts,
                        void* nlist, uint16_t* ncount, uint32_t* ctaRows, int first, int count,
                        const uint32_t* dev = nullptr /* slab mode: {first, end, ...} read on the device; count = launch bound */);
void sph_launch_force(const SphLaunch& L, const SphPairConfig& cfg, const SimParams& par,
                      const float4* posP, const float4* velD, const float4* velS, const uint32_t* keyS,
                      const uint32_t* cellStart, const uint32_t* maxCount, const void* nlist, const uint16_t* ncount,
                      const uint32_t* ctaRows, float4* velOut, int first, int count,
                      int ctaFirst = 0, int ctaCount = -1 /* L1 / rm variants: only CTAs [ctaFirst, ctaFirst+ctaCount) of the range */,
                      const uint32_t* dev = nullptr, int part = 0 /* with dev: 1 = CTAs without ghost neighbours, 2 = the others */,
                      int part2Blocks = 0 /* part 2: CTAs to launch (an upper bound of those that touch the two boundary layers) */);

// ---- sph_extras_kernels.cu --------------------------------------------------------------------
bool sph_needs_obstacles(const SimParams& par);        // height map or rotor configured
void sph_launch_obstacles(const SphLaunch& L, const SimParams& par, const float4* posP, const float4* velD, float4* velNew,
                          int first, int count, const uint32_t* dev = nullptr);
void sph_launch_color_dye(const SphLaunch& L, const SimParams& par, const float4* posS, const float4* velS, const float4* velD,
                          const float4* velNew, const uint32_t* keyS, const uint32_t* cellStart, const uint32_t* idx,
                          float4* clr, float* dye, int first, int count);
