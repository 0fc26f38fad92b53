/* TEST INFRASTRUCTURE ONLY -- the C interface shared by the two CPU oracles:
 *   oracle/_ref/libsphref.so   the reference's own kernel text, host-compiled (build_ref.sh)
 *   oracle/libsphport.so       this repo's plain C++ restatement (sph_port.cpp)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load either of them.  The product library never does.
 *
 * All arrays are host arrays.  pos/vel are float4 AoS (xyzw), hash pairs are
 * (cellHash, particleIndex) uint32 pairs as in the reference's uint2 particleHash.
 */
#ifndef SPH_ORACLE_API_H
#define SPH_ORACLE_API_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* "reference" or "port" */
const char* orc_kind(void);
/* sizeof(SimParams) the oracle was built with (560) */
int  orc_sizeof_params(void);
void orc_set_threads(int n);            /* OpenMP threads for the block loops; <=0: all */
int  orc_get_threads(void);

/* --- stage-level entry points (one per reference kernel) ------------------- */
void orc_set_params(const void* simParams);
void orc_integrate(const float* oldPos, const float* oldVel, float* newPos, float* newVel, int n);
void orc_calc_hash(const float* pos, uint32_t* pairs, int n);
void orc_sort_pairs(uint32_t* pairs, int n);     /* stable by .x  (== the reference RadixSort) */
void orc_reorder(const uint32_t* pairs, uint32_t* cellStart, const float* oldPos, const float* oldVel,
                 float* sortedPos, float* sortedVel, int n, int numCells);
void orc_density(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                 float* pressure, float* density, int n, int numCells);
/* neighbour count as SURVEY.md Q5 defines it: visited j != i with r2 < h2 in the density walk */
void orc_neighbor_counts(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                         uint32_t* counts, int n, int numCells);
/* newVel/clr/dye are indexed by ORIGINAL particle index (pairs[i].y), as in the reference */
void orc_force(const float* sortedPos, const float* sortedVel, const float* pressure, const float* density,
               const uint32_t* pairs, const uint32_t* cellStart,
               float* newVel, float* clr, float* dyeColor, int n, int numCells);

/* --- whole-system object following cSPH::Update (SPH_Update.cpp:12-81) ----- */
typedef struct orc_system orc_system;
orc_system* orc_create(const void* simParams);
void orc_destroy(orc_system*);
void orc_sys_set_params(orc_system*, const void* simParams);
void orc_sys_set_array(orc_system*, int which /*0 pos, 1 vel*/, const float* xyzw, int start, int count);
void orc_sys_get_array(orc_system*, int which, float* xyzw, int start, int count);
void orc_sys_step(orc_system*, int nsteps);
/* scratch of the LAST step; what: 0 pairs(2n u32) 1 cellStart(numCells u32) 2 sortedPos 3 sortedVel
 * 4 pressure 5 density 6 neighbour counts(n u32, computed on demand) 7 colour(4n) 8 dye(n) */
void orc_sys_dump(orc_system*, int what, void* out);

#ifdef __cplusplus
}
#endif
#endif
