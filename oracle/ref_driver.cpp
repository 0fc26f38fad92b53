// TEST INFRASTRUCTURE ONLY -- builds into oracle/_ref/libsphref.so (see build_ref.sh).
//
// Host driver around the REFERENCE's own kernel text.  build_ref.sh extracts
// source/CUDA/System.cu:11-548 (the block the file itself labels "KERNEL.CU") into
// oracle/_ref/gen/ref_kernels.inc at build time; that block textually includes
// source/CUDA/Kernel_Cell.cui straight from /root/reference.  Nothing of the reference is
// stored in this repository.  This file only emulates <<<grid,block>>> launches with
// sequential ascending thread order (which satisfies reorderD's single __syncthreads,
// Kernel_Cell.cui:50-61) and supplies a bounds-checked FETCH that returns 0 out of range,
// as tex1Dfetch does on a linear texture (needed for the unclamped neighbour cells,
// Kernel_Cell.cui:9-11,146-147).
#include "shim/ref_shim.h"
#include "oracle_api.h"
#include <omp.h>
#include <cstdio>

// --- bounds-checked FETCH ---------------------------------------------------
// Params.cuh defines FETCH(t,i) as t[i] in emulation mode; include it first (it has
// #pragma once) and then replace the macro before the kernel text is seen.
#include "Params.cuh"
#undef FETCH

struct RefBound { const void* p; long long n; };
static RefBound g_bounds[8];
static int g_nbounds = 0;
static void ref_bounds_clear() { g_nbounds = 0; }
static void ref_bound(const void* p, long long n) { g_bounds[g_nbounds].p = p; g_bounds[g_nbounds].n = n; g_nbounds++; }

template <class T>
static inline T ref_fetch(const T* base, unsigned int idx)
{
    long long i = (int)idx;              // tex1Dfetch takes an int coordinate
    for (int k = 0; k < g_nbounds; k++)
        if (g_bounds[k].p == (const void*)base) {
            if (i < 0 || i >= g_bounds[k].n) { T z; memset(&z, 0, sizeof(T)); return z; }
            return base[i];
        }
    return base[i];
}
#define FETCH(t, i) ref_fetch(t, (unsigned int)(i))

// --- the reference kernels ---------------------------------------------------
#include "ref_kernels.inc"

static int g_threads = 0;

// Emulated launch: blocks are independent (they write disjoint outputs), threads of a
// block run in ascending order on one host thread.
#define REF_LAUNCH(nblocks, nthreads, CALL)                                        \
    {                                                                              \
        int _nb = (nblocks), _nt = (nthreads);                                     \
        _Pragma("omp parallel for schedule(static) num_threads(orc_get_threads())") \
        for (int _b = 0; _b < _nb; _b++) {                                         \
            blockDim.x = _nt;  blockIdx.x = _b;                                    \
            for (int _t = 0; _t < _nt; _t++) { threadIdx.x = _t;  CALL; }          \
        }                                                                          \
    }

// computeGridSize (System.cu:641-647)
static void ref_grid(int n, int blockSize, int& nb, int& nt)
{
    nt = blockSize < n ? blockSize : n;
    nb = (n % nt != 0) ? n / nt + 1 : n / nt;
}

extern "C" const char* orc_kind(void) { return "reference"; }
extern "C" int orc_sizeof_params(void) { return (int)sizeof(SimParams); }
extern "C" void orc_set_threads(int n) { g_threads = n; }
extern "C" int orc_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }

extern "C" void orc_set_params(const void* p) { memcpy(&par, p, sizeof(SimParams)); }

// NB the reference kernels do not bounds-check the thread index: n must be a
// multiple of the block size, as it is for every ParticlesK scene.  The driver asserts that.
static void ref_check_n(int n, int bs)
{
    if (n >= bs && n % bs != 0) { fprintf(stderr, "oracle(ref): n=%d not a multiple of %d\n", n, bs); abort(); }
}

extern "C" void orc_integrate(const float* oldPos, const float* oldVel, float* newPos, float* newVel, int n)
{
    int nb, nt;  ref_grid(n, 256, nb, nt);  ref_check_n(n, 256);          // System.cu:653-661
    REF_LAUNCH(nb, nt, integrateD((float4*)newPos, (float4*)newVel, (float4*)oldPos, (float4*)oldVel));
}

extern "C" void orc_calc_hash(const float* pos, uint32_t* pairs, int n)
{
    int nb, nt;  ref_grid(n, 512, nb, nt);  ref_check_n(n, 512);          // System.cu:671-679
    REF_LAUNCH(nb, nt, calcHashD((float4*)pos, (uint2*)pairs));
}

extern "C" void orc_sort_pairs(uint32_t* pairs, int n)
{
    // RadixSort (radixsort_kernel.cu:445-472) is an LSD radix sort whose every pass is stable,
    // so its result is exactly a stable sort by key.
    uint2* p = (uint2*)pairs;
    std::stable_sort(p, p + n, [](const uint2& a, const uint2& b) { return a.x < b.x; });
}

extern "C" void orc_reorder(const uint32_t* pairs, uint32_t* cellStart, const float* oldPos, const float* oldVel,
                            float* sortedPos, float* sortedVel, int n, int numCells)
{
    int nb, nt;  ref_grid(n, 256, nb, nt);  ref_check_n(n, 256);          // System.cu:689-704
    memset(cellStart, 0xff, sizeof(uint32_t) * (size_t)numCells);
    ref_bounds_clear();  ref_bound(oldPos, n);  ref_bound(oldVel, n);
    REF_LAUNCH(nb, nt, reorderD((uint2*)pairs, cellStart, (float4*)oldPos, (float4*)oldVel,
                                (float4*)sortedPos, (float4*)sortedVel));
}

extern "C" void orc_density(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                            float* pressure, float* density, int n, int numCells)
{
    int nb, nt;  ref_grid(n, 64, nb, nt);  ref_check_n(n, 64);            // System.cu:736-740
    ref_bounds_clear();  ref_bound(sortedPos, n);  ref_bound(pairs, n);  ref_bound(cellStart, numCells);
    REF_LAUNCH(nb, nt, computeDensityD(0, (float4*)sortedPos, pressure, density, (uint2*)pairs, (uint*)cellStart));
}

// The reference never stores a neighbour count.  This walk follows compDensCell
// (Kernel_Cell.cui:142-173) with the reference's own calcGridPos/calcGridHash/FETCH and counts
// the candidates that pass its "r2 < par.h2" test.
static void ref_count_one(uint index, float4* oldPos, uint2* particleHash, uint* cellStart, uint32_t* counts)
{
    float4 pos = FETCH(oldPos, index);
    int3 gridPos = calcGridPos(pos);
    uint32_t cnt = 0;
    for (int z = -1; z <= 1; z++)
    for (int y = -1; y <= 1; y++)
    for (int x = -1; x <= 1; x++) {
        uint gridHash = calcGridHash(gridPos + make_int3(x, y, z));
        uint bucketStart = FETCH(cellStart, gridHash);
        if (bucketStart == 0xffffffff) continue;
        for (uint i = 0; i < par.maxParInCell; i++) {
            uint index2 = bucketStart + i;
            uint2 cellData = FETCH(particleHash, index2);
            if (cellData.x != gridHash) break;
            if (index2 != index) {
                float4 pos2 = FETCH(oldPos, index2);
                float4 p = pos - pos2;
                float r2 = p.x * p.x + p.y * p.y + p.z * p.z;
                if (r2 < par.h2) cnt++;
            }
        }
    }
    counts[index] = cnt;
}

extern "C" void orc_neighbor_counts(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                                    uint32_t* counts, int n, int numCells)
{
    ref_bounds_clear();  ref_bound(sortedPos, n);  ref_bound(pairs, n);  ref_bound(cellStart, numCells);
    #pragma omp parallel for schedule(static) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++)
        ref_count_one((uint)i, (float4*)sortedPos, (uint2*)pairs, (uint*)cellStart, counts);
}

extern "C" void orc_force(const float* sortedPos, const float* sortedVel, const float* pressure, const float* density,
                          const uint32_t* pairs, const uint32_t* cellStart,
                          float* newVel, float* clr, float* dyeColor, int n, int numCells)
{
    int nb, nt;  ref_grid(n, 64, nb, nt);  ref_check_n(n, 64);            // System.cu:744
    ref_bounds_clear();
    ref_bound(sortedPos, n);  ref_bound(sortedVel, n);  ref_bound(pressure, n);  ref_bound(density, n);
    ref_bound(pairs, n);  ref_bound(cellStart, numCells);  ref_bound(dyeColor, n);
    REF_LAUNCH(nb, nt, computeForceD(0, (float4*)newVel, (float4*)sortedPos, (float4*)sortedVel, (float4*)clr,
                                     (float*)pressure, (float*)density, dyeColor, (uint2*)pairs, (uint*)cellStart));
}

static void orc_sys_read_dims(orc_system* s);
#include "oracle_system.inc"
static void orc_sys_read_dims(orc_system* s)
{
    const SimParams* p = (const SimParams*)s->par;
    s->n = (int)p->numParticles;  s->numCells = (int)p->numCells;
}
