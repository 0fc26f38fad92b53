// sph_pair_kernels.cu -- density/pressure and pair-force kernels for sm_100a.
//
// One CTA owns a run of T consecutive SORTED particles.  Because the cell hash is linear and z-major
// (reference Kernel_Cell.cui:15-19), the 3x3x3 neighbourhood of that run is nine contiguous ranges
// of the sorted arrays -- one per (dy,dz) grid row -- so the whole candidate set is staged into
// shared memory by at most nine 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on
// one mbarrier; float4 elements keep every source offset 16-byte aligned.  Overlapping ranges are
// merged first so no candidate is staged twice.
//
// density: each thread walks, for its own particle, the nine 3-cell runs [cellStart[h-1],
//   cellStart[h+2]) out of shared memory, tests every candidate against h, accumulates the Poly6 sum
//   and appends the shared-memory SLOT of every hit to a per-thread neighbour list held in shared
//   memory; the CTA's list block then leaves in one TMA bulk store (shared -> global).
// force:   stages the identical candidate layout (positions+pressure, velocities+density) plus the
//   CTA's list block with the same mbarrier, and evaluates only the listed neighbours (about 25 of
//   the 82 candidates), so there is no divergence on the range test and no global load in the loop.
//
// Reference semantics kept:
//   * search radius is always +-1 cell, cells are addressed by unclamped hash arithmetic, and
//     hashes outside [0,numCells) contribute nothing (Kernel_Cell.cui:146-155);
//   * at most maxParInCell entries of a cell are visited (Kernel_Cell.cui:151,245).  The scan
//     records the largest cell; only if it exceeds maxParInCell do the kernels take the per-cell
//     truncating walk, otherwise three cells are walked as one run;
//   * density: sum (h2-r2)^3 over r2<h2, j!=i; rho = sum*Poly6*mass; p = (rho-rho0)*k
//     (Kernel_Cell.cui:142-199).  The r2<h2 predicate is evaluated without FMA contraction so the
//     neighbour set is bit-identical to the CPU oracle's;
//   * force: compForcePair / compForceCell (Kernel_Cell.cui:210-261), then F*mass*dt, sphere
//     collider, accelerators, newVel = vel + dv (System.cu:247-250,373-402).
#include "sph_device.cuh"
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace {

constexpr int kRows = 9;
constexpr uint32_t kListInvalid = 0xFFFFu;     // ncount value: particle has no usable neighbour list

struct StageTable {
    uint32_t g0[kRows], g1[kRows];          // clipped sorted-index range of each (dy,dz) row
    int      segOf[kRows];                  // row -> merged segment (or -1)
    uint32_t segG0[kRows], segG1[kRows];    // merged segment: global range
    uint32_t segS0[kRows];                  // merged segment: first shared-memory slot
    int      nseg;
    uint32_t total;                         // staged candidates
    int      staged;                        // 0: candidate set larger than the staging buffer
    uint32_t txBytes;                       // bytes the mbarrier waits for (0: nothing in flight)
    uint32_t listRows;                      // force: rows of the CTA's neighbour-list block
    unsigned int ctaMax;                    // density: longest valid list of the CTA
    unsigned long long bar;                 // mbarrier
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dstSmem, const void* srcGlobal, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 1-D TMA bulk copy shared -> global; returns when the shared source has been read
__device__ __forceinline__ void tma_bulk_s2g_and_wait(void* dstGlobal, const void* srcSmem, uint32_t bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dstGlobal), "r"(smem_u32(srcSmem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Threads 0..8 look up the nine row ranges; thread 0 merges them and issues the bulk copies.
// Returns after a __syncthreads with st filled in.  NARR arrays of float4 are staged back to back
// (array a occupies slots [a*cap, a*cap+total)); `listSrc` (force only) is the CTA's neighbour-list
// block, copied to `listDst` with the same barrier.
template <int NARR>
__device__ __forceinline__ void stage_candidates(StageTable& st, float4* sbuf, int cap, const SimParams& par,
                                                 const uint32_t* __restrict__ keyS, const uint32_t* __restrict__ cellStart,
                                                 const float4* __restrict__ arr0, const float4* __restrict__ arr1,
                                                 int p0, int p1,
                                                 const uint16_t* listSrc, uint16_t* listDst, const uint32_t* __restrict__ ctaRows)
{
    const int tid = threadIdx.x;
    if (tid < kRows) {
        const long long kLo = keyS[p0], kHi = keyS[p1 - 1];
        const int dz = tid / 3 - 1, dy = tid % 3 - 1;
        const long long off = (long long)dz * par.gridSize_yx + (long long)dy * par.gridSize.x;
        long long lo = kLo + off - 1, hi = kHi + off + 1;
        if (lo < 0) lo = 0;
        if (hi > (long long)par.numCells - 1) hi = (long long)par.numCells - 1;
        uint32_t a = 0, e = 0;
        if (lo <= hi) { a = __ldg(cellStart + lo); e = __ldg(cellStart + hi + 1); }
        st.g0[tid] = a;  st.g1[tid] = e;
    }
    if (tid == kRows) st.listRows = ctaRows ? __ldg(ctaRows + blockIdx.x) : 0u;
    if (tid == 0) { mbar_init(&st.bar, 1);  st.ctaMax = 0; }
    __syncthreads();
    if (tid == 0) {
        int nseg = 0;  uint32_t total = 0;
        for (int r = 0; r < kRows; r++) {
            uint32_t a = st.g0[r], e = st.g1[r];
            if (e <= a) { st.segOf[r] = -1; continue; }
            if (nseg > 0 && a <= st.segG1[nseg - 1]) {
                if (e > st.segG1[nseg - 1]) st.segG1[nseg - 1] = e;
            } else {
                st.segG0[nseg] = a;  st.segG1[nseg] = e;  nseg++;
            }
            st.segOf[r] = nseg - 1;
        }
        for (int s = 0; s < nseg; s++) { st.segS0[s] = total;  total += st.segG1[s] - st.segG0[s]; }
        st.nseg = nseg;  st.total = total;
        st.staged = (total <= (uint32_t)cap) ? 1 : 0;
        uint32_t tx = 0;
        if (st.staged) {
            const uint32_t listBytes = listSrc ? st.listRows * blockDim.x * 2u : 0u;
            tx = total * 16u * NARR + listBytes;
            if (tx > 0) {
                mbar_expect_tx(&st.bar, tx);
                if (listBytes) tma_bulk_g2s(listDst, listSrc, listBytes, &st.bar);
            }
        }
        st.txBytes = tx;
    }
    __syncthreads();
    // one thread per merged segment issues that segment's bulk copies
    if (tid < st.nseg && st.staged) {
        const uint32_t len = st.segG1[tid] - st.segG0[tid];
        tma_bulk_g2s(sbuf + st.segS0[tid], arr0 + st.segG0[tid], len * 16u, &st.bar);
        if (NARR > 1) tma_bulk_g2s(sbuf + cap + st.segS0[tid], arr1 + st.segG0[tid], len * 16u, &st.bar);
    }
    __syncthreads();
}

// sorted-index bounds [a,e) of the 3-cell run centred on hash hb (clipped to the grid); false if empty
__device__ __forceinline__ bool run_bounds(const uint32_t* __restrict__ cellStart, long long hb, long long C,
                                           uint32_t& a, uint32_t& e)
{
    long long lo = hb - 1, hi = hb + 1;
    if (lo < 0) lo = 0;
    if (hi > C - 1) hi = C - 1;
    if (lo > hi) { a = e = 0;  return false; }
    a = __ldg(cellStart + lo);
    e = __ldg(cellStart + hi + 1);
    return true;
}

__device__ __forceinline__ long long row_hash(const SimParams& par, uint32_t key, int r)
{
    const int dz = r / 3 - 1, dy = r % 3 - 1;
    return (long long)key + (long long)dz * par.gridSize_yx + (long long)dy * par.gridSize.x;
}

// Truncating mode (some cell holds more than maxParInCell particles): the three cells of a row are
// still walked as ONE run unless one of them overflows; only then each cell is cut to its first maxParInCell entries.
// b[0..ncell] = cellStart of the (clipped) cells hb-1..hb+1 and one past; returns ncell (0: nothing to walk).
__device__ __forceinline__ int row_cells(const uint32_t* __restrict__ cellStart, long long hb, long long C, uint32_t maxPar,
                                         uint32_t b[4], bool& overflow)
{
    long long c0 = hb - 1, c1 = hb + 1;
    if (c0 < 0) c0 = 0;
    if (c1 > C - 1) c1 = C - 1;
    if (c0 > c1) return 0;
    const int ncell = (int)(c1 - c0) + 1;
    #pragma unroll
    for (int k = 0; k < 4; k++) b[k] = (k <= ncell) ? __ldg(cellStart + c0 + k) : 0u;
    overflow = false;
    #pragma unroll
    for (int k = 0; k < 3; k++) if (k < ncell && b[k + 1] - b[k] > maxPar) overflow = true;
    return ncell;
}

// the same table entries held in registers (no indexed array, which would live in local memory)
struct RowCells { uint32_t b0, b1, b2, b3;  int ncell; };

__device__ __forceinline__ RowCells load_row_cells(const uint32_t* __restrict__ cellStart, long long hb, long long C)
{
    RowCells rc;
    long long c0 = hb - 1, c1 = hb + 1;
    if (c0 < 0) c0 = 0;
    if (c1 > C - 1) c1 = C - 1;
    rc.ncell = c0 > c1 ? 0 : (int)(c1 - c0) + 1;
    rc.b0 = rc.ncell > 0 ? __ldg(cellStart + c0) : 0u;
    rc.b1 = rc.ncell > 0 ? __ldg(cellStart + c0 + 1) : 0u;
    rc.b2 = rc.ncell > 1 ? __ldg(cellStart + c0 + 2) : 0u;
    rc.b3 = rc.ncell > 2 ? __ldg(cellStart + c0 + 3) : 0u;
    return rc;
}

// ---- density -------------------------------------------------------------------------------------

// r2 exactly as the CPU evaluates "p.x*p.x + p.y*p.y + p.z*p.z" (no contraction)
__device__ __forceinline__ float dist2_exact(float dx, float dy, float dz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// The same value with the x and y lanes evaluated by the packed f32x2 pipe (FADD2 / FMUL2: per-lane IEEE
// round-to-nearest, so bit-identical): 4 issue slots instead of 6 for the differences and squares.  pxy = {p.x, p.y}.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}

__device__ __forceinline__ float dist2_exact_packed(unsigned long long pxy, float pz, const float4& q)
{
    unsigned long long d, m;
    float xx, yy;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pxy), "l"(pack_f32x2(q.x, q.y)));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(m) : "l"(d));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(xx), "=f"(yy) : "l"(m));
    const float dz = __fsub_rn(pz, q.z);
    return __fadd_rn(__fadd_rn(xx, yy), __fmul_rn(dz, dz));
}

// Per-thread neighbour list under construction: column `tid` of a [kMax][T] uint16 array in shared
// memory.  `next` walks down the column; writes stop at `end` but the hit count keeps counting.
struct ListCursor { uint16_t* next; uint16_t* end; uint32_t strideElems; };

// test candidates [a,e) (sorted indices; smem slot = index + shift)
template <bool LIST>
__device__ __forceinline__ void density_span(const float4* __restrict__ cand, int shift, uint32_t a, uint32_t e,
                                             float3 pi, float h2, float& sum, uint32_t& cnt, ListCursor& lc)
{
    #pragma unroll 4
    for (uint32_t g = a; g < e; g++) {
        const int slot = (int)g + shift;
        float4 q = cand[slot];
        float r2 = dist2_exact(pi.x - q.x, pi.y - q.y, pi.z - q.z);
        if (r2 < h2) {
            float c = h2 - r2;
            sum += c * c * c;
            cnt++;
            if (LIST) {
                if (lc.next < lc.end) *lc.next = (uint16_t)slot;
                lc.next += lc.strideElems;
            }
        }
    }
}

template <bool LIST>
__device__ __forceinline__ void density_run(const float4* __restrict__ cand, int shift, uint32_t a, uint32_t e,
                                            uint32_t self, float3 pi, float h2, float& sum, uint32_t& cnt, ListCursor& lc)
{
    if (self - a < e - a) {                 // a <= self < e: skip self
        density_span<LIST>(cand, shift, a, self, pi, h2, sum, cnt, lc);
        density_span<LIST>(cand, shift, self + 1, e, pi, h2, sum, cnt, lc);
    } else {
        density_span<LIST>(cand, shift, a, e, pi, h2, sum, cnt, lc);
    }
}

template <bool STAGED>
__device__ __forceinline__ void density_particle(const StageTable& st, const float4* __restrict__ sbuf,
                                                 const float4* __restrict__ posS, const uint32_t* __restrict__ cellStart,
                                                 const SimParams& par, bool trunc, uint32_t i, uint32_t key, float3 pi,
                                                 float& sum, uint32_t& cnt, ListCursor& lc)
{
    const float h2 = par.h2;
    const long long C = par.numCells;
    if (!trunc) {
        // bounds of the next row are fetched while the current one is walked
        uint32_t a, e, an = 0, en = 0;
        bool ok = run_bounds(cellStart, row_hash(par, key, 0), C, a, e), okn = false;
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            if (r + 1 < kRows) okn = run_bounds(cellStart, row_hash(par, key, r + 1), C, an, en);
            const float4* cand = posS;  int shift = 0;  bool have = ok;
            if (STAGED) {
                int sg = st.segOf[r];
                have = ok && sg >= 0;
                cand = sbuf;
                if (have) shift = (int)st.segS0[sg] - (int)st.segG0[sg];
            }
            if (have) {
                if (r == 4) density_run<STAGED>(cand, shift, a, e, i, pi, h2, sum, cnt, lc);
                else        density_span<STAGED>(cand, shift, a, e, pi, h2, sum, cnt, lc);
            }
            a = an;  e = en;  ok = okn;
        }
    } else {
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            const long long hb = row_hash(par, key, r);
            const float4* cand = posS;  int shift = 0;
            if (STAGED) {
                int sg = st.segOf[r];
                if (sg < 0) continue;
                cand = sbuf;  shift = (int)st.segS0[sg] - (int)st.segG0[sg];
            }
            uint32_t b[4];  bool over;
            const int ncell = row_cells(cellStart, hb, C, par.maxParInCell, b, over);
            if (ncell == 0) continue;
            if (!over) { density_run<STAGED>(cand, shift, b[0], b[ncell], i, pi, h2, sum, cnt, lc);  continue; }
            #pragma unroll
            for (int x = 0; x < 3; x++)
                if (x < ncell) density_run<STAGED>(cand, shift, b[x], min(b[x + 1], b[x] + par.maxParInCell), i, pi, h2, sum, cnt, lc);
        }
    }
}

__global__ void __launch_bounds__(256)
k_density(const __grid_constant__ SimParams par, int cap, int kMax,
          const float4* __restrict__ posS, const float4* __restrict__ velS, const uint32_t* __restrict__ keyS,
          const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
          float4* __restrict__ posP, float4* __restrict__ velD, uint32_t* __restrict__ neighborCounts,
          uint16_t* __restrict__ nlist, uint16_t* __restrict__ ncount, uint32_t* __restrict__ ctaRows, int first, int n)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    float4* sbuf = reinterpret_cast<float4*>(smemRaw);
    uint16_t* slist = reinterpret_cast<uint16_t*>(smemRaw + (size_t)cap * 16);
    __shared__ StageTable st;

    const int T = blockDim.x;
    const int p0 = first + blockIdx.x * T;     // particles [first, n) are processed
    const int p1 = min(n, p0 + T);
    const int i = p0 + threadIdx.x;
    const bool active = i < p1;
    const int il = min(i, p1 - 1);
    const float4 p4 = posS[il];                 // issued before the staging prologue: latency overlaps it
    const float4 v4 = velS[il];
    const uint32_t key = keyS[il];
    const bool trunc = __ldg(maxCount) > par.maxParInCell;

    stage_candidates<1>(st, sbuf, cap, par, keyS, cellStart, posS, nullptr, p0, p1, nullptr, nullptr, nullptr);

    uint32_t validLen = 0;
    if (active) {
        const float3 pi = make_float3(p4.x, p4.y, p4.z);

        ListCursor lc;
        lc.next = slist + threadIdx.x;
        lc.end = slist + (size_t)kMax * T;
        lc.strideElems = (uint32_t)T;

        float sum = 0.f;  uint32_t cnt = 0;
        bool listOk = false;
        if (st.staged) {
            if (st.txBytes > 0) mbar_wait(&st.bar, 0);
            density_particle<true>(st, sbuf, posS, cellStart, par, trunc, (uint32_t)i, key, pi, sum, cnt, lc);
            listOk = cnt <= (uint32_t)kMax;
        } else {
            density_particle<false>(st, sbuf, posS, cellStart, par, trunc, (uint32_t)i, key, pi, sum, cnt, lc);
        }

        const float dens = sum * par.Poly6Kern * par.particleMass;        // Kernel_Cell.cui:194-195
        const float pres = (dens - par.restDensity) * par.stiffness;
        posP[i] = make_float4(p4.x, p4.y, p4.z, pres);
        velD[i] = make_float4(v4.x, v4.y, v4.z, dens);
        ncount[i] = listOk ? (uint16_t)cnt : (uint16_t)kListInvalid;
        if (neighborCounts) neighborCounts[i] = cnt;
        validLen = listOk ? cnt : 0u;
    }

    // the CTA's list block leaves in one bulk store: rows [0, longest valid list)
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, validLen);
    if ((threadIdx.x & 31) == 0 && wmax > 0) atomicMax(&st.ctaMax, wmax);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t rows = st.staged ? min(st.ctaMax, (unsigned int)kMax) : 0u;
        ctaRows[blockIdx.x] = rows;
        if (rows > 0) tma_bulk_s2g_and_wait(nlist + (size_t)blockIdx.x * kMax * T, slist, rows * T * 2u);
    }
}

// ---- force ---------------------------------------------------------------------------------------

struct ForceConsts { float h, minDist, invMinDist, spiky, vterm, minDens; };

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One pair (reference compForcePair, Kernel_Cell.cui:210-228, with d12 from compForceCell :256).
// r, 1/r and 1/(rho_i rho_j) use the SFU approximations (rsqrt / rcp, <= 2 ulp): the parity bar for
// velocities is 1e-5 relative and these stay two orders of magnitude inside it, at a third of the
// instruction count of IEEE sqrt + two divisions.  Coincident particles (r2 = 0) clamp to minDist as
// in the reference: fmaxf drops the NaN of 0*inf, fminf the inf of 1/0.
__device__ __forceinline__ void force_pair(float4 q, float4 u, float3 pi, float3 vi, float presI, float densI,
                                           const ForceConsts& k, float3& f)
{
    float dx = pi.x - q.x, dy = pi.y - q.y, dz = pi.z - q.z;
    float r2 = dx * dx + dy * dy + dz * dz;
    float inv = rsqrtf(r2);
    float r = fmaxf(k.minDist, r2 * inv);
    float invr = fminf(k.invMinDist, inv);
    if (r < k.h) {
        float c = k.h - r;
        float pterm = c * k.spiky * (presI + q.w) * invr;
        float d12 = fminf(k.minDens, rcp_approx(densI * u.w));
        float s = c * d12;
        f.x += (pterm * dx + k.vterm * (u.x - vi.x)) * s;
        f.y += (pterm * dy + k.vterm * (u.y - vi.y)) * s;
        f.z += (pterm * dz + k.vterm * (u.z - vi.z)) * s;
    }
}

__device__ __forceinline__ void force_span(const float4* __restrict__ cpp, const float4* __restrict__ cvd,
                                           uint32_t a, uint32_t e, float3 pi, float3 vi, float presI, float densI,
                                           const ForceConsts& k, float3& f)
{
    #pragma unroll 2
    for (uint32_t g = a; g < e; g++) force_pair(cpp[g], cvd[g], pi, vi, presI, densI, k, f);
}

__device__ __forceinline__ void force_run(const float4* __restrict__ cpp, const float4* __restrict__ cvd, int shift,
                                          uint32_t a, uint32_t e, uint32_t self, float3 pi, float3 vi,
                                          float presI, float densI, const ForceConsts& k, float3& f)
{
    const float4* c0 = cpp + shift;
    const float4* c1 = cvd + shift;
    if (self - a < e - a) {
        force_span(c0, c1, a, self, pi, vi, presI, densI, k, f);
        force_span(c0, c1, self + 1, e, pi, vi, presI, densI, k, f);
    } else {
        force_span(c0, c1, a, e, pi, vi, presI, densI, k, f);
    }
}

// the filtering walk over all candidates (used when a particle has no neighbour list)
template <bool STAGED>
__device__ __forceinline__ float3 force_particle_walk(const StageTable& st, const float4* __restrict__ sbuf, int cap,
                                                      const float4* __restrict__ posP, const float4* __restrict__ velD,
                                                      const uint32_t* __restrict__ cellStart, const SimParams& par, bool trunc,
                                                      uint32_t i, uint32_t key, float4 pp, float4 vd, const ForceConsts& k)
{
    const float3 pi = make_float3(pp.x, pp.y, pp.z), vi = make_float3(vd.x, vd.y, vd.z);
    const long long C = par.numCells;
    float3 f = make_float3(0.f, 0.f, 0.f);
    #pragma unroll 1
    for (int r = 0; r < kRows; r++) {
        const long long hb = row_hash(par, key, r);
        const float4 *c0, *c1;  int shift;
        if (STAGED) {
            int sg = st.segOf[r];
            if (sg < 0) continue;
            c0 = sbuf;  c1 = sbuf + cap;  shift = (int)st.segS0[sg] - (int)st.segG0[sg];
        } else { c0 = posP;  c1 = velD;  shift = 0; }
        if (!trunc) {
            uint32_t a, e;
            if (!run_bounds(cellStart, hb, C, a, e)) continue;
            force_run(c0, c1, shift, a, e, i, pi, vi, pp.w, vd.w, k, f);
        } else {
            uint32_t b[4];  bool over;
            const int ncell = row_cells(cellStart, hb, C, par.maxParInCell, b, over);
            if (ncell == 0) continue;
            if (!over) { force_run(c0, c1, shift, b[0], b[ncell], i, pi, vi, pp.w, vd.w, k, f);  continue; }
            #pragma unroll
            for (int x = 0; x < 3; x++)
                if (x < ncell) force_run(c0, c1, shift, b[x], min(b[x + 1], b[x] + par.maxParInCell), i, pi, vi, pp.w, vd.w, k, f);
        }
    }
    return f;
}

// DEM sphere contact (Kernel_Cell.cui:78-96): relPos/relVel are collider minus particle
__device__ __forceinline__ float3 sphere_contact(const SimParams& par, float3 relPos, float3 relVel, float radiusAB)
{
    float dist = sqrtf(relPos.x * relPos.x + relPos.y * relPos.y + relPos.z * relPos.z);
    float3 force = make_float3(0.f, 0.f, 0.f);
    if (dist < radiusAB) {
        float inv = 1.0f / dist;
        float3 nrm = make_float3(relPos.x * inv, relPos.y * inv, relPos.z * inv);
        float dn = relVel.x * nrm.x + relVel.y * nrm.y + relVel.z * nrm.z;
        float3 tanVel = make_float3(relVel.x - dn * nrm.x, relVel.y - dn * nrm.y, relVel.z - dn * nrm.z);
        float sp = par.spring * (dist - radiusAB);
        force.x = sp * nrm.x + par.damping * relVel.x + par.shear * tanVel.x;
        force.y = sp * nrm.y + par.damping * relVel.y + par.shear * tanVel.y;
        force.z = sp * nrm.z + par.damping * relVel.z + par.shear * tanVel.z;
    }
    return force;
}

// F -> dv = F*mass*dt, then sphere collider and accelerators, newVel = vel + dv   (System.cu:247-250,373-402)
__device__ __forceinline__ float4 finish_velocity(const SimParams& par, float4 pp, float4 vd, float velW, float3 f)
{
    const float md = par.particleMass * par.timeStep;                     // System.cu:250
    float3 dv = make_float3(f.x * md, f.y * md, f.z * md);

    if (par.rotType == 0) {                                               // System.cu:373-375
        float3 c = sphere_contact(par, make_float3(par.collPos.x - pp.x, par.collPos.y - pp.y, par.collPos.z - pp.z),
                                  make_float3(-vd.x, -vd.y, -vd.z), par.particleR + par.collR);
        dv.x += c.x;  dv.y += c.y;  dv.z += c.z;
    }

    #pragma unroll
    for (int a = 0; a < SPH_NUM_ACC; a++) {                               // System.cu:379-399
        const Accel& ac = par.acc[a];
        if (ac.type == ACC_Off) continue;
        float3 rel = make_float3(pp.x - ac.pos.x, pp.y - ac.pos.y, pp.z - ac.pos.z);
        if (ac.type == ACC_Box) {
            if (fabsf(rel.x) < ac.size.x && fabsf(rel.y) < ac.size.y && fabsf(rel.z) < ac.size.z) {
                dv.x += ac.acc.x * par.timeStep;  dv.y += ac.acc.y * par.timeStep;  dv.z += ac.acc.z * par.timeStep;
            }
        } else if (ac.type == ACC_CylY) {
            float ex = rel.x / ac.size.x, ez = rel.z / ac.size.z;
            float rr = sqrtf(ex * ex + ez * ez);
            if (fabsf(rel.y) < ac.size.y && rr < 1.f) {
                dv.x += ac.acc.x * par.timeStep;  dv.y += ac.acc.y * par.timeStep;  dv.z += ac.acc.z * par.timeStep;
            }
        } else if (ac.type == ACC_CylYsm) {
            float rr = sqrtf(rel.x * rel.x + rel.z * rel.z);
            if (fabsf(rel.y) < ac.size.y && rr < ac.size.x) {
                dv.x += ac.acc.x * (1.f - rr / ac.size.z) * par.timeStep;
                dv.y += ac.acc.y * (1.f - rr / ac.size.z) * par.timeStep;
                dv.z += ac.acc.z * (1.f - rr / ac.size.z) * par.timeStep;
            }
        }
    }

    return make_float4(vd.x + dv.x, vd.y + dv.y, vd.z + dv.z, velW + 0.0f);   // System.cu:402
}


__global__ void __launch_bounds__(256)
k_force(const __grid_constant__ SimParams par, int cap, int kMax,
        const float4* __restrict__ posP, const float4* __restrict__ velD, const float4* __restrict__ velS,
        const uint32_t* __restrict__ keyS, const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
        const uint16_t* __restrict__ nlist, const uint16_t* __restrict__ ncount, const uint32_t* __restrict__ ctaRows,
        float4* __restrict__ velOut, int first, int n)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    float4* sbuf = reinterpret_cast<float4*>(smemRaw);
    uint16_t* slist = reinterpret_cast<uint16_t*>(smemRaw + (size_t)cap * 32);
    __shared__ StageTable st;

    const int T = blockDim.x;
    const int p0 = first + blockIdx.x * T;     // particles [first, n) are processed
    const int p1 = min(n, p0 + T);
    const int i = p0 + threadIdx.x;
    const int il = min(i, p1 - 1);              // tail threads load a valid element and discard it
    // own-particle loads are issued first so that their latency overlaps the staging prologue
    const float4 pp = posP[il];
    const float4 vd = velD[il];
    const float velW = velS[il].w;
    const uint32_t key = keyS[il];
    const uint32_t cnt = ncount[il];
    const bool trunc = __ldg(maxCount) > par.maxParInCell;

    stage_candidates<2>(st, sbuf, cap, par, keyS, cellStart, posP, velD, p0, p1,
                        nlist + (size_t)blockIdx.x * kMax * T, slist, ctaRows);
    if (i >= p1) return;

    ForceConsts k;
    k.h = par.h;  k.minDist = par.minDist;  k.invMinDist = 1.0f / par.minDist;  k.spiky = par.SpikyKern;
    k.vterm = par.LapKern * par.viscosity;  k.minDens = par.minDens;

    float3 f = make_float3(0.f, 0.f, 0.f);
    if (st.staged) {
        if (st.txBytes > 0) mbar_wait(&st.bar, 0);
        if (cnt != kListInvalid) {
            const float3 pi = make_float3(pp.x, pp.y, pp.z), vi = make_float3(vd.x, vd.y, vd.z);
            const uint16_t* lst = slist + threadIdx.x;
            const float4* sPP = sbuf;
            const float4* sVD = sbuf + cap;
            #pragma unroll 4
            for (uint32_t t = 0; t < cnt; t++) {
                const uint32_t slot = lst[t * T];
                force_pair(sPP[slot], sVD[slot], pi, vi, pp.w, vd.w, k, f);
            }
        } else {
            f = force_particle_walk<true>(st, sbuf, cap, posP, velD, cellStart, par, trunc, (uint32_t)i, key, pp, vd, k);
        }
    } else {
        f = force_particle_walk<false>(st, sbuf, cap, posP, velD, cellStart, par, trunc, (uint32_t)i, key, pp, vd, k);
    }

    velOut[i] = finish_velocity(par, pp, vd, velW, f);
}

// Slab mode with device-resident ranges: dev = {first owned, end of owned, end of the first owned layer, start of the last
// owned layer}.  part 1 = CTAs whose particles all lie strictly between the two boundary layers (no ghost neighbours),
// part 2 = the others, part 0 = all.
__device__ __forceinline__ bool pair_cta_in_part(int p0, int perCta, int n, const uint32_t* __restrict__ dev, int part)
{
    if (part == 0) return true;
    const int p1 = min(n, p0 + perCta);
    const bool interior = p0 >= (int)__ldg(dev + 2) && p1 <= (int)__ldg(dev + 3);
    return part == 1 ? interior : !interior;
}

// part 2 is launched over a COMPACT grid (the two boundary layers are a small fraction of the slab): block b is the b-th
// CTA that touches the first owned layer, or, past those, the CTAs from the one that reaches into the last owned layer on.
__device__ __forceinline__ int pair_part2_cta(int block, int first, int perCta, const uint32_t* __restrict__ dev)
{
    const int nLo = ((int)__ldg(dev + 2) - first + perCta - 1) / perCta;
    const int hiStart = max(nLo, ((int)__ldg(dev + 3) - first) / perCta);
    return block < nLo ? block : hiStart + (block - nLo);
}

// ---- L1-cached variant ------------------------------------------------------------------------------
// Same walk and same arithmetic, but candidates are read straight from the sorted arrays through L1
// (LDG.128 on the read-only path) instead of being staged by TMA: no shared memory, no CTA prologue,
// register-limited occupancy.  Neighbour lists hold GLOBAL sorted indices (uint32), CTA-blocked
// [cta][k][thread] like the staged variant.  Which variant runs is a launch-time choice
// (SphPairConfig::mode); profiles/ holds the ncu evidence for the default.

// The candidate loop is branch-free: a hit adds c^3 with one predicated FFMA and its sorted index goes to the thread's
// column of the CTA's list block [k][thread] with one predicated 4-byte store (row k of a warp = one 128-byte
// segment).  Left to the compiler the store becomes a branch around five address instructions per candidate.
// (Measured alternative, profiles/: lists built in shared memory and written with one TMA bulk store need fewer
// instructions but 25 KB of shared memory per CTA; the lost occupancy costs more than the instructions save.)
// 8 CTAs of 256 threads = every warp slot of an SM: the walk is latency-bound, so the register budget is 32
#ifndef SPH_DENSITY_MIN_BLOCKS
#define SPH_DENSITY_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(256, SPH_DENSITY_MIN_BLOCKS)
k_density_l1(const __grid_constant__ SimParams par, int kMax,
             const float4* __restrict__ posS, const float4* __restrict__ velS, const uint32_t* __restrict__ keyS,
             const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
             float4* __restrict__ posP, float4* __restrict__ velD, uint32_t* __restrict__ neighborCounts,
             uint32_t* __restrict__ nlist, uint16_t* __restrict__ ncount, int first, int n, const uint32_t* __restrict__ dev)
{
    const int T = blockDim.x;
    if (dev) { first = (int)__ldg(dev);  n = (int)__ldg(dev + 1); }      // slab mode: the owned range lives on the device
    const int i = first + blockIdx.x * T + threadIdx.x;
    if (i >= n) return;
    const float4 p4 = posS[i];
    const uint32_t key = keyS[i];
    const bool trunc = __ldg(maxCount) > par.maxParInCell;
    const float3 pi = make_float3(p4.x, p4.y, p4.z);
    const float h2 = par.h2;
    const long long C = par.numCells;
    // this thread's column of the CTA's list block [k][thread]: byte offset `off` advances one row per hit
    char* const lb = reinterpret_cast<char*>(nlist + (size_t)blockIdx.x * kMax * T);
    const uint32_t rowBytes = (uint32_t)T * 4u;
    const uint32_t endOff = (uint32_t)kMax * rowBytes;
    uint32_t off = threadIdx.x * 4u;
    const unsigned long long pxy = pack_f32x2(pi.x, pi.y);

    float sum = 0.f;
    auto span = [&](uint32_t a, uint32_t e) {
        #pragma unroll 4
        for (uint32_t g = a; g < e; g++) {
            const float4 q = __ldg(posS + g);
            const float r2 = dist2_exact_packed(pxy, pi.z, q);
            const bool hit = r2 < h2;
            const float c = __fsub_rn(h2, r2);
            sum = hit ? fmaf(c * c, c, sum) : sum;              // one predicated FFMA
            // store and advance are predicated, not branched around: the address is formed unconditionally
            asm volatile("{ .reg .pred h, k;  setp.ne.u32 h, %3, 0;  setp.lt.and.u32 k, %0, %4, h;\n"
                         "  @k st.global.u32 [%1], %2;  @h add.u32 %0, %0, %5; }"
                         : "+r"(off) : "l"(lb + off), "r"(g), "r"((uint32_t)hit), "r"(endOff), "r"(rowBytes));
        }
    };
    auto run = [&](uint32_t a, uint32_t e) {
        if ((uint32_t)i - a < e - a) { span(a, (uint32_t)i);  span((uint32_t)i + 1, e); }
        else span(a, e);
    };
    if (!trunc) {
        uint32_t a, e, an = 0, en = 0;
        bool ok = run_bounds(cellStart, row_hash(par, key, 0), C, a, e), okn = false;
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            if (r + 1 < kRows) okn = run_bounds(cellStart, row_hash(par, key, r + 1), C, an, en);
            if (ok) { if (r == 4) run(a, e); else span(a, e); }
            a = an;  e = en;  ok = okn;
        }
    } else {
        // some cell holds more than maxParInCell particles: a row is still one run unless one of ITS cells overflows.
        // The next row's four table entries are in flight while this row is walked.
        const uint32_t mp = par.maxParInCell;
        RowCells cur = load_row_cells(cellStart, row_hash(par, key, 0), C), nxt = cur;
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            if (r + 1 < kRows) nxt = load_row_cells(cellStart, row_hash(par, key, r + 1), C);
            if (cur.ncell > 0) {
                const bool over = cur.b1 - cur.b0 > mp || (cur.ncell > 1 && cur.b2 - cur.b1 > mp) || (cur.ncell > 2 && cur.b3 - cur.b2 > mp);
                if (!over) {
                    const uint32_t e = cur.ncell == 3 ? cur.b3 : cur.ncell == 2 ? cur.b2 : cur.b1;
                    if (r == 4) run(cur.b0, e); else span(cur.b0, e);
                } else {
                    run(cur.b0, min(cur.b1, cur.b0 + mp));
                    if (cur.ncell > 1) run(cur.b1, min(cur.b2, cur.b1 + mp));
                    if (cur.ncell > 2) run(cur.b2, min(cur.b3, cur.b2 + mp));
                }
            }
            cur = nxt;
        }
    }
    const float dens = sum * par.Poly6Kern * par.particleMass;            // Kernel_Cell.cui:194-195
    const float pres = (dens - par.restDensity) * par.stiffness;
    const float4 v4 = velS[i];                  // not needed before: keeping it live across the walk costs registers
    posP[i] = make_float4(p4.x, p4.y, p4.z, pres);
    velD[i] = make_float4(v4.x, v4.y, v4.z, dens);
    const uint32_t cnt = (off - threadIdx.x * 4u) / rowBytes;
    ncount[i] = cnt <= (uint32_t)kMax ? (uint16_t)cnt : (uint16_t)kListInvalid;
    if (neighborCounts) neighborCounts[i] = cnt;
}

// list entries are read once: kStream loads them around L1 (no allocation) so that they do not evict the gathered
// particle records, which is what the kernel is bound by
template <bool kStream>
__device__ __forceinline__ uint32_t load_list_entry(const uint32_t* p)
{
    if (!kStream) return __ldg(p);
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool kStream>
__global__ void __launch_bounds__(256)
k_force_l1(const __grid_constant__ SimParams par, int kMax,
           const float4* __restrict__ posP, const float4* __restrict__ velD, const float4* __restrict__ velS,
           const uint32_t* __restrict__ keyS, const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
           const uint32_t* __restrict__ nlist, const uint16_t* __restrict__ ncount,
           float4* __restrict__ velOut, int first, int n, int ctaFirst, const uint32_t* __restrict__ dev, int part)
{
    __shared__ StageTable st;                  // only read by the (never staged) filtering walk
    const int T = blockDim.x;
    int cta = blockIdx.x + ctaFirst;           // a launch may cover a sub-range of the CTAs the density launch used
    if (dev) {
        first = (int)__ldg(dev);  n = (int)__ldg(dev + 1);
        if (part == 2) cta = pair_part2_cta(blockIdx.x, first, T, dev);
        if (!pair_cta_in_part(first + cta * T, T, n, dev, part)) return;
    }
    const int i = first + cta * T + threadIdx.x;
    if (i >= n) return;
    const float4 pp = posP[i];
    const float4 vd = velD[i];
    const float velW = velS[i].w;
    const uint32_t cnt = ncount[i];

    ForceConsts k;
    k.h = par.h;  k.minDist = par.minDist;  k.invMinDist = 1.0f / par.minDist;  k.spiky = par.SpikyKern;
    k.vterm = par.LapKern * par.viscosity;  k.minDens = par.minDens;

    float3 f = make_float3(0.f, 0.f, 0.f);
    if (cnt != kListInvalid) {
        const float3 pi = make_float3(pp.x, pp.y, pp.z), vi = make_float3(vd.x, vd.y, vd.z);
        // list entries are consumed four at a time; the next four are in flight while these are evaluated
        const uint32_t* lst = nlist + (size_t)cta * kMax * T + threadIdx.x;
        const uint32_t groups = (cnt + 3) >> 2;
        auto load4 = [&](uint32_t q) {
            const uint32_t t0 = 4 * q;
            uint4 e;
            e.x = t0 < cnt ? load_list_entry<kStream>(lst + (size_t)t0 * T) : 0u;
            e.y = t0 + 1 < cnt ? load_list_entry<kStream>(lst + (size_t)(t0 + 1) * T) : 0u;
            e.z = t0 + 2 < cnt ? load_list_entry<kStream>(lst + (size_t)(t0 + 2) * T) : 0u;
            e.w = t0 + 3 < cnt ? load_list_entry<kStream>(lst + (size_t)(t0 + 3) * T) : 0u;
            return e;
        };
        uint4 cur = load4(0);
        for (uint32_t q = 0; q < groups; q++) {
            const uint4 nxt = load4(q + 1);
            const uint32_t left = cnt - 4 * q;              // >= 1
            const uint32_t g0 = cur.x, g1 = left > 1 ? cur.y : cur.x, g2 = left > 2 ? cur.z : cur.x, g3 = left > 3 ? cur.w : cur.x;
            const float4 q0 = __ldg(posP + g0), u0 = __ldg(velD + g0);
            const float4 q1 = __ldg(posP + g1), u1 = __ldg(velD + g1);
            const float4 q2 = __ldg(posP + g2), u2 = __ldg(velD + g2);
            const float4 q3 = __ldg(posP + g3), u3 = __ldg(velD + g3);
            force_pair(q0, u0, pi, vi, pp.w, vd.w, k, f);
            if (left > 1) force_pair(q1, u1, pi, vi, pp.w, vd.w, k, f);
            if (left > 2) force_pair(q2, u2, pi, vi, pp.w, vd.w, k, f);
            if (left > 3) force_pair(q3, u3, pi, vi, pp.w, vd.w, k, f);
            cur = nxt;
        }
    } else {
        const bool trunc = __ldg(maxCount) > par.maxParInCell;
        f = force_particle_walk<false>(st, nullptr, 0, posP, velD, cellStart, par, trunc, (uint32_t)i, keyS[i], pp, vd, k);
    }
    velOut[i] = finish_velocity(par, pp, vd, velW, f);
}


// ---- row-mask variant ("rm") -----------------------------------------------------------------------
// One thread per sorted particle, candidates through L1 like the L1 variant, but the hits are not written as one index
// per neighbour.  The density kernel keeps the hits of up to 32 consecutive candidates as a bit mask in a register and
// emits one 8-byte record {mask, first sorted index} per non-empty word: nine coalesced 8-byte stores per particle
// instead of ~25 scattered 4-byte stores (ncu, profiles/: the index-list stores were 40 % of the density kernel's L1
// wavefronts and 1.1 GB of DRAM writes per step at 8M).  The force kernel walks the records of its particle with a
// cursor (lowest set bit first, so neighbours are visited in the candidate order of the reference) and prefetches the
// next record; only the refill is divergent.  There is no cap on the number of NEIGHBOURS any more, only on the number of
// records (cap; dense cells need one record per 32 candidates of a row), so dense scenes keep their lists.
constexpr uint32_t kRmInvalid = 0xFFFFu;       // nrec value: the record stream overflowed, the force kernel walks

// MINB = resident 128-thread CTAs per SM the register allocation aims at: 12 (40 registers) or 16 (32, a few spills)
template <int MINB, int UNROLL>
__global__ void __launch_bounds__(128, MINB)
k_density_rm(const __grid_constant__ SimParams par, int recCap,
             const float4* __restrict__ posS, const float4* __restrict__ velS, const uint32_t* __restrict__ keyS,
             const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
             float4* __restrict__ posP, float4* __restrict__ velD, uint32_t* __restrict__ neighborCounts,
             uint2* __restrict__ records, uint16_t* __restrict__ nrec, int first, int n, const uint32_t* __restrict__ dev)
{
    if (dev) { first = (int)__ldg(dev);  n = (int)__ldg(dev + 1); }      // slab mode: the owned range lives on the device
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p4 = posS[i];
    const int key = (int)keyS[i];
    const bool trunc = __ldg(maxCount) > par.maxParInCell;
    const float h2 = par.h2;
    const int C = (int)par.numCells;                                      // sph_create keeps numCells below 2^30
    const unsigned long long pxy = pack_f32x2(p4.x, p4.y);
    const float pz = p4.z;

    float sum = 0.f;
    uint32_t w = 0u, cnt = 0u;
    // this thread's column of the CTA's record block [record][thread]; advances one row per record
    uint2* rec = records + (size_t)blockIdx.x * recCap * blockDim.x + threadIdx.x;
    // window [a, e): one record per 32 candidates; bit k of the mask = candidate base + k.  centre = the centre row,
    // where the walk steps over the particle itself (Kernel_Cell.cui:166).  The tail of a candidate is predicated
    // instructions from one PTX block (hit test, c^3 into the sum, bit into the mask): no branch, no select.
    auto window = [&](uint32_t a, uint32_t e, auto centreTag) {
        constexpr bool centre = decltype(centreTag)::value;
        for (uint32_t base = a; base < e; base += 32u) {
            const uint32_t g1 = min(e, base + 32u);
            uint32_t mask = 0u, bit = 1u;
            #pragma unroll (UNROLL)
            for (uint32_t g = base; g < g1; g++, bit += bit) {
                const float4 q = __ldg(posS + g);
                const float r2 = dist2_exact_packed(pxy, pz, q);
                const float c = __fsub_rn(h2, r2);
                const float cc = c * c;
                if (centre)
                    asm("{ .reg .pred h;  setp.lt.f32 h, %2, %3;  setp.ne.and.u32 h, %7, %8, h;\n"
                        "  @h fma.rn.f32 %0, %4, %5, %0;  @h or.b32 %1, %1, %6; }"
                        : "+f"(sum), "+r"(mask) : "f"(r2), "f"(h2), "f"(cc), "f"(c), "r"(bit), "r"(g), "r"((uint32_t)i));
                else
                    asm("{ .reg .pred h;  setp.lt.f32 h, %2, %3;\n"
                        "  @h fma.rn.f32 %0, %4, %5, %0;  @h or.b32 %1, %1, %6; }"
                        : "+f"(sum), "+r"(mask) : "f"(r2), "f"(h2), "f"(cc), "f"(c), "r"(bit));
            }
            if (mask != 0u) {
                if (w < (uint32_t)recCap) *rec = make_uint2(mask, base);
                rec += blockDim.x;
                w++;
                cnt += __popc(mask);
            }
        }
    };
    auto row = [&](uint32_t a, uint32_t e, int r) {
        if (r == 4) window(a, e, std::true_type{});  else window(a, e, std::false_type{});
    };
    // hash of the centre cell of row r: rows advance by one y line, every third row by one z plane (32-bit arithmetic)
    const int gx = (int)par.gridSize.x, gyx = (int)par.gridSize_yx;
    int hb = key - gyx - gx;
    if (!trunc) {
        // sorted-index bounds of the three cells around hb, clipped to the grid; the next row's are in flight
        auto bounds = [&](int h, uint32_t& a, uint32_t& e) -> bool {
            const int lo = max(h - 1, 0), hi = min(h + 1, C - 1);
            if (lo > hi) { a = e = 0u;  return false; }
            a = __ldg(cellStart + lo);  e = __ldg(cellStart + hi + 1);
            return true;
        };
        uint32_t a, e, an = 0, en = 0;
        bool ok = bounds(hb, a, e), okn = false;
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            hb += (r % 3 == 2) ? gyx - 2 * gx : gx;
            if (r + 1 < kRows) okn = bounds(hb, an, en);
            if (ok) row(a, e, r);
            a = an;  e = en;  ok = okn;
        }
    } else {
        // some cell holds more than maxParInCell particles: a row is still one run unless one of ITS cells overflows
        const uint32_t mp = par.maxParInCell;
        #pragma unroll 1
        for (int r = 0; r < kRows; r++) {
            const RowCells cur = load_row_cells(cellStart, (long long)hb, (long long)C);
            hb += (r % 3 == 2) ? gyx - 2 * gx : gx;
            if (cur.ncell == 0) continue;
            const bool over = cur.b1 - cur.b0 > mp || (cur.ncell > 1 && cur.b2 - cur.b1 > mp) || (cur.ncell > 2 && cur.b3 - cur.b2 > mp);
            if (!over) {
                row(cur.b0, cur.ncell == 3 ? cur.b3 : cur.ncell == 2 ? cur.b2 : cur.b1, r);
            } else {
                #pragma unroll 1
                for (int x = 0; x < cur.ncell; x++) {
                    const uint32_t c0 = x == 0 ? cur.b0 : x == 1 ? cur.b1 : cur.b2, c1 = x == 0 ? cur.b1 : x == 1 ? cur.b2 : cur.b3;
                    row(c0, min(c1, c0 + mp), r);
                }
            }
        }
    }
    const float dens = sum * par.Poly6Kern * par.particleMass;            // Kernel_Cell.cui:194-195
    const float pres = (dens - par.restDensity) * par.stiffness;
    const float4 v4 = velS[i];                  // not needed before: keeping it live across the walk costs registers
    posP[i] = make_float4(p4.x, p4.y, p4.z, pres);
    velD[i] = make_float4(v4.x, v4.y, v4.z, dens);
    nrec[i] = w <= (uint32_t)recCap ? (uint16_t)w : (uint16_t)kRmInvalid;
    if (neighborCounts) neighborCounts[i] = cnt;
}

// records are read once: around L1, so that they do not evict the gathered particle rows
__device__ __forceinline__ uint2 load_record(const uint2* p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// Walks the records of one particle: next() returns the sorted index of the next neighbour.  The refill is predicated,
// not branched (lanes run dry at different iterations, so a branch would be taken by some lane almost every time and
// split the warp around the pair arithmetic): an empty mask takes the prefetched record and starts the load of the one
// after it.  Stored records are never empty, so a mask that is still empty after the refill means the stream is done.
struct RmCursor {
    const uint2* pn;            // record after `nxt`
    uint32_t strideBytes, left; // records not yet loaded
    uint32_t mask, base;  uint2 nxt;
    __device__ __forceinline__ void open(const uint2* col, uint32_t stride, uint32_t n)
    {
        strideBytes = stride * (uint32_t)sizeof(uint2);
        const uint2 cur = n > 0u ? load_record(col) : make_uint2(0u, 0u);
        nxt = n > 1u ? load_record(col + stride) : make_uint2(0u, 0u);
        pn = col + 2 * (size_t)stride;
        left = n > 2u ? n - 2u : 0u;
        mask = cur.x;  base = cur.y;
    }
    __device__ __forceinline__ bool next(uint32_t& g)
    {
        const bool empty = mask == 0u;
        mask = empty ? nxt.x : mask;
        base = empty ? nxt.y : base;
        const bool more = empty && left != 0u;
        if (empty) nxt = make_uint2(0u, 0u);
        if (more) {                              // compiles to predicated instructions (one load, three adds)
            nxt = load_record(pn);
            pn = reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(pn) + strideBytes);
            left--;
        }
        const bool ok = mask != 0u;
        const uint32_t b = (uint32_t)__ffs((int)mask) - 1u;
        g = ok ? base + b : base;
        mask &= mask - 1u;
        return ok;
    }
};

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_force_rm(const __grid_constant__ SimParams par, int recCap,
           const float4* __restrict__ posP, const float4* __restrict__ velD, const float4* __restrict__ velS,
           const uint32_t* __restrict__ keyS, const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ maxCount,
           const uint2* __restrict__ records, const uint16_t* __restrict__ nrec,
           float4* __restrict__ velOut, int first, int n, int ctaFirst, const uint32_t* __restrict__ dev, int part)
{
    __shared__ StageTable st;                  // only read by the (never staged) filtering walk
    const int T = blockDim.x;
    int cta = blockIdx.x + ctaFirst;           // a launch may cover a sub-range of the CTAs the density launch used
    if (dev) {
        first = (int)__ldg(dev);  n = (int)__ldg(dev + 1);
        if (part == 2) cta = pair_part2_cta(blockIdx.x, first, T, dev);
        if (!pair_cta_in_part(first + cta * T, T, n, dev, part)) return;
    }
    const int i = first + cta * T + threadIdx.x;
    if (i >= n) return;
    const float4 pp = posP[i];
    const float4 vd = velD[i];
    const float velW = velS[i].w;
    const uint32_t nw = nrec[i];

    ForceConsts k;
    k.h = par.h;  k.minDist = par.minDist;  k.invMinDist = 1.0f / par.minDist;  k.spiky = par.SpikyKern;
    k.vterm = par.LapKern * par.viscosity;  k.minDens = par.minDens;

    float3 f = make_float3(0.f, 0.f, 0.f);
    if (nw != kRmInvalid) {
        const float3 pi = make_float3(pp.x, pp.y, pp.z), vi = make_float3(vd.x, vd.y, vd.z);
        RmCursor cur;
        cur.open(records + (size_t)cta * recCap * T + threadIdx.x, (uint32_t)T, nw);
        // two neighbours per iteration so that two gathers are in flight
        for (;;) {
            uint32_t ga, gb;
            if (!cur.next(ga)) break;
            const bool two = cur.next(gb);
            if (!two) gb = ga;
            const float4 qa = __ldg(posP + ga), ua = __ldg(velD + ga);
            const float4 qb = __ldg(posP + gb), ub = __ldg(velD + gb);
            force_pair(qa, ua, pi, vi, pp.w, vd.w, k, f);
            if (two) force_pair(qb, ub, pi, vi, pp.w, vd.w, k, f);
        }
    } else {
        const bool trunc = __ldg(maxCount) > par.maxParInCell;
        f = force_particle_walk<false>(st, nullptr, 0, posP, velD, cellStart, par, trunc, (uint32_t)i, keyS[i], pp, vd, k);
    }
    velOut[i] = finish_velocity(par, pp, vd, velW, f);
}


inline size_t density_smem(const SphPairConfig& c) { return (size_t)c.cap * 16 + (size_t)c.kMax * c.threads * 2; }
inline size_t force_smem(const SphPairConfig& c) { return (size_t)c.cap * 32 + (size_t)c.kMax * c.threads * 2; }

}  // namespace

#define SPH_COUNT(L) do { if ((L).launches) ++*(L).launches; } while (0)

void sph_pair_default_config(SphPairConfig* cfg)
{
    // rm: cap = records per particle.  A row of three cells is one record per 32 candidates, or one per cell where the
    // maxParInCell truncation cuts it: 27 at most in ordinary scenes; beyond the cap a particle falls back to the walk
    cfg->mode = SPH_PAIR_RM;  cfg->threads = 128;  cfg->cap = 32;  cfg->kMax = 48;
}

const char* sph_pair_mode_name(int mode) { return mode == SPH_PAIR_TMA ? "tma" : mode == SPH_PAIR_RM ? "rm" : "l1"; }

int sph_pair_particles_per_cta(const SphPairConfig& cfg) { return cfg.threads; }

size_t sph_pair_blocks(const SphPairConfig& cfg, int n)
{
    const size_t per = (size_t)sph_pair_particles_per_cta(cfg);
    return ((size_t)n + per - 1) / per;
}

// rm: `cap` records of 8 bytes per particle; the others: kMax list entries per particle
size_t sph_pair_list_bytes(const SphPairConfig& cfg, int n)
{
    if (cfg.mode == SPH_PAIR_RM) return sph_pair_blocks(cfg, n) * (size_t)cfg.cap * cfg.threads * sizeof(uint2);
    return sph_pair_blocks(cfg, n) * cfg.kMax * cfg.threads * (cfg.mode == SPH_PAIR_TMA ? 2 : 4);
}

cudaError_t sph_pair_prepare(const SphPairConfig& cfg)
{
    if (cfg.mode != SPH_PAIR_TMA) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_density, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)density_smem(cfg));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_force, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)force_smem(cfg));
}

void sph_launch_density(const SphLaunch& L, const SphPairConfig& cfg, const SimParams& par,
                        const float4* posS, const float4* velS, const uint32_t* keyS, const uint32_t* cellStart,
                        const uint32_t* maxCount, float4* posP, float4* velD, uint32_t* neighborCounts,
                        void* nlist, uint16_t* ncount, uint32_t* ctaRows, int first, int count, const uint32_t* dev)
{
    if (count <= 0) return;
    const int n = first + count;
    int blocks = (int)sph_pair_blocks(cfg, count);
    if (cfg.mode == SPH_PAIR_RM) {
        // SPH_B200_RM_OCC=12|16 (tuning aid): resident 128-thread CTAs per SM the density kernel is compiled for
        static const int occ = [] { const char* e = getenv("SPH_B200_RM_OCC");  return e && atoi(e) == 16 ? 16 : 12; }();
        static const int unr = [] { const char* e = getenv("SPH_B200_RM_UNROLL");  return e && atoi(e) == 2 ? 2 : 4; }();
        auto launch = [&](auto kern) {
            kern<<<blocks, cfg.threads, 0, L.stream>>>(par, cfg.cap, posS, velS, keyS, cellStart, maxCount,
                                                       posP, velD, neighborCounts, (uint2*)nlist, ncount, first, n, dev);
        };
        if (occ == 16) { if (unr == 2) launch(k_density_rm<16, 2>);  else launch(k_density_rm<16, 4>); }
        else           { if (unr == 2) launch(k_density_rm<12, 2>);  else launch(k_density_rm<12, 4>); }
    } else if (cfg.mode == SPH_PAIR_TMA)
        k_density<<<blocks, cfg.threads, density_smem(cfg), L.stream>>>(par, cfg.cap, cfg.kMax, posS, velS, keyS, cellStart, maxCount,
                                                                        posP, velD, neighborCounts, (uint16_t*)nlist, ncount, ctaRows, first, n);
    else
        k_density_l1<<<blocks, cfg.threads, 0, L.stream>>>(par, cfg.kMax, posS, velS, keyS, cellStart, maxCount,
                                                           posP, velD, neighborCounts, (uint32_t*)nlist, ncount, first, n, dev);
    SPH_COUNT(L);
}

// SPH_B200_FORCE_LISTS=l1|stream (tuning aid): how k_force_l1 reads its neighbour lists
static bool force_list_streaming()
{
    static const int mode = [] { const char* e = getenv("SPH_B200_FORCE_LISTS");  return e && strcmp(e, "l1") == 0 ? 0 : 1; }();
    return mode != 0;
}

void sph_launch_force(const SphLaunch& L, const SphPairConfig& cfg, const SimParams& par,
                      const float4* posP, const float4* velD, const float4* velS, const uint32_t* keyS,
                      const uint32_t* cellStart, const uint32_t* maxCount, const void* nlist, const uint16_t* ncount,
                      const uint32_t* ctaRows, float4* velOut, int first, int count, int ctaFirst, int ctaCount,
                      const uint32_t* dev, int part, int part2Blocks)
{
    if (count <= 0) return;
    const int n = first + count;
    int blocks = (int)sph_pair_blocks(cfg, count);
    if (dev && part == 2 && part2Blocks > 0 && part2Blocks < blocks) blocks = part2Blocks;     // compact grid, see pair_part2_cta
    if (cfg.mode != SPH_PAIR_TMA && ctaCount >= 0) {            // sub-range of the CTAs (slab mode: interior / boundary)
        if (ctaFirst < 0 || ctaFirst + ctaCount > blocks) ctaCount = blocks - ctaFirst;
        if (ctaCount <= 0) return;
        blocks = ctaCount;
    } else ctaFirst = 0;
    if (cfg.mode == SPH_PAIR_RM) {
        static const int occ = [] { const char* e = getenv("SPH_B200_RM_FORCE_OCC");  const int v = e ? atoi(e) : 10;  return v == 8 || v == 12 ? v : 10; }();
        auto launch = [&](auto kern) {
            kern<<<blocks, cfg.threads, 0, L.stream>>>(par, cfg.cap, posP, velD, velS, keyS, cellStart, maxCount,
                                                       (const uint2*)nlist, ncount, velOut, first, n, ctaFirst, dev, part);
        };
        if (occ == 8) launch(k_force_rm<8>);  else if (occ == 12) launch(k_force_rm<12>);  else launch(k_force_rm<10>);
    }
    else if (cfg.mode == SPH_PAIR_TMA)
        k_force<<<blocks, cfg.threads, force_smem(cfg), L.stream>>>(par, cfg.cap, cfg.kMax, posP, velD, velS, keyS, cellStart, maxCount,
                                                                    (const uint16_t*)nlist, ncount, ctaRows, velOut, first, n);
    else if (force_list_streaming())
        k_force_l1<true><<<blocks, cfg.threads, 0, L.stream>>>(par, cfg.kMax, posP, velD, velS, keyS, cellStart, maxCount,
                                                               (const uint32_t*)nlist, ncount, velOut, first, n, ctaFirst, dev, part);
    else
        k_force_l1<false><<<blocks, cfg.threads, 0, L.stream>>>(par, cfg.kMax, posP, velD, velS, keyS, cellStart, maxCount,
                                                                (const uint32_t*)nlist, ncount, velOut, first, n, ctaFirst, dev, part);
    SPH_COUNT(L);
}
