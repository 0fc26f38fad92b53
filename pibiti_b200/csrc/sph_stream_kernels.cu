// sph_stream_kernels.cu -- the HBM-streaming stages of the SPH step for sm_100a:
//   integrate + boundary + cell hash + cell histogram   (one pass, 72 B/particle)
//   cell-table exclusive scan                            (8 B/cell)
//   deterministic stable counting sort + reorder         (bucket, then rank-by-index + gather)
//   original-order accessors
//
// Built with -fmad=false: every float expression here is evaluated with the same IEEE
// operations, in the same order, as the reference source evaluates them when it is compiled for
// a CPU (which is what the parity oracle is).  That makes positions, velocities and therefore
// cell hashes bit-exact against the oracle for every boundary type that uses no libm call.
// These kernels are HBM-bound; the extra multiplies cost nothing.
//
// Reference behaviour restated here (nothing is copied; see DESIGN.md for the mapping):
//   boundary()            source/CUDA/System.cu:41-159
//   integrateD            source/CUDA/System.cu:165-205
//   calcGridPos/Hash      source/CUDA/Kernel_Cell.cui:5-19
//   RadixSort + reorderD  source/CUDA/radixsort_kernel.cu:445-472, Kernel_Cell.cui:40-67
#include "sph_device.cuh"

namespace {

constexpr float kBndEps = 0.00001f;      // System.cu:51
constexpr uint32_t kDeadIndex = 0xFFFFFFFFu;   // original-index value of a retired slot (slab mode)

__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// One soft-wall impulse:  acc = stiff*diff - damp*(n.v);  v += acc*n*dt      (System.cu:52-53)
__device__ __forceinline__ void wall_push(float3& v, float3 n, float diff, float stiff, float damp, float dt)
{
    float acc = stiff * diff - damp * dot3(n, v);
    v.x += (acc * n.x) * dt;
    v.y += (acc * n.y) * dt;
    v.z += (acc * n.z) * dt;
}

struct BoundaryCtx {
    float waveShift;        // rTwist*(1+sinf(rAngle)), evaluated on the host so that it is the same
                            // libm result the CPU oracle sees (System.cu:60)
    float pumpXs, pumpZs;   // pump outlet box: sinf(angOut*s3)*rad and cosf(angOut*s3)*rad*s4 (System.cu:117-119),
                            // parameter-only values, from the host's libm for the same reason
};

inline BoundaryCtx boundary_ctx(const SimParams& par)
{
    BoundaryCtx c;
    c.waveShift = par.rTwist * (1.f + sinf(par.rAngle));
    const float rad = par.worldMax.x;
    c.pumpXs = sinf(par.angOut * par.s3) * rad;
    c.pumpZs = cosf(par.angOut * par.s3) * rad * par.s4;
    return c;
}

// Soft penalty boundaries.  pos may be teleported (wrap / cycle / pump exit).
__device__ __forceinline__ void soft_boundary(const SimParams& par, const BoundaryCtx& ctx, float3& pos, float3& vel)
{
    float3 wmin = par.worldMin, wmax = par.worldMax;
    const float b = par.distBndSoft, stiff = par.bndStiff, dampB = par.bndDamp, dampC = par.bndDampC;
    const float dt = par.timeStep;
    const BndType t = par.bndType;
    const bool cylY = t == BND_CYL_Y, cylZ = t == BND_CYL_Z;
    const bool wave = par.bndEffZ == BND_EFF_WAVE, noEff = par.bndEffZ == BND_EFF_NONE,
               cycle = par.bndEffZ == BND_EFF_CYCLE;
    float diff;

    if (wave) {                                                     // System.cu:55-61
        float sl = -par.r2Angle;
        diff = b - (pos.y - wmin.y) - (pos.z - wmin.z) * sl;
        if (diff > kBndEps) wall_push(vel, make_float3(0.f, 1.f - sl, sl), diff, stiff, dampB, dt);
        wmin.z += ctx.waveShift;
    }

    if (t != BND_SPHERE) {                                          // System.cu:64-77
        if (!cylY) {
            if (noEff || wave) {
                diff = b - pos.z + wmin.z;
                if (diff > kBndEps) wall_push(vel, make_float3(0.f, 0.f, 1.f), diff, stiff, dampC, dt);
            }
            if (!cycle) {
                diff = b + pos.z - wmax.z;
                if (diff > kBndEps) wall_push(vel, make_float3(0.f, 0.f, -1.f), diff, stiff, dampC, dt);
            }
        }
        if (!cylY && !cylZ) {
            diff = b - pos.x + wmin.x;
            if (diff > kBndEps) wall_push(vel, make_float3(1.f, 0.f, 0.f), diff, stiff, dampB, dt);
            diff = b + pos.x - wmax.x;
            if (diff > kBndEps) wall_push(vel, make_float3(-1.f, 0.f, 0.f), diff, stiff, dampB, dt);
        }
        if (!cylZ) {
            diff = b - pos.y + wmin.y;
            if (diff > kBndEps) wall_push(vel, make_float3(0.f, 1.f, 0.f), diff, stiff, dampB, dt);
            diff = b + pos.y - wmax.y;
            if (diff > kBndEps) wall_push(vel, make_float3(0.f, -1.f, 0.f), diff, stiff, dampB, dt);
        }
    } else {                                                        // System.cu:78-81
        float len = sqrtf(dot3(pos, pos));
        diff = b + len + wmin.y;
        if (diff > kBndEps) wall_push(vel, make_float3(-pos.x / len, -pos.y / len, -pos.z / len), diff, stiff, dampC, dt);
    }

    if (cylY || t == BND_CYL_YZ) {                                  // System.cu:84-86
        float len = sqrtf(pos.x * pos.x + pos.z * pos.z);
        diff = b + len - wmax.x;
        if (diff > kBndEps) wall_push(vel, make_float3(-pos.x / len, 0.f, -pos.z / len), diff, stiff, dampC, dt);
    }
    if (cylZ || t == BND_CYL_YZ) {                                  // System.cu:89-91
        float len = sqrtf(pos.x * pos.x + pos.y * pos.y);
        diff = b + len + wmin.y;
        if (diff > kBndEps) wall_push(vel, make_float3(-pos.x / len, -pos.y / len, 0.f), diff, stiff, dampC, dt);
    }

    if (!wave && !noEff) {                                          // wrap / cycle in Z, System.cu:94-97
        float dr = 1.f * par.particleR;
        if (cycle && vel.z > par.rVexit && pos.z > wmax.z - b - dr)  pos.z -= wmax.z - wmin.z - 2 * b - dr;
        else if (vel.z < -par.rVexit && pos.z < wmin.z + b + dr)     pos.z += wmax.z - wmin.z - 2 * b - dr;
    }

    if (t == BND_PUMP_Y) {                                          // System.cu:101-158
        const float rad = wmax.x, ang = par.angOut, hc = par.hClose, rin = rad * par.radIn;
        float len = sqrtf(pos.x * pos.x + pos.y * pos.y);
        diff = b + len - rad;
        if (diff > kBndEps) {                                       // cylinder frame
            float a = atanf(pos.x / pos.y);
            bool hit = ang < 0.5f ? (a < -ang || a > ang || pos.y < 0)
                                  : (pos.y < 0 || (len < rad * par.s5 && a < ang));
            if (hit) wall_push(vel, make_float3(-pos.x / len, -pos.y / len, 0.f), diff, stiff, dampB, dt);
        }
        float xs;                                                   // outlet box
        if (ang < 0.5f) {
            xs = ctx.pumpXs;
            const float zs = ctx.pumpZs;
            if (pos.y > zs) {
                diff = b - pos.x - xs;
                if (diff > kBndEps) wall_push(vel, make_float3(1.f, 0.f, 0.f), diff, stiff, dampB, dt);
                diff = b + pos.x - xs;
                if (diff > kBndEps) wall_push(vel, make_float3(-1.f, 0.f, 0.f), diff, stiff, dampB, dt);
            }
        } else {
            xs = 0.09f * par.s4;
            if (len >= rad * par.s6) {
                diff = b - pos.x + xs;
                if (diff > kBndEps) wall_push(vel, make_float3(1.f, 0.f, 0.f), diff, stiff, dampB, dt);
            }
        }
        if (pos.z > hc - b * par.s1) {                              // inlet hole
            diff = b + len - rin;
            if (diff > kBndEps) wall_push(vel, make_float3(-pos.x / len, -pos.y / len, 0.f), diff, stiff, dampB, dt);
        }
        if (pos.z < hc - b * par.s2) {
            diff = b + pos.z - hc;
            if (diff > kBndEps) wall_push(vel, make_float3(0.f, 0.f, -1.f), diff, stiff, dampB, dt);
        }
        diff = pos.y - wmax.y + par.rDexit;                         // exit -> inlet teleport
        if (diff > kBndEps && vel.y > par.rVexit) {
            float aa, rr;
            float zz = fabsf(hc - wmin.z);
            if (ang < 0.5f) {
                float xx = xs * 2;
                rr = (pos.x + xx / 2) / xx * 0.7f;
                aa = (pos.z - zz / 2) / zz * 1.6f;
            } else {
                rr = (wmax.x - pos.x) / xs * 0.45f;
                aa = (pos.z - zz / 2) / zz * 1.8f;
            }
            rr *= rin;  aa *= 2.f * PI;
            float x = cosf(aa) * rr, y = sinf(aa) * rr;
            // the reference multiplies by the double literal 0.01 here (System.cu:154)
            float z = (float)((double)(wmax.z - b) - (double)fabsf(vel.y - par.rVexit) * 0.01);
            pos = make_float3(x, y, z);
            vel = make_float3(vel.x, vel.z, -vel.y);
        }
    }
}

// calcGridPos + calcGridHash: true division per component, floor, z-major linear hash.
__device__ __forceinline__ uint32_t cell_hash(const SimParams& par, float3 p)
{
    int gx = (int)floorf((p.x - par.worldMin.x) / par.cellSize.x);
    int gy = (int)floorf((p.y - par.worldMin.y) / par.cellSize.y);
    int gz = (int)floorf((p.z - par.worldMin.z) / par.cellSize.z);
    return (uint32_t)(gz * (int)par.gridSize_yx + gy * (int)par.gridSize.x + gx);
}

// boundary impulse -> gravity -> damping -> position -> hard clamp (step order Q6)
__device__ __forceinline__ void integrate_particle(const SimParams& par, const BoundaryCtx& ctx, float3& p, float3& v)
{
    soft_boundary(par, ctx, p, v);

    const float dt = par.timeStep;                                  // System.cu:177-179
    v.x += par.gravity.x * dt;  v.y += par.gravity.y * dt;  v.z += par.gravity.z * dt;
    v.x *= par.globalDamping;   v.y *= par.globalDamping;   v.z *= par.globalDamping;
    p.x += v.x * dt;            p.y += v.y * dt;            p.z += v.z * dt;

    const float hb = par.distBndHard;                               // System.cu:193-200
    if (p.x > par.worldMax.x - hb) p.x = par.worldMax.x - hb;
    if (p.x < par.worldMin.x + hb) p.x = par.worldMin.x + hb;
    if (p.y > par.worldMax.y - hb) p.y = par.worldMax.y - hb;
    if (p.y < par.worldMin.y + hb) p.y = par.worldMin.y + hb;
    if (p.z > par.worldMax.z - hb) p.z = par.worldMax.z - hb;
    if (p.z < par.worldMin.z + hb) p.z = par.worldMin.z + hb;

}

__global__ void __launch_bounds__(256)
k_integrate_hash(const __grid_constant__ SimParams par, const BoundaryCtx ctx,
                 float4* __restrict__ pos, float4* __restrict__ vel,
                 uint32_t* __restrict__ keyU, uint32_t* __restrict__ rankU,
                 uint32_t* __restrict__ cellCount, int first, int n)
{
    int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p4 = pos[i], v4 = vel[i];
    float3 p = make_float3(p4.x, p4.y, p4.z), v = make_float3(v4.x, v4.y, v4.z);

    integrate_particle(par, ctx, p, v);

    pos[i] = make_float4(p.x, p.y, p.z, p4.w);
    vel[i] = make_float4(v.x, v.y, v.z, v4.w);
    if (!cellCount) return;                 // slab mode: hashing happens after migration and halo exchange

    uint32_t key = cell_hash(par, p);
    // The hard clamp keeps every finite particle inside the grid.  A NaN position is undefined
    // behaviour in the reference (out-of-bounds cellStart write); here it lands in the last cell.
    if (key >= par.numCells) key = par.numCells - 1;
    keyU[i] = key;
    rankU[i] = atomicAdd(&cellCount[key], 1u);
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan of the cell histogram: reduce per tile -> scan the tile sums -> scan per tile.
// The last pass also zeroes the histogram for the next step and records the largest cell.

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix,
// total in *total
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warpSums[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) warpSums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < 8 ? warpSums[lane] : 0u;
        uint32_t si = warp_incl_scan(s);
        if (lane < 8) warpSums[lane] = si - s;
        if (lane == 7) *total = si;
    }
    __syncthreads();
    uint32_t r = incl - v + warpSums[w];
    __syncthreads();
    return r;
}

// Occupied-range bound (slab mode).  A slab's cell table spans its whole z range, of which the last slab of a long
// tank uses a fraction.  The histogram kernel records the largest live key; tiles that lie entirely above
// bound = largest key + guard (two layers: every neighbour lookup stays below it) hold only zero counts and are not
// read, zeroed or written.  The last tile (the dummy cell and cellStart[numCells]) is always processed.  cellStart
// above the bound is stale; nothing on the device reads it and the host substitutes the live total (sph_slab_sort).
__device__ __forceinline__ bool scan_tile_skipped(const uint32_t* __restrict__ boundCells)
{
    if (!boundCells) return false;
    return (uint32_t)blockIdx.x * SPH_SCAN_TILE >= __ldg(boundCells) && blockIdx.x != gridDim.x - 1;
}

// A tile is four chunks of 1024 cells; thread t owns cells [4t, 4t+4) of every chunk, so that a warp reads and writes 512
// contiguous bytes per instruction (sixteen consecutive cells per thread, the round-1 layout, made every 128-bit access of a
// warp straddle sixteen lines: the scan ran at a third of the HBM rate).
__global__ void __launch_bounds__(256)
k_scan_reduce(const uint32_t* __restrict__ cnt, uint32_t* __restrict__ tileSums, int numCells,
              const uint32_t* __restrict__ boundCells)
{
    __shared__ uint32_t total;
    if (scan_tile_skipped(boundCells)) { if (threadIdx.x == 0) tileSums[blockIdx.x] = 0;  return; }
    uint32_t s = 0;
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int base = blockIdx.x * SPH_SCAN_TILE + j * 1024 + threadIdx.x * 4;
        if (base + 4 <= numCells) {
            const uint4 q = *reinterpret_cast<const uint4*>(cnt + base);
            s += q.x + q.y + q.z + q.w;
        } else {
            for (int k = 0; k < 4; k++) if (base + k < numCells) s += cnt[base + k];
        }
    }
    block_excl_scan_256(s, &total);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place, 16 per thread per round; tileSums[numTiles] = grand total
__global__ void __launch_bounds__(256)
k_scan_tiles(uint32_t* __restrict__ tileSums, int numTiles, uint32_t* __restrict__ maxCount)
{
    __shared__ uint32_t total;
    uint32_t carry = 0;
    if (threadIdx.x == 0) { maxCount[0] = 0;  maxCount[kBigCount] = 0; }
    for (int base = 0; base < numTiles; base += 4096) {
        const int i0 = base + threadIdx.x * 16;
        uint32_t v[16], s = 0;
        #pragma unroll
        for (int k = 0; k < 16; k++) { v[k] = i0 + k < numTiles ? tileSums[i0 + k] : 0u;  s += v[k]; }
        uint32_t ex = carry + block_excl_scan_256(s, &total);
        #pragma unroll
        for (int k = 0; k < 16; k++) { if (i0 + k < numTiles) tileSums[i0 + k] = ex;  ex += v[k]; }
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tileSums[numTiles] = carry;
}

__global__ void __launch_bounds__(256)
k_scan_apply(uint32_t* __restrict__ cnt, uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ tileSums,
             uint32_t* __restrict__ maxCount, int numCells, int maxCells, const uint32_t* __restrict__ boundCells)
{
    // exclusive offsets of (chunk j, warp w) in cell order: one warp scan over the 4 x 8 warp totals
    __shared__ uint32_t warpExcl[32];
    if (scan_tile_skipped(boundCells)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int tile0 = blockIdx.x * SPH_SCAN_TILE;
    uint32_t c[4][4], incl[4], mx = 0;
    // all four loads first (the atomics of the big-cell list below would otherwise fence them into a chain)
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int base = tile0 + j * 1024 + threadIdx.x * 4;
        if (base + 4 <= numCells) {
            const uint4 q = *reinterpret_cast<const uint4*>(cnt + base);
            c[j][0] = q.x;  c[j][1] = q.y;  c[j][2] = q.z;  c[j][3] = q.w;
        } else {
            #pragma unroll
            for (int k = 0; k < 4; k++) c[j][k] = base + k < numCells ? cnt[base + k] : 0u;
        }
    }
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int base = tile0 + j * 1024 + threadIdx.x * 4;
        if (base + 4 <= numCells) *reinterpret_cast<uint4*>(cnt + base) = make_uint4(0u, 0u, 0u, 0u);     // clean for the next step
        else {
            #pragma unroll
            for (int k = 0; k < 4; k++) if (base + k < numCells) cnt[base + k] = 0;
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            if (base + k < maxCells) {
                mx = max(mx, c[j][k]);
                // a cell too full for the per-entry counting of k_rank_gather goes on the list k_rank_big_cells works off
                if (c[j][k] > (uint32_t)kBigCell) {
                    const uint32_t at = atomicAdd(maxCount + kBigCount, 1u);
                    if (at < (uint32_t)kBigCap) maxCount[kBigList + at] = (uint32_t)(base + k);
                }
            }
        }
        incl[j] = warp_incl_scan(c[j][0] + c[j][1] + c[j][2] + c[j][3]);
        if (lane == 31) warpExcl[j * 8 + w] = incl[j];
    }
    __syncthreads();
    if (w == 0) {
        const uint32_t v = warpExcl[lane], iv = warp_incl_scan(v);
        warpExcl[lane] = iv - v;
    }
    __syncthreads();
    const uint32_t tileBase = tileSums[blockIdx.x];
    #pragma unroll
    for (int j = 0; j < 4; j++) {
        const int base = tile0 + j * 1024 + threadIdx.x * 4;
        uint32_t ex = tileBase + warpExcl[j * 8 + w] + incl[j] - (c[j][0] + c[j][1] + c[j][2] + c[j][3]);
        uint32_t o[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) { o[k] = ex;  ex += c[j][k]; }
        if (base + 4 <= numCells) *reinterpret_cast<uint4*>(cellStart + base) = make_uint4(o[0], o[1], o[2], o[3]);
        else {
            #pragma unroll
            for (int k = 0; k < 4; k++) if (base + k < numCells) cellStart[base + k] = o[k];
        }
        // cellStart[numCells] = n : the thread that owns the last cell writes it
        if (base <= numCells - 1 && numCells - 1 < base + 4) cellStart[numCells] = o[numCells - 1 - base] + c[j][numCells - 1 - base];
    }

    // largest cell population (decides whether the neighbour walk must truncate)
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    // only when it would raise the value: same-address atomics serialise in L2
    if ((threadIdx.x & 31) == 0 && mx > *reinterpret_cast<volatile uint32_t*>(maxCount)) atomicMax(maxCount, mx);
}

// ------------------------------------------------------------------------------------------------
// Counting sort, made deterministic and stable.
//  bucket:       slot j goes to cellStart[key] + (arrival rank from the histogram atomic).  The order
//                inside a cell is arbitrary at this point.
//  rank_gather:  every bucketed entry counts how many entries of its cell carry a smaller ORIGINAL
//                index; that count is its stable rank.  The result is exactly the order a stable
//                sort by cell hash of the original-order particle list produces, i.e. the
//                reference's RadixSort output, with no multi-pass radix sort.

__global__ void __launch_bounds__(256)
k_bucket(const uint32_t* __restrict__ keyU, const uint32_t* __restrict__ rankU, const uint32_t* __restrict__ idxIn,
         const uint32_t* __restrict__ cellStart, uint2* __restrict__ pairT, int n, const uint32_t* __restrict__ nDev)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (nDev ? (int)*nDev : n)) return;
    uint32_t dst = cellStart[keyU[j]] + rankU[j];
    pairT[dst] = make_uint2((uint32_t)j, idxIn[j]);
}

__global__ void __launch_bounds__(256)
k_rank_gather(const uint2* __restrict__ pairT, const uint32_t* __restrict__ keyU, const uint32_t* __restrict__ cellStart,
              const float4* __restrict__ posIn, const float4* __restrict__ velIn,
              float4* __restrict__ posOut, float4* __restrict__ velOut,
              uint32_t* __restrict__ idxOut, uint32_t* __restrict__ keyS, int n, const uint32_t* __restrict__ nDev,
              const uint32_t* __restrict__ big, int realCells)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= (nDev ? (int)*nDev : n)) return;
    uint2 me = pairT[d];
    uint32_t key = keyU[me.x];
    uint32_t s = cellStart[key], e = cellStart[key + 1];
    // cells on the big-cell list (more than kBigCell entries) are ranked by a whole CTA each, k_rank_big_cells
    if (big && e - s > (uint32_t)kBigCell && key < (uint32_t)realCells && __ldg(big + kBigCount) <= (uint32_t)kBigCap) return;
    uint32_t r = 0;
    if (me.y == kDeadIndex) r = (uint32_t)d - s;            // slab mode: retired slots, order irrelevant
    else for (uint32_t t = s; t < e; t++) r += (pairT[t].y < me.y) ? 1u : 0u;
    uint32_t f = s + r;
    float4 p = posIn[me.x], v = velIn[me.x];
    posOut[f] = p;
    velOut[f] = v;
    idxOut[f] = me.y;
    keyS[f] = key;
}

// The same stable rank + gather for the cells on the big-cell list: one CTA per cell, the original indices of the cell
// staged through shared memory in tiles, every thread counting for one entry at a time.  Counting costs n^2 / 256 per
// thread instead of n^2 / 1 (a cell holds a handful of particles in a healthy fluid; this is for the pile-up that a
// diverging run or a degenerate initial state produces, where the per-entry loop would take seconds).
__global__ void __launch_bounds__(256)
k_rank_big_cells(const uint2* __restrict__ pairT, const uint32_t* __restrict__ cellStart,
                 const float4* __restrict__ posIn, const float4* __restrict__ velIn,
                 float4* __restrict__ posOut, float4* __restrict__ velOut,
                 uint32_t* __restrict__ idxOut, uint32_t* __restrict__ keyS, const uint32_t* __restrict__ big)
{
    __shared__ uint32_t tile[1024];
    const uint32_t listed = min(__ldg(big + kBigCount), (uint32_t)kBigCap);
    if (__ldg(big + kBigCount) > (uint32_t)kBigCap) return;             // list overflowed: k_rank_gather did everything
    for (uint32_t b = blockIdx.x; b < listed; b += gridDim.x) {
        const uint32_t key = __ldg(big + kBigList + b);
        const uint32_t s = cellStart[key], e = cellStart[key + 1];
        for (uint32_t chunk = s; chunk < e; chunk += blockDim.x) {
            const uint32_t d = chunk + threadIdx.x;
            const bool have = d < e;
            const uint2 me = have ? pairT[d] : make_uint2(0u, 0u);
            uint32_t r = 0;
            for (uint32_t t0 = s; t0 < e; t0 += 1024u) {
                const uint32_t m = min(1024u, e - t0);
                __syncthreads();
                for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) tile[k] = pairT[t0 + k].y;
                __syncthreads();
                if (have) for (uint32_t k = 0; k < m; k++) r += (tile[k] < me.y) ? 1u : 0u;
            }
            if (have) {
                const uint32_t f = s + r;
                posOut[f] = posIn[me.x];
                velOut[f] = velIn[me.x];
                idxOut[f] = me.y;
                keyS[f] = key;
            }
        }
    }
}

__global__ void k_iota(uint32_t* idx, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

__global__ void k_unpermute4(const float4* __restrict__ src, const uint32_t* __restrict__ idx, float4* __restrict__ out,
                             int start, int count, int n)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t o = idx[j] - (uint32_t)start;
    if (o < (uint32_t)count) out[o] = src[j];
}

__global__ void k_unpermute_w(const float4* __restrict__ src, const uint32_t* __restrict__ idx, float* __restrict__ out,
                              int start, int count, int n)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t o = idx[j] - (uint32_t)start;
    if (o < (uint32_t)count) out[o] = src[j].w;
}

__global__ void k_permute4(float4* __restrict__ dst, const uint32_t* __restrict__ idx, const float4* __restrict__ in,
                           int start, int count, int n)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t o = idx[j] - (uint32_t)start;
    if (o < (uint32_t)count) dst[j] = in[o];
}

__global__ void k_cell_table_dump(const uint32_t* __restrict__ cellStart, uint32_t* __restrict__ outStart,
                                  uint32_t* __restrict__ outEnd, int numCells)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= numCells) return;
    uint32_t s = cellStart[c], e = cellStart[c + 1];
    if (outStart) outStart[c] = (s == e) ? 0xffffffffu : s;       // reference cellStart: memset 0xff, set at run starts
    if (outEnd) outEnd[c] = e;
}

__global__ void k_pack_pairs(const uint32_t* __restrict__ keyS, const uint32_t* __restrict__ idx, uint2* __restrict__ out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uint2(keyS[i], idx[i]);
}

// ------------------------------------------------------------------------------------------------
// Slab decomposition (one handle per GPU owns the z-cell layers [zLo, zHi)).
// Records travelling between ranks are 48 bytes: position, velocity, (original index, 0, 0, 0).

struct SlabRecord { float4 pos; float4 vel; uint4 meta; };

__device__ __forceinline__ int z_cell(const SimParams& par, float z)
{
    return (int)floorf((z - par.worldMin.z) / par.cellSize.z);
}

// Append to a message section: one atomic per warp instead of one per record (the boundary layers are contiguous
// runs of the sorted state, so whole warps append together and same-address atomics serialise in L2).  Every lane
// of the warp must call it.
__device__ __forceinline__ uint32_t warp_append_slot(uint32_t* ctr, bool pred)
{
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0u) return 0u;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// owned particles [first, n): those whose z-cell left [zLo, zHi) are copied out and retired
__global__ void __launch_bounds__(256)
k_slab_take_leavers(const __grid_constant__ SimParams par, const float4* __restrict__ pos, const float4* __restrict__ vel,
                    uint32_t* __restrict__ idx, int first, int n, int zLo, int zHi, int hasLower, int hasUpper,
                    SlabRecord* __restrict__ down, int capDown, SlabRecord* __restrict__ up, int capUp,
                    uint32_t* __restrict__ ctrDown, uint32_t* __restrict__ ctrUp)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n && idx[i] != kDeadIndex;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    bool goDown = false, goUp = false;
    if (live) {
        p = pos[i];
        const int zc = z_cell(par, p.z);
        goDown = zc < zLo && hasLower;
        goUp = zc >= zHi && hasUpper;
    }
    const uint32_t slotDown = warp_append_slot(ctrDown, goDown), slotUp = warp_append_slot(ctrUp, goUp);
    if (!(goDown || goUp)) return;
    SlabRecord r;  r.pos = p;  r.vel = vel[i];  r.meta = make_uint4(idx[i], 0u, 0u, 0u);
    if (goDown) { if (slotDown < (uint32_t)capDown) down[slotDown] = r; }
    else if (slotUp < (uint32_t)capUp) up[slotUp] = r;
    idx[i] = kDeadIndex;
}

// copies of the live owned particles in the first / last owned layer (the neighbours' ghosts)
__global__ void __launch_bounds__(256)
k_slab_boundary(const __grid_constant__ SimParams par, const float4* __restrict__ pos, const float4* __restrict__ vel,
                const uint32_t* __restrict__ idx, int n, int zLo, int zHi, int hasLower, int hasUpper,
                SlabRecord* __restrict__ down, int capDown, SlabRecord* __restrict__ up, int capUp,
                uint32_t* __restrict__ ctrDown, uint32_t* __restrict__ ctrUp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n && idx[i] != kDeadIndex;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    bool toDown = false, toUp = false;
    if (live) {
        p = pos[i];
        const int zc = z_cell(par, p.z);            // a ghost left over in the work set (outside [zLo,zHi)) is not ours to export
        toDown = zc == zLo && hasLower;
        toUp = zc == zHi - 1 && hasUpper && zc >= zLo;
    }
    const uint32_t slotDown = warp_append_slot(ctrDown, toDown), slotUp = warp_append_slot(ctrUp, toUp);
    if (!(toDown || toUp)) return;
    SlabRecord r;  r.pos = p;  r.vel = vel[i];  r.meta = make_uint4(idx[i], 0u, 0u, 0u);
    if (toDown && slotDown < (uint32_t)capDown) down[slotDown] = r;
    if (toUp && slotUp < (uint32_t)capUp) up[slotUp] = r;
}

// sph_slab_integrate + sph_slab_pack in one pass over the work set [0, work): slots outside the owned range
// [first, first+count) are retired (last step's ghosts); owned particles are integrated, and on the way out a particle
// whose z cell left [zLo, zHi) moves into the neighbour's leaver section (and is retired here), while one in the first
// or last owned layer leaves a copy in the neighbour's boundary section.
__global__ void __launch_bounds__(256)
k_slab_integrate_pack(const __grid_constant__ SimParams par, const BoundaryCtx ctx, float4* __restrict__ pos,
                      float4* __restrict__ vel, uint32_t* __restrict__ idx, int first, int count, int work,
                      int zLo, int zHi, int hasLower, int hasUpper,
                      SlabRecord* __restrict__ leavDown, SlabRecord* __restrict__ leavUp, int capL,
                      SlabRecord* __restrict__ bndDown, SlabRecord* __restrict__ bndUp, int capB,
                      uint32_t* __restrict__ headDown, uint32_t* __restrict__ headUp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool owned = i >= first && i < first + count;
    if (i < work && !owned) idx[i] = kDeadIndex;
    uint32_t id = kDeadIndex;
    float4 pOut = make_float4(0.f, 0.f, 0.f, 0.f), vOut = pOut;
    bool goDown = false, goUp = false, copyDown = false, copyUp = false;
    if (owned) {
        id = idx[i];
        const float4 p4 = pos[i], v4 = vel[i];
        float3 p = make_float3(p4.x, p4.y, p4.z), v = make_float3(v4.x, v4.y, v4.z);
        integrate_particle(par, ctx, p, v);
        pOut = make_float4(p.x, p.y, p.z, p4.w);
        vOut = make_float4(v.x, v.y, v.z, v4.w);
        pos[i] = pOut;
        vel[i] = vOut;
        if (id != kDeadIndex) {
            const int zc = z_cell(par, p.z);
            goDown = zc < zLo && hasLower;
            goUp = zc >= zHi && hasUpper;
            copyDown = zc == zLo && hasLower;
            copyUp = zc == zHi - 1 && hasUpper && zc >= zLo;
        }
    }
    const uint32_t sLeavDown = warp_append_slot(headDown + 0, goDown), sLeavUp = warp_append_slot(headUp + 0, goUp);
    const uint32_t sBndDown = warp_append_slot(headDown + 1, copyDown), sBndUp = warp_append_slot(headUp + 1, copyUp);
    if (!(goDown || goUp || copyDown || copyUp)) return;
    SlabRecord r;  r.pos = pOut;  r.vel = vOut;  r.meta = make_uint4(id, 0u, 0u, 0u);
    if (goDown) { if (sLeavDown < (uint32_t)capL) leavDown[sLeavDown] = r; }
    else if (goUp) { if (sLeavUp < (uint32_t)capL) leavUp[sLeavUp] = r; }
    if (goDown || goUp) idx[i] = kDeadIndex;
    if (copyDown && sBndDown < (uint32_t)capB) bndDown[sBndDown] = r;
    if (copyUp && sBndUp < (uint32_t)capB) bndUp[sBndUp] = r;
}

__global__ void __launch_bounds__(256)
k_slab_append(const SlabRecord* __restrict__ recs, int count, float4* __restrict__ pos, float4* __restrict__ vel,
              uint32_t* __restrict__ idx, int at)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    SlabRecord r = recs[i];
    pos[at + i] = r.pos;  vel[at + i] = r.vel;  idx[at + i] = r.meta.x;
}

__global__ void __launch_bounds__(256)
k_slab_export(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint32_t* __restrict__ idx,
              const float4* __restrict__ posP, const float4* __restrict__ velD, int first, int count, SlabRecord* __restrict__ recs)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    SlabRecord r;  r.pos = pos[first + i];  r.vel = vel[first + i];
    // after a step the spare words carry density and pressure (ignored when a record is imported)
    r.meta = make_uint4(idx[first + i], posP ? __float_as_uint(velD[first + i].w) : 0u, posP ? __float_as_uint(posP[first + i].w) : 0u, 0u);
    recs[i] = r;
}

// Message layout (both directions): row 0 = header {nLeavers, nBoundary, 0...} (uint32 words), rows [1, 1+capL) =
// particles that leave towards the receiver (they become OWNED there), rows [1+capL, 1+capL+capB) = copies of the
// sender's boundary layer (they become GHOSTS there).  One launch appends, behind slot work0:
//   leavers from below, leavers from above                       -> owned arrivals
//   boundary copies from below / above, and this rank's own leavers (now sitting in the neighbour's
//   boundary layer, i.e. in this rank's ghost layer)             -> ghosts
// Counts stay on the device; dev[0] receives the new work-set size, dev[1] an overflow flag.
__global__ void __launch_bounds__(256)
k_slab_unpack(const SlabRecord* __restrict__ inBelow, const SlabRecord* __restrict__ inAbove,
              const SlabRecord* __restrict__ ownDown, const SlabRecord* __restrict__ ownUp, int capL, int capB,
              float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ idx, int work0, int capacity,
              uint32_t* __restrict__ dev)
{
    const SlabRecord* msg[6] = {inBelow, inAbove, inBelow, inAbove, ownDown, ownUp};
    const int word[6] = {0, 0, 1, 1, 0, 0};                 // header word holding the section's count
    const int row0[6] = {1, 1, 1 + capL, 1 + capL, 1, 1};   // first row of the section
    const int cap[6]  = {capL, capL, capB, capB, capL, capL};
    const int sec = blockIdx.y;
    uint32_t cnt[6], off = 0, total = 0, over = 0;
    #pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t c = msg[k] ? reinterpret_cast<const uint32_t*>(msg[k])[word[k]] : 0u;
        if (c > (uint32_t)cap[k]) { over = 1;  c = (uint32_t)cap[k]; }
        cnt[k] = c;
        if (k < sec) off += c;
        total += c;
    }
    if ((long long)work0 + total > capacity) over = 1;
    if (sec == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        dev[0] = (uint32_t)min((long long)work0 + total, (long long)capacity);
        if (over) dev[1] = 1u;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mine = 0;
    #pragma unroll
    for (int k = 0; k < 6; k++) if (k == sec) mine = cnt[k];
    if (i >= (int)mine) return;
    const long long dst = (long long)work0 + off + i;
    if (dst >= capacity) return;
    const SlabRecord* src = nullptr;  int r0 = 0;
    #pragma unroll
    for (int k = 0; k < 6; k++) if (k == sec) { src = msg[k];  r0 = row0[k]; }
    const SlabRecord r = src[r0 + i];
    pos[dst] = r.pos;  vel[dst] = r.vel;  idx[dst] = r.meta.x;
}

__global__ void k_fill_u32(uint32_t* p, uint32_t v, int first, int n)
{
    int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// hash + histogram of the work set [0, n): local key = global hash - keyOffset; retired slots and anything
// outside the local table go to the dummy cell numCellsLocal, which sorts behind every real cell
__global__ void __launch_bounds__(256)
k_slab_hash_hist(const __grid_constant__ SimParams par, const float4* __restrict__ pos, const uint32_t* __restrict__ idx,
                 uint32_t* __restrict__ keyU, uint32_t* __restrict__ rankU, uint32_t* __restrict__ cellCount,
                 const uint32_t* __restrict__ nDev, long long keyOffset, int numCellsLocal, uint32_t* __restrict__ keyMaxSlots)
{
    __shared__ uint32_t blockMax;
    if (threadIdx.x == 0) blockMax = 0;
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t live1 = 0;                     // live key + 1
    if (i < (int)*nDev) {
        uint32_t key = (uint32_t)numCellsLocal;
        if (idx[i] != kDeadIndex) {
            const float4 p = pos[i];
            long long k = (long long)cell_hash(par, make_float3(p.x, p.y, p.z)) - keyOffset;
            if (k >= 0 && k < (long long)numCellsLocal) { key = (uint32_t)k;  live1 = key + 1; }
            // a LIVE particle outside the local table moved further than the one-layer halo in a step (or was teleported):
            // the decomposition cannot follow it, which is reported (word nDev[2], read back by sph_slab_sort), not hidden
            else atomicOr(const_cast<uint32_t*>(nDev) + 2, 1u);
        }
        keyU[i] = key;
        rankU[i] = atomicAdd(&cellCount[key], 1u);
    }
    live1 = __reduce_max_sync(0xffffffffu, live1);
    if ((threadIdx.x & 31) == 0 && live1) atomicMax(&blockMax, live1);
    __syncthreads();
    if (threadIdx.x == 0 && blockMax) atomicMax(&keyMaxSlots[blockIdx.x & (kKeyMaxSlots - 1)], blockMax);
}

// one warp: bound = min(numCellsLocal, largest live key + 1 + guard); the slots are cleared for the next step
__global__ void k_slab_scan_bound(uint32_t* __restrict__ keyMaxSlots, uint32_t guard, uint32_t numCellsLocal)
{
    uint32_t m = 0;
    for (int k = threadIdx.x; k < kKeyMaxSlots; k += 32) { m = max(m, keyMaxSlots[k]);  keyMaxSlots[k] = 0; }
    m = __reduce_max_sync(0xffffffffu, m);
    if (threadIdx.x == 0) {
        unsigned long long b = (unsigned long long)m + guard;
        keyMaxSlots[kKeyMaxSlots] = b < numCellsLocal ? (uint32_t)b : numCellsLocal;
    }
}

// ------------------------------------------------------------------------------------------------
// Slab mode with device-resident bookkeeping (the multi-GPU driver, sph_capi.cu "multi-GPU driver").  The ranges the
// host used to read back after every sort live in the state words st[] (SlabDevWord, sph_device.cuh); kernels are
// launched over upper bounds and find their own ranges, so a step never synchronises with the host.

// Phase A: integrate the first two and the last two owned layers only -- a particle moves at most one layer per step, so
// these are all that can leave the slab or land in a boundary layer -- and pack what the neighbours need from them
// (leavers, copies of the boundary layers).  Thread t < nLo handles slot first + t, the next nHi threads the slots from
// bHi on.
__global__ void __launch_bounds__(256)
k_slab_boundary_integrate_pack(const __grid_constant__ SimParams par, const BoundaryCtx ctx, float4* __restrict__ pos,
                               float4* __restrict__ vel, uint32_t* __restrict__ idx, uint32_t* __restrict__ st,
                               int zLo, int zHi, int hasLower, int hasUpper,
                               SlabRecord* __restrict__ leavDown, SlabRecord* __restrict__ leavUp, int capL,
                               SlabRecord* __restrict__ bndDown, SlabRecord* __restrict__ bndUp, int capB,
                               uint32_t* __restrict__ headDown, uint32_t* __restrict__ headUp,
                               SlabRecord* __restrict__ peerDown, SlabRecord* __restrict__ peerUp)
{
    // peerDown / peerUp (peer-store exchange): the neighbours' INBOXES, written straight over NVLink -- row 0 is the header
    // (published by k_slab_publish_headers once the counts are final), leaver rows from 1, boundary rows from 1 + capL
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g0 = st[SD_FIRST], g1 = st[SD_END], bLo = st[SD_BLO2], bHi = st[SD_BHI2];
    const uint32_t nLo = bLo - g0, nHi = g1 - bHi;
    // the launch is sized for layers that fit the messages; fuller ones do not fit the messages either
    if (t == 0 && nLo + nHi > gridDim.x * blockDim.x) st[SD_OVERFLOW] = 1u;
    const bool active = t < nLo + nHi;
    const uint32_t i = t < nLo ? g0 + t : bHi + (t - nLo);
    uint32_t id = kDeadIndex;
    float4 pOut = make_float4(0.f, 0.f, 0.f, 0.f), vOut = pOut;
    bool goDown = false, goUp = false, copyDown = false, copyUp = false;
    if (active) {
        id = idx[i];
        const float4 p4 = pos[i], v4 = vel[i];
        float3 p = make_float3(p4.x, p4.y, p4.z), v = make_float3(v4.x, v4.y, v4.z);
        integrate_particle(par, ctx, p, v);
        pOut = make_float4(p.x, p.y, p.z, p4.w);
        vOut = make_float4(v.x, v.y, v.z, v4.w);
        pos[i] = pOut;
        vel[i] = vOut;
        if (id != kDeadIndex) {
            const int zc = z_cell(par, p.z);
            goDown = zc < zLo && hasLower;
            goUp = zc >= zHi && hasUpper;
            copyDown = zc == zLo && hasLower;
            copyUp = zc == zHi - 1 && hasUpper && zc >= zLo;
        }
    }
    const uint32_t sLeavDown = warp_append_slot(headDown + 0, goDown), sLeavUp = warp_append_slot(headUp + 0, goUp);
    const uint32_t sBndDown = warp_append_slot(headDown + 1, copyDown), sBndUp = warp_append_slot(headUp + 1, copyUp);
    if (!(goDown || goUp || copyDown || copyUp)) return;
    SlabRecord r;  r.pos = pOut;  r.vel = vOut;  r.meta = make_uint4(id, 0u, 0u, 0u);
    // leavers stay in the local message too: they are this slab's own ghosts on that side (k_slab_unpack_hist)
    if (goDown) { if (sLeavDown < (uint32_t)capL) { leavDown[sLeavDown] = r;  if (peerDown) peerDown[1 + sLeavDown] = r; } }
    else if (goUp) { if (sLeavUp < (uint32_t)capL) { leavUp[sLeavUp] = r;  if (peerUp) peerUp[1 + sLeavUp] = r; } }
    if (goDown || goUp) idx[i] = kDeadIndex;
    if (copyDown && sBndDown < (uint32_t)capB) { if (peerDown) peerDown[1 + capL + sBndDown] = r; else bndDown[sBndDown] = r; }
    if (copyUp && sBndUp < (uint32_t)capB) { if (peerUp) peerUp[1 + capL + sBndUp] = r; else bndUp[sBndUp] = r; }
}

// peer-store exchange: the final counts of this slab's two messages go into the header rows of the neighbours' inboxes
__global__ void k_slab_publish_headers(const uint32_t* __restrict__ headDown, const uint32_t* __restrict__ headUp,
                                       uint32_t* __restrict__ peerDown, uint32_t* __restrict__ peerUp)
{
    if (threadIdx.x < 2) {
        if (peerDown) peerDown[threadIdx.x] = headDown[threadIdx.x];
        if (peerUp) peerUp[threadIdx.x] = headUp[threadIdx.x];
    }
}

// local key of a live particle: global hash minus the slab's offset; anything outside the local table goes to the dummy
// cell numCellsLocal, which sorts behind every real cell.  An OWNED particle outside the owned key range was moved by more
// than one cell layer in a step (or teleported): the decomposition cannot follow it, which is reported, not hidden.
__device__ __forceinline__ uint32_t slab_local_key(const SimParams& par, float4 p, long long keyOffset, int numCellsLocal,
                                                   bool mustBeOwned, uint32_t ownedLo, uint32_t ownedHi, uint32_t* __restrict__ st)
{
    const long long k = (long long)cell_hash(par, make_float3(p.x, p.y, p.z)) - keyOffset;
    const bool inTable = k >= 0 && k < (long long)numCellsLocal;
    if (mustBeOwned && !(inTable && (uint32_t)k >= ownedLo && (uint32_t)k < ownedHi)) atomicOr(&st[SD_LOST], 1u);
    return inTable ? (uint32_t)k : (uint32_t)numCellsLocal;
}

// Phase B, over the whole work set while exchange 1 is in flight: integrate the owned particles between the two boundary
// layers (split != 0), retire every slot that is not owned (last step's ghosts and dead slots), and hash + count all slots.
__global__ void __launch_bounds__(256)
k_slab_interior_hist(const __grid_constant__ SimParams par, const BoundaryCtx ctx, float4* __restrict__ pos, float4* __restrict__ vel,
                     uint32_t* __restrict__ idx, uint32_t* __restrict__ keyU, uint32_t* __restrict__ rankU,
                     uint32_t* __restrict__ cellCount, uint32_t* __restrict__ st, long long keyOffset, int numCellsLocal,
                     uint32_t ownedLo, uint32_t ownedHi, uint32_t* __restrict__ keyMaxSlots, int split)
{
    __shared__ uint32_t blockMax;
    if (threadIdx.x == 0) blockMax = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // the work set of this step starts as [0, end of the owned range): the slots behind it (last step's ghosts above and
    // retired slots, all sorted to the back) are dropped, so the work set does not grow from step to step
    const uint32_t work = st[SD_END];
    if (i == 0) { st[SD_WORK0] = work;  st[SD_WORK] = work; }       // the unpack kernel appends behind this
    uint32_t live1 = 0;
    bool isDummy = false;
    if (i < work) {
        const uint32_t g0 = st[SD_FIRST], g1 = st[SD_END], bLo = st[SD_BLO2], bHi = st[SD_BHI2];
        const bool owned = i >= g0 && i < g1;
        uint32_t id = idx[i];
        if (!owned && id != kDeadIndex) { idx[i] = kDeadIndex;  id = kDeadIndex; }
        uint32_t key = (uint32_t)numCellsLocal;
        if (id != kDeadIndex) {
            float4 p4 = pos[i];
            if (split && i >= bLo && i < bHi) {
                const float4 v4 = vel[i];
                float3 p = make_float3(p4.x, p4.y, p4.z), v = make_float3(v4.x, v4.y, v4.z);
                integrate_particle(par, ctx, p, v);
                p4 = make_float4(p.x, p.y, p.z, p4.w);
                pos[i] = p4;
                vel[i] = make_float4(v.x, v.y, v.z, v4.w);
            }
            key = slab_local_key(par, p4, keyOffset, numCellsLocal, true, ownedLo, ownedHi, st);
            if (key != (uint32_t)numCellsLocal) live1 = key + 1;
        }
        keyU[i] = key;
        isDummy = key == (uint32_t)numCellsLocal;
        if (!isDummy) rankU[i] = atomicAdd(&cellCount[key], 1u);
    }
    // retired slots (last step's ghosts, leavers) all count into the ONE dummy cell: one atomic per warp, not per slot --
    // same-address atomics serialise in L2, and there is a boundary layer's worth of them every step
    const uint32_t slot = warp_append_slot(&cellCount[numCellsLocal], isDummy);
    if (isDummy) rankU[i] = slot;
    live1 = __reduce_max_sync(0xffffffffu, live1);
    if ((threadIdx.x & 31) == 0 && live1) atomicMax(&blockMax, live1);
    __syncthreads();
    if (threadIdx.x == 0 && blockMax) atomicMax(&keyMaxSlots[blockIdx.x & (kKeyMaxSlots - 1)], blockMax);
}

// k_slab_unpack with the work-set size on the device and the hash + count of every appended record fused in
__global__ void __launch_bounds__(256)
k_slab_unpack_hist(const __grid_constant__ SimParams par, const SlabRecord* __restrict__ inBelow, const SlabRecord* __restrict__ inAbove,
                   const SlabRecord* __restrict__ ownDown, const SlabRecord* __restrict__ ownUp, int capL, int capB,
                   float4* __restrict__ pos, float4* __restrict__ vel, uint32_t* __restrict__ idx, int capacity,
                   uint32_t* __restrict__ keyU, uint32_t* __restrict__ rankU, uint32_t* __restrict__ cellCount,
                   uint32_t* __restrict__ st, long long keyOffset, int numCellsLocal, uint32_t ownedLo, uint32_t ownedHi,
                   uint32_t* __restrict__ keyMaxSlots)
{
    const SlabRecord* msg[6] = {inBelow, inAbove, inBelow, inAbove, ownDown, ownUp};
    const int word[6] = {0, 0, 1, 1, 0, 0};
    const int row0[6] = {1, 1, 1 + capL, 1 + capL, 1, 1};
    const int cap[6]  = {capL, capL, capB, capB, capL, capL};
    const int sec = blockIdx.y;
    const uint32_t work0 = st[SD_WORK0];
    uint32_t cnt[6], off = 0, total = 0, over = 0;
    #pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t c = msg[k] ? reinterpret_cast<const uint32_t*>(msg[k])[word[k]] : 0u;
        if (c > (uint32_t)cap[k]) { over = 1;  c = (uint32_t)cap[k]; }
        cnt[k] = c;
        if (k < sec) off += c;
        total += c;
    }
    if ((long long)work0 + total > capacity) over = 1;
    if (sec == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        st[SD_WORK] = (uint32_t)min((long long)work0 + total, (long long)capacity);
        if (over) st[SD_OVERFLOW] = 1u;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mine = 0;
    #pragma unroll
    for (int k = 0; k < 6; k++) if (k == sec) mine = cnt[k];
    uint32_t live1 = 0;
    const long long dst = (long long)work0 + off + i;
    if (i < (int)mine && dst < capacity) {
        const SlabRecord* src = nullptr;  int r0 = 0;
        #pragma unroll
        for (int k = 0; k < 6; k++) if (k == sec) { src = msg[k];  r0 = row0[k]; }
        const SlabRecord r = src[r0 + i];
        pos[dst] = r.pos;  vel[dst] = r.vel;  idx[dst] = r.meta.x;
        // sections 0 and 1 are arrivals: they are owned here from now on
        const uint32_t key = slab_local_key(par, r.pos, keyOffset, numCellsLocal, sec < 2, ownedLo, ownedHi, st);
        if (key != (uint32_t)numCellsLocal) live1 = key + 1;
        keyU[dst] = key;
        rankU[dst] = atomicAdd(&cellCount[key], 1u);
    }
    live1 = __reduce_max_sync(0xffffffffu, live1);
    if ((threadIdx.x & 31) == 0 && live1) atomicMax(&keyMaxSlots[(blockIdx.x + 7 * blockIdx.y) & (kKeyMaxSlots - 1)], live1);
}

// After the scan: the sorted ranges (ghosts below | owned | ghosts above) and the two boundary layers, from five cell-table
// entries.  Entries at or above the scan bound were not written this step: no live key is that large, so they equal the
// live total, which is the start of the dummy cell (always scanned).
__global__ void k_slab_bounds(const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ scanBound, uint32_t* __restrict__ st,
                              int c0, int c1, int c2, int c3, int c4, int c5, int c6, int hasLower, int hasUpper)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int cells[7] = {c0, c1, c2, c3, c4, c5, c6};
    const int bound = (int)*scanBound, lastTileStart = c2 / SPH_SCAN_TILE * SPH_SCAN_TILE;
    uint32_t v[7];
    const uint32_t total = cellStart[c2];
    for (int k = 0; k < 7; k++) v[k] = (cells[k] >= bound && cells[k] < lastTileStart) ? total : cellStart[cells[k]];
    st[SD_FIRST] = v[0];  st[SD_END] = v[1];  st[SD_G2] = v[2];
    st[SD_BLO] = hasLower ? v[3] : v[0];        // no neighbour on a side: no boundary layer there
    st[SD_BHI] = hasUpper ? v[4] : v[1];
    st[SD_BLO2] = hasLower ? v[5] : v[0];       // cells c5 / c6: two layers in from each end (the host clamps them to
    st[SD_BHI2] = hasUpper ? max(v[6], st[SD_BLO2]) : v[1];      // the owned range when the slab is thinner than that)
}

// rho,p rows of the two boundary layers: [0] = {count, 0, 0, 0}, then (x,y,z,p), (vx,vy,vz,rho) per particle
__global__ void __launch_bounds__(256)
k_slab_pack_dp(const float4* __restrict__ posP, const float4* __restrict__ velD, uint32_t* __restrict__ st,
               float4* __restrict__ dpDown, float4* __restrict__ dpUp, int capRows)
{
    const uint32_t g0 = st[SD_FIRST], g1 = st[SD_END], bLo = st[SD_BLO], bHi = st[SD_BHI];
    const uint32_t nDown = bLo - g0, nUp = g1 - bHi;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        dpDown[0] = make_float4(__uint_as_float(min(nDown, (uint32_t)capRows)), 0.f, 0.f, 0.f);
        dpUp[0] = make_float4(__uint_as_float(min(nUp, (uint32_t)capRows)), 0.f, 0.f, 0.f);
        if (nDown > (uint32_t)capRows || nUp > (uint32_t)capRows) st[SD_OVERFLOW] = 1u;
    }
    if (blockIdx.y == 0) { if (t < nDown && t < (uint32_t)capRows) { dpDown[1 + 2 * t] = posP[g0 + t];  dpDown[2 + 2 * t] = velD[g0 + t]; } }
    else                 { if (t < nUp && t < (uint32_t)capRows)   { dpUp[1 + 2 * t] = posP[bHi + t];   dpUp[2 + 2 * t] = velD[bHi + t]; } }
}

// the neighbours' rows become the (x,y,z,p) / (v,rho) of this rank's ghosts, which sort in the same order on both sides
__global__ void __launch_bounds__(256)
k_slab_unpack_dp(const float4* __restrict__ dpBelow, const float4* __restrict__ dpAbove, uint32_t* __restrict__ st,
                 float4* __restrict__ posP, float4* __restrict__ velD)
{
    const uint32_t g0 = st[SD_FIRST], g1 = st[SD_END], g2 = st[SD_G2];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.y == 0) {
        const uint32_t n = dpBelow ? __float_as_uint(dpBelow[0].x) : 0u;
        if (t == 0 && n != g0) st[SD_DPERR] = 1u;               // ghost sets out of step between the ranks
        if (t < n && t < g0) { posP[t] = dpBelow[1 + 2 * t];  velD[t] = dpBelow[2 + 2 * t]; }
    } else {
        const uint32_t n = dpAbove ? __float_as_uint(dpAbove[0].x) : 0u;
        if (t == 0 && n != g2 - g1) st[SD_DPERR] = 1u;
        if (t < n && t < g2 - g1) { posP[g1 + t] = dpAbove[1 + 2 * t];  velD[g1 + t] = dpAbove[2 + 2 * t]; }
    }
}

inline int blocks_for(int n, int t) { return (n + t - 1) / t; }

}  // namespace

#define SPH_COUNT(L) do { if ((L).launches) ++*(L).launches; } while (0)

void sph_launch_integrate_hash(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel,
                               uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, int first, int count)
{
    if (count <= 0) return;
    const BoundaryCtx ctx = boundary_ctx(par);
    k_integrate_hash<<<blocks_for(count, 256), 256, 0, L.stream>>>(par, ctx, pos, vel, keyU, rankU, cellCount, first, first + count);
    SPH_COUNT(L);
}

void sph_launch_slab_integrate_pack(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                    int first, int count, int work, int zLo, int zHi, int hasLower, int hasUpper,
                                    void* leavDown, void* leavUp, int capL, void* bndDown, void* bndUp, int capB,
                                    uint32_t* headDown, uint32_t* headUp)
{
    if (work <= 0) return;
    const BoundaryCtx ctx = boundary_ctx(par);
    k_slab_integrate_pack<<<blocks_for(work, 256), 256, 0, L.stream>>>(par, ctx, pos, vel, idx, first, count, work, zLo, zHi,
                                                                        hasLower, hasUpper, (SlabRecord*)leavDown, (SlabRecord*)leavUp, capL,
                                                                        (SlabRecord*)bndDown, (SlabRecord*)bndUp, capB, headDown, headUp);
    SPH_COUNT(L);
}

void sph_launch_scan(const SphLaunch& L, uint32_t* cellCount, uint32_t* cellStart, uint32_t* tileSums,
                     uint32_t* maxCount, int numCells, int maxCells, const uint32_t* boundCells)
{
    int tiles = blocks_for(numCells, SPH_SCAN_TILE);
    k_scan_reduce<<<tiles, 256, 0, L.stream>>>(cellCount, tileSums, numCells, boundCells);   SPH_COUNT(L);
    k_scan_tiles<<<1, 256, 0, L.stream>>>(tileSums, tiles, maxCount);                   SPH_COUNT(L);
    k_scan_apply<<<tiles, 256, 0, L.stream>>>(cellCount, cellStart, tileSums, maxCount, numCells, maxCells, boundCells);  SPH_COUNT(L);
}

void sph_launch_bucket(const SphLaunch& L, const uint32_t* keyU, const uint32_t* rankU, const uint32_t* idxIn,
                       const uint32_t* cellStart, uint2* pairT, int n, const uint32_t* nDev)
{
    k_bucket<<<blocks_for(n, 256), 256, 0, L.stream>>>(keyU, rankU, idxIn, cellStart, pairT, n, nDev);  SPH_COUNT(L);
}

void sph_launch_rank_gather(const SphLaunch& L, const uint2* pairT, const uint32_t* keyU, const uint32_t* cellStart,
                            const float4* posIn, const float4* velIn,
                            float4* posOut, float4* velOut, uint32_t* idxOut, uint32_t* keyS, int n, const uint32_t* nDev,
                            const uint32_t* big, int realCells)
{
    k_rank_gather<<<blocks_for(n, 256), 256, 0, L.stream>>>(pairT, keyU, cellStart, posIn, velIn, posOut, velOut, idxOut, keyS, n, nDev,
                                                            big, realCells);
    SPH_COUNT(L);
    if (big) {
        k_rank_big_cells<<<296, 256, 0, L.stream>>>(pairT, cellStart, posIn, velIn, posOut, velOut, idxOut, keyS, big);
        SPH_COUNT(L);
    }
}

void sph_launch_iota(const SphLaunch& L, uint32_t* idx, int n)
{
    k_iota<<<blocks_for(n, 256), 256, 0, L.stream>>>(idx, n);  SPH_COUNT(L);
}

void sph_launch_unpermute4(const SphLaunch& L, const float4* src, const uint32_t* idx, float4* out, int start, int count, int n)
{
    k_unpermute4<<<blocks_for(n, 256), 256, 0, L.stream>>>(src, idx, out, start, count, n);  SPH_COUNT(L);
}

void sph_launch_unpermute_w(const SphLaunch& L, const float4* src, const uint32_t* idx, float* out, int start, int count, int n)
{
    k_unpermute_w<<<blocks_for(n, 256), 256, 0, L.stream>>>(src, idx, out, start, count, n);  SPH_COUNT(L);
}

void sph_launch_permute4(const SphLaunch& L, float4* dst, const uint32_t* idx, const float4* in, int start, int count, int n)
{
    k_permute4<<<blocks_for(n, 256), 256, 0, L.stream>>>(dst, idx, in, start, count, n);  SPH_COUNT(L);
}

void sph_launch_cell_table_dump(const SphLaunch& L, const uint32_t* cellStart, uint32_t* outStart, uint32_t* outEnd, int numCells)
{
    k_cell_table_dump<<<blocks_for(numCells, 256), 256, 0, L.stream>>>(cellStart, outStart, outEnd, numCells);  SPH_COUNT(L);
}

void sph_launch_pack_pairs(const SphLaunch& L, const uint32_t* keyS, const uint32_t* idx, uint2* out, int n)
{
    k_pack_pairs<<<blocks_for(n, 256), 256, 0, L.stream>>>(keyS, idx, out, n);  SPH_COUNT(L);
}

// ---- slab mode --------------------------------------------------------------------------------
void sph_launch_slab_take_leavers(const SphLaunch& L, const SimParams& par, const float4* pos, const float4* vel, uint32_t* idx,
                                  int first, int count, int zLo, int zHi, int hasLower, int hasUpper,
                                  void* down, int capDown, void* up, int capUp, uint32_t* ctrDown, uint32_t* ctrUp)
{
    if (count <= 0) return;
    k_slab_take_leavers<<<blocks_for(count, 256), 256, 0, L.stream>>>(par, pos, vel, idx, first, first + count, zLo, zHi, hasLower, hasUpper,
                                                                      (SlabRecord*)down, capDown, (SlabRecord*)up, capUp, ctrDown, ctrUp);
    SPH_COUNT(L);
}

void sph_launch_slab_boundary(const SphLaunch& L, const SimParams& par, const float4* pos, const float4* vel, const uint32_t* idx,
                              int n, int zLo, int zHi, int hasLower, int hasUpper,
                              void* down, int capDown, void* up, int capUp, uint32_t* ctrDown, uint32_t* ctrUp)
{
    if (n <= 0) return;
    k_slab_boundary<<<blocks_for(n, 256), 256, 0, L.stream>>>(par, pos, vel, idx, n, zLo, zHi, hasLower, hasUpper,
                                                              (SlabRecord*)down, capDown, (SlabRecord*)up, capUp, ctrDown, ctrUp);
    SPH_COUNT(L);
}

void sph_launch_slab_append(const SphLaunch& L, const void* recs, int count, float4* pos, float4* vel, uint32_t* idx, int at)
{
    if (count <= 0) return;
    k_slab_append<<<blocks_for(count, 256), 256, 0, L.stream>>>((const SlabRecord*)recs, count, pos, vel, idx, at);
    SPH_COUNT(L);
}

void sph_launch_slab_export(const SphLaunch& L, const float4* pos, const float4* vel, const uint32_t* idx,
                            const float4* posP, const float4* velD, int first, int count, void* recs)
{
    if (count <= 0) return;
    k_slab_export<<<blocks_for(count, 256), 256, 0, L.stream>>>(pos, vel, idx, posP, velD, first, count, (SlabRecord*)recs);
    SPH_COUNT(L);
}

void sph_launch_fill_u32(const SphLaunch& L, uint32_t* p, uint32_t v, int first, int count)
{
    if (count <= 0) return;
    k_fill_u32<<<blocks_for(count, 256), 256, 0, L.stream>>>(p, v, first, first + count);
    SPH_COUNT(L);
}

void sph_launch_slab_hash_hist(const SphLaunch& L, const SimParams& par, const float4* pos, const uint32_t* idx,
                               uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, int nMax, const uint32_t* nDev,
                               long long keyOffset, int numCellsLocal, uint32_t* keyMaxSlots, uint32_t guardCells)
{
    if (nMax > 0) {
        k_slab_hash_hist<<<blocks_for(nMax, 256), 256, 0, L.stream>>>(par, pos, idx, keyU, rankU, cellCount, nDev, keyOffset,
                                                                      numCellsLocal, keyMaxSlots);
        SPH_COUNT(L);
    }
    k_slab_scan_bound<<<1, 32, 0, L.stream>>>(keyMaxSlots, guardCells, (uint32_t)numCellsLocal);
    SPH_COUNT(L);
}

void sph_launch_slab_unpack(const SphLaunch& L, const void* inBelow, const void* inAbove, const void* ownDown, const void* ownUp,
                            int capL, int capB, float4* pos, float4* vel, uint32_t* idx, int work0, int capacity, uint32_t* dev)
{
    int m = capL > capB ? capL : capB;
    dim3 grid(blocks_for(m, 256), 6);
    k_slab_unpack<<<grid, 256, 0, L.stream>>>((const SlabRecord*)inBelow, (const SlabRecord*)inAbove, (const SlabRecord*)ownDown,
                                              (const SlabRecord*)ownUp, capL, capB, pos, vel, idx, work0, capacity, dev);
    SPH_COUNT(L);
}

// ---- slab mode, device-resident bookkeeping -------------------------------------------------------
void sph_launch_slab_boundary_integrate_pack(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                             uint32_t* st, int bound, int zLo, int zHi, int hasLower, int hasUpper,
                                             void* leavDown, void* leavUp, int capL, void* bndDown, void* bndUp, int capB,
                                             uint32_t* headDown, uint32_t* headUp, void* peerDown, void* peerUp)
{
    const BoundaryCtx ctx = boundary_ctx(par);
    k_slab_boundary_integrate_pack<<<blocks_for(bound, 256), 256, 0, L.stream>>>(par, ctx, pos, vel, idx, st, zLo, zHi, hasLower, hasUpper,
                                                                                  (SlabRecord*)leavDown, (SlabRecord*)leavUp, capL,
                                                                                  (SlabRecord*)bndDown, (SlabRecord*)bndUp, capB, headDown, headUp,
                                                                                  (SlabRecord*)peerDown, (SlabRecord*)peerUp);
    SPH_COUNT(L);
    if (peerDown || peerUp) {
        k_slab_publish_headers<<<1, 32, 0, L.stream>>>(headDown, headUp, (uint32_t*)peerDown, (uint32_t*)peerUp);
        SPH_COUNT(L);
    }
}

void sph_launch_slab_interior_hist(const SphLaunch& L, const SimParams& par, float4* pos, float4* vel, uint32_t* idx,
                                   uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, uint32_t* st, int bound,
                                   long long keyOffset, int numCellsLocal, uint32_t ownedLo, uint32_t ownedHi,
                                   uint32_t* keyMaxSlots, int split)
{
    const BoundaryCtx ctx = boundary_ctx(par);
    k_slab_interior_hist<<<blocks_for(bound, 256), 256, 0, L.stream>>>(par, ctx, pos, vel, idx, keyU, rankU, cellCount, st, keyOffset,
                                                                      numCellsLocal, ownedLo, ownedHi, keyMaxSlots, split);
    SPH_COUNT(L);
}

void sph_launch_slab_unpack_hist(const SphLaunch& L, const SimParams& par, const void* inBelow, const void* inAbove,
                                 const void* ownDown, const void* ownUp, int capL, int capB, float4* pos, float4* vel, uint32_t* idx,
                                 int capacity, uint32_t* keyU, uint32_t* rankU, uint32_t* cellCount, uint32_t* st,
                                 long long keyOffset, int numCellsLocal, uint32_t ownedLo, uint32_t ownedHi, uint32_t* keyMaxSlots)
{
    const int m = capL > capB ? capL : capB;
    dim3 grid(blocks_for(m, 256), 6);
    k_slab_unpack_hist<<<grid, 256, 0, L.stream>>>(par, (const SlabRecord*)inBelow, (const SlabRecord*)inAbove, (const SlabRecord*)ownDown,
                                                   (const SlabRecord*)ownUp, capL, capB, pos, vel, idx, capacity, keyU, rankU, cellCount,
                                                   st, keyOffset, numCellsLocal, ownedLo, ownedHi, keyMaxSlots);
    SPH_COUNT(L);
}

void sph_launch_slab_scan_bound(const SphLaunch& L, uint32_t* keyMaxSlots, uint32_t guardCells, int numCellsLocal)
{
    k_slab_scan_bound<<<1, 32, 0, L.stream>>>(keyMaxSlots, guardCells, (uint32_t)numCellsLocal);
    SPH_COUNT(L);
}

void sph_launch_slab_bounds(const SphLaunch& L, const uint32_t* cellStart, const uint32_t* scanBound, uint32_t* st,
                            const int cells[7], int hasLower, int hasUpper)
{
    k_slab_bounds<<<1, 32, 0, L.stream>>>(cellStart, scanBound, st, cells[0], cells[1], cells[2], cells[3], cells[4], cells[5], cells[6],
                                          hasLower, hasUpper);
    SPH_COUNT(L);
}

void sph_launch_slab_pack_dp(const SphLaunch& L, const float4* posP, const float4* velD, uint32_t* st, float4* dpDown, float4* dpUp, int capRows)
{
    dim3 grid(blocks_for(capRows, 256), 2);
    k_slab_pack_dp<<<grid, 256, 0, L.stream>>>(posP, velD, st, dpDown, dpUp, capRows);
    SPH_COUNT(L);
}

void sph_launch_slab_unpack_dp(const SphLaunch& L, const float4* dpBelow, const float4* dpAbove, uint32_t* st, float4* posP, float4* velD, int capRows)
{
    dim3 grid(blocks_for(capRows, 256), 2);
    k_slab_unpack_dp<<<grid, 256, 0, L.stream>>>(dpBelow, dpAbove, st, posP, velD);
    SPH_COUNT(L);
}
