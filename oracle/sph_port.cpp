// TEST INFRASTRUCTURE ONLY -- builds into oracle/libsphport.so.  Never linked into the product.
//
// Plain scalar C++ restatement of the reference's SPH solver step, one function per reference
// kernel, written to be read next to the reference.  Every function cites the lines it follows.
// Parity status: PINNED -- tests/test_oracle_golden.py checks this port bit-for-bit against
// vectors produced by the reference's own kernel text compiled for the host (oracle/_ref, see
// build_ref.sh and tests/golden/make_golden.py); where /root/reference exists the two libraries
// are also compared live on whole scenes.
//
// Arithmetic conventions that matter, all visible below:
//   * compiled with -ffp-contract=off: no fused multiply-add, same as the reference text under g++;
//   * float3 / float3 is a true division, float3 / float multiplies by the reciprocal
//     (source/external/cutil_math.h:354-362);
//   * pow(c, 3) on a float and an int promotes to double: the density term is accumulated as
//     float(double(dens) + pow(double(c), 3.0))   (Kernel_Cell.cui:168 under C++11 <cmath>);
//   * `x * 0.01` and `x * 0.5` with double literals are double products (System.cu:154,487);
//   * out-of-range FETCH returns 0 as tex1Dfetch does (needed for the unclamped neighbour cells).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>
#include <omp.h>
#include "sph_params.h"
#include "oracle_api.h"

namespace {

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline void operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
inline void operator*=(V3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
inline V3 div_scalar(V3 a, float s) { float inv = 1.0f / s; return a * inv; }        // cutil_math.h:358-362
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline float length2d(float a, float b) { return sqrtf(a * a + b * b); }               // length(make_float2(a,b))
inline V3 from3(float3 f) { return v3(f.x, f.y, f.z); }

SimParams par;           // the reference's __constant__ par (System.cu:33)
int g_threads = 0;

// ---- boundary(): System.cu:41-159 ----------------------------------------------------------------
const float EPS = 0.00001f;

inline void push(V3& vel, V3 norm, float diff, float stiff, float damp)
{   // addB()/addC(): System.cu:52-53
    float accBnd = stiff * diff - damp * dot(norm, vel);
    vel += accBnd * norm * par.timeStep;
}

void boundary(V3& pos, V3& vel)
{
    V3 wmin = from3(par.worldMin), wmax = from3(par.worldMax);
    float b = par.distBndSoft, stiff = par.bndStiff, damp = par.bndDamp, damp2 = par.bndDampC;
    float diff;
    BndType t = par.bndType;
    bool bCylY = t == BND_CYL_Y, bCylZ = t == BND_CYL_Z,
         bWave = par.bndEffZ == BND_EFF_WAVE, bNoEff = par.bndEffZ == BND_EFF_NONE, bCycle = par.bndEffZ == BND_EFF_CYCLE;

    if (bWave) {                                                                        // :55-61
        float sl = -par.r2Angle;
        diff = b - (pos.y - wmin.y) - (pos.z - wmin.z) * sl;
        if (diff > EPS) push(vel, v3(0, 1 - sl, sl), diff, stiff, damp);
        wmin.z += par.rTwist * (1.f + sinf(par.rAngle));
    }
    if (t != BND_SPHERE) {                                                              // :64-77
        if (!bCylY) {
            if (bNoEff || bWave) { diff = b - pos.z + wmin.z;  if (diff > EPS) push(vel, v3(0, 0, 1), diff, stiff, damp2); }
            if (!bCycle)         { diff = b + pos.z - wmax.z;  if (diff > EPS) push(vel, v3(0, 0, -1), diff, stiff, damp2); }
        }
        if (!bCylY && !bCylZ) {
            diff = b - pos.x + wmin.x;  if (diff > EPS) push(vel, v3(1, 0, 0), diff, stiff, damp);
            diff = b + pos.x - wmax.x;  if (diff > EPS) push(vel, v3(-1, 0, 0), diff, stiff, damp);
        }
        if (!bCylZ) {
            diff = b - pos.y + wmin.y;  if (diff > EPS) push(vel, v3(0, 1, 0), diff, stiff, damp);
            diff = b + pos.y - wmax.y;  if (diff > EPS) push(vel, v3(0, -1, 0), diff, stiff, damp);
        }
    } else {                                                                            // :78-81
        float len = length(pos);  diff = b + len + wmin.y;
        if (diff > EPS) push(vel, v3(-pos.x / len, -pos.y / len, -pos.z / len), diff, stiff, damp2);
    }
    if (bCylY || t == BND_CYL_YZ) {                                                     // :84-86
        float len = length2d(pos.x, pos.z);  diff = b + len - wmax.x;
        if (diff > EPS) push(vel, v3(-pos.x / len, 0, -pos.z / len), diff, stiff, damp2);
    }
    if (bCylZ || t == BND_CYL_YZ) {                                                     // :89-91
        float len = length2d(pos.x, pos.y);  diff = b + len + wmin.y;
        if (diff > EPS) push(vel, v3(-pos.x / len, -pos.y / len, 0), diff, stiff, damp2);
    }
    if (!bWave && !bNoEff) {                                                            // :94-97
        float dr = 1.f * par.particleR;
        if (bCycle && vel.z > par.rVexit && pos.z > wmax.z - b - dr) { pos.z -= wmax.z - wmin.z - 2 * b - dr; }
        else if (vel.z < -par.rVexit && pos.z < wmin.z + b + dr)     { pos.z += wmax.z - wmin.z - 2 * b - dr; }
    }
    if (t == BND_PUMP_Y) {                                                              // :101-158
        float rad = wmax.x, ang = par.angOut, hc = par.hClose, rin = rad * par.radIn;
        float len = length2d(pos.x, pos.y);  diff = b + len - rad;
        if (diff > EPS) {
            float a = atanf(pos.x / pos.y);
            if (ang < 0.5f) { if (a < -ang || a > ang || pos.y < 0) push(vel, v3(-pos.x / len, -pos.y / len, 0), diff, stiff, damp); }
            else { if (pos.y < 0 || (len < rad * par.s5 && a < ang)) push(vel, v3(-pos.x / len, -pos.y / len, 0), diff, stiff, damp); }
        }
        float xs;
        if (ang < 0.5f) {
            xs = sinf(ang * par.s3) * rad;
            float zs = cosf(ang * par.s3) * rad * par.s4;
            if (pos.y > zs) {
                diff = b - pos.x - xs;  if (diff > EPS) push(vel, v3(1, 0, 0), diff, stiff, damp);
                diff = b + pos.x - xs;  if (diff > EPS) push(vel, v3(-1, 0, 0), diff, stiff, damp);
            }
        } else {
            xs = 0.09f * par.s4;
            if (len >= rad * par.s6) { diff = b - pos.x + xs;  if (diff > EPS) push(vel, v3(1, 0, 0), diff, stiff, damp); }
        }
        if (pos.z > hc - b * par.s1) { diff = b + len - rin;  if (diff > EPS) push(vel, v3(-pos.x / len, -pos.y / len, 0), diff, stiff, damp); }
        if (pos.z < hc - b * par.s2) { diff = b + pos.z - hc; if (diff > EPS) push(vel, v3(0, 0, -1), diff, stiff, damp); }
        diff = pos.y - wmax.y + par.rDexit;
        if (diff > EPS && vel.y > par.rVexit) {
            float aa, rr;
            if (ang < 0.5f) { float xx = xs * 2, zz = fabsf(hc - wmin.z);  rr = (pos.x + xx / 2) / xx * 0.7f;  aa = (pos.z - zz / 2) / zz * 1.6f; }
            else            { float zz = fabsf(hc - wmin.z);               rr = (wmax.x - pos.x) / xs * 0.45f; aa = (pos.z - zz / 2) / zz * 1.8f; }
            rr *= rin;  aa *= PI2;
            float x = cosf(aa) * rr, y = sinf(aa) * rr;
            float z = (float)((double)(wmax.z - b) - (double)fabsf(vel.y - par.rVexit) * 0.01);
            pos = v3(x, y, z);
            vel = v3(vel.x, vel.z, -vel.y);
        }
    }
}

// ---- integrateD: System.cu:165-205 ----------------------------------------------------------------
void integrate_one(const float* oldPos, const float* oldVel, float* newPos, float* newVel, int i)
{
    V3 pos = v3(oldPos[4 * i], oldPos[4 * i + 1], oldPos[4 * i + 2]);
    V3 vel = v3(oldVel[4 * i], oldVel[4 * i + 1], oldVel[4 * i + 2]);
    boundary(pos, vel);
    vel += from3(par.gravity) * par.timeStep;
    vel *= par.globalDamping;
    pos += vel * par.timeStep;
    float b = par.distBndHard;
    V3 wmin = from3(par.worldMin), wmax = from3(par.worldMax);
    if (pos.x > wmax.x - b) pos.x = wmax.x - b;
    if (pos.x < wmin.x + b) pos.x = wmin.x + b;
    if (pos.y > wmax.y - b) pos.y = wmax.y - b;
    if (pos.y < wmin.y + b) pos.y = wmin.y + b;
    if (pos.z > wmax.z - b) pos.z = wmax.z - b;
    if (pos.z < wmin.z + b) pos.z = wmin.z + b;
    newPos[4 * i] = pos.x;  newPos[4 * i + 1] = pos.y;  newPos[4 * i + 2] = pos.z;  newPos[4 * i + 3] = oldPos[4 * i + 3];
    newVel[4 * i] = vel.x;  newVel[4 * i + 1] = vel.y;  newVel[4 * i + 2] = vel.z;  newVel[4 * i + 3] = oldVel[4 * i + 3];
}

// ---- calcGridPos / calcGridHash: Kernel_Cell.cui:5-19 -----------------------------------------------
struct I3 { int x, y, z; };
inline I3 grid_pos(float px, float py, float pz)
{
    I3 g;
    g.x = (int)floorf((px - par.worldMin.x) / par.cellSize.x);
    g.y = (int)floorf((py - par.worldMin.y) / par.cellSize.y);
    g.z = (int)floorf((pz - par.worldMin.z) / par.cellSize.z);
    return g;
}
inline uint32_t grid_hash(I3 g)
{
    return (uint32_t)(g.z * (int)par.gridSize_yx + g.y * (int)par.gridSize.x + g.x);
}

// ---- FETCH with tex1Dfetch's out-of-range behaviour (index is an int; outside -> 0) -----------------
struct Tables {
    const float* pos;  const float* vel;  const float* pressure;  const float* density;
    const uint32_t* pairs;  const uint32_t* cellStart;  const float* dye;
    long long n, numCells;
};
inline bool in_range(uint32_t idx, long long n) { long long i = (int)idx; return i >= 0 && i < n; }
inline uint32_t fetch_cell_start(const Tables& T, uint32_t h) { return in_range(h, T.numCells) ? T.cellStart[(int)h] : 0u; }
inline uint32_t fetch_hash(const Tables& T, uint32_t i) { return in_range(i, T.n) ? T.pairs[2 * (size_t)(int)i] : 0u; }

// ---- compDensCell / computeDensityD: Kernel_Cell.cui:142-199 -----------------------------------------
template <bool COUNT>
float dens_cell(const Tables& T, I3 g, uint32_t index, const float* pos, uint32_t* cnt)
{
    float dens = 0.0f;
    uint32_t gridHash = grid_hash(g);
    uint32_t bucketStart = fetch_cell_start(T, gridHash);
    if (bucketStart == 0xffffffffu) return dens;
    for (uint32_t i = 0; i < par.maxParInCell; i++) {
        uint32_t index2 = bucketStart + i;
        if (fetch_hash(T, index2) != gridHash) break;
        if (index2 != index) {
            const float* p2 = T.pos + 4 * (size_t)index2;
            float px = pos[0] - p2[0], py = pos[1] - p2[1], pz = pos[2] - p2[2];
            float r2 = px * px + py * py + pz * pz;
            if (r2 < par.h2) {
                float c = par.h2 - r2;
                dens = (float)((double)dens + pow((double)c, 3.0));     // dens += pow(c, 3)
                if (COUNT) (*cnt)++;
            }
        }
    }
    return dens;
}

template <bool COUNT>
void density_one(const Tables& T, float* pressure, float* density, uint32_t* counts, uint32_t index)
{
    const float* pos = T.pos + 4 * (size_t)index;
    I3 g = grid_pos(pos[0], pos[1], pos[2]);
    float sum = 0.0f;  uint32_t cnt = 0;
    for (int z = -1; z <= 1; z++)
        for (int y = -1; y <= 1; y++)
            for (int x = -1; x <= 1; x++) {
                I3 c = {g.x + x, g.y + y, g.z + z};
                sum += dens_cell<COUNT>(T, c, index, pos, &cnt);
            }
    if (COUNT) { counts[index] = cnt; return; }
    float dens = sum * par.Poly6Kern * par.particleMass;
    float pres = (dens - par.restDensity) * par.stiffness;
    pressure[index] = pres;
    density[index] = dens;
}

// ---- compForcePair / compForceCell: Kernel_Cell.cui:210-261 -------------------------------------------
inline V3 force_pair(V3 relPos, V3 relVel, float p1_add_p2, float d1_mul_d2)
{
    float r = std::max(par.minDist, length(relPos));
    V3 fcur = v3(0, 0, 0);
    if (r < par.h) {
        float c = par.h - r;
        float pterm = c * par.SpikyKern * p1_add_p2 / r;
        float vterm = par.LapKern * par.viscosity;
        fcur = pterm * relPos + vterm * relVel;
        fcur *= c * d1_mul_d2;
    }
    return fcur;
}

V3 force_cell(const Tables& T, I3 g, uint32_t index, const float* pos, const float* vel, float pres, float dens)
{
    V3 force = v3(0, 0, 0);
    uint32_t gridHash = grid_hash(g);
    uint32_t bucketStart = fetch_cell_start(T, gridHash);
    if (bucketStart == 0xffffffffu) return force;
    for (uint32_t i = 0; i < par.maxParInCell; i++) {
        uint32_t index2 = bucketStart + i;
        if (fetch_hash(T, index2) != gridHash) break;
        if (index2 != index) {
            const float* p2 = T.pos + 4 * (size_t)index2;
            const float* v2 = T.vel + 4 * (size_t)index2;
            float pres2 = T.pressure[index2], dens2 = T.density[index2];
            float d12 = std::min(par.minDens, 1.0f / (dens * dens2));
            force += force_pair(v3(pos[0] - p2[0], pos[1] - p2[1], pos[2] - p2[2]),
                                v3(v2[0] - vel[0], v2[1] - vel[1], v2[2] - vel[2]), pres + pres2, d12);
        }
    }
    return force;
}

// ---- collideSpheres / collideSpheresR: Kernel_Cell.cui:78-115 -----------------------------------------
V3 collide_spheres(V3 relPos, V3 relVel, float radiusAB)
{
    float dist = length(relPos);
    V3 force = v3(0, 0, 0);
    if (dist < radiusAB) {
        V3 norm = div_scalar(relPos, dist);
        V3 tanVel = relVel - (dot(relVel, norm) * norm);
        force = par.spring * (dist - radiusAB) * norm;
        force += par.damping * relVel;
        force += par.shear * tanVel;
    }
    return force;
}
V3 collide_spheres_r(V3 posAB, V3 relVel, float radiusAB)
{
    float dist = length(posAB);
    V3 force = v3(0, 0, 0);
    if (dist < radiusAB) {
        V3 norm = div_scalar(posAB, dist);
        force = par.spring * (dist - radiusAB) * norm;
        force += par.damping * relVel;
    }
    return force;
}

inline float Wfoam(float x, float h) { return x <= h ? (1 - x) / h : (float)0.0; }                 // System.cu:217-222
inline float phi_clamp(float I, float Tmin, float Tmax) { return (std::min(I, Tmax) - std::min(I, Tmin)) / (Tmax - Tmin); }

// ---- computeForceD: System.cu:228-547 -----------------------------------------------------------------
void force_one(const Tables& T, float* newVel, float* clr, float* dyeColor, uint32_t index)
{
    const float* pos = T.pos + 4 * (size_t)index;
    const float* vel = T.vel + 4 * (size_t)index;
    float pres = T.pressure[index], dens = T.density[index];
    I3 gridPos = grid_pos(pos[0], pos[1], pos[2]);
    V3 P = v3(pos[0], pos[1], pos[2]), Vv = v3(vel[0], vel[1], vel[2]);

    V3 addVel = v3(0, 0, 0);
    for (int z = -1; z <= 1; z++)
        for (int y = -1; y <= 1; y++)
            for (int x = -1; x <= 1; x++) {
                I3 c = {gridPos.x + x, gridPos.y + y, gridPos.z + z};
                addVel += force_cell(T, c, index, pos, vel, pres, dens);
            }
    uint32_t si = T.pairs[2 * (size_t)index + 1];
    addVel *= par.particleMass * par.timeStep;                                          // :250

    if (par.iHmap > 0) {                                                                // :254-309
        V3 vel3 = Vv * -1;
        float rr = par.particleR + par.rotR;
        const int hz = 2, hy = 1;
        if (par.iHmap == 1) {
            int ix = (int)((P.x - par.worldMin.x) / par.rotSpc);  float xf = ix * par.rotSpc + par.worldMin.x;
            int iz = (int)((P.z - par.worldMin.z) / par.rotSpc);  float zf = iz * par.rotSpc + par.worldMin.z;
            for (int j = -hz; j <= hz; j++)
                for (int i = -hz; i <= hz; i++) {
                    float xh = xf + i * par.rotSpc, zh = zf + j * par.rotSpc;
                    for (int k = 0; k <= hy; k++) {
                        float xn = xh / par.worldSizeD.x * PI2, zn = zh / par.worldSizeD.z * PI2;
                        float ss = sinf(par.s1 * xn + par.s2 * PI / 180.f) * sinf(par.s3 * zn + par.s4 * PI / 180.f);
                        if (par.s5 <= -1.f || ss > par.s5) {
                            float yf = ss * par.hClose + par.r2Angle + par.worldMin.y;
                            float yh = yf - k * par.rotSpc;
                            addVel += collide_spheres_r(v3(xh - P.x, yh - P.y, zh - P.z), vel3, rr);
                        }
                    }
                }
        } else {
            int iy = (int)((P.y - par.worldMin.y) / par.rotSpc);  float yf = iy * par.rotSpc + par.worldMin.y;
            int iz = (int)((P.z - par.worldMin.z) / par.rotSpc);  float zf = iz * par.rotSpc + par.worldMin.z;
            for (int j = -hz; j <= hz; j++)
                for (int i = -hz; i <= hz; i++) {
                    float yh = yf + i * par.rotSpc, zh = zf + j * par.rotSpc;
                    for (int k = 0; k <= hy; k++) {
                        float yn = yh / par.worldSizeD.y * PI2, zn = zh / par.worldSizeD.z * PI2;
                        float ss = sinf(par.s1 * yn + par.s2 * PI / 180.f) * sinf(par.s3 * zn + par.s4 * PI / 180.f);
                        if (par.s5 <= -1.f || ss > par.s5) {
                            float xf = ss * par.hClose + par.r2Angle + par.worldMin.x;
                            float xh = xf - k * par.rotSpc;
                            addVel += collide_spheres_r(v3(xh - P.x, yh - P.y, zh - P.z), vel3, rr);
                        }
                    }
                }
        }
    }

    V3 cpos = v3(par.collPos.x, par.collPos.y, par.collPos.z);
    if (par.rotType > 0) {                                                              // :312-372
        int sx = par.rotSize.x, sy = par.rotSize.y, sz = par.rotSize.z, cb = par.rotBlades;
        float r = par.rotR, sp = par.rotSpc, ca = PI2 / cb, x2 = sx * 0.5f, y2 = sy * 0.5f;
        float rr = par.particleR + r;
        V3 vel3 = Vv * -1;
        switch (par.rotType) {
        case 1:
            if (P.z > par.collPos.z - rr && P.z < par.collPos.z + rr + sp * sy)
                for (int z = 1; z <= sz; z++)
                    for (int x = 0; x <= sx; x++)
                        for (int c = 0; c < cb; c++) {
                            float a = -par.rAngle + c * ca + z * par.rTwist, cs = cosf(a) * sp, sn = -sinf(a) * sp;
                            for (int y = 0; y <= sy; y++) {
                                V3 pc = v3((x - x2) * cs - z * sn, (x - x2) * sn + z * cs, sp * y);
                                addVel += collide_spheres_r(cpos + pc - P, vel3, rr);
                            }
                        }
            break;
        case 2:
            if (P.y > par.collPos.y - rr && P.y < par.collPos.y + rr + sp * sy)
                for (int z = 1; z <= sz; z++)
                    for (int x = 0; x <= sx; x++)
                        for (int c = 0; c < cb; c++) {
                            float a = -par.rAngle + c * ca + z * par.rTwist, cs = cosf(a) * sp, sn = -sinf(a) * sp;
                            for (int y = 0; y <= sy; y++) {
                                V3 pc = v3((x - x2) * cs - z * sn, sp * y, (x - x2) * sn + z * cs);
                                addVel += collide_spheres_r(cpos + pc - P, vel3, rr);
                            }
                        }
            break;
        case 3:
            if (P.z > par.collPos.z - rr - sp * sz / 2.f && P.z < par.collPos.z + rr + sp * sz / 2.f) {
                V3 rotPos = cpos;
                float aa = par.rAngle, tw = par.rTwist;
                if (par.r2Dist > 0.f) {
                    if (P.x > 0.f) { rotPos.x += par.r2Dist * 0.5f; }
                    else { rotPos.x -= par.r2Dist * 0.5f;  aa = par.r2Angle;  tw *= par.r2twist; }
                }
                for (int c = 0; c < cb; c++) {
                    float a = aa + c * ca;
                    for (int h = 0; h <= sy; h++) {
                        float dh = 0;
                        if (h == sy - 1) dh = 0.5f; else if (h == sy) dh = 1.2f;
                        for (int x = 0; x <= sz; x++) {
                            float d = dh;
                            if (x == sz && d == 0) d = 0.4f;
                            float k = cosf((x - x2) * 0.2f + PI * 0.6f);
                            float at = tw * (d * -0.05f - (h - y2) * k);
                            float ac = a + at, cs = cosf(ac) * sp, sn = -sinf(ac) * sp;
                            V3 pc = v3(x * cs, x * sn, sp * (-d + h - y2) * k);
                            r = par.rotR * fabsf(1 - d);
                            addVel += collide_spheres_r(rotPos + pc - P, vel3, par.particleR + r);
                        }
                    }
                }
            }
            break;
        }
    } else {                                                                            // :374-375
        addVel += collide_spheres(cpos - P, Vv * -1.f, par.particleR + par.collR);
    }

    for (int i = 0; i < NumAcc; i++) {                                                  // :379-399
        const Accel& ac = par.acc[i];
        if (ac.type == ACC_Off) continue;
        V3 rel = P - from3(ac.pos);
        switch (ac.type) {
        case ACC_Box:
            if (fabsf(rel.x) < ac.size.x && fabsf(rel.y) < ac.size.y && fabsf(rel.z) < ac.size.z)
                addVel += from3(ac.acc) * par.timeStep;
            break;
        case ACC_CylY: {
            float r = length2d(rel.x / ac.size.x, rel.z / ac.size.z);
            if (fabsf(rel.y) < ac.size.y && r < 1.f) addVel += from3(ac.acc) * par.timeStep;
        } break;
        case ACC_CylYsm: {
            float r = length2d(rel.x, rel.z);
            if (fabsf(rel.y) < ac.size.y && r < ac.size.x) addVel += from3(ac.acc) * (1.f - r / ac.size.z) * par.timeStep;
        } break;
        default: break;
        }
    }

    float* nv = newVel + 4 * (size_t)si;                                                // :402
    nv[0] = vel[0] + addVel.x;  nv[1] = vel[1] + addVel.y;  nv[2] = vel[2] + addVel.z;  nv[3] = vel[3] + 0.0f;

    // ---- colouring (visual only): System.cu:406-515 ----
    V3 color = v3((float)0.2, (float)0.5, 1);
    float intens = 0.f;
    switch (par.clrType) {
    case CLR_None: return;
    case CLR_VelAcc: {
        float v = 2.5f * length(Vv);
        float f = 0.02f * length(addVel) / par.timeStep;
        float clrV = par.brightness + par.contrast * v, clrF = par.contrast * f;
        color = color * clrV + v3(0.7f, 0.35f, 0) * clrF;
    } break;
    case CLR_DensAcc: {
        float d = 4.f * (dens - par.restDensity) / par.restDensity;
        float f = 0.02f * length(addVel) / par.timeStep;
        float clrD = par.brightness + par.contrast * d, clrF = par.contrast * f;
        color = color * clrD + v3(0.7f, 0.7f, 0) * clrF;
    } break;
    case CLR_Vel: {
        float v = 2.5f * length(Vv);
        float clrV = par.brightness + par.contrast * v;  intens = clrV;
        // foam / trapped-air potential (:429-498).  NB the loop order is x,y,z outer and the slot index
        // innermost, and the normalised sums use '+' -- restated as written.
        float minvThreshold = 5.0f, maxvThreshold = 20.0f, vdiff = 0.0f, h = 3.0f;
        for (int x = -1; x < 2; x++)
            for (int y = -1; y < 2; y++)
                for (int z = -1; z < 2; z++)
                    for (uint32_t i = 0; i < par.maxParInCell; i++) {
                        I3 c = {gridPos.x + x, gridPos.y + y, gridPos.z + z};
                        uint32_t gridHash = grid_hash(c);
                        uint32_t bucketStart = fetch_cell_start(T, gridHash);
                        uint32_t index2 = bucketStart + i;
                        if (fetch_hash(T, index2) != gridHash) break;
                        if (index2 != index) {
                            bool ok = in_range(index2, T.n);
                            float v2[4] = {0, 0, 0, 0}, p2[4] = {0, 0, 0, 0};
                            if (ok) { memcpy(v2, T.vel + 4 * (size_t)index2, 16);  memcpy(p2, T.pos + 4 * (size_t)index2, 16); }
                            // float4 normalize(a+b): all four components
                            float sv[4] = {vel[0] + v2[0], vel[1] + v2[1], vel[2] + v2[2], vel[3] + v2[3]};
                            float spv[4] = {pos[0] + p2[0], pos[1] + p2[1], pos[2] + p2[2], pos[3] + p2[3]};
                            float ilv = 1.0f / sqrtf(sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2] + sv[3] * sv[3]);
                            float ilp = 1.0f / sqrtf(spv[0] * spv[0] + spv[1] * spv[1] + spv[2] * spv[2] + spv[3] * spv[3]);
                            float nd = (sv[0] * ilv) * (spv[0] * ilp) + (sv[1] * ilv) * (spv[1] * ilp) +
                                       (sv[2] * ilv) * (spv[2] * ilp) + (sv[3] * ilv) * (spv[3] * ilp);
                            float dv[4] = {vel[0] - v2[0], vel[1] - v2[1], vel[2] - v2[2], vel[3] - v2[3]};
                            float dp[4] = {pos[0] - p2[0], pos[1] - p2[1], pos[2] - p2[2], pos[3] - p2[3]};
                            float lv = sqrtf(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2] + dv[3] * dv[3]);
                            float lp = sqrtf(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2] + dp[3] * dp[3]);
                            vdiff += lv * (1.0f - nd) * Wfoam(lp, h);
                        }
                    }
        float TrappedAirPot = phi_clamp(vdiff, minvThreshold, maxvThreshold);
        float Iwc = 0;
        float kEnergy = (float)((double)(par.particleMass * 1000 * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2] + vel[3] * vel[3])) * 0.5);
        float kPot = phi_clamp(kEnergy, 0.1f, 1.0f);
        float Nd = kPot * (3.0f * TrappedAirPot + 3.0f * Iwc);
        color = v3(Nd, Nd, Nd);
    } break;
    case CLR_VelRGB:
        color = v3(0.5f, 0.5f, 0.5f) + par.contrast * 2.5f * Vv;
        break;
    case CLR_Accel: {
        float f = 0.02f * length(addVel) / par.timeStep;
        float clrF = par.brightness + par.contrast * f;  intens = clrF;
        color *= clrF;
    } break;
    case CLR_Dens: {
        float d = 4.f * (dens - par.restDensity) / par.restDensity;
        float clrD = par.brightness + par.contrast * d;  intens = clrD;
        color *= clrD;
    } break;
    default: break;
    }
    if (par.iHue == 1) { color.x = intens;  color.y = 0.f; }

    // ---- dye: System.cu:519-545 ----
    if (par.dyeClear > 0) dyeColor[si] = 0.f;
    else if (par.dyeType > 0) {
        float dyeCl = in_range(si, T.n) ? T.dye[si] : 0.f;
        V3 rel = P - from3(par.dyePos);
        switch (par.dyeType) {
        case 1: if (fabsf(rel.x) < par.dyeSize.x && fabsf(rel.y) < par.dyeSize.y && fabsf(rel.z) < par.dyeSize.z) dyeCl = 1.f; break;
        case 2: if (length(rel) < par.dyeSize.y) dyeCl = 1.f; break;
        }
        dyeCl -= par.timeStep * par.dyeFade;
        if (dyeCl < 0.f) dyeCl = 0.f;
        if (dyeCl >= 0.f) dyeColor[si] = dyeCl;
        if (par.iHue == 0) color += v3(0.9f, 0.9f, 1) * dyeCl;
        else color.y = dyeCl;
    }
    float* c4 = clr + 4 * (size_t)si;
    c4[0] = color.x;  c4[1] = color.y;  c4[2] = color.z;  c4[3] = 1.f;
}

}  // namespace

// ---- C interface ---------------------------------------------------------------------------------------
extern "C" const char* orc_kind(void) { return "port"; }
extern "C" int orc_sizeof_params(void) { return (int)sizeof(SimParams); }
extern "C" void orc_set_threads(int n) { g_threads = n; }
extern "C" int orc_get_threads(void) { return g_threads > 0 ? g_threads : omp_get_max_threads(); }
extern "C" void orc_set_params(const void* p) { memcpy(&par, p, sizeof(SimParams)); }

extern "C" void orc_integrate(const float* oldPos, const float* oldVel, float* newPos, float* newVel, int n)
{
    #pragma omp parallel for schedule(static) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) integrate_one(oldPos, oldVel, newPos, newVel, i);
}

extern "C" void orc_calc_hash(const float* pos, uint32_t* pairs, int n)
{   // calcHashD: Kernel_Cell.cui:23-34
    #pragma omp parallel for schedule(static) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) {
        pairs[2 * (size_t)i] = grid_hash(grid_pos(pos[4 * (size_t)i], pos[4 * (size_t)i + 1], pos[4 * (size_t)i + 2]));
        pairs[2 * (size_t)i + 1] = (uint32_t)i;
    }
}

extern "C" void orc_sort_pairs(uint32_t* pairs, int n)
{   // RadixSort (radixsort_kernel.cu:445-472): LSD radix sort, every pass stable => stable sort by key
    struct KV { uint32_t key, value; };
    KV* p = (KV*)pairs;
    std::stable_sort(p, p + n, [](const KV& a, const KV& b) { return a.key < b.key; });
}

extern "C" void orc_reorder(const uint32_t* pairs, uint32_t* cellStart, const float* oldPos, const float* oldVel,
                            float* sortedPos, float* sortedVel, int n, int numCells)
{   // memset + reorderD: System.cu:694, Kernel_Cell.cui:40-67
    memset(cellStart, 0xff, sizeof(uint32_t) * (size_t)numCells);
    for (int i = 0; i < n; i++) {
        uint32_t h = pairs[2 * (size_t)i];
        if (i == 0 || h != pairs[2 * (size_t)(i - 1)]) cellStart[h] = (uint32_t)i;
    }
    #pragma omp parallel for schedule(static) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) {
        uint32_t src = pairs[2 * (size_t)i + 1];
        if (in_range(src, n)) {
            memcpy(sortedPos + 4 * (size_t)i, oldPos + 4 * (size_t)src, 16);
            memcpy(sortedVel + 4 * (size_t)i, oldVel + 4 * (size_t)src, 16);
        } else {
            memset(sortedPos + 4 * (size_t)i, 0, 16);  memset(sortedVel + 4 * (size_t)i, 0, 16);
        }
    }
}

static Tables make_tables(const float* pos, const float* vel, const float* pressure, const float* density,
                          const uint32_t* pairs, const uint32_t* cellStart, const float* dye, int n, int numCells)
{
    Tables T;
    T.pos = pos;  T.vel = vel;  T.pressure = pressure;  T.density = density;  T.pairs = pairs;
    T.cellStart = cellStart;  T.dye = dye;  T.n = n;  T.numCells = numCells;
    return T;
}

extern "C" void orc_density(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                            float* pressure, float* density, int n, int numCells)
{
    Tables T = make_tables(sortedPos, nullptr, nullptr, nullptr, pairs, cellStart, nullptr, n, numCells);
    #pragma omp parallel for schedule(static, 64) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) density_one<false>(T, pressure, density, nullptr, (uint32_t)i);
}

extern "C" void orc_neighbor_counts(const float* sortedPos, const uint32_t* pairs, const uint32_t* cellStart,
                                    uint32_t* counts, int n, int numCells)
{
    Tables T = make_tables(sortedPos, nullptr, nullptr, nullptr, pairs, cellStart, nullptr, n, numCells);
    #pragma omp parallel for schedule(static, 64) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) density_one<true>(T, nullptr, nullptr, counts, (uint32_t)i);
}

extern "C" void orc_force(const float* sortedPos, const float* sortedVel, const float* pressure, const float* density,
                          const uint32_t* pairs, const uint32_t* cellStart,
                          float* newVel, float* clr, float* dyeColor, int n, int numCells)
{
    Tables T = make_tables(sortedPos, sortedVel, pressure, density, pairs, cellStart, dyeColor, n, numCells);
    #pragma omp parallel for schedule(static, 64) num_threads(orc_get_threads())
    for (int i = 0; i < n; i++) force_one(T, newVel, clr, dyeColor, (uint32_t)i);
}

static void orc_sys_read_dims(orc_system* s);
#include "oracle_system.inc"
static void orc_sys_read_dims(orc_system* s)
{
    const SimParams* p = (const SimParams*)s->par;
    s->n = (int)p->numParticles;  s->numCells = (int)p->numCells;
}
