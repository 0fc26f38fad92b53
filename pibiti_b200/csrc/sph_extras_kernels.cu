// sph_extras_kernels.cu -- the parts of the reference's computeForceD that only some scenes use, kept
// out of the hot kernels so that they cost no registers there:
//   k_obstacles   height-map sphere lattice (System.cu:254-309) and rotor / propeller sphere sets
//                 (System.cu:312-372); launched only when iHmap > 0 or rotType > 0
//   k_color_dye   particle colour (System.cu:406-515, incl. the foam / trapped-air potential of CLR_Vel) and
//                 dye (System.cu:519-545); launched only when the visual outputs are enabled (sph_set_visual)
// Both run after the force kernel on the sorted arrays.  The obstacle impulses are added to the new velocity
// afterwards ((vel + dv) + dvObstacle instead of vel + (dv + dvObstacle)): last-bit differences only.
#include "sph_device.cuh"

namespace {

// collideSpheresR (Kernel_Cell.cui:101-115): DEM spring + damping, no shear
__device__ __forceinline__ float3 sphere_contact_r(const SimParams& par, float3 posAB, float3 relVel, float radiusAB)
{
    float dist = sqrtf(posAB.x * posAB.x + posAB.y * posAB.y + posAB.z * posAB.z);
    float3 force = make_float3(0.f, 0.f, 0.f);
    if (dist < radiusAB) {
        float inv = 1.0f / dist;
        float sp = par.spring * (dist - radiusAB);
        force.x = sp * (posAB.x * inv) + par.damping * relVel.x;
        force.y = sp * (posAB.y * inv) + par.damping * relVel.y;
        force.z = sp * (posAB.z * inv) + par.damping * relVel.z;
    }
    return force;
}

__device__ __forceinline__ void add3(float3& a, float3 b) { a.x += b.x;  a.y += b.y;  a.z += b.z; }

__global__ void __launch_bounds__(128)
k_obstacles(const __grid_constant__ SimParams par, const float4* __restrict__ posP, const float4* __restrict__ velD,
            float4* __restrict__ velNew, int first, int n, const uint32_t* __restrict__ dev)
{
    if (dev) { first = (int)__ldg(dev);  n = (int)__ldg(dev + 1); }
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 pp = posP[i];
    const float4 vd = velD[i];
    const float3 pos = make_float3(pp.x, pp.y, pp.z);
    const float3 vel3 = make_float3(-vd.x, -vd.y, -vd.z);
    float3 add = make_float3(0.f, 0.f, 0.f);
    const float twoPi = 2.f * PI;

    if (par.iHmap > 0) {                                                    // System.cu:254-309
        const float rr = par.particleR + par.rotR;
        const int hz = 2, hy = 1;
        const bool onXZ = par.iHmap == 1;
        // the lattice is indexed along (x or y) and z; the third axis carries the height
        const float u = onXZ ? pos.x : pos.y, uMin = onXZ ? par.worldMin.x : par.worldMin.y;
        const float uSize = onXZ ? par.worldSizeD.x : par.worldSizeD.y;
        const int iu = (int)((u - uMin) / par.rotSpc);           const float uf = iu * par.rotSpc + uMin;
        const int iz = (int)((pos.z - par.worldMin.z) / par.rotSpc);  const float zf = iz * par.rotSpc + par.worldMin.z;
        for (int j = -hz; j <= hz; j++)
            for (int ii = -hz; ii <= hz; ii++) {
                const float uh = uf + ii * par.rotSpc, zh = zf + j * par.rotSpc;
                const float un = uh / uSize * 2.f * PI, zn = zh / par.worldSizeD.z * 2.f * PI;    // "x / size * PI2", PI2 = 2.f*PI unparenthesised
                const float ss = sinf(par.s1 * un + par.s2 * PI / 180.f) * sinf(par.s3 * zn + par.s4 * PI / 180.f);
                if (par.s5 <= -1.f || ss > par.s5) {
                    const float hf = ss * par.hClose + par.r2Angle + (onXZ ? par.worldMin.y : par.worldMin.x);
                    for (int k = 0; k <= hy; k++) {
                        const float hh = hf - k * par.rotSpc;
                        const float3 c = onXZ ? make_float3(uh - pos.x, hh - pos.y, zh - pos.z)
                                              : make_float3(hh - pos.x, uh - pos.y, zh - pos.z);
                        add3(add, sphere_contact_r(par, c, vel3, rr));
                    }
                }
            }
    }

    if (par.rotType > 0) {                                                  // System.cu:312-372
        const int sx = par.rotSize.x, sy = par.rotSize.y, sz = par.rotSize.z, cb = par.rotBlades;
        const float sp = par.rotSpc, ca = twoPi / cb, x2 = sx * 0.5f, y2 = sy * 0.5f;
        const float rr = par.particleR + par.rotR;
        const float3 cpos = make_float3(par.collPos.x, par.collPos.y, par.collPos.z);
        if (par.rotType == 1 || par.rotType == 2) {
            const bool axisZ = par.rotType == 1;
            const float along = axisZ ? pos.z : pos.y, c0 = axisZ ? cpos.z : cpos.y;
            if (along > c0 - rr && along < c0 + rr + sp * sy)
                for (int z = 1; z <= sz; z++)
                    for (int x = 0; x <= sx; x++)
                        for (int c = 0; c < cb; c++) {
                            const float a = -par.rAngle + c * ca + z * par.rTwist, cs = cosf(a) * sp, sn = -sinf(a) * sp;
                            const float p0 = (x - x2) * cs - z * sn, p1 = (x - x2) * sn + z * cs;
                            for (int y = 0; y <= sy; y++) {
                                const float3 pc = axisZ ? make_float3(p0, p1, sp * y) : make_float3(p0, sp * y, p1);
                                add3(add, sphere_contact_r(par, make_float3(cpos.x + pc.x - pos.x, cpos.y + pc.y - pos.y, cpos.z + pc.z - pos.z), vel3, rr));
                            }
                        }
        } else if (par.rotType == 3) {
            if (pos.z > cpos.z - rr - sp * sz / 2.f && pos.z < cpos.z + rr + sp * sz / 2.f) {
                float3 rotPos = cpos;
                float aa = par.rAngle, tw = par.rTwist;
                if (par.r2Dist > 0.f) {
                    if (pos.x > 0.f) rotPos.x += par.r2Dist * 0.5f;
                    else { rotPos.x -= par.r2Dist * 0.5f;  aa = par.r2Angle;  tw *= par.r2twist; }
                }
                for (int c = 0; c < cb; c++) {
                    const float a = aa + c * ca;
                    for (int h = 0; h <= sy; h++) {
                        float dh = 0.f;
                        if (h == sy - 1) dh = 0.5f; else if (h == sy) dh = 1.2f;
                        for (int x = 0; x <= sz; x++) {
                            float d = dh;
                            if (x == sz && d == 0.f) d = 0.4f;
                            const float k = cosf((x - x2) * 0.2f + PI * 0.6f);
                            const float at = tw * (d * -0.05f - (h - y2) * k);
                            const float ac = a + at, cs = cosf(ac) * sp, sn = -sinf(ac) * sp;
                            const float3 pc = make_float3(x * cs, x * sn, sp * (-d + h - y2) * k);
                            const float r = par.rotR * fabsf(1 - d);
                            add3(add, sphere_contact_r(par, make_float3(rotPos.x + pc.x - pos.x, rotPos.y + pc.y - pos.y, rotPos.z + pc.z - pos.z),
                                                       vel3, par.particleR + r));
                        }
                    }
                }
            }
        }
    }

    if (add.x != 0.f || add.y != 0.f || add.z != 0.f) {
        float4 v = velNew[i];
        v.x += add.x;  v.y += add.y;  v.z += add.z;
        velNew[i] = v;
    }
}

// ---- colour + dye ----------------------------------------------------------------------------------

__device__ __forceinline__ float foam_w(float x, float h) { return x <= h ? (1.f - x) / h : 0.f; }     // System.cu:217-222
__device__ __forceinline__ float phi_clamp(float I, float tmin, float tmax) { return (fminf(I, tmax) - fminf(I, tmin)) / (tmax - tmin); }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// posS/velS: sorted positions / post-integration velocities WITH their w components (the foam term
// normalises float4 sums); velD.w = density; velNew = velocities after the force step.
__global__ void __launch_bounds__(128)
k_color_dye(const __grid_constant__ SimParams par, const float4* __restrict__ posS, const float4* __restrict__ velS,
            const float4* __restrict__ velD, const float4* __restrict__ velNew, const uint32_t* __restrict__ keyS,
            const uint32_t* __restrict__ cellStart, const uint32_t* __restrict__ idx,
            float4* __restrict__ clr, float* __restrict__ dye, int first, int n)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (par.clrType == CLR_None) return;                                    // System.cu:409: no colour, no dye
    const float4 pos = posS[i], vel = velS[i], vn = velNew[i];
    const float dens = velD[i].w;
    const uint32_t si = idx[i];
    const float3 v3 = make_float3(vel.x, vel.y, vel.z);
    const float3 addVel = make_float3(vn.x - vel.x, vn.y - vel.y, vn.z - vel.z);
    float3 color = make_float3(0.2f, 0.5f, 1.f);
    float intens = 0.f;

    switch (par.clrType) {
    case CLR_VelAcc: {
        float v = 2.5f * len3(v3), f = 0.02f * len3(addVel) / par.timeStep;
        float clrV = par.brightness + par.contrast * v, clrF = par.contrast * f;
        color = make_float3(color.x * clrV + 0.7f * clrF, color.y * clrV + 0.35f * clrF, color.z * clrV);
    } break;
    case CLR_DensAcc: {
        float d = 4.f * (dens - par.restDensity) / par.restDensity, f = 0.02f * len3(addVel) / par.timeStep;
        float clrD = par.brightness + par.contrast * d, clrF = par.contrast * f;
        color = make_float3(color.x * clrD + 0.7f * clrF, color.y * clrD + 0.7f * clrF, color.z * clrD);
    } break;
    case CLR_Vel: {
        intens = par.brightness + par.contrast * 2.5f * len3(v3);
        // foam / trapped-air potential (System.cu:429-498): x outermost, '+' inside normalize() as written there
        const uint32_t key = keyS[i];
        const long long C = par.numCells;
        float vdiff = 0.f;
        for (int x = -1; x < 2; x++)
            for (int y = -1; y < 2; y++)
                for (int z = -1; z < 2; z++) {
                    const long long h = (long long)key + (long long)z * par.gridSize_yx + (long long)y * par.gridSize.x + x;
                    if (h < 0 || h >= C) continue;
                    const uint32_t a = __ldg(cellStart + h);
                    uint32_t e = __ldg(cellStart + h + 1);
                    if (e - a > par.maxParInCell) e = a + par.maxParInCell;
                    for (uint32_t g = a; g < e; g++) {
                        if (g == (uint32_t)i) continue;
                        const float4 v2 = __ldg(velS + g), p2 = __ldg(posS + g);
                        float4 sv = make_float4(vel.x + v2.x, vel.y + v2.y, vel.z + v2.z, vel.w + v2.w);
                        float4 sp = make_float4(pos.x + p2.x, pos.y + p2.y, pos.z + p2.z, pos.w + p2.w);
                        const float ilv = 1.0f / sqrtf(dot4(sv, sv)), ilp = 1.0f / sqrtf(dot4(sp, sp));
                        sv = make_float4(sv.x * ilv, sv.y * ilv, sv.z * ilv, sv.w * ilv);
                        sp = make_float4(sp.x * ilp, sp.y * ilp, sp.z * ilp, sp.w * ilp);
                        const float4 dv = make_float4(vel.x - v2.x, vel.y - v2.y, vel.z - v2.z, vel.w - v2.w);
                        const float4 dp = make_float4(pos.x - p2.x, pos.y - p2.y, pos.z - p2.z, pos.w - p2.w);
                        vdiff += sqrtf(dot4(dv, dv)) * (1.0f - dot4(sv, sp)) * foam_w(sqrtf(dot4(dp, dp)), 3.0f);
                    }
                }
        const float trapped = phi_clamp(vdiff, 5.0f, 20.0f);
        const float kEnergy = (float)((double)(par.particleMass * 1000 * dot4(vel, vel)) * 0.5);
        const float nd = phi_clamp(kEnergy, 0.1f, 1.0f) * (3.0f * trapped + 3.0f * 0.f);
        color = make_float3(nd, nd, nd);
    } break;
    case CLR_VelRGB:
        color = make_float3(0.5f + par.contrast * 2.5f * v3.x, 0.5f + par.contrast * 2.5f * v3.y, 0.5f + par.contrast * 2.5f * v3.z);
        break;
    case CLR_Accel: {
        intens = par.brightness + par.contrast * (0.02f * len3(addVel) / par.timeStep);
        color = make_float3(color.x * intens, color.y * intens, color.z * intens);
    } break;
    case CLR_Dens: {
        intens = par.brightness + par.contrast * (4.f * (dens - par.restDensity) / par.restDensity);
        color = make_float3(color.x * intens, color.y * intens, color.z * intens);
    } break;
    default: break;
    }
    if (par.iHue == 1) { color.x = intens;  color.y = 0.f; }

    if (par.dyeClear > 0) dye[si] = 0.f;                                    // System.cu:519-545
    else if (par.dyeType > 0) {
        float dyeCl = dye[si];
        const float3 rel = make_float3(pos.x - par.dyePos.x, pos.y - par.dyePos.y, pos.z - par.dyePos.z);
        if (par.dyeType == 1) { if (fabsf(rel.x) < par.dyeSize.x && fabsf(rel.y) < par.dyeSize.y && fabsf(rel.z) < par.dyeSize.z) dyeCl = 1.f; }
        else if (par.dyeType == 2) { if (len3(rel) < par.dyeSize.y) dyeCl = 1.f; }
        dyeCl -= par.timeStep * par.dyeFade;
        if (dyeCl < 0.f) dyeCl = 0.f;
        dye[si] = dyeCl;
        if (par.iHue == 0) { color.x += 0.9f * dyeCl;  color.y += 0.9f * dyeCl;  color.z += dyeCl; }
        else color.y = dyeCl;
    }
    clr[si] = make_float4(color.x, color.y, color.z, 1.f);
}

}  // namespace

#define SPH_COUNT(L) do { if ((L).launches) ++*(L).launches; } while (0)

bool sph_needs_obstacles(const SimParams& par) { return par.iHmap > 0 || par.rotType > 0; }

void sph_launch_obstacles(const SphLaunch& L, const SimParams& par, const float4* posP, const float4* velD, float4* velNew,
                          int first, int count, const uint32_t* dev)
{
    if (count <= 0) return;
    k_obstacles<<<(count + 127) / 128, 128, 0, L.stream>>>(par, posP, velD, velNew, first, first + count, dev);
    SPH_COUNT(L);
}

void sph_launch_color_dye(const SphLaunch& L, const SimParams& par, const float4* posS, const float4* velS, const float4* velD,
                          const float4* velNew, const uint32_t* keyS, const uint32_t* cellStart, const uint32_t* idx,
                          float4* clr, float* dye, int first, int count)
{
    if (count <= 0) return;
    k_color_dye<<<(count + 127) / 128, 128, 0, L.stream>>>(par, posS, velS, velD, velNew, keyS, cellStart, idx, clr, dye, first, first + count);
    SPH_COUNT(L);
}
