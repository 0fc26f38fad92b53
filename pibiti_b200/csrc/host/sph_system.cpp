// sph_system.cpp -- the cSPH-shaped system object on top of the C ABI.
//
// Re-states, headless, what the reference's SPH layer does around the solver:
//   cSPH::cSPH / Reset / Drop          source/SPH/SPH_Init.cpp:8-118
//   cSPH::_InitMem / _FreeMem          source/SPH/SPH_Mem.cpp:11-82      (-> sph_create / sph_destroy)
//   cSPH::Update                       source/SPH/SPH_Update.cpp:12-81   (-> sph_set_params + sph_step)
//   cSPH::getArray / setArray          source/SPH/SPH_Util.cpp:44-71     (-> sph_get_array / sph_set_array)
//   InitScene / LoadScenes / Next/Prev source/SPH/SPH_Scenes.cpp:9-111
//   App::UpdateEmitter                 source/App/Update.cpp:9-97        (per-step host prologue)
// Random numbers come from the C library's rand(), never seeded, exactly as in the reference
// (pch/header.h:139-146) so that "random" scenes reproduce.
#include "sph_host.h"
#include "xml_lite.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>

namespace {

inline float3 f3(float x, float y, float z) { float3 v; v.x = x; v.y = y; v.z = z; return v; }
inline float4 f4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }

const float kRandMaxInv = 1.f / float(RAND_MAX);
inline float frand() { return rand() * kRandMaxInv; }
inline float random_between(float a, float b) { return (b - a) * float(rand()) * kRandMaxInv + a; }
inline float len3(float x, float y, float z) { return sqrtf(x * x + y * y + z * z); }

}  // namespace

// ---- timer (pch/timer.cpp:6-65, on a portable clock) ------------------------------------------------

static double now_seconds()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

Timer::Timer() : iFR(0), dt(0.), FR(0.), iv(0.), iv1(0.4)
{
    t = now_seconds();
    st = t;  st1 = t;
}

bool Timer::update(bool updFR)
{
    t = now_seconds();
    dt = t - st;
    if (dt < iv) return false;          // not yet one interval
    st = t;
    if (!updFR) return true;
    iFR++;
    const double dt1 = t - st1;
    if (dt1 >= iv1) { FR = iFR / dt1;  iFR = 0;  st1 = t; }
    return true;
}

cSPH::cSPH(const char* scenesXmlPath, int dev)
    : bInitialized(false), curScene(0), hPos(nullptr), hVel(nullptr), hCounters(nullptr), colorVbo(0),
      curPosRead(0), curPosWrite(1), xmlPath(scenesXmlPath ? scenesXmlPath : "Scenes.xml"), device(dev), sys(nullptr),
      msys(nullptr), multiDirty(false), memParticles(0), memCells(0)
{
    posVbo[0] = posVbo[1] = 0;
    DropPos = f3(0, 0, 0);
    LoadScenes();
}

cSPH::cSPH(const char* scenesXmlPath, const int* devs, int ndev)
    : bInitialized(false), curScene(0), hPos(nullptr), hVel(nullptr), hCounters(nullptr), colorVbo(0),
      curPosRead(0), curPosWrite(1), xmlPath(scenesXmlPath ? scenesXmlPath : "Scenes.xml"), device(ndev > 0 ? devs[0] : -1), sys(nullptr),
      devices(devs, devs + (ndev > 0 ? ndev : 0)), msys(nullptr), multiDirty(false), memParticles(0), memCells(0)
{
    posVbo[0] = posVbo[1] = 0;
    DropPos = f3(0, 0, 0);
    LoadScenes();
}

cSPH::~cSPH()
{
    _FreeMem();
    if (sys) { sph_destroy(sys);  sys = nullptr; }
    if (msys) { sph_multi_destroy(msys);  msys = nullptr; }
}

// ---- multi-GPU mode: the slabs are authoritative after a step, the host mirrors after a write -----

int cSPH::multiSync()
{
    if (!msys || multiDirty) return SPH_OK;                 // the mirrors already hold the newest state
    int rc = sph_multi_get_state(msys, (float*)hPos, (float*)hVel, nullptr, nullptr, (int)scn.params.numParticles, nullptr);
    if (rc != SPH_OK) err = sph_multi_last_error(msys);
    return rc;
}

int cSPH::multiFlush()
{
    if (!msys || !multiDirty) return SPH_OK;
    int rc = sph_multi_set_params(msys, &scn.params);
    if (rc == SPH_OK) rc = sph_multi_set_state(msys, (const float*)hPos, (const float*)hVel, (int)scn.params.numParticles, nullptr);
    if (rc != SPH_OK) err = sph_multi_last_error(msys); else multiDirty = false;
    return rc;
}

// ---- memory ------------------------------------------------------------------------------------

void cSPH::_InitMem()
{
    if (bInitialized) return;
    bInitialized = true;
    const size_t npar = scn.params.numParticles;
    hPos = new float4[npar];  memset(hPos, 0, npar * sizeof(float4));
    hVel = new float4[npar];  memset(hVel, 0, npar * sizeof(float4));
    hCounters = new int[10];  memset(hCounters, 0, 10 * sizeof(int));     // SPH_Mem.cpp:24
    if (device < 0) return;
    if (devices.size() > 1) {
        // one slab per device; a slab holds its share of the particles plus ghosts, arrivals and drift of the balance
        if (msys) { sph_multi_destroy(msys);  msys = nullptr; }
        const size_t cap = npar / devices.size() * 3 / 2 + 65536;
        int rc = sph_multi_create(&scn.params, (int)devices.size(), devices.data(), (int)(cap < npar + 4096 ? cap : npar + 4096), &msys);
        if (rc != SPH_OK) { err = sph_multi_last_error(nullptr);  msys = nullptr;  fprintf(stderr, "cSPH: %s\n", err.c_str()); }
        multiDirty = true;                                  // the first Update pushes whatever Reset / setArray put in the mirrors
        return;
    }
    // A scene that fits the device buffers of the previous one keeps them (the reference frees and reallocates
    // everything on every scene switch, SPH_Scenes.cpp:9-13 -> SPH_Mem.cpp:11-82): only the parameters change.
    if (sys && npar <= memParticles && scn.params.numCells <= memCells) {
        if (sph_set_params(sys, &scn.params) == SPH_OK && sph_reset_state(sys) == SPH_OK) return;
        err = sph_last_error(sys);
    }
    if (sys) { sph_destroy(sys);  sys = nullptr;  posVbo[0] = posVbo[1] = colorVbo = 0; }
    int rc = sph_create(&scn.params, device, &sys);
    if (rc != SPH_OK) { err = sph_last_error(nullptr);  sys = nullptr;  memParticles = memCells = 0;  fprintf(stderr, "cSPH: %s\n", err.c_str()); }
    else { memParticles = npar;  memCells = scn.params.numCells; }
}

// Host mirrors go; the device buffers stay with the handle until the object dies or a larger scene needs new ones.
void cSPH::_FreeMem()
{
    if (!bInitialized) return;
    bInitialized = false;
    delete[] hPos;  hPos = nullptr;
    delete[] hVel;  hVel = nullptr;
    delete[] hCounters;  hCounters = nullptr;
}

// ---- particle initialisers ---------------------------------------------------------------------

void cSPH::Reset(int type)
{
    SimParams& p = scn.params;
    const float r = p.particleR, spc = scn.spacing, b = p.distBndSoft;
    const float3 cp = f3(p.collPos.x, p.collPos.y, p.collPos.z);
    float3 wMin = p.worldMinD, wSize = p.worldSizeD;
    float3 imin = scn.initMin, imax = scn.initMax;
    const float4 vel0 = f4(0, 0, 0, 0);
    imin.y += r;
    wMin.x += r;  wMin.y += r;  wMin.z += r;
    wSize.x -= 2 * r;  wSize.y -= 2 * r;  wSize.z -= 2 * r;
    const uint n = p.numParticles;
    uint i = 0;

    if (type == 1) {                    // uniform random in the (inset) world box
        for (i = 0; i < n; ++i) {
            hPos[i].x = wMin.x + wSize.x * frand();
            hPos[i].y = wMin.y + wSize.y * frand();
            hPos[i].z = wMin.z + wSize.z * frand();
            hPos[i].w = 1.f;
            hVel[i] = vel0;
        }
    } else {                            // lattice walk through the init volume
        float4 pos = f4(imin.x, imin.y, imin.z, 1);
        // advance one lattice step; axis order a (fastest), b, c.  On overflow of the last axis the
        // walker wraps to its start, as the reference does (SPH_Init.cpp:41-48).
        auto step3 = [&](float& pa, float amin, float amax, float& pb, float bmin, float bmax, float& pc, float cmin, float cmax) {
            pa += spc;
            if (pa >= amax) {
                pa = amin;  pb += spc;
                if (pb >= bmax) {
                    pb = bmin;  pc += spc;
                    if (pc >= cmax) pc = cmin;
                }
            }
        };
        auto advance = [&]() {
            switch (scn.initLast) {
            default:
            case 1: step3(pos.x, imin.x, imax.x, pos.z, imin.z, imax.z, pos.y, imin.y, imax.y); break;
            case 0: step3(pos.y, imin.y, imax.y, pos.z, imin.z, imax.z, pos.x, imin.x, imax.x); break;
            case 2: step3(pos.x, imin.x, imax.x, pos.y, imin.y, imax.y, pos.z, imin.z, imax.z); break;
            }
        };
        // a lattice that can never satisfy the acceptance test would spin forever in the reference;
        // give up after visiting far more sites than any volume can hold
        unsigned long long guard = 0;
        const unsigned long long guardMax = 4000000000ull;

        if (p.bndType == BND_PUMP_Y) {
            const float rad = scn.initMax.x, hc = p.hClose - b * 0.90f, rin = rad * p.radIn - b / 2,
                        xs = sinf(p.angOut * p.s3) * rad - b / 2 - p.particleR;
            while (i < n && guard++ < guardMax) {
                bool in =
                    (p.angOut < 0.5f && pos.y > 0 && pos.z < hc && pos.x < xs && pos.x > -xs) ||
                    (p.angOut >= 0.5f && pos.y > 0 && pos.z < hc && pos.x > 0.09f * p.s4) ||
                    sqrtf(pos.x * pos.x + pos.y * pos.y) * 1.01f < (pos.z < hc ? rad : rin);
                if (in) { hPos[i] = pos;  hVel[i] = vel0;  i++; }
                advance();
            }
        } else {
            while (i < n && guard++ < guardMax) {
                bool outsideCollider = len3(pos.x - cp.x, pos.y - cp.y, pos.z - cp.z) > p.collR || p.rotType > 0;
                if (outsideCollider) {
                    bool in = p.bndType == BND_BOX || scn.initType >= 9 || p.bndType == BND_CYL_YZ ||
                              (p.bndType == BND_CYL_Y && sqrtf(pos.x * pos.x + pos.z * pos.z) < imax.x) ||
                              (p.bndType == BND_CYL_Z && sqrtf(pos.x * pos.x + pos.y * pos.y) < -imin.y) ||
                              (p.bndType == BND_SPHERE && len3(pos.x, pos.y, pos.z) < -imin.y);
                    if (in) { hPos[i] = pos;  hVel[i] = vel0;  i++; }
                }
                advance();
            }
        }
        if (i < n) {
            err = "Reset: init volume accepts no lattice site; remaining particles left at zero";
            fprintf(stderr, "cSPH: %s\n", err.c_str());
        }
    }
    setArray(0, hPos, 0, (int)n);
    setArray(1, hVel, 0, (int)n);
}

void cSPH::Drop(bool bRandom)
{
    SimParams& p = scn.params;
    const float r = scn.dropR, spc = scn.spacing, r2 = r * r, db = (r + 1) * spc, d = 0.5f;
    const float4 vel0 = f4(0, 0, 0, 0);
    float3 pos = f3(0.5f, 0.5f, 0.5f);

    if (bRandom) {
        BndType t = p.bndType;
        pos.y = (t == BND_BOX || t == BND_CYL_Y) ? 1.0f : 0.75f;
        if (t == BND_BOX || t == BND_CYL_Z) {
            pos.x = frand();
            pos.z = frand();
        } else {
            float rr = frand(), k = random_between(0, 2 * PI);
            pos.x = d + d * cosf(k) * rr;
            pos.z = d - d * sinf(k) * rr;
        }
    }
    const float3 posw = f3(p.worldMinD.x + db + pos.x * (p.worldSizeD.x - db * 2),
                           p.worldMinD.y + db + pos.y * (p.worldSizeD.y - db * 2),
                           p.worldMinD.z + db + pos.z * (p.worldSizeD.z - db * 2));
    DropPos = posw;

    uint size = 0, id = (uint)app.emitId, a = id;
    const int ir = (int)r;
    for (int z = -ir; z <= r; z++)
        for (int y = -ir; y <= r; y++)
            for (int x = -ir; x <= r; x++)
                if (x * x + y * y + z * z <= r2 && a < p.numParticles) {
                    // the reference memcpy()s 16 bytes out of a 12-byte float3 here (SPH_Init.cpp:109),
                    // leaving w undefined; w = 1 like every other particle
                    hPos[a] = f4((float)x * spc + posw.x, (float)y * spc + posw.y, (float)z * spc + posw.z, 1.f);
                    hVel[a] = vel0;
                    size++;  a++;
                }
    setArray(0, &hPos[id], (int)id, (int)size);
    setArray(1, &hVel[id], (int)id, (int)size);

    const uint num = p.numParticles;
    app.emitId += (int)size;
    if ((uint)app.emitId >= num) app.emitId -= (int)num;
}

// ---- stepping ----------------------------------------------------------------------------------

int cSPH::Update() { return Update(1); }

int cSPH::Update(int nsteps)
{
    if (!bInitialized) return SPH_ERR_STATE;
    if (msys) {
        tim.update(true);
        if (int rc = multiFlush()) return rc;
        if (app.bChangedAny) {
            app.bChangedAny = false;
            if (sph_multi_set_params(msys, &scn.params) != SPH_OK) { err = sph_multi_last_error(msys);  return SPH_ERR_PARAMS; }
        }
        int rc = sph_multi_step(msys, nsteps);
        if (rc != SPH_OK) err = sph_multi_last_error(msys);
        return rc;
    }
    if (!sys) { if (err.empty()) err = "cSPH::Update: no solver (constructed without a device)";  return SPH_ERR_STATE; }
    tim.update(true);                                       // SPH_Update.cpp:16
    if (app.bChangedAny) {                                  // SPH_Update.cpp:19-27
        app.bChangedAny = false;
        int rc = sph_set_params(sys, &scn.params);
        if (rc != SPH_OK) { err = sph_last_error(sys);  return rc; }
    }
    int rc = sph_step(sys, nsteps);
    if (rc == SPH_OK && (posVbo[0] || colorVbo)) rc = sph_gl_update(sys);      // the renderer's buffers follow the step
    if (rc != SPH_OK) err = sph_last_error(sys);
    return rc;
}

int cSPH::registerGLBuffers(uint positionsVbo, uint colorsVbo)
{
    if (msys) { err = "registerGLBuffers: not available with several GPUs (positions live in per-device slabs)";  return SPH_ERR_STATE; }
    if (!sys) { err = "registerGLBuffers: no solver";  return SPH_ERR_STATE; }
    int rc = sph_gl_register(sys, SPH_POS, positionsVbo);
    if (rc == SPH_OK) { posVbo[0] = posVbo[1] = positionsVbo; }
    if (rc == SPH_OK && colorsVbo) {
        rc = sph_set_visual(sys, 1);
        if (rc == SPH_OK) rc = sph_gl_register(sys, SPH_COLOR, colorsVbo);
        if (rc == SPH_OK) colorVbo = colorsVbo;
    }
    if (rc != SPH_OK) err = sph_last_error(sys);
    return rc;
}

// The per-step host prologue the reference runs before cSPH::Update (App::Simulate,
// source/App/Render.cpp:8-14).  The emitter orientation is the GL matrix Rx(rot.x)*Ry(-rot.y) that
// the reference builds with glRotatef and reads back (Update.cpp:73-76), written out here.
void cSPH::UpdateEmitter()
{
    Scene& sc = scn;
    SimParams* p = &sc.params;

    if (sc.rVel != 0.f) {                                   // rotor / wave phase
        p->rAngle += sc.rVel * p->timeStep;
        if (p->r2Dist > 0.f) {
            p->r2Angle += sc.rVel * sc.r2Vel * p->timeStep;
            p->r2twist = sc.r2Vel < 0.f ? -1.f : 1.f;
        }
        app.bChangedAny = true;
    }
    if (p->dyeClear > 0) { p->dyeClear--;  app.bChangedAny = true; }

    const float mind = 0.000707f, mind2 = mind * mind;
    const float inertia = app.inertia;
    {                                                       // collider follows its target
        float4 d = f4(app.colliderPos.x - p->collPos.x, app.colliderPos.y - p->collPos.y,
                      app.colliderPos.z - p->collPos.z, app.colliderPos.w - p->collPos.w);
        if (fabsf(d.x) > mind || fabsf(d.y) > mind || fabsf(d.z) > mind) {
            for (int iter = 0; iter < 3; ++iter) {
                p->collPos.x += d.x * inertia;  p->collPos.y += d.y * inertia;
                p->collPos.z += d.z * inertia;  p->collPos.w += d.w * inertia;
                d = f4(app.colliderPos.x - p->collPos.x, app.colliderPos.y - p->collPos.y,
                       app.colliderPos.z - p->collPos.z, app.colliderPos.w - p->collPos.w);
            }
            app.bChangedAny = true;
        }
    }
    {                                                       // current accelerator follows its target
        Accel& ac = p->acc[sc.ca];
        float3 ad = f3(sc.accPos[sc.ca].x - ac.pos.x, sc.accPos[sc.ca].y - ac.pos.y, sc.accPos[sc.ca].z - ac.pos.z);
        if (ad.x * ad.x + ad.y * ad.y + ad.z * ad.z > mind2) {
            for (int iter = 0; iter < 3; ++iter) {
                ac.pos.x += ad.x * inertia;  ac.pos.y += ad.y * inertia;  ac.pos.z += ad.z * inertia;
                ad = f3(sc.accPos[sc.ca].x - ac.pos.x, sc.accPos[sc.ca].y - ac.pos.y, sc.accPos[sc.ca].z - ac.pos.z);
            }
            app.bChangedAny = true;
        }
    }
    {                                                       // dye source follows its target
        float3 dd = f3(app.dyePos.x - p->dyePos.x, app.dyePos.y - p->dyePos.y, app.dyePos.z - p->dyePos.z);
        if (fabsf(dd.x) > mind || fabsf(dd.y) > mind || fabsf(dd.z) > mind) {
            for (int iter = 0; iter < 3; ++iter) {
                p->dyePos.x += dd.x * inertia;  p->dyePos.y += dd.y * inertia;  p->dyePos.z += dd.z * inertia;
                dd = f3(app.dyePos.x - p->dyePos.x, app.dyePos.y - p->dyePos.y, app.dyePos.z - p->dyePos.z);
            }
            app.bChangedAny = true;
        }
    }

    for (int e = 0; e < NumEmit; e++) {                     // emitters recycle ring slots
        Emitter& em = sc.emit[e];
        if (em.size <= 0) continue;
        const int eX = em.size, eY = (em.size2 == 0) ? eX : em.size2;
        const int size = std::min(eX * eY, 100);            // the reference's static buffers hold 100
        const float spc = sc.spacing;
        float4 pos[100], vel[100];

        // glRotatef evaluates sine and cosine of angle*pi/180 in double and rounds to float (OpenGL's fixed-function
        // matrix code, e.g. Mesa's _math_matrix_rotate); the same here so that the emitted particles are bit-identical
        const double ax = (double)em.rotLag.x * (M_PI / 180.0), ay = (double)(-em.rotLag.y) * (M_PI / 180.0);
        const float ca = (float)cos(ax), sa = (float)sin(ax), cb = (float)cos(ay), sb = (float)sin(ay);
        // columns of M = Rx(ax) * Ry(ay); the reference multiplies a vector by the columns
        const float c0[3] = {cb, sa * sb, -ca * sb}, c1[3] = {0.f, ca, sa}, c2[3] = {sb, -sa * cb, ca * cb};
        auto mulTr = [&](const float* v, float* rr) {
            rr[0] = v[0] * c0[0] + v[1] * c0[1] + v[2] * c0[2];
            rr[1] = v[0] * c1[0] + v[1] * c1[1] + v[2] * c1[2];
            rr[2] = v[0] * c2[0] + v[1] * c2[1] + v[2] * c2[2];
        };
        const float ev[3] = {0.f, 0.f, em.vel};
        float4 vel4 = f4(0, 0, 0, 0);
        mulTr(ev, &vel4.x);

        int i = 0;
        const float z = (eX - 1) * 0.5f, z2 = (eY - 1) * 0.5f;
        for (int y = 0; y < eY && i < size; y++)
            for (int x = 0; x < eX && i < size; x++, i++) {
                const float pp[3] = {(x - z) * spc, (y - z2) * spc, -spc};
                pos[i] = f4(0, 0, 0, 1.f);
                mulTr(pp, &pos[i].x);
                pos[i].x += em.posLag.x;  pos[i].y += em.posLag.y;  pos[i].z += em.posLag.z;
                vel[i] = vel4;
            }
        // a batch never runs past the end of the particle array
        const int num = (int)sc.params.numParticles;
        const int cnt = std::min(size, num - app.emitId);
        setArray(0, pos, app.emitId, cnt);
        setArray(1, vel, app.emitId, cnt);
        app.emitId += size;
        if (app.emitId >= num) app.emitId -= num;
    }

    if (sc.rain > 0) {                                      // rain
        app.cntRain++;
        if (app.cntRain >= sc.rain) { app.cntRain = 0;  Drop(true); }
    }
    app.fSimTime += p->timeStep;
}

// ---- accessors ---------------------------------------------------------------------------------

float4* cSPH::getArray(bool pos)
{
    if (!bInitialized) return nullptr;
    float4* hdata = !pos ? hPos : hVel;                     // SPH_Util.cpp:48-52: false -> positions
    if (msys) { multiSync();  return hdata; }
    if (sys) {
        int rc = sph_get_array(sys, !pos ? SPH_POS : SPH_VEL, (float*)hdata, 0, (int)scn.params.numParticles);
        if (rc != SPH_OK) err = sph_last_error(sys);
    }
    return hdata;
}

void cSPH::setArray(bool pos, const float4* data, int start, int count)
{
    if (!bInitialized || count <= 0) return;
    if (msys) {
        // a partial write needs the rest of the mirrors current first; the slabs see the result at the next Update
        if (!(start == 0 && count == (int)scn.params.numParticles)) multiSync();
        float4* mirror = !pos ? hPos : hVel;
        if (data != mirror + start) memcpy(mirror + start, data, (size_t)count * sizeof(float4));
        if (start == 0 && count == (int)scn.params.numParticles && !multiDirty) {
            // the other array must be current as well before both are pushed
            float4* keep = !pos ? hVel : hPos;
            sph_multi_get_state(msys, !pos ? nullptr : (float*)keep, !pos ? (float*)keep : nullptr, nullptr, nullptr, count, nullptr);
        }
        multiDirty = true;
        return;
    }
    if (sys) {
        int rc = sph_set_array(sys, !pos ? SPH_POS : SPH_VEL, (const float*)data, start, count);
        if (rc != SPH_OK) err = sph_last_error(sys);
    }
    // keep the host mirror coherent when the caller passed its own buffer
    float4* mirror = !pos ? hPos : hVel;
    if (data != mirror + start) memcpy(mirror + start, data, (size_t)count * sizeof(float4));
}

int cSPH::exchangeArrays(float4* outPos, float4* outVel, const float4* inPos, const float4* inVel)
{
    if (!bInitialized) return SPH_ERR_STATE;
    const size_t bytes = (size_t)scn.params.numParticles * sizeof(float4);
    if (msys || !sys) {                 // several GPUs (or no solver): the plain accessors, one after the other
        float4* p = getArray(false);
        if (p != outPos) memcpy(outPos, p, bytes);
        float4* v = getArray(true);
        if (v != outVel) memcpy(outVel, v, bytes);
        setArray(false, inPos, 0, (int)scn.params.numParticles);
        setArray(true, inVel, 0, (int)scn.params.numParticles);
        return SPH_OK;
    }
    int rc = sph_exchange_arrays(sys, (float*)outPos, (float*)outVel, (const float*)inPos, (const float*)inVel);
    if (rc != SPH_OK) { err = sph_last_error(sys);  return rc; }
    if (inPos != hPos) memcpy(hPos, inPos, bytes);          // the mirrors follow what the device now holds
    if (inVel != hVel) memcpy(hVel, inVel, bytes);
    return SPH_OK;
}

const float4* cSPH::getPosDevice() const
{
    const float* d = nullptr;
    if (sys) sph_device_buffers(sys, &d, nullptr, nullptr, nullptr);
    return (const float4*)d;
}

// ---- checkpoint --------------------------------------------------------------------------------

namespace {
struct CheckpointHeader {
    char magic[8];              // "SPHB200\0"
    unsigned version, numParticles;
    unsigned sizeofScene, sizeofParams;     // layout guard: a file written by another build is rejected
    int emitId, cntRain;
    float fSimTime;
    int curScene;
    float4 colliderPos;         // App::colliderPos / App::dyePos: the targets the per-step prologue drags
    float3 dyePos;              //   scn.params.collPos / dyePos towards (App/Update.cpp:28-61)
    float3 camPosLag, camRotLag;
};
const unsigned kCheckpointVersion = 2;
}

int cSPH::SaveState(const char* path)
{
    if (!bInitialized) return SPH_ERR_STATE;
    const float4* pos = getArray(false);
    const float4* vel = getArray(true);
    FILE* f = fopen(path, "wb");
    if (!f) { err = std::string("cannot write ") + path;  return SPH_ERR_ARG; }
    CheckpointHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "SPHB200", 8);
    h.version = kCheckpointVersion;  h.numParticles = scn.params.numParticles;
    h.sizeofScene = (unsigned)sizeof(Scene);  h.sizeofParams = (unsigned)sizeof(SimParams);
    h.emitId = app.emitId;  h.cntRain = app.cntRain;  h.fSimTime = app.fSimTime;  h.curScene = curScene;
    h.colliderPos = app.colliderPos;  h.dyePos = app.dyePos;  h.camPosLag = app.camPosLag;  h.camRotLag = app.camRotLag;
    const size_t n = scn.params.numParticles;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(&scn, sizeof(Scene), 1, f) == 1 &&
              fwrite(pos, sizeof(float4), n, f) == n && fwrite(vel, sizeof(float4), n, f) == n;
    if (ok && sys) {            // dye concentrations, when the visual outputs are on
        std::vector<float> dye(n);
        unsigned hasDye = sph_get_array(sys, SPH_DYE, dye.data(), 0, (int)n) == SPH_OK ? 1u : 0u;
        ok = fwrite(&hasDye, sizeof hasDye, 1, f) == 1 && (!hasDye || fwrite(dye.data(), sizeof(float), n, f) == n);
    } else if (ok) {
        unsigned hasDye = 0;
        ok = fwrite(&hasDye, sizeof hasDye, 1, f) == 1;
    }
    fclose(f);
    if (!ok) { err = std::string("short write to ") + path;  return SPH_ERR_ARG; }
    return SPH_OK;
}

int cSPH::LoadState(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot read ") + path;  return SPH_ERR_ARG; }
    CheckpointHeader h;
    Scene saved;
    auto reject = [&](const char* why) { fclose(f);  err = std::string(why) + ": " + path;  return SPH_ERR_ARG; };
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SPHB200", 8) != 0) return reject("not a checkpoint");
    if (h.version != kCheckpointVersion || h.sizeofScene != sizeof(Scene) || h.sizeofParams != sizeof(SimParams))
        return reject("checkpoint written by an incompatible build");
    if (fread(&saved, sizeof(Scene), 1, f) != 1 || saved.params.numParticles != h.numParticles) return reject("not a checkpoint");
    // the raw Scene block is only trusted after its indices and sizes have been checked
    if (saved.ca < 0 || saved.ca >= SPH_NUM_ACC || saved.ce < 0 || saved.ce >= NumEmit || saved.params.numParticles == 0 ||
        (unsigned long long)saved.params.gridSize.x * saved.params.gridSize.y * saved.params.gridSize.z != saved.params.numCells)
        return reject("corrupt checkpoint (scene block)");
    for (int e = 0; e < NumEmit; e++)
        if (saved.emit[e].size < 0 || saved.emit[e].size > 10 || saved.emit[e].size2 < 0 || saved.emit[e].size2 > 10)
            return reject("corrupt checkpoint (emitter sizes)");
    saved.title[sizeof saved.title - 1] = 0;
    scn = saved;
    _FreeMem();  _InitMem();                                // buffers for the saved particle / cell counts
    const size_t n = h.numParticles;
    bool ok = fread(hPos, sizeof(float4), n, f) == n && fread(hVel, sizeof(float4), n, f) == n;
    unsigned hasDye = 0;
    std::vector<float> dye;
    if (ok && fread(&hasDye, sizeof hasDye, 1, f) == 1 && hasDye) {
        dye.resize(n);
        ok = fread(dye.data(), sizeof(float), n, f) == n;
    }
    fclose(f);
    if (!ok) { err = std::string("truncated checkpoint: ") + path;  return SPH_ERR_ARG; }
    setArray(0, hPos, 0, (int)n);
    setArray(1, hVel, 0, (int)n);
    if (sys && hasDye) {
        if (sph_set_visual(sys, 1) != SPH_OK || sph_set_dye(sys, dye.data(), 0, (int)n) != SPH_OK) { err = sph_last_error(sys);  return SPH_ERR_CUDA; }
    }
    app.emitId = h.emitId;  app.cntRain = h.cntRain;  app.fSimTime = h.fSimTime;
    app.colliderPos = h.colliderPos;  app.dyePos = h.dyePos;  app.camPosLag = h.camPosLag;  app.camRotLag = h.camRotLag;
    if (h.curScene >= 0 && h.curScene < (int)scenes.size()) curScene = h.curScene;
    app.bChangedAny = false;                                // _InitMem uploaded scn.params
    return sys || msys || device < 0 ? SPH_OK : SPH_ERR_CUDA;
}

// ---- scenes ------------------------------------------------------------------------------------

void cSPH::InitScene()
{
    _FreeMem();
    _InitMem();
    Reset(scn.initType);

    app.camPosLag = scn.camPos;
    app.camRotLag = scn.camRot;
    for (int i = 0; i < NumEmit; i++) {
        scn.emit[i].posLag = scn.emit[i].pos;
        scn.emit[i].rotLag = scn.emit[i].rot;
    }
    for (int i = 0; i < SPH_NUM_ACC; i++) scn.accPos[i] = scn.params.acc[i].pos;
    app.colliderPos = scn.params.collPos;
    scn.params.dyePos = app.dyePos;                         // SPH_Scenes.cpp:23
    app.emitId = 0;
}

void cSPH::UpdScene()
{
    scn = scenes[curScene];
    InitScene();
}

void cSPH::NextScene(bool chapter)
{
    do { curScene++;  if (curScene >= (int)scenes.size()) curScene = 0; } while (chapter && !scenes[curScene].bChapter);
    UpdScene();
}

void cSPH::PrevScene(bool chapter)
{
    do { curScene--;  if (curScene < 0) curScene = (int)scenes.size() - 1; } while (chapter && !scenes[curScene].bChapter);
    UpdScene();
}

void cSPH::LoadScenes()
{
    scenes.clear();
    sphxml::Document file;
    file.LoadFile(xmlPath.c_str());
    const sphxml::Element* root = file.RootElement();
    const sphxml::Element* s = nullptr;
    if (!root) {
        err = "cannot load " + xmlPath + ": " + file.error;
        fprintf(stderr, "\nError!  Can't load %s (%s)\n", xmlPath.c_str(), file.error.c_str());
    } else {
        s = root->FirstChildElement("Scene");
        if (!s) fprintf(stderr, "Warning:  No <Scene> in xml.\n");
    }

    int i = -1, ch = 0;
    curScene = 0;
    while (s) {
        if (s->Attribute("default") || s->Attribute("def")) curScene = i + 1;
        Scene sc(s);
        scenes.push_back(sc);
        if (sc.bChapter) ch++;
        s = s->NextSiblingElement("Scene");
        i++;
    }
    if (i == -1) { Scene sc;  scenes.push_back(sc);  scn = sc; }
    else scn = scenes[curScene];
    scenes[0].bChapter = true;
    InitScene();
}

SphOptions cSPH::LoadOptions(const char* scenesXmlPath)
{
    SphOptions o;
    sphxml::Document file;
    file.LoadFile(scenesXmlPath ? scenesXmlPath : "Scenes.xml");
    const sphxml::Element* root = file.RootElement();
    const sphxml::Element* opt = root ? root->FirstChildElement("Options") : nullptr;
    if (!opt) return o;
    const char* a;
    auto toInt = [](const char* str) { return (int)strtol(str, nullptr, 0); };
    if ((a = opt->Attribute("Windowed")))  o.bWindowed = toInt(a) > 0;
    if ((a = opt->Attribute("WSizeX")))    o.WSizeX = toInt(a);
    if ((a = opt->Attribute("WSizeY")))    o.WSizeY = toInt(a);
    if ((a = opt->Attribute("VSyncOff")))  o.bVsyncOff = toInt(a) > 0;
    if ((a = opt->Attribute("timAvgCnt"))) o.timAvgCnt = toInt(a);
    if ((a = opt->Attribute("barsScale"))) o.barsScale = (float)toInt(a);
    if ((a = opt->Attribute("showInfo")))  o.bShowInfo = toInt(a) > 0;
    return o;
}
