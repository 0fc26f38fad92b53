// host_capi.cpp -- flat C entry points over the C++ host layer (cSPH / Scene), so that the
// Python tests and bench.py can drive the same objects the reference's App layer would.
// Declared in include/sph_host_c.h.
#include "sph_host.h"
#include "sph_host_c.h"
#include "xml_lite.h"
#include <cstring>
#include <cstdlib>

extern "C" {

sphh_t* sphh_create(const char* scenesXmlPath, int device)
{
    return reinterpret_cast<sphh_t*>(new cSPH(scenesXmlPath, device));
}

sphh_t* sphh_create_multi(const char* scenesXmlPath, const int* devices, int ndev)
{
    return reinterpret_cast<sphh_t*>(new cSPH(scenesXmlPath, devices, ndev));
}

void sphh_destroy(sphh_t* h) { delete reinterpret_cast<cSPH*>(h); }

static cSPH* S(sphh_t* h) { return reinterpret_cast<cSPH*>(h); }

sph_multi_t* sphh_multi_solver(sphh_t* h) { return S(h)->multiSolver(); }
int sphh_exchange_arrays(sphh_t* h, float* outPos, float* outVel, const float* inPos, const float* inVel)
{
    return S(h)->exchangeArrays((float4*)outPos, (float4*)outVel, (const float4*)inPos, (const float4*)inVel);
}
int sphh_num_scenes(sphh_t* h) { return (int)S(h)->scenes.size(); }
int sphh_cur_scene(sphh_t* h) { return S(h)->curScene; }
const char* sphh_last_error(sphh_t* h) { return S(h)->lastError(); }
const char* sphh_scene_title(sphh_t* h, int idx) { return S(h)->scenes[idx].title; }

void sphh_scene_params(sphh_t* h, int idx, struct SimParams* out) { *out = S(h)->scenes[idx].params; }

static void scene_extra(const Scene& s, float* o)
{
    int k = 0;
    o[k++] = s.initMin.x; o[k++] = s.initMin.y; o[k++] = s.initMin.z;
    o[k++] = s.initMax.x; o[k++] = s.initMax.y; o[k++] = s.initMax.z;
    o[k++] = (float)s.initType; o[k++] = (float)s.initLast; o[k++] = s.spacing; o[k++] = s.fCellSize;
    o[k++] = s.dropR; o[k++] = (float)s.rain; o[k++] = s.rVel; o[k++] = s.r2Vel;
    o[k++] = s.camPos.x; o[k++] = s.camPos.y; o[k++] = s.camPos.z; o[k++] = s.camRot.x; o[k++] = s.camRot.y;
    o[k++] = s.bChapter ? 1.f : 0.f;
    for (int e = 0; e < NumEmit; e++) {
        const Emitter& m = s.emit[e];
        o[k++] = m.pos.x; o[k++] = m.pos.y; o[k++] = m.pos.z; o[k++] = m.rot.x; o[k++] = m.rot.y;
        o[k++] = m.vel; o[k++] = (float)m.size; o[k++] = (float)m.size2;
    }
    while (k < 64) o[k++] = 0.f;
}
void sphh_scene_extra(sphh_t* h, int idx, float* out64) { scene_extra(S(h)->scenes[idx], out64); }
void sphh_live_extra(sphh_t* h, float* out64) { scene_extra(S(h)->scn, out64); }

void sphh_live_params(sphh_t* h, struct SimParams* out) { *out = S(h)->scn.params; }
void sphh_set_live_params(sphh_t* h, const struct SimParams* in)
{
    S(h)->scn.params = *in;
    S(h)->app.bChangedAny = true;
}

int sphh_select_scene(sphh_t* h, int idx)
{
    cSPH* s = S(h);
    if (idx < 0 || idx >= (int)s->scenes.size()) return -1;
    s->curScene = idx;
    s->UpdScene();
    return (int)s->scn.params.numParticles;
}
// append a scene built from a <Scene .../> element given as text (synthetic scale-ups go through
// the same Scene path: every constant is re-derived by Scene::Update); returns its index
int sphh_add_scene_xml(sphh_t* h, const char* sceneElementXml)
{
    cSPH* s = S(h);
    sphxml::Document doc;
    if (!doc.Parse(sceneElementXml) || !doc.RootElement()) return -1;
    Scene sc(doc.RootElement());
    s->scenes.push_back(sc);
    return (int)s->scenes.size() - 1;
}
void sphh_next_scene(sphh_t* h, int chapter) { S(h)->NextScene(chapter != 0); }
void sphh_prev_scene(sphh_t* h, int chapter) { S(h)->PrevScene(chapter != 0); }

void sphh_host_arrays(sphh_t* h, float* pos, float* vel)
{
    cSPH* s = S(h);
    size_t n = s->scn.params.numParticles;
    if (pos) memcpy(pos, s->hPos, n * sizeof(float4));
    if (vel) memcpy(vel, s->hVel, n * sizeof(float4));
}
void sphh_reset(sphh_t* h, int type) { S(h)->Reset(type); }
int sphh_drop(sphh_t* h, int bRandom) { S(h)->Drop(bRandom != 0);  return S(h)->app.emitId; }
int sphh_emit_id(sphh_t* h) { return S(h)->app.emitId; }
void sphh_srand(unsigned seed) { srand(seed); }
void sphh_update_emitter(sphh_t* h) { S(h)->UpdateEmitter(); }
int sphh_update(sphh_t* h, int nsteps) { return S(h)->Update(nsteps); }
void sphh_mark_changed(sphh_t* h) { S(h)->app.bChangedAny = true; }

int sphh_get_array(sphh_t* h, int velocities, float* out)
{
    cSPH* s = S(h);
    float4* p = s->getArray(velocities != 0);
    if (!p) return -1;
    memcpy(out, p, (size_t)s->scn.params.numParticles * sizeof(float4));
    return 0;
}
void sphh_set_array(sphh_t* h, int velocities, const float* data, int start, int count)
{
    S(h)->setArray(velocities != 0, (const float4*)data, start, count);
}
sph_t* sphh_solver(sphh_t* h) { return S(h)->solver(); }
int sphh_save_state(sphh_t* h, const char* path) { return S(h)->SaveState(path); }
int sphh_load_state(sphh_t* h, const char* path) { return S(h)->LoadState(path); }

// the targets the per-step prologue drags collPos / dyePos / acc[ca].pos towards (the reference's mouse handlers
// write them, App/Input.cpp); null pointers leave a target alone
void sphh_set_targets(sphh_t* h, const float* collider4, const float* dye3, const float* acc3)
{
    cSPH* s = S(h);
    if (collider4) { s->app.colliderPos.x = collider4[0]; s->app.colliderPos.y = collider4[1]; s->app.colliderPos.z = collider4[2]; s->app.colliderPos.w = collider4[3]; }
    if (dye3) { s->app.dyePos.x = dye3[0]; s->app.dyePos.y = dye3[1]; s->app.dyePos.z = dye3[2]; }
    if (acc3) { float3& a = s->scn.accPos[s->scn.ca];  a.x = acc3[0]; a.y = acc3[1]; a.z = acc3[2]; }
}
void sphh_set_emitter(sphh_t* h, int e, const float* posLag3, const float* rotLag2, float vel, int size, int size2)
{
    if (e < 0 || e >= NumEmit) return;
    Emitter& em = S(h)->scn.emit[e];
    em.posLag.x = posLag3[0]; em.posLag.y = posLag3[1]; em.posLag.z = posLag3[2];
    em.rotLag.x = rotLag2[0]; em.rotLag.y = rotLag2[1]; em.rotLag.z = 0.f;
    em.vel = vel;  em.size = size;  em.size2 = size2;
}
int sphh_cnt_rain(sphh_t* h) { return S(h)->app.cntRain; }
int sphh_changed_flag(sphh_t* h, int clear) { int v = S(h)->app.bChangedAny ? 1 : 0;  if (clear) S(h)->app.bChangedAny = false;  return v; }
int sphh_register_gl(sphh_t* h, unsigned posVbo, unsigned colorVbo) { return S(h)->registerGLBuffers(posVbo, colorVbo); }
unsigned sphh_pos_buffer(sphh_t* h) { return S(h)->getPosBuffer(); }
double sphh_timer_fps(sphh_t* h) { return S(h)->tim.FR; }

void sphh_load_options(const char* scenesXmlPath, int* out7)
{
    SphOptions o = cSPH::LoadOptions(scenesXmlPath);
    out7[0] = o.bWindowed; out7[1] = o.WSizeX; out7[2] = o.WSizeY; out7[3] = o.bVsyncOff;
    out7[4] = o.timAvgCnt; out7[5] = (int)o.barsScale; out7[6] = o.bShowInfo;
}

}  // extern "C"
