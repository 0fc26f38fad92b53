/* sph_host_c.h -- flat C view of the C++ host layer in sph_host.h (class cSPH / Scene), for callers
 * that cannot include C++ (ctypes, cgo, JNI).  One sphh_t == one cSPH object == what the reference's
 * App layer holds in App::psys (source/App/App.h:44).
 *
 * Reference interface each group replaces:
 *   sphh_create/destroy              new cSPH() / delete            source/App/App.cpp:67,95
 *   sphh_num_scenes .. prev_scene    cSPH::scenes, curScene, UpdScene, Next/PrevScene   source/SPH/SPH_Scenes.cpp:30-84
 *   sphh_scene_params / live_params  Scene::params (scenes[i] / scn)                    source/SPH/Scene.h:33
 *   sphh_reset / sphh_drop           cSPH::Reset / cSPH::Drop                           source/SPH/SPH_Init.cpp:23-118
 *   sphh_update_emitter              App::UpdateEmitter                                 source/App/Update.cpp:9-97
 *   sphh_update                      cSPH::Update                                       source/SPH/SPH_Update.cpp:12-81
 *   sphh_get_array / set_array       cSPH::getArray / setArray (same inverted flag)     source/SPH/SPH_Util.cpp:44-71
 *   sphh_load_options                cSPH::LoadOptions                                  source/SPH/SPH_Scenes.cpp:89-111
 */
#ifndef SPH_HOST_C_H
#define SPH_HOST_C_H
#include "sph_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sphh_system sphh_t;

/* device >= 0: solver on that GPU; device < 0: scene / initialiser layer only (no GPU touched) */
sphh_t* sphh_create(const char* scenesXmlPath, int device);
/* several GPUs of one box (z slabs, multi-GPU driver of sph_b200.h); same interface, results identical to one GPU */
sphh_t* sphh_create_multi(const char* scenesXmlPath, const int* devices, int ndev);
sph_multi_t* sphh_multi_solver(sphh_t* h);                         /* NULL unless created with a device list */
void sphh_destroy(sphh_t* h);
const char* sphh_last_error(sphh_t* h);

int sphh_num_scenes(sphh_t* h);
int sphh_cur_scene(sphh_t* h);
const char* sphh_scene_title(sphh_t* h, int idx);
void sphh_scene_params(sphh_t* h, int idx, struct SimParams* out);
/* 64 floats: initMin3 initMax3 initType initLast spacing fCellSize dropR rain rVel r2Vel camPos3 camRot2
 * bChapter, then per emitter pos3 rot2 vel size size2 */
void sphh_scene_extra(sphh_t* h, int idx, float* out64);
void sphh_live_extra(sphh_t* h, float* out64);
void sphh_live_params(sphh_t* h, struct SimParams* out);
void sphh_set_live_params(sphh_t* h, const struct SimParams* in);   /* scn.params = *in; marks changed */
int  sphh_select_scene(sphh_t* h, int idx);                         /* curScene = idx; UpdScene(); returns numParticles */
int  sphh_add_scene_xml(sphh_t* h, const char* sceneElementXml);    /* returns the new scene's index */
void sphh_next_scene(sphh_t* h, int chapter);
void sphh_prev_scene(sphh_t* h, int chapter);

void sphh_host_arrays(sphh_t* h, float* pos, float* vel);           /* copies of hPos / hVel */
void sphh_reset(sphh_t* h, int type);
int  sphh_drop(sphh_t* h, int bRandom);                             /* returns the ring index after the drop */
int  sphh_emit_id(sphh_t* h);
void sphh_srand(unsigned seed);
void sphh_update_emitter(sphh_t* h);
int  sphh_update(sphh_t* h, int nsteps);
void sphh_mark_changed(sphh_t* h);                                  /* ParamBase::bChangedAny = true */

int  sphh_get_array(sphh_t* h, int velocities, float* out);
/* cSPH::exchangeArrays: current positions + velocities out, new ones in, downloads overlapping uploads */
int  sphh_exchange_arrays(sphh_t* h, float* outPos, float* outVel, const float* inPos, const float* inVel);
void sphh_set_array(sphh_t* h, int velocities, const float* data, int start, int count);
sph_t* sphh_solver(sphh_t* h);
int  sphh_save_state(sphh_t* h, const char* path);                  /* checkpoint: params, pos, vel, ring counters */
int  sphh_load_state(sphh_t* h, const char* path);
void sphh_load_options(const char* scenesXmlPath, int* out7);

/* state the reference keeps in App:: statics and the UI writes (App/Input.cpp): drag targets of the per-step
 * prologue, emitter lag state, rain counter, ParamBase::bChangedAny */
void sphh_set_targets(sphh_t* h, const float* collider4, const float* dye3, const float* acc3);
void sphh_set_emitter(sphh_t* h, int e, const float* posLag3, const float* rotLag2, float vel, int size, int size2);
int  sphh_cnt_rain(sphh_t* h);
int  sphh_changed_flag(sphh_t* h, int clear);
/* renderer hand-over: psys->colorVbo / getPosBuffer() (App/App.cpp:70, App/Render.cpp:12); psys->tim.FR */
int  sphh_register_gl(sphh_t* h, unsigned posVbo, unsigned colorVbo);
unsigned sphh_pos_buffer(sphh_t* h);
double sphh_timer_fps(sphh_t* h);

#ifdef __cplusplus
}
#endif
#endif
