// sph_scene.cpp -- scene defaults, Scenes.xml attributes -> SimParams, derived constants.
//
// Behaviour follows the reference's scene layer, because every derived constant feeds the kernels
// and must come out bit-identical (tests/test_scene_host.py pins all scenes of the reference's
// Scenes.xml against vectors produced by the reference code itself):
//   Scene::InitDefault   source/SPH/Scene.cpp:7-66
//   Scene::_UpdatePar    source/SPH/Scene.cpp:76-94
//   Scene::_UpdateGrid   source/SPH/Scene.cpp:98-113
//   Scene::_FromXML      source/SPH/Scene_Load.cpp:11-112
// The float/double promotion of each expression is spelled out where it matters.
#include "sph_host.h"
#include "xml_lite.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

inline float3 f3(float x, float y, float z) { float3 v; v.x = x; v.y = y; v.z = z; return v; }
inline float4 f4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline float3 scaled(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }

// "x y z" -> float3; components that are not present keep their previous (here: zero) value.
// (The reference's toVec3, pch/header.h:148-153, leaves them uninitialised.)
float3 parse_vec3(const char* str)
{
    float3 v = f3(0.f, 0.f, 0.f);
    sscanf(str, "%f %f %f", &v.x, &v.y, &v.z);
    return v;
}
inline float parse_float(const char* str) { return (float)atof(str); }          // header.h:155
inline int parse_int(const char* str) { return (int)strtol(str, nullptr, 0); }  // header.h:156

}  // namespace

Emitter::Emitter() : vel(1.5f), size(0), size2(0)
{
    pos = f3(0.f, 0.f, 0.f);
    rot = f3(0.f, 0.f, 0.f);
    posLag = pos;
    rotLag = rot;
}

SphAppState::SphAppState()
{
    camPosLag = camRotLag = dyePos = f3(0.f, 0.f, 0.f);
    colliderPos = f4(0.f, 0.f, 0.f, 0.f);
}

Scene::Scene()
{
    InitDefault();
    Update();
}

Scene::Scene(const sphxml::Element* s)
{
    InitDefault();
    _FromXML(s);
    Update();
}

void Scene::InitDefault()
{
    memset(&params, 0, sizeof params);      // the reference leaves unset fields (ff2, padding) undefined
    SimParams& p = params;

    // simulation
    p.numParticles = 7 * 8 * 1024;
    p.maxParInCell = 16;
    p.timeStep = 0.0026f;
    p.globalDamping = 1.00f;
    p.gravity = f3(0, (float)-9.81, 0);

    // SPH
    p.particleR = (float)0.004;
    p.minDist = 1.0f;           // x particleR
    p.h = (float)0.01;
    spacing = (float)1.38;      // x particleR
    p.restDensity = 1000;
    p.minDens = 1.0f;           // x restDensity
    p.stiffness = 3.0f;
    p.viscosity = 0.5f;

    // world
    float3 w = f3((float)0.2, (float)0.25, (float)0.2);
    p.worldMin = f3(-w.x, -w.y, -w.z);
    p.worldMax = w;
    initMin = p.worldMin;
    initMax = w;
    fCellSize = (float)((double)p.particleR * 2.0);
    dropR = 5;
    rain = 0;
    initType = 0;
    initLast = 1;

    // boundary (distances x particleR)
    p.distBndSoft = 8;
    p.distBndHard = 1;
    p.bndStiff = 30000;
    p.bndDamp = 256;
    p.bndDampC = 60;
    p.bndType = BND_BOX;
    p.bndEffZ = BND_EFF_NONE;
    p.iHmap = 0;

    // collider; x = -77 means "place it from the world box" (see _UpdateGrid)
    p.collR = (float)0.05;
    p.collPos = f4(-77.f, 0.f, 0.f, 0.f);
    p.spring = 80;
    p.damping = (float)0.02;
    p.shear = (float)0.1;

    // pump
    p.angOut = (float)0.7;  p.hClose = (float)-0.03;  p.radIn = (float)0.4;
    p.s1 = (float)0.90;  p.s2 = (float)0.90;  p.s3 = (float)1.198;
    p.s4 = (float)0.7782;  p.s5 = (float)0.8696;  p.s6 = (float)0.843;
    p.rVexit = (float)0.03;  p.rDexit = (float)0.04;
    // rotor
    p.rotType = 0;  p.rAngle = 0;
    p.rotBlades = 4;  p.rTwist = (float)0.04;
    p.rotSize.x = 0;  p.rotSize.y = 4;  p.rotSize.z = 12;
    p.rotR = 3;  p.rotSpc = (float)0.7;
    p.r2Dist = 0;  p.r2Angle = 0;  p.r2twist = 1;
    rVel = 0;  r2Vel = 1;

    // camera
    strcpy(title, "No name");
    bChapter = false;
    camPos = f3(0, (float)0.04, (float)-0.44);
    camRot = f3(5, 0, 0);
    collidPos = f4(0, 0, 0, 0);

    ce = 0;  ca = 0;
    for (int i = 0; i < SPH_NUM_ACC; i++) {
        p.acc[i].pos = f3(0, 0, 0);
        p.acc[i].size = f3((float)0.005, (float)0.010, (float)0.005);
        p.acc[i].acc = f3(0, 10, 0);
        p.acc[i].type = ACC_Off;
        accPos[i] = f3(0, 0, 0);
    }

    // visual
    p.clrType = CLR_VelAcc;  p.brightness = (float)0.3;  p.contrast = (float)0.3;  p.iHue = 0;
    p.dyeType = 0;  p.dyeClear = 2;  p.dyeFade = 1.f;
    p.dyePos = f3(0, (float)-0.21, 0);
    p.dyeSize = f3((float)0.008, (float)0.012, (float)0.008);
}

void Scene::Update()
{
    _UpdatePar();
    _UpdateGrid();
}

void Scene::_UpdatePar()
{
    SimParams& p = params;
    p.minDist *= p.particleR;
    p.h2 = p.h * p.h;
    spacing *= p.particleR;

    // pow(float,int) promotes to double: the kernel norms are evaluated in double and rounded once
    const float pi = PI;
    p.Poly6Kern = (float)((double)315.0f / ((double)(64.0f * pi) * pow((double)p.h, 9.0)));
    p.SpikyKern = (float)((double)(-0.5f * -45.0f) / ((double)pi * pow((double)p.h, 6.0)));
    p.LapKern   = (float)((double)45.0f / ((double)pi * pow((double)p.h, 6.0)));

    // pow(float,float) stays float
    p.minDens = 1.f / powf(p.minDens * p.restDensity, 2.f);
    p.particleMass = p.restDensity * 4.f / 3.f * pi * powf(p.particleR, 3.f);

    p.distBndSoft *= p.particleR;
    p.distBndHard *= p.particleR;
    p.rotR *= p.particleR;
    p.rotSpc *= p.rotR;
}

void Scene::_UpdateGrid()
{
    SimParams& p = params;
    const float b = p.distBndSoft - p.particleR;
    p.worldMinD = f3(p.worldMin.x + b, p.worldMin.y + b, p.worldMin.z + b);
    p.worldMaxD = f3(p.worldMax.x - b, p.worldMax.y - b, p.worldMax.z - b);
    initMin = f3(initMin.x + b, initMin.y + b, initMin.z + b);
    initMax = f3(initMax.x - b, initMax.y - b, initMax.z - b);
    p.worldSize = f3(p.worldMax.x - p.worldMin.x, p.worldMax.y - p.worldMin.y, p.worldMax.z - p.worldMin.z);
    p.worldSizeD = f3(p.worldMaxD.x - p.worldMinD.x, p.worldMaxD.y - p.worldMinD.y, p.worldMaxD.z - p.worldMinD.z);

    p.cellSize = f3(fCellSize, fCellSize, fCellSize);
    p.gridSize.x = (uint)ceilf(p.worldSize.x / p.cellSize.x);
    p.gridSize.y = (uint)ceilf(p.worldSize.y / p.cellSize.y);
    p.gridSize.z = (uint)ceilf(p.worldSize.z / p.cellSize.z);
    p.gridSize_yx = p.gridSize.y * p.gridSize.x;
    p.numCells = p.gridSize.x * p.gridSize.y * p.gridSize.z;

    if (p.collPos.x == -77.f)
        p.collPos = f4(p.worldMin.x + b - p.collR * 1.2f, p.worldMin.y + b + p.collR * 1.f, 0, 1);
}

void Scene::_FromXML(const sphxml::Element* s)
{
    SimParams& p = params;
    const char* a;

    auto F = [&](const sphxml::Element* e, const char* key, float& dst) { if ((a = e->Attribute(key))) dst = parse_float(a); };
    auto I = [&](const sphxml::Element* e, const char* key, int& dst)   { if ((a = e->Attribute(key))) dst = parse_int(a); };
    auto U = [&](const sphxml::Element* e, const char* key, uint& dst)  { if ((a = e->Attribute(key))) dst = (uint)parse_int(a); };
    auto V = [&](const sphxml::Element* e, const char* key, float3& dst){ if ((a = e->Attribute(key))) dst = parse_vec3(a); };

    // world / init volume: "World" sets both, "Init" only the init volume, Min/Max override
    if ((a = s->Attribute("World"))) {
        float3 sw = parse_vec3(a);
        p.worldMin = scaled(sw, -0.5f);  p.worldMax = scaled(sw, 0.5f);
        initMin = scaled(sw, -0.5f);     initMax = scaled(sw, 0.5f);
    }
    if ((a = s->Attribute("Init"))) {
        float3 si = parse_vec3(a);
        initMin = scaled(si, -0.5f);     initMax = scaled(si, 0.5f);
    }
    V(s, "WorldMin", p.worldMin);  V(s, "InitMin", initMin);
    V(s, "WorldMax", p.worldMax);  V(s, "InitMax", initMax);
    I(s, "InitType", initType);    I(s, "InitLast", initLast);
    F(s, "CellSize", fCellSize);

    // title: the reference reads "name" on Windows only (Scene_Load.cpp:37-39); harmless to keep
    if ((a = s->Attribute("name"))) { strncpy(title, a, sizeof(title) - 1); title[sizeof(title) - 1] = 0; }
    if (s->Attribute("chapter")) bChapter = true;
    V(s, "CamPos", camPos);  V(s, "CamRot", camRot);
    F(s, "dropR", dropR);    I(s, "rain", rain);

    // emitter 0 as attributes, then up to NumEmit <Emitter> children
    I(s, "EmitSize", emit[0].size);    V(s, "EmitPos", emit[0].pos);  F(s, "EmitVel", emit[0].vel);
    I(s, "EmitSize2", emit[0].size2);  V(s, "EmitRot", emit[0].rot);
    {
        int i = 0;
        for (const sphxml::Element* e = s->FirstChildElement("Emitter"); e && i < NumEmit; e = e->NextSiblingElement("Emitter"), i++) {
            I(e, "size", emit[i].size);    V(e, "pos", emit[i].pos);  F(e, "vel", emit[i].vel);
            I(e, "size2", emit[i].size2);  V(e, "rot", emit[i].rot);
        }
    }
    // up to NumAcc <Accel> children
    {
        int i = 0;
        for (const sphxml::Element* e = s->FirstChildElement("Accel"); e && i < SPH_NUM_ACC; e = e->NextSiblingElement("Accel"), i++) {
            if ((a = e->Attribute("type"))) p.acc[i].type = (AccType)parse_int(a);
            V(e, "pos", p.acc[i].pos);  V(e, "size", p.acc[i].size);  V(e, "acc", p.acc[i].acc);
        }
    }

    // collider
    F(s, "ColliderR", p.collR);
    if ((a = s->Attribute("ColliderPos"))) { float3 v = parse_vec3(a); p.collPos = f4(v.x, v.y, v.z, 1); }
    F(s, "spring", p.spring);  F(s, "damping", p.damping);  F(s, "shear", p.shear);

    // simulation
    U(s, "Particles", p.numParticles);
    if ((a = s->Attribute("ParticlesK"))) p.numParticles = (uint)(parse_int(a) * 1024);
    U(s, "maxParInCell", p.maxParInCell);
    F(s, "TimeStep", p.timeStep);  F(s, "globalDamping", p.globalDamping);
    V(s, "Gravity", p.gravity);

    // SPH
    F(s, "particleR", p.particleR);  F(s, "minDist", p.minDist);
    F(s, "particleH", p.h);          F(s, "spacing", spacing);
    F(s, "RestDensity", p.restDensity);  F(s, "minDens", p.minDens);
    F(s, "Stiffness", p.stiffness);      F(s, "Viscosity", p.viscosity);

    // boundary
    F(s, "distBndSoft", p.distBndSoft);  F(s, "bndStiff", p.bndStiff);
    F(s, "distBndHard", p.distBndHard);  F(s, "bndDamp", p.bndDamp);  F(s, "bndDampC", p.bndDampC);
    if ((a = s->Attribute("bndType"))) p.bndType = (BndType)parse_int(a);
    if ((a = s->Attribute("bndEffZ"))) p.bndEffZ = (BndEff)parse_int(a);
    I(s, "HmapType", p.iHmap);

    // pump
    F(s, "PumpAngOut", p.angOut);  F(s, "PumpHClose", p.hClose);  F(s, "PumpRadIn", p.radIn);
    F(s, "ExitVel", p.rVexit);     F(s, "ExitDist", p.rDexit);
    F(s, "s1", p.s1);  F(s, "s2", p.s2);  F(s, "s3", p.s3);  F(s, "s4", p.s4);  F(s, "s5", p.s5);  F(s, "s6", p.s6);
    // rotor
    if ((a = s->Attribute("RotorSizes"))) { float3 v = parse_vec3(a); p.rotSize.x = (int)v.x; p.rotSize.y = (int)v.y; p.rotSize.z = (int)v.z; }
    I(s, "RotorType", p.rotType);  F(s, "Rotor2Dist", p.r2Dist);
    F(s, "RotorVel", rVel);        F(s, "Rotor2Vel", r2Vel);
    if ((a = s->Attribute("RotorAngle")))  p.rAngle  = parse_float(a) * PI / 180.f;
    if ((a = s->Attribute("Rotor2Angle"))) p.r2Angle = parse_float(a) * PI / 180.f;
    F(s, "colParR", p.rotR);       F(s, "colParSpc", p.rotSpc);
    I(s, "RotorBlades", p.rotBlades);  F(s, "RotorTwist", p.rTwist);

    // waves: aliases onto the rotor fields (Scene_Load.cpp:105-106)
    if ((a = s->Attribute("WaveSpeed"))) { rVel = parse_float(a);  p.rAngle = -PI / 2.f; }
    F(s, "WaveAmpl", p.rTwist);  F(s, "HSlope", p.r2Angle);
    // height map: aliases (Scene_Load.cpp:109-111)
    F(s, "Hheight", p.r2Angle);  F(s, "Hscale", p.hClose);
    F(s, "HxFq", p.s1);  F(s, "HxOfs", p.s2);
    F(s, "HzFq", p.s3);  F(s, "HzOfs", p.s4);  F(s, "HholeR", p.s5);
}
