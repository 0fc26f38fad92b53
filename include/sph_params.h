/* sph_params.h -- the solver parameter block shared by host and device code.
 *
 * This is an ABI struct: its field names, order, types and 560-byte layout are those of the
 * reference's `struct SimParams` (reference source/CUDA/Params.cuh:53-114) because the layers
 * above the solver (sliders, scene loader, emitter prologue) read and write the fields by name
 * through `cSPH::scn.params`, and because the parity oracle memcpy()s the block.  The
 * static_asserts at the end pin the layout.  Enumerator values follow Params.cuh:24-42.
 *
 * Plain C / C++ / CUDA.  Needs only <vector_types.h> from the CUDA toolkit for float3/float4
 * (float4 is 16-byte aligned; collPos sits at offset 208 and the block is padded to 560).
 */
#ifndef SPH_PARAMS_H
#define SPH_PARAMS_H

#include <vector_types.h>
#include <stddef.h>

#ifndef SPH_NO_REFERENCE_NAMES      /* names the reference's upper layers use unqualified */
typedef unsigned int uint;
#ifndef PI
#define PI   3.141592654f           /* Params.cuh:18 -- deliberately the float literal */
#endif
#ifndef PI2
#define PI2  2.f*PI
#endif
#endif

/* boundary shape (SimParams::bndType) */
enum BndType {
    BND_BOX = 0, BND_CYL_Y, BND_CYL_Z, BND_CYL_YZ, BND_SPHERE, BND_PUMP_Y,
    BND_ALL, BND_DW = 0xFFFFffff
};
/* what happens at the Z ends (SimParams::bndEffZ) */
enum BndEff {
    BND_EFF_NONE = 0, BND_EFF_WRAP, BND_EFF_CYCLE, BND_EFF_WAVE,
    BEF_ALL, BEF_DW = 0xFFFFffff
};
/* colouring mode of the (visual-only) colour output */
enum ClrType {
    CLR_Dens = 0, CLR_Accel, CLR_DensAcc, CLR_Vel, CLR_VelAcc, CLR_VelRGB, CLR_None,
    CLR_ALL, CLR_DW = 0xFFFFffff
};
/* accelerator volume shape */
enum AccType {
    ACC_Off = 0, ACC_Box, ACC_CylY, ACC_CylYsm,
    ACC_ALL, ACC_DW = 0xFFFFffff
};

#define SPH_NUM_ACC 4
#ifdef __cplusplus
static const int NumAcc = SPH_NUM_ACC;
#endif

struct Accel {
    float3 pos;             /* centre                                  */
    float3 size;            /* half extents / radii                    */
    float3 acc;             /* acceleration added inside the volume    */
    enum AccType type;
};

struct SimParams {
    /* -- simulation ------------------------------------------------------------- */
    float timeStep;
    uint  numParticles;
    uint  maxParInCell;             /* neighbour walk visits at most this many per cell */
    float3 gravity;
    float globalDamping;

    /* -- uniform grid ----------------------------------------------------------- */
    uint3  gridSize;
    float3 cellSize;
    uint   gridSize_yx;             /* gridSize.y * gridSize.x                   */
    uint   numCells;

    /* -- world box; the *D variants are inset by the soft boundary (drawing) ---- */
    float3 worldMin, worldMax, worldSize;
    float3 worldMinD, worldMaxD, worldSizeD;

    /* -- SPH kernel constants --------------------------------------------------- */
    float particleR;
    float h, h2;                    /* smoothing radius and its square           */
    float SpikyKern, LapKern, Poly6Kern;

    /* -- fluid ------------------------------------------------------------------ */
    float particleMass, restDensity;
    float stiffness, viscosity;
    float minDens, minDist;         /* stability clamps                          */

    /* -- boundary --------------------------------------------------------------- */
    float distBndHard, distBndSoft;
    float bndDamp, bndStiff, bndDampC;
    enum BndType bndType;
    enum BndEff  bndEffZ;

    /* -- sphere collider -------------------------------------------------------- */
    float4 collPos;
    float  collR;
    float  spring, damping, shear;

    /* -- visual (colour output only) -------------------------------------------- */
    enum ClrType clrType;
    int   iHue;
    float brightness, contrast;

    /* -- dye -------------------------------------------------------------------- */
    int   dyeType, dyeClear;
    float dyeFade;
    float3 dyePos, dyeSize;

    /* -- accelerators, height map ------------------------------------------------ */
    struct Accel acc[SPH_NUM_ACC];
    int   iHmap;

    /* -- pump boundary ----------------------------------------------------------- */
    float angOut, hClose, radIn;
    float rVexit, rDexit;
    float s1, s2, s3, s4, s5, s6;

    /* -- rotor / propeller (also aliased by the wave and height-map scenes) ------- */
    float rAngle, rTwist;
    int   rotType, rotBlades;
    int3  rotSize;
    float rotR, rotSpc;
    float r2Dist, r2Angle, r2twist, ff2;
};

#ifdef __cplusplus
static_assert(sizeof(struct Accel) == 40, "Accel layout");
static_assert(sizeof(struct SimParams) == 560, "SimParams must stay 560 bytes (reference ABI)");
static_assert(offsetof(struct SimParams, gridSize)   == 28,  "SimParams layout");
static_assert(offsetof(struct SimParams, worldMin)   == 60,  "SimParams layout");
static_assert(offsetof(struct SimParams, particleR)  == 132, "SimParams layout");
static_assert(offsetof(struct SimParams, collPos)    == 208, "SimParams layout");
static_assert(offsetof(struct SimParams, acc)        == 292, "SimParams layout");
static_assert(offsetof(struct SimParams, rAngle)     == 500, "SimParams layout");
#endif

#endif /* SPH_PARAMS_H */
