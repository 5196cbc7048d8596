/*
 * pt_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the Monte-Carlo estimator of Zydak/Vulkan-Path-Tracer
 * (reference paths relative to /root/reference):
 *   PathTracer/Shaders/{Defines,Sampler,RTCommon,Surface,Material,ClosestHit,Miss,RayGen}.slang
 *   PathTracer/Shaders/PostProcess/{BloomDownSample,BloomUpSample,Tonemap}.slang
 *   PathTracer/PathTracer.cpp:122-156 (dispatch bookkeeping), :1137-1332 (env alias table)
 *   PathTracer/PostProcessor.cpp:128-246, PathTracer/FlyCamera.cpp:84-140
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (vulkan-path-tracer_b200/) never links or calls it.
 *
 * PARITY PIN STATUS: "parity unpinned" against a Vulkan RENDER.  The reference ships no tests,
 * golden images or fixtures for this path and cannot be built here.
 * What pins the oracle instead:
 *   (a) reference-produced data: the shipped energy-compensation tables Assets/LookupTables/[*].bin were
 *       baked by the reference's own shaders from the hot-path functions GGXSampleAnisotopic /
 *       EvaluateReflection / EvaluateRefraction / DielectricFresnel / UniformFloat; the oracle's restated
 *       baker (orc_bake_*_texel) reproduces them within Monte-Carlo noise (tests/test_oracle_kat.py);
 *   (b) analytic known-answer tests derived from the reference source (furnace mode, PCG sequences,
 *       Fresnel/ACES fixed points, bloom mip sizes, alias-table invariants, LUT ranges, fixture facts).
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float pos[3]; float nrm[3]; float uv[2]; } OrcVertex;       /* SH/Bindings.slang:7-12  (32 B) */

typedef struct {                                                             /* PT/PathTracer.h:12-34  (112 B) */
    float BaseColor[3], EmissiveColor[3], SpecularColor[3], MediumColor[3], MediumEmissiveColor[3];
    float Metallic, Roughness, IOR, Transmission, Anisotropy, AnisotropyRotation;
    float MediumDensity, MediumAnisotropy;
    uint32_t BaseColorTextureIndex, NormalTextureIndex, RoughnessTextureIndex, MetallicTextureIndex, EmissiveTextureIndex;
} OrcMaterial;

typedef struct { uint32_t width, height, channels; uint32_t _pad; const uint8_t *data; } OrcTexture; /* channels 4 (RGBA8) or 1 (R8) */
typedef struct { const OrcVertex *verts; const uint32_t *indices; uint32_t nverts, nindices; } OrcMesh;
typedef struct { float transform[16]; uint32_t mesh, material; } OrcInstance; /* column-major (glm) */
typedef struct { uint32_t Alias; float Importance; } OrcAliasEntry;           /* SH/Bindings.slang:1-5 */

typedef struct {
    const OrcMesh *meshes;         uint32_t nmeshes;   uint32_t _p0;
    const OrcMaterial *materials;  uint32_t nmaterials; uint32_t _p1;
    const OrcTexture *textures;    uint32_t ntextures; uint32_t _p2;
    const OrcInstance *instances;  uint32_t ninstances; uint32_t _p3;
    const float *env_rgba;         /* envW*envH*4, alpha = pdf (orc_build_env_alias) */
    const OrcAliasEntry *env_alias;
    uint32_t envW, envH;
    const float *lut_reflect;      /* 64x64x32   [layer][y][x]  PT/PathTracer.cpp:199 */
    const float *lut_refract_out;  /* 128x128x32               :200 */
    const float *lut_refract_in;   /* 128x128x32               :201 */
} OrcSceneDesc;

/* Homogeneous AABB volume: the fields of VolumeGPU (PT/PathTracer.h:341-395, SH/Volume.slang:20-52) that a volume without NanoVDB
 * density data (m_DensityDataIndex == -1) reads. */
typedef struct {
    float CornerMin[3], CornerMax[3];
    float Color[3], EmissiveColor[3];
    float Density, Anisotropy, Alpha, DropletSize;
    uint32_t ApproximatedScattering;     /* m_ApproximatedScattering: anisotropy decays with the volume depth (SH/Volume.slang:150-156) */
    uint32_t _pad;
} OrcVolume;

typedef struct {                                                             /* PT/PathTracer.h:271-309 subset */
    float ViewInverse[16];        /* column-major */
    float ProjectionInverse[16];
    uint32_t SampleCount;         /* SamplesPerFrame */
    uint32_t MaxDepth;
    float MaxLuminance, FocusDistance, DepthOfFieldStrength;
    float SkyRotationAzimuth, SkyRotationAltitude, EnvironmentIntensity;
    float EmissiveMeshSamplingPDFBias;
    uint32_t ScreenSplitCount;
    /* shader #defines, PT/PathTracer.cpp:621-654 */
    uint32_t EnableSkyMIS, EnableMeshMIS, ShowEnvMapDirectly, UseOnlyGeometryNormals, UseEnergyCompensation, FurnaceTestMode;
    /* volumes: SH/RayGen.slang:162-380, SH/Volume.slang (homogeneous part); PHASE_FUNCTION_* define, PT/PathTracer.cpp:639-650 */
    uint32_t PhaseFunction;       /* 0 Henyey-Greenstein, 1 Draine, 2 Henyey-Greenstein + Draine (PT/PathTracer.h:76-81) */
    uint32_t VolumesCount;
    const OrcVolume *Volumes;
    /* atmosphere: ENABLE_ATMOSPHERE define (PT/PathTracer.cpp:636-637) + UBO fields SH/Bindings.slang:26-37, defaults PT/PathTracer.h:221-232 */
    uint32_t EnableAtmosphere; uint32_t _padA;
    float PlanetPosition[3], PlanetRadius, AtmosphereHeight;
    float RayleighScatteringCoefficientMultiplier[3], MieScatteringCoefficientMultiplier[3], OzoneAbsorptionCoefficientMultiplier[3];
    float RayleighDensityFalloff, MieDensityFalloff, OzoneDensityFalloff, OzonePeak;
    float SunColor[3];
} OrcConfig;

typedef struct { uint64_t paths, segments, surface_hits, misses, shadow_rays, medium_events; } OrcCounters;   /* medium_events: scattering events inside a mesh medium or an AABB volume */

typedef struct OrcScene OrcScene;

/* ---- KAT helpers ---- */
uint32_t orc_pcg_hash(uint32_t seed);                                  /* SH/Sampler.slang:4-9 */
void     orc_rng_floats(uint32_t seed, uint32_t n, float *out);        /* SH/Sampler.slang:38-43 */
float    orc_dielectric_fresnel(float cosI, float eta);                /* SH/Material.slang:434-449 */
void     orc_aces_fitted(const float in[3], float out[3]);             /* SH/PostProcess/Tonemap.slang:20-55 */

/* one texel of the reference's energy-compensation LUT baker (SH/LookupReflect.slang, SH/LookupRefract.slang) */
float    orc_bake_reflect_texel(uint32_t tx, uint32_t ty, uint32_t tz, uint32_t samples, uint32_t seed);
void orc_draine_cos_theta(float rx, float g, float a, float *cos32, double *cos64);
float orc_bake_reflect_params(float viewCosine, float roughness, float anisotropy, float lz_min, uint32_t samples, uint32_t seed);
float    orc_bake_refract_texel(uint32_t tx, uint32_t ty, uint32_t tz, int above_surface, uint32_t samples, uint32_t seed);
float    orc_bake_lut_texel(int kind, uint32_t SX, uint32_t SY, uint32_t SZ, uint32_t tx, uint32_t ty, uint32_t tz, uint32_t sample_count, uint32_t seed);

/* ---- host-side one-offs ---- */
/* PT/PathTracer.cpp:1137-1332. rgba is modified in place (alpha <- pdf). Returns importance sum. */
float orc_build_env_alias(float *rgba, uint32_t w, uint32_t h, OrcAliasEntry *alias_out);
/* PT/Editor.cpp:45-48,1042-1051 + PT/FlyCamera.cpp:84-140: the matrices the shader really receives. */
void  orc_camera_from_view(const float view[16], float aspect, float viewInv_out[16], float projInv_out[16]);
void  orc_default_config(OrcConfig *cfg);
/* Radiance .hdr -> float RGBA (alpha 1), stbi_loadf semantics; free with orc_free. NULL on error. */
float *orc_load_hdr(const char *path, uint32_t *w_out, uint32_t *h_out);
void   orc_free(void *p);

/* ---- scene ---- */
OrcScene *orc_scene_create(const OrcSceneDesc *desc);   /* copies nothing: desc arrays must outlive the scene */
void      orc_scene_destroy(OrcScene *s);
uint32_t  orc_scene_triangle_count(const OrcScene *s);
uint32_t  orc_scene_emissive_count(const OrcScene *s);
/* world-space triangles in (instance, primitive) order: 9 floats each */
void      orc_scene_world_triangles(const OrcScene *s, float *out9, uint32_t *inst_out, uint32_t *prim_out);

/* closest hit, semantics of SH/RTCommon.slang:47-117 / TraceRay (tmin < t < tmax, no culling, ties -> lowest id).
 * use_bvh=0 -> brute force over all triangles. miss: t=-1, prim=inst=0xFFFFFFFF. */
void orc_trace_closest(const OrcScene *s, uint32_t n, const float *org3, const float *dir3, float tmin, float tmax,
                       int use_bvh, float *t_out, uint32_t *prim_out, uint32_t *inst_out, float *uv_out2);

/* One reference dispatch sequence: frames [frame0, frame0+nframes) with Seed_f = PCG_HASH(base_seed + f)
 * (the reference's own seed is wall-clock, PT/PathTracer.cpp:127-140), running-mean accumulate
 * into image (RGBA32F, W*H*4) exactly as SH/RayGen.slang:130-159.  Only rows y with
 * ((y / band_rows) % world) == rank are rendered (image-tile partition); pass world=1 for all.
 * nthreads<=0 -> hardware concurrency. */
void orc_render(const OrcScene *s, const OrcConfig *cfg, uint32_t W, uint32_t H,
                uint32_t frame0, uint32_t nframes, uint32_t base_seed,
                uint32_t rank, uint32_t world, uint32_t band_rows,
                float *image, int nthreads, OrcCounters *counters);

/* Single path sample of pixel (x,y) for frame seed `seed` (push-constant Seed): returns pathLight of the
 * first of SampleCount samples; used by fine-grained parity tests. */
void orc_sample_pixel(const OrcScene *s, const OrcConfig *cfg, uint32_t W, uint32_t H,
                      uint32_t x, uint32_t y, uint32_t seed, float out_rgb[3], uint32_t *segments_out);

/* ---- post chain: PT/PostProcessor.cpp:128-246 + SH/PostProcess/{Bloom,Tonemap}.slang ---- */
typedef struct { float Exposure, Gamma, BloomThreshold, BloomStrength, FalloffRange; uint32_t MipCount; } OrcPostConfig;
void     orc_default_post_config(OrcPostConfig *c);
uint32_t orc_bloom_mip_sizes(uint32_t W, uint32_t H, uint32_t *wh_out20); /* returns level count (<=10) */
/* hdr: W*H*4 float; ldr_out: W*H*4 u8; bloom0_out (optional, W*H*4 float) = mip 0 after the up pass */
void     orc_post_process(const float *hdr, uint32_t W, uint32_t H, const OrcPostConfig *c,
                          uint8_t *ldr_out, float *bloom0_out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
