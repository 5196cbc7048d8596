// volumes.cuh -- AABB volumes: homogeneous, and heterogeneous over a dense copy of the density data (DevGrid).
// Reference: SH/Volume.slang, SH/RayGen.slang:162-380 (ScatteredInVolume,
// EvaluateVolumeScatteringEvent), phase functions SH/RTCommon.slang:214-228 with their samplers SH/Sampler.slang:169-295,
// host API PT/PathTracer.h:36-81,157-166, PT/PathTracer.cpp:1334-1345.
// Wavefront placement:
//   k_volume_decide  before k_extend: the reference's per-segment distance query (SH/RTCommon.slang:86-100: un-normalised direction,
//                    tmin 1e-5) + the free-flight draws of ScatteredInVolume; a path that scatters is flagged in so.hit[i]
//                    (w = VOLUME_EVENT, x = distance, y = volume index) and k_extend queues it without tracing;
//   k_shade_volume   EvaluateVolumeScatteringEvent for the flagged entries of the hit queue (k_shade_hit skips them): emission,
//                    sky / light NEE requests weighted by the phase function and the analytic transmittance, phase-function sample;
//   k_shade_hit      multiplies its NEE requests by the transmittance from the new path origin (SH/ClosestHit.slang:332-364);
//   k_connect        joins and compacts volume events like surface hits (their Depth grows by one, SH/RayGen.slang:377).
#pragma once
#include "shading.cuh"

namespace b200pt {

constexpr uint32_t VOLUME_EVENT = 0xFFFFFFFEu;      // so.hit[i].w of a path that scattered inside a volume this segment
constexpr int MAX_VOLUMES = 100;                     // RayGen.slang:165-166 (float distances[100]; int indices[100]): the per-ray sort arrays live in local memory

// The walks over the majorant blocks end on a comparison of two quantities that are EQUAL in exact arithmetic (the last block's far face reached by
// stepping, against the box's far distance): how many random numbers a walk consumes there is decided by rounding.  So the geometry of the volume code is
// evaluated with individually rounded IEEE operations (no fused multiply-add, no approximate division / reciprocal, whatever the build flags say),
// which makes a collision-free walk take bit for bit the steps the oracle's takes (profiles/r02_het_walks.txt).
__device__ __forceinline__ float ie_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float ie_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ie_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ie_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float3 ie_at(float3 o, float3 d, float s) { return f3(ie_add(o.x, ie_mul(d.x, s)), ie_add(o.y, ie_mul(d.y, s)), ie_add(o.z, ie_mul(d.z, s))); }   // o + d * s
__device__ __forceinline__ float3 ie_rel(float3 p, float3 mn, float3 mx) {                        // (p - mn) / (mx - mn)
    return f3(ie_div(ie_sub(p.x, mn.x), ie_sub(mx.x, mn.x)), ie_div(ie_sub(p.y, mn.y), ie_sub(mx.y, mn.y)), ie_div(ie_sub(p.z, mn.z), ie_sub(mx.z, mn.z)));
}
struct VolIsect { float Near, Far; };
// SH/Volume.slang:188-211 (the x/y/z mix-up of the max / min chains is the reference's)
__device__ __forceinline__ VolIsect vol_intersect(float3 o, float3 d, float3 mn, float3 mx) {
    const float3 inv = f3(ie_div(1.0f, d.x), ie_div(1.0f, d.y), ie_div(1.0f, d.z));
    const float3 t0 = f3(ie_mul(ie_sub(mn.x, o.x), inv.x), ie_mul(ie_sub(mn.y, o.y), inv.y), ie_mul(ie_sub(mn.z, o.z), inv.z));
    const float3 t1 = f3(ie_mul(ie_sub(mx.x, o.x), inv.x), ie_mul(ie_sub(mx.y, o.y), inv.y), ie_mul(ie_sub(mx.z, o.z), inv.z));
    const float3 ts = f3(fminf(t0.x, t1.x), fminf(t0.y, t1.y), fminf(t0.z, t1.z)), tb = f3(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y), fmaxf(t0.z, t1.z));
    const float tmin = fmaxf(fmaxf(ts.x, ts.y), fmaxf(ts.x, ts.z));
    const float tmax = fminf(fminf(tb.x, tb.y), fminf(tb.x, tb.z));
    VolIsect r; r.Near = tmin; r.Far = tmax;
    if (tmax < 0.0f || tmin > tmax) { r.Near = -1.0f; r.Far = -1.0f; }
    return r;
}
__device__ __forceinline__ VolIsect vol_intersect(const DevVolume &v, float3 o, float3 d) { return vol_intersect(o, d, f3(v.mn_density), f3(v.mx_g)); }

// SH/Volume.slang:150-156
__device__ __forceinline__ float vol_effective_anisotropy(const DevVolume &v, float rayDepth) {
    const float g = v.mx_g.w;
    if (v.flags.x != 0u) { const float sg = g > 0.0f ? 1.0f : (g < 0.0f ? -1.0f : 0.0f); return pt_pow(fabsf(g), 1.0f + rayDepth) * sg; }
    return g;
}
// SH/RTCommon.slang:214-221
__device__ __forceinline__ float phase_hg(float3 V, float3 L, float g) {
    if (g == 0.0f) return 1.0f / (4.0f * PT_PI);
    const float c = dot(V, L);
    return (1.0f / (4.0f * PT_PI)) * ((1.0f - g * g) / pt_pow(1.0f + g * g - 2.0f * g * c, 1.5f));
}
// SH/RTCommon.slang:223-228
__device__ __forceinline__ float phase_draine(float3 V, float3 L, float g, float a) {
    const float c = dot(V, L);
    return ((1.0f - g * g) * (1.0f + a * c * c)) / (4.0f * (1.0f + (a * (1.0f + 2.0f * g * g)) / 3.0f) * PT_PI * pt_pow(1.0f + g * g - 2.0f * g * c, 1.5f));
}
__device__ __forceinline__ float3 phase_frame(float3 incident, float cosTheta, float phi) {      // tail of SH/Sampler.slang:183-192 / :265-274
    const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float sp, cp; pt_sincos(phi, &sp, &cp);
    const float3 nd = f3(sinTheta * cp, sinTheta * sp, cosTheta);
    const float3 up = fabsf(incident.y) < 0.9999999f ? f3(0, 1, 0) : f3(0, 0, 1);
    const float3 tangent = normalize(cross(up, incident));
    const float3 bitangent = cross(incident, tangent);
    return normalize((tangent * nd.x + bitangent * nd.y) + incident * nd.z);
}
// SH/Sampler.slang:219-276
static __device__ __noinline__ float3 sample_draine(Rng &r, float3 incident, float g, float a) {
    const float rx = r.next(), ry = r.next();
    float cosTheta;
    if (fabsf(g) < 1e-5f) cosTheta = 2.0f * rx - 1.0f;
    else if (fabsf(a) < 1e-5f) { const float sq = (1.0f - g * g) / (1.0f - g + 2.0f * g * rx); cosTheta = (1.0f + g * g - sq * sq) / (2.0f * g); }
    else {
        const float g2 = g * g, g3 = g * g2, g4 = g2 * g2, g6 = g2 * g4;
        const float pgp1_2 = (1.0f + g2) * (1.0f + g2);
        const float T1a = -a + a * g4;
        const float T1a3 = T1a * T1a * T1a;
        const float T2 = -1296.0f * (-1.0f + g2) * (a - a * g2) * (T1a) * (4.0f * g2 + a * pgp1_2);
        const float T3 = 3.0f * g2 * (1.0f + g * (-1.0f + 2.0f * rx)) + a * (2.0f + g2 + g3 * (1.0f + 2.0f * g2) * (-1.0f + 2.0f * rx));
        const float T4a = 432.0f * T1a3 + T2 + 432.0f * (a - a * g2) * T3 * T3;
        const float T4b = -144.0f * a * g2 + 288.0f * a * g4 - 144.0f * a * g6;
        const float T4b3 = T4b * T4b * T4b;
        const float T4 = T4a + sqrtf(-4.0f * T4b3 + T4a * T4a);
        const float T4p3 = pt_pow(T4, 1.0f / 3.0f);
        const float cbrt2 = pt_pow(2.0f, 1.0f / 3.0f);
        const float T6 = (2.0f * T1a + (48.0f * cbrt2 * (-(a * g2) + 2.0f * a * g4 - a * g6)) / T4p3 + T4p3 / (3.0f * cbrt2)) / (a - a * g2);
        const float T5 = 6.0f * (1.0f + g2) + T6;
        const float q = -0.5f * sqrtf(T5) + sqrtf(6.0f * (1.0f + g2) - (8.0f * T3) / (a * (-1.0f + g2) * sqrtf(T5)) - T6) / 2.0f;
        cosTheta = (1.0f + g2 - q * q) / (2.0f * g);
    }
    return phase_frame(incident, cosTheta, 2.0f * PT_PI * ry);
}
struct HgDraineFit { float GHG, GD, AD, WD; };
__device__ __forceinline__ HgDraineFit hg_draine_fit(float d) {                                  // SH/Volume.slang:389-395, SH/Sampler.slang:281-284
    HgDraineFit f;
    f.GHG = expf(-(0.0990567f / (d - 1.67154f)));
    f.GD = expf(-(2.20679f / (d + 3.91029f)) - 0.428934f);
    f.AD = expf(3.62489f - (8.29288f / (d + 5.52825f)));
    f.WD = expf(-(0.599085f / (d - 0.641583f)) - 0.665888f);
    return f;
}
// Volume::GetScatteringDirection, SH/Volume.slang:354-371
static __device__ __noinline__ float3 vol_scatter_direction(uint32_t phase_function, const DevVolume &v, Rng &rng, float3 incident, int rayDepth) {
    if (phase_function == 0u) return sample_henyey_greenstein(rng, incident, vol_effective_anisotropy(v, (float)rayDepth));
    if (phase_function == 1u) return sample_draine(rng, incident, vol_effective_anisotropy(v, (float)rayDepth), v.color_alpha.w);
    HgDraineFit f = hg_draine_fit(v.emis_droplet.w);                                             // SH/Sampler.slang:278-295
    f.GHG = pt_pow(fmaxf(f.GHG, 0.0f), 1.0f + (float)rayDepth);
    f.GD = pt_pow(fmaxf(f.GD, 0.0f), 1.0f + (float)rayDepth);
    const float u = rng.next();
    if (u < f.WD) return sample_henyey_greenstein(rng, incident, f.GHG);
    return sample_draine(rng, incident, f.GD, f.AD);
}
// Volume::EvaluatePhaseFunction, SH/Volume.slang:373-401 (the HG + Draine evaluation ignores the depth, unlike its sampler)
static __device__ __noinline__ float vol_phase(uint32_t phase_function, const DevVolume &v, float3 V, float3 L, int rayDepth) {
    if (phase_function == 0u) return phase_hg(V, L, vol_effective_anisotropy(v, (float)rayDepth));
    if (phase_function == 1u) return phase_draine(V, L, vol_effective_anisotropy(v, (float)rayDepth), v.color_alpha.w);
    const HgDraineFit f = hg_draine_fit(v.emis_droplet.w);
    return mixf(phase_hg(V, L, f.GHG), phase_draine(V, L, f.GD, f.AD), f.WD);
}
// ---- heterogeneous volumes: SH/Volume.slang:54-166,230-252,291-352,448-517 over DevGrid (a dense copy of what the reference's NanoVDB buffer returns) ----
constexpr int GRID_DIM = 32;                                                                     // MAX_DENSITY_GRID_DIM, SH/Volume.slang:11
// SampleNanoVDBBuffer, SH/Volume.slang:69-117: box -> [0, 1]^3 (Y flipped) -> the grid's integer world box -> index space (float inverse map) -> floor,
// three raw PCG draws jitter the voxel by -1 / 0 / +1, clamp to the root bbox, read, normalise
static __device__ __noinline__ float grid_sample(const DevVolume &v, const DevGrid &g, Rng &rng, float3 x) {
    float3 np = ie_rel(x, f3(v.mn_density), f3(v.mx_g));
    np.y = 1.0f - np.y;
    const float3 gp = f3(ie_add(ie_mul(np.x, g.wext[0]), g.wmin[0]), ie_add(ie_mul(np.y, g.wext[1]), g.wmin[1]), ie_add(ie_mul(np.z, g.wext[2]), g.wmin[2]));
    const float3 ip = f3(ie_mul(ie_sub(gp.x, g.trans[0]), g.inv_vs[0]), ie_mul(ie_sub(gp.y, g.trans[1]), g.inv_vs[1]), ie_mul(ie_sub(gp.z, g.trans[2]), g.inv_vs[2]));
    int c[3] = { (int)floorf(ip.x), (int)floorf(ip.y), (int)floorf(ip.z) };
    c[0] = (int)((uint32_t)c[0] + (rng.next_u32() % 3u - 1u));
    c[1] = (int)((uint32_t)c[1] + (rng.next_u32() % 3u - 1u));
    c[2] = (int)((uint32_t)c[2] + (rng.next_u32() % 3u - 1u));
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] = max(g.imin[k], min(g.imin[k] + g.dim[k] - 1, c[k]));
    const float value = __ldg(g.values + ((size_t)(c[2] - g.imin[2]) * g.dim[1] + (size_t)(c[1] - g.imin[1])) * g.dim[0] + (size_t)(c[0] - g.imin[0]));
    return clampf(ie_mul(ie_div(value, v.kelvin.z), v.tparams.w), 0.0f, 1.0f);
}
__device__ __forceinline__ float vol_effective_density(const DevVolume &v, float base, float rayDepth) {   // SH/Volume.slang:159-166
    if (v.flags.x != 0u) return ie_mul(base, pt_pow(v.tparams.z, rayDepth));
    return base;
}
struct VolCtx { float3 blockSize; float epsilon, tEnter, tExit; };
struct VolBlock { int blockIndex; float3 minCorner, maxCorner; };
__device__ __forceinline__ VolCtx vol_ctx(const DevVolume &v, VolIsect is) {                     // CreateTraversalContext, :119-128
    VolCtx c;
    const float3 ext = f3(v.mx_g) - f3(v.mn_density);
    c.blockSize = f3(ie_div(ext.x, (float)GRID_DIM), ie_div(ext.y, (float)GRID_DIM), ie_div(ext.z, (float)GRID_DIM));
    c.epsilon = ie_mul(0.0001f, fmaxf(ext.x, fmaxf(ext.y, ext.z)));
    c.tEnter = fmaxf(is.Near, 0.0f); c.tExit = is.Far;
    return c;
}
__device__ __forceinline__ VolBlock vol_block(const DevVolume &v, float3 p, const VolCtx &c) {  // CalculateBlockInfo, :131-147
    const float3 mn = f3(v.mn_density);
    const float3 rel = ie_rel(p, mn, f3(v.mx_g));
    const int ix = max(0, min(GRID_DIM - 1, (int)ie_mul(rel.x, (float)GRID_DIM))), iy = max(0, min(GRID_DIM - 1, (int)ie_mul(rel.y, (float)GRID_DIM))),
              iz = max(0, min(GRID_DIM - 1, (int)ie_mul(rel.z, (float)GRID_DIM)));
    VolBlock b;
    b.blockIndex = ix + iy * GRID_DIM + iz * GRID_DIM * GRID_DIM;
    b.minCorner = f3(ie_add(mn.x, ie_mul(c.blockSize.x, (float)ix)), ie_add(mn.y, ie_mul(c.blockSize.y, (float)iy)), ie_add(mn.z, ie_mul(c.blockSize.z, (float)iz)));
    b.maxCorner = b.minCorner + c.blockSize;
    return b;
}
// ProcessHeterogeneousVolumeScattering (:291-352, delta tracking, SCATTER) / ProcessHeterogeneousVolumeTransmittance (:448-517, ratio tracking with
// Russian roulette) share the walk over the majorant blocks
template <bool SCATTER>
static __device__ __noinline__ float vol_grid_walk(const DevVolume &v, const DevGrid &g, Rng &rng, float3 o, float3 d, float rayDepth, VolIsect is) {
    const VolCtx c = vol_ctx(v, is);
    VolBlock b = vol_block(v, ie_at(o, d, c.tEnter + c.epsilon), c);
    float T = 1.0f, t = 0.0f;
    for (int i = 0; i < (SCATTER ? 10000 : 1000); i++) {
        const float3 cur = ie_at(o, d, c.tEnter + t + c.epsilon);
        const VolIsect bi = vol_intersect(cur, d, b.minCorner, b.maxCorner);
        const float maxDensity = vol_effective_density(v, ie_mul(__ldg(g.max_densities + b.blockIndex), v.mn_density.w), rayDepth);
        const float sampled = ie_div(-logf(rng.next()), maxDensity);
        if (bi.Far <= 0.0f) {                                                                    // the ray misses the block (precision): creep forward
            t += c.epsilon;
            if (c.tEnter + t > c.tExit) return SCATTER ? -1.0f : T;
            b = vol_block(v, ie_at(o, d, c.tEnter + t + c.epsilon), c);
            continue;
        }
        const float toExit = bi.Far - fmaxf(bi.Near, 0.0f);
        if (sampled > toExit) {                                                                  // next block
            t += toExit + c.epsilon;
            if (c.tEnter + t > c.tExit) return SCATTER ? -1.0f : T;
            b = vol_block(v, ie_at(o, d, c.tEnter + t + c.epsilon), c);
            continue;
        }
        t += sampled;
        if (c.tEnter + t > c.tExit) return SCATTER ? -1.0f : T;
        const float dens = vol_effective_density(v, ie_mul(grid_sample(v, g, rng, ie_at(o, d, c.tEnter + t)), v.mn_density.w), rayDepth);
        if (SCATTER) {
            if (ie_div(dens, maxDensity) < rng.next()) continue;                                 // null collision
            return c.tEnter + t;
        } else {
            T = ie_mul(T, 1.0f - ie_div(dens, maxDensity));
            const float p = T;
            if (rng.next() > p) return 0.0f;
            T = ie_div(T, p);
        }
    }
    return SCATTER ? -1.0f : T;
}
// Volume::CalculateVolumesTransmittance, SH/Volume.slang:419-446, scenes WITHOUT heterogeneous volumes: analytic, no random numbers (shading kernels)
static __device__ __noinline__ float volumes_transmittance(const DevScene &sc, float3 o, float3 d) {
    float T = 1.0f;
    for (uint32_t i = 0; i < sc.n_volumes; i++) {
        const DevVolume &v = sc.volumes[i];
        VolIsect is = vol_intersect(v, o, d);
        is.Near = fmaxf(is.Near, 0.0f);
        const float len = is.Far - is.Near;
        if (len > 0.0f) T *= expf(-v.mn_density.w * len);
    }
    return clampf(T, 0.0f, 1.0f);
}
// the same with heterogeneous volumes present (k_connect, after the visibility query, on the path's own stream): a ratio-tracking walk per grid volume crossed
static __device__ __noinline__ float volumes_transmittance_walk(const DevScene &sc, Rng &rng, float3 o, float3 d, float rayDepth) {
    float T = 1.0f;
    for (uint32_t i = 0; i < sc.n_volumes; i++) {
        const DevVolume &v = sc.volumes[i];
        VolIsect is = vol_intersect(v, o, d);
        is.Near = fmaxf(is.Near, 0.0f);
        if (v.flags.w != 0xFFFFFFFFu && is.Far >= 0.0f) {
            T *= vol_grid_walk<false>(v, sc.grids[v.flags.w], rng, o, d, rayDepth, is);
            if (T <= 0.0f) return 0.0f;
        } else {
            const float len = is.Far - is.Near;
            if (len > 0.0f) T *= expf(-v.mn_density.w * len);
        }
    }
    return clampf(T, 0.0f, 1.0f);
}
// NEE terms of the shading kernels: with grid volumes in the scene the factor is applied by k_connect instead (volumes_transmittance_walk)
__device__ __forceinline__ float volumes_transmittance_shade(const DevScene &sc, float3 o, float3 d) { return sc.n_grids ? 1.0f : volumes_transmittance(sc, o, d); }
// Volume::DoesRayScatterInVolume, SH/Volume.slang:254-289 (homogeneous branch: one random number when the ray crosses the box)
__device__ __forceinline__ float vol_scatter_distance(const DevScene &sc, const DevVolume &v, float3 o, float3 d, Rng &rng, float rayDepth, float ignoreIfFartherThan) {
    const VolIsect is = vol_intersect(v, o, d);
    if (is.Far < 0.0f) return -1.0f;
    if (ignoreIfFartherThan >= 0.0f && is.Near > ignoreIfFartherThan) return -1.0f;
    const float inside = is.Far - fmaxf(is.Near, 0.0f);
    if (inside <= 0.0f) return -1.0f;
    if (v.flags.w != 0xFFFFFFFFu) return vol_grid_walk<true>(v, sc.grids[v.flags.w], rng, o, d, rayDepth, is);
    const float sampled = ie_div(-logf(rng.next()), v.mn_density.w);                             // SH/Sampler.slang:425-428
    if (sampled < inside) return fmaxf(is.Near, 0.0f) + sampled;
    return -1.0f;
}
// The free-flight part of ScatteredInVolume (SH/RayGen.slang:164-209): volumes visited in order of their (clamped) entry distance; rayDepth is
// payload.Depth (:199), not VolumeDepth.  Returns the scatter distance (< 0: none) and the index of the volume that scattered.
__device__ __forceinline__ float volumes_free_flight(const DevScene &sc, float3 o, float3 d, Rng &rng, float rayDepth, int &scattered) {
    float distances[MAX_VOLUMES]; int indices[MAX_VOLUMES];
    const int n = (int)min(sc.n_volumes, (uint32_t)MAX_VOLUMES);
    for (int i = 0; i < n; i++) { const VolIsect is = vol_intersect(sc.volumes[i], o, d); distances[i] = fmaxf(0.0f, is.Near); indices[i] = i; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (distances[j] < distances[i]) { const float td = distances[i]; const int ti = indices[i]; distances[i] = distances[j]; indices[i] = indices[j]; distances[j] = td; indices[j] = ti; }
    float scatterDistance = -1.0f; scattered = -1;
    for (int i = 0; i < n; i++) {
        const float t = vol_scatter_distance(sc, sc.volumes[indices[i]], o, d, rng, rayDepth, scatterDistance);
        if (t >= 0.0f && (t < scatterDistance || scatterDistance < 0.0f)) { scatterDistance = t; scattered = indices[i]; }
    }
    return scatterDistance;
}
// Blackbody, SH/RTCommon.slang:139-172
__device__ __forceinline__ float3 blackbody(float temperature) {
    const float temp = temperature / 100.0f;
    float r, g, b;
    if (temp <= 66.0f) r = 255.0f; else r = 329.698727446f * pt_pow(temp - 60.0f, -0.1332047592f);
    if (temp <= 66.0f) g = 99.4708025861f * logf(temp) - 161.1195681661f; else g = 288.1221695283f * pt_pow(temp - 60.0f, -0.0755148492f);
    if (temp >= 66.0f) b = 255.0f; else if (temp <= 19.0f) b = 0.0f; else b = 138.5177312231f * logf(temp - 10.0f) - 305.0447927307f;
    return f3(clampf(r / 255.0f, 0.0f, 1.0f), clampf(g / 255.0f, 0.0f, 1.0f), clampf(b / 255.0f, 0.0f, 1.0f));
}
// Volume::GetEmissionFromTemperatureAtPoint, SH/Volume.slang:230-252 (the temperature is read from the DENSITY buffer, :235)
static __device__ __noinline__ float3 vol_temperature_emission(const DevScene &sc, const DevVolume &v, Rng &rng, float3 x) {
    if (v.flags.y == 0u) return f3(0.0f);
    const float tn = grid_sample(v, sc.grids[v.flags.w], rng, x);
    const float3 color = v.flags.z ? blackbody(tn * v.kelvin.y + v.kelvin.x) : f3(v.tcol_gamma);
    const float intensity = pt_pow(tn, v.tcol_gamma.w) * v.tparams.x;
    return f3(intensity * pt_pow(color.x, v.tparams.y), intensity * pt_pow(color.y, v.tparams.y), intensity * pt_pow(color.z, v.tparams.y));
}

} // namespace b200pt
