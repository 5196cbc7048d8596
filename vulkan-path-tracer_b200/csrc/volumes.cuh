// volumes.cuh -- homogeneous AABB volumes (the part that needs no NanoVDB grid).
// Reference: SH/Volume.slang (the m_DensityDataIndex == -1 paths), SH/RayGen.slang:162-380 (ScatteredInVolume,
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

struct VolIsect { float Near, Far; };
// SH/Volume.slang:188-211 (the x/y/z mix-up of the max / min chains is the reference's)
__device__ __forceinline__ VolIsect vol_intersect(float3 o, float3 d, float3 mn, float3 mx) {
    const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const float3 t0 = (mn - o) * inv, t1 = (mx - o) * inv;
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
// Volume::CalculateVolumesTransmittance, SH/Volume.slang:419-446: analytic for homogeneous volumes, no random numbers
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
// Volume::DoesRayScatterInVolume, SH/Volume.slang:254-289 (homogeneous branch: one random number when the ray crosses the box)
__device__ __forceinline__ float vol_scatter_distance(const DevVolume &v, float3 o, float3 d, Rng &rng, float ignoreIfFartherThan) {
    const VolIsect is = vol_intersect(v, o, d);
    if (is.Far < 0.0f) return -1.0f;
    if (ignoreIfFartherThan >= 0.0f && is.Near > ignoreIfFartherThan) return -1.0f;
    const float inside = is.Far - fmaxf(is.Near, 0.0f);
    if (inside <= 0.0f) return -1.0f;
    const float sampled = -logf(rng.next()) / v.mn_density.w;                                    // SH/Sampler.slang:425-428
    if (sampled < inside) return fmaxf(is.Near, 0.0f) + sampled;
    return -1.0f;
}
// The free-flight part of ScatteredInVolume (SH/RayGen.slang:164-209): volumes visited in order of their (clamped) entry distance.
// Returns the scatter distance (< 0: none) and the index of the volume that scattered.
__device__ __forceinline__ float volumes_free_flight(const DevScene &sc, float3 o, float3 d, Rng &rng, int &scattered) {
    float distances[MAX_VOLUMES]; int indices[MAX_VOLUMES];
    const int n = (int)min(sc.n_volumes, (uint32_t)MAX_VOLUMES);
    for (int i = 0; i < n; i++) { const VolIsect is = vol_intersect(sc.volumes[i], o, d); distances[i] = fmaxf(0.0f, is.Near); indices[i] = i; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (distances[j] < distances[i]) { const float td = distances[i]; const int ti = indices[i]; distances[i] = distances[j]; indices[i] = indices[j]; distances[j] = td; indices[j] = ti; }
    float scatterDistance = -1.0f; scattered = -1;
    for (int i = 0; i < n; i++) {
        const float t = vol_scatter_distance(sc.volumes[indices[i]], o, d, rng, scatterDistance);
        if (t >= 0.0f && (t < scatterDistance || scatterDistance < 0.0f)) { scatterDistance = t; scattered = indices[i]; }
    }
    return scatterDistance;
}

} // namespace b200pt
