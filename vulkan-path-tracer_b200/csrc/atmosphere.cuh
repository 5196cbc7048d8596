// atmosphere.cuh -- the reference's atmosphere (ENABLE_ATMOSPHERE): SH/Atmosphere.slang, SH/RayGen.slang:76-84,212-255,382-471,
// SH/Sampler.slang:196-215,430-476, SH/RTCommon.slang:174-211, SH/Miss.slang:11-14; host block PT/PathTracer.h:129-144,170-181,221-232.
// With the atmosphere on, a miss emits nothing: the sky is sun light scattered in by Rayleigh / Mie events found by delta tracking on ONE colour
// channel (the path is "split" at its first atmosphere event), the sky NEE samples the sun disk, and every sky NEE term is attenuated by a
// ratio-tracked transmittance that CONSUMES RANDOM NUMBERS -- only when the shadow query found the sun visible.  Wavefront placement:
//   k_volume_decide  kills paths below the planet surface (RayGen.slang:76-84), draws the channel pick and the delta-tracking walk after the volumes'
//                    free-flight draws (the reference's order), flags atmosphere events in the hit record (y = -2 - component) and stores the chosen
//                    channel in the path's Depth word (bits 27-28);
//   k_shade_volume   EvaluateAtmosphereScatteringEvent: new direction (Rayleigh / HG 0.85 / unchanged for ozone), sun-disk sample, sky request;
//   k_shade_hit / k_shade_volume   always emit the sky request (even with a zero contribution) so that k_connect can decide visibility;
//   k_connect        after the sky query: if visible, the transmittance walk(s) on the path's own RNG stream (three while unsplit, one after),
//                    contribution * T, then the Russian-roulette draw -- the reference's draw order;
//   k_shade_miss     no emission; every path epilogue adds only the path's channel once it is split (RayGen.slang:116-128).
#pragma once
#include "volumes.cuh"

namespace b200pt {

constexpr uint32_t DEAD_EVENT = 0xFFFFFFFDu;        // so.hit[i].w of a path that k_volume_decide finished (below the planet surface): k_extend queues nothing
// Depth word of a path (thr_depth.w / e0.w): [31] InMedium  [30] light request [29] sky request (e0.w only)  [28:27] colour channel + 1 (0 = unsplit)  [26:0] Depth
constexpr uint32_t DEPTH_MASK = 0x07FFFFFFu, CHANNEL_SHIFT = 27u, CHANNEL_MASK = 3u << 27;
__device__ __forceinline__ int dflags_channel(uint32_t dflags) { return (int)((dflags & CHANNEL_MASK) >> CHANNEL_SHIFT) - 1; }

__device__ __forceinline__ float2 intersect_sphere(float3 ro, float3 rd, float3 center, float radius) {      // RTCommon.slang:174-193
    ro = ro - center;
    const float a = dot(rd, rd);
    const float b = 2.0f * dot(ro, rd);
    const float c = dot(ro, ro) - radius * radius;
    const float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f) return make_float2(-1.0f, -1.0f);
    return make_float2((-b - sqrtf(disc)) / (2.0f * a), (-b + sqrtf(disc)) / (2.0f * a));
}
__device__ __forceinline__ float3 atm_planet(const DevConfig &c) { return f3(c.PlanetPosition[0], c.PlanetPosition[1], c.PlanetPosition[2]); }
__device__ __forceinline__ float atm_height(const DevConfig &c, float3 p) { return length(p - atm_planet(c)) - c.PlanetRadius; }   // Atmosphere.slang:13-16
__device__ __forceinline__ float atm_rayleigh_density(const DevConfig &c, float h) { return expf(-h / c.RayleighDensityFalloff); }
__device__ __forceinline__ float atm_mie_density(const DevConfig &c, float h) { return expf(-h / c.MieDensityFalloff); }
__device__ __forceinline__ float atm_ozone_density(const DevConfig &c, float h) { return expf(-(fabsf(h - c.OzonePeak) / c.OzoneDensityFalloff)); }
struct AtmCoef { float kr, km, ko, majorant; };
__device__ __forceinline__ AtmCoef atm_coefficients(const DevConfig &c, int ch) {                             // Atmosphere.slang:7-11,52-60
    const float C_RAYLEIGH = ch == 0 ? 5.802f * 1e-6f : (ch == 1 ? 13.558f * 1e-6f : 33.100f * 1e-6f);
    const float C_OZONE = ch == 0 ? 0.650f * 1e-6f : (ch == 1 ? 1.881f * 1e-6f : 0.085f * 1e-6f);
    const float C_MIE = 3.996f * 1e-6f + 4.40f * 1e-6f;
    AtmCoef k;
    k.kr = C_RAYLEIGH * c.RayleighMult[ch]; k.km = C_MIE * c.MieMult[ch]; k.ko = C_OZONE * c.OzoneMult[ch];
    k.majorant = atm_rayleigh_density(c, 0.0f) * k.kr + atm_mie_density(c, 0.0f) * k.km + atm_ozone_density(c, c.OzonePeak) * k.ko;
    return k;
}
// CalculateTransmittanceThroughAtmosphere, Atmosphere.slang:33-104 (one channel; the caller places the value)
static __device__ __noinline__ float atm_transmittance(const DevConfig &c, Rng &rng, float3 ro, float3 rd, int ch) {
    const float2 pi = intersect_sphere(ro, rd, atm_planet(c), c.PlanetRadius);
    if (pi.y > 0.0f) return 0.0f;                                                                             // occluded by the planet
    const float2 ai = intersect_sphere(ro, rd, atm_planet(c), c.PlanetRadius + c.AtmosphereHeight);
    const float tMin = fmaxf(ai.x, 0.0f), tMax = ai.y;
    if (tMax < 0.0f) return 1.0f;
    const AtmCoef k = atm_coefficients(c, ch);
    if (k.majorant <= 0.0f) return 1.0f;
    float t = 0.0f, T = 1.0f;
    for (int i = 0; i < 1000; i++) {
        const float deltaT = -logf(1.0f - rng.next()) / k.majorant;
        t += deltaT;
        if (t >= tMax - tMin) break;
        const float h = atm_height(c, ro + rd * (t + tMin));
        if (h < 0.0f) break;
        const float dr = atm_rayleigh_density(c, h) * k.kr, dm = atm_mie_density(c, h) * k.km, dz = atm_ozone_density(c, h) * k.ko;
        T *= 1.0f - (dr + dm + dz) / k.majorant;
        const float p = T;
        if (rng.next() > p) { T = 0.0f; break; }
        T /= p;
    }
    return T;
}
// the NEE transmittance of ClosestHit.slang:335-349 / RayGen.slang:328-342: three walks (r, g, b in this order) while the path is unsplit, one after --
// a split path's vector is zero in the other two channels
__device__ __forceinline__ float3 atm_transmittance_nee(const DevConfig &c, Rng &rng, float3 ro, float3 rd, int channel) {
    if (channel < 0) {
        float3 T;
        T.x = atm_transmittance(c, rng, ro, rd, 0); T.y = atm_transmittance(c, rng, ro, rd, 1); T.z = atm_transmittance(c, rng, ro, rd, 2);
        return T;
    }
    const float t = atm_transmittance(c, rng, ro, rd, channel);
    return f3(channel == 0 ? t : 0.0f, channel == 1 ? t : 0.0f, channel == 2 ? t : 0.0f);
}
// SampleAtmosphereScatterDistance, Atmosphere.slang:114-201: delta tracking; component: -1 none, 0 Rayleigh, 1 Mie, 2 ozone
static __device__ __noinline__ float atm_sample_scatter_distance(const DevConfig &c, Rng &rng, float3 ro, float3 rd, int ch, int &component) {
    const float2 ai = intersect_sphere(ro, rd, atm_planet(c), c.PlanetRadius + c.AtmosphereHeight);
    const float tMinA = fmaxf(ai.x, 0.0f), tMaxA = ai.y;
    component = -1;
    const float2 pi = intersect_sphere(ro, rd, atm_planet(c), c.PlanetRadius);
    const float tMinPlanet = pi.x;
    if (tMaxA < 0.0f) return -1.0f;
    const AtmCoef k = atm_coefficients(c, ch);
    if (k.majorant <= 0.0f) return -1.0f;
    float t = tMinA;
    for (int i = 0; i < 1000; i++) {
        const float deltaT = -logf(1.0f - rng.next()) / k.majorant;
        t += deltaT;
        if (t >= tMaxA) break;
        if (tMinPlanet > 0.0f && t >= tMinPlanet) break;
        const float h = atm_height(c, ro + rd * t);
        const float dr = atm_rayleigh_density(c, h) * k.kr, dm = atm_mie_density(c, h) * k.km, dz = atm_ozone_density(c, h) * k.ko;
        const float density = dr + dm + dz;
        if (density / k.majorant < rng.next()) continue;                                                      // null collision
        const float pr = dr / density, pm = dm / density;
        const float x = rng.next();
        component = x <= pr ? 0 : (x <= pr + pm ? 1 : 2);
        return t;
    }
    return -1.0f;
}
__device__ __forceinline__ float phase_rayleigh(float3 V, float3 L) { const float ct = dot(V, L); return (3.0f / (16.0f * PT_PI)) * (1.0f + ct * ct); }   // RTCommon.slang:197-201
__device__ __forceinline__ float phase_mie_approx(float3 V, float3 L, float g) {                              // RTCommon.slang:204-211
    const float ct = dot(V, L);
    g = fminf(g, 0.9381f);
    const float k = 1.55f * g - 0.55f * g * g * g;
    const float kc = k * ct;
    return (1.0f - k * k) / ((4.0f * PT_PI) * (1.0f - kc) * (1.0f - kc));
}
__device__ __forceinline__ float3 sample_rayleigh(Rng &r, float3 incident) {                                  // Sampler.slang:196-215
    const float rx = r.next(), ry = r.next();
    const float a = 2.0f * rx - 1.0f;
    const float u = -pt_pow(2.0f * a + sqrtf(4.0f * pt_pow(a, 2.0f) + 1.0f), 1.0f / 3.0f);
    const float cosTheta = u - (1.0f / u);
    return phase_frame(incident, cosTheta, 2.0f * PT_PI * ry);
}
// SampleSunDisk(0.004675), Sampler.slang:430-463 -- ImportanceSampleSky under ENABLE_ATMOSPHERE (:465-476); cos / sin of the sky angles come from the host
__device__ __forceinline__ void sample_sun_disk(const DevConfig &c, Rng &rng, float3 &toLight, float4 &colorPDF) {
    float3 sunDir = rotate3_cs(f3(0.0f, 0.0f, -1.0f), f3(1.0f, 0.0f, 0.0f), c.cosAl, c.sinAl);
    sunDir = rotate3_cs(sunDir, f3(0.0f, 1.0f, 0.0f), c.cosAz, c.sinAz);
    const float cosThetaMax = c.cosSunTheta;
    const float phi = 2.0f * PT_PI * rng.next();
    const float cosTheta = mixf(cosThetaMax, 1.0f, rng.next());
    const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float sp, cp; pt_sincos(phi, &sp, &cp);
    const float3 local = f3(cp * sinTheta, sp * sinTheta, cosTheta);
    const float3 w = normalize(sunDir);
    const float3 up = fabsf(w.z) < 0.999f ? f3(0, 0, 1) : f3(1, 0, 0);
    const float3 u = normalize(cross(up, w));
    const float3 v = cross(w, u);
    toLight = (u * local.x + v * local.y) + w * local.z;
    const float solidAngle = 2.0f * PT_PI * (1.0f - cosThetaMax);
    colorPDF.w = 1.0f / solidAngle;
    colorPDF.x = 2e5f * c.SunColor[0] * c.EnvironmentIntensity; colorPDF.y = 2e5f * c.SunColor[1] * c.EnvironmentIntensity; colorPDF.z = 2e5f * c.SunColor[2] * c.EnvironmentIntensity;
}

} // namespace b200pt
