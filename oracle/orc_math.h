/* orc_math.h -- tiny fp32 vector helpers for the CPU oracle (test infrastructure only).
 * All arithmetic is IEEE fp32, compiled with -ffp-contract=off so no FMA contraction happens.
 * The shading-language intrinsics are restated as the SPIR-V GLSL.std.450 definitions. */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <stdint.h>

/* SH/Defines.slang:1-17 */
#define ORC_PI          3.1415926535897F
#define ORC_2PI         6.2831853071795F
#define ORC_1_OVER_PI   0.3183098861837F
#define ORC_MAX_DEPTH   1000000u

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3s(float s) { return V3(s, s, s); }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3div(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 v3scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 v3divs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 v3neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float v3dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 v3cross(v3 a, v3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float v3length(v3 a) { return sqrtf(v3dot(a, a)); }
/* normalize(v) = v * (1 / sqrt(dot(v,v))) */
static inline v3 v3normalize(v3 a) { float inv = 1.0f / sqrtf(v3dot(a, a)); return v3scale(a, inv); }
static inline float orc_max3(v3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
static inline float orc_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float orc_saturate(float x) { return orc_clamp(x, 0.0f, 1.0f); }
/* FMix: x*(1-a) + y*a */
static inline float orc_lerp(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline v3 v3lerp(v3 x, v3 y, float a) { return V3(orc_lerp(x.x, y.x, a), orc_lerp(x.y, y.y, a), orc_lerp(x.z, y.z, a)); }
/* reflect(i, n) = i - 2 dot(n, i) n */
static inline v3 v3reflect(v3 i, v3 n) { float d = v3dot(n, i); return v3sub(i, v3scale(n, 2.0f * d)); }
/* refract(i, n, eta): k = 1 - eta^2 (1 - dot(n,i)^2); k<0 -> 0 */
static inline v3 v3refract(v3 i, v3 n, float eta) {
    float d = v3dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return V3(0.0f, 0.0f, 0.0f);
    return v3sub(v3scale(i, eta), v3scale(n, eta * d + sqrtf(k)));
}
static inline float orc_smoothstep(float e0, float e1, float x) {
    float t = orc_saturate((x - e0) / (e1 - e0));
    return t * t * (3.0f - 2.0f * t);
}
/* SH/RTCommon.slang:37-45 (and SH/Sampler.slang:11-19): Rodrigues rotation */
static inline v3 orc_rotate(v3 v, v3 axis, float theta) {
    float c = cosf(theta), s = sinf(theta);
    v3 n = v3normalize(axis);
    v3 a = v3scale(v, c);
    v3 b = v3scale(v3cross(n, v), s);
    v3 d = v3scale(v3scale(n, v3dot(n, v)), 1.0f - c);
    return v3add(v3add(a, b), d);
}
#endif
