// lut_baker.cu -- energy-compensation lookup-table baker.
// Replaces LookupTableCalculator::CalculateTable (PT/LookupTableCalculator.cpp:44-157) and its two compute shaders
// (SH/LookupReflect.slang:25-84, SH/LookupRefract.slang:23-102).  The bake integrates the SAME device functions the
// path tracer's closest-hit shading uses (ggx_sample_vndf, ggx_d, ggx_g1, dielectric_fresnel, Rng from shading.cuh):
// regenerating the shipped Assets/LookupTables/*.bin is therefore an independent check of rows C7-C9 of SURVEY 8a.
//
// Reference structure kept: one thread per texel; sampleCount/20 "dispatches" of 20 samples; dispatch i seeds the
// texel's sampler with ty + tx*tx + Seed_i and adds finalValue/20 to the table in fp32; the sum is divided by the
// dispatch count at the end.  The reference derives Seed_i from the wall clock (:104-105); here the caller's `seed`
// takes the place of the clock reading so a bake is reproducible.
// B200 mapping: the bake is pure ALU/MUFU work with 4 B of output per 10^7 samples (no memory roofline); the only
// knob is filling 148 SMs, so the dispatch range is cut into `slices` (grid.y) whose partial sums are combined in a
// fixed order by k_bake_finish -- deterministic for a given slice count; slices == 1 is the reference's summation order.
#include "shading.cuh"
#include "kernels.h"

namespace b200pt {

struct BakeMat { float Ax, Ay, Eta, viewCosine; };

__device__ __forceinline__ BakeMat bake_texel_params(int kind, uint32_t SX, uint32_t SY, uint32_t SZ, uint32_t tx, uint32_t ty, uint32_t tz) {
    BakeMat b;
    if (kind == 0) {   // LookupReflect.slang:36-44
        b.viewCosine = clampf((float)tx / (float)SX, 0.05f, 0.999f);
        const float roughness = clampf((float)ty / (float)SY, 0.0001f, 1.0f);
        const float anisotropy = (float)tz / (float)SZ;
        const float aspect = sqrtf(1.0f - sqrtf(anisotropy) * 0.9f);
        b.Ax = fmaxf(0.0001f, roughness / aspect);
        b.Ay = fmaxf(0.0001f, roughness * aspect);
        b.Eta = 0.0f;
    } else {           // LookupRefract.slang:34-49
        const float vc = (float)tx / ((float)SX - 1.0f);
        b.viewCosine = clampf(vc * vc, 0.01f, 0.9999f);
        const float roughness = clampf((float)ty / ((float)SY - 1.0f), 0.01f, 1.0f);
        const float ior = 1.0f + clampf((float)tz / ((float)SZ - 1.0f), 0.0001f, 1.0f);
        b.Ax = roughness; b.Ay = roughness;
        b.Eta = kind == 1 ? (1.0f / ior) : ior;
    }
    return b;
}

// Material.EvaluateReflection(V, L, F = 1) (SH/Material.slang:331-351): BxDF.x and PDF; same expression order as the
// reflect branch of eval_bsdf.  The caller applies the baker's own PDF / NaN / inf rejections.
__device__ __forceinline__ void bake_reflection(const Mat &m, const BsdfCtx &c, float3 V, float3 L, float &bx, float &pdf) {
    bx = 0.0f; pdf = 0.0f;
    if (L.z <= 1e-5f) return;
    const float3 H = normalize(V + L);
    const float VdotH = dot(V, H);
    const float D = ggx_d(c, H), GL = ggx_g1(m, L);
    pdf = (c.GV * fmaxf(VdotH, 0.0f) * D / V.z) / (4.0f * VdotH);
    bx = (((1.0f * D) * c.GV) * GL) / c.fourVz;
}
// Material.EvaluateRefraction(V, L, F = 1) (SH/Material.slang:359-387)
__device__ __forceinline__ void bake_refraction(const Mat &m, const BsdfCtx &c, float3 V, float3 L, float &bx, float &pdf) {
    bx = 0.0f; pdf = 0.0f;
    if (L.z >= 1e-5f) return;
    float3 H = normalize(V * m.Eta + L);
    if (H.z < 0.0f) H = -H;
    const float VdotH = dot(V, H), LdotH = dot(L, H);
    const float D = ggx_d(c, H), GL = ggx_g1(m, L);
    const float G = c.GV * GL;
    const float den = LdotH + m.Eta * VdotH;
    const float den2 = den * den;
    const float eta2 = m.Eta * m.Eta;
    const float jac = (eta2 * fabsf(LdotH)) / den2;
    pdf = (c.GV * fabsf(VdotH) * D / V.z) * jac;
    const float k = fabsf(VdotH) * fabsf(LdotH) / fabsf(V.z);
    bx = ((((1.0f * D) * G) * eta2) / den2) * k;
}

// grid.x covers the texels, grid.y the dispatch slices; partial[slice * n_texels + texel]
template <int KIND>
__global__ void __launch_bounds__(128) k_bake_lut(float *__restrict__ partial, uint32_t SX, uint32_t SY, uint32_t SZ, uint32_t sample_count,
                                                  uint32_t hashed_seed, uint32_t loops, uint32_t loops_per_slice) {
    const uint32_t n = SX * SY * SZ;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const uint32_t tx = idx % SX, ty = (idx / SX) % SY, tz = idx / (SX * SY);
    const BakeMat bm = bake_texel_params(KIND, SX, SY, SZ, tx, ty, tz);
    Mat m; m.Ax = bm.Ax; m.Ay = bm.Ay; m.Eta = bm.Eta;
    BsdfCtx c; c.ax2 = m.Ax * m.Ax; c.ay2 = m.Ay * m.Ay; c.dnorm = PT_PI * m.Ax * m.Ay;
    const float viewCosine = bm.viewCosine;
    const uint32_t l0 = blockIdx.y * loops_per_slice, l1 = min(loops, l0 + loops_per_slice);
    float table = 0.0f;
    for (uint32_t i = l0; i < l1; i++) {
        Rng rng; rng.s = ty + tx * tx + pcg_hash(i * 2u + sample_count + hashed_seed);
        float finalValue = 0.0f;
        for (int k = 0; k < 20; k++) {
            const float xy = sqrtf(1.0f - viewCosine * viewCosine);
            const float phi = rng.next() * (2.0f * PT_PI);
            float sp, cp; pt_sincos(phi, &sp, &cp);
            const float3 V = normalize(f3(xy * cp, xy * sp, viewCosine));
            const float3 H = ggx_sample_vndf(rng, V, m.Ax, m.Ay);
            c.GV = ggx_g1(m, V); c.fourVz = 4.0f * V.z;
            float bx, pdf;
            if (KIND == 0) {
                const float3 L = normalize(reflect3(-V, H));
                if (L.z <= 0.0f) continue;
                bake_reflection(m, c, V, L, bx, pdf);
                if (pdf <= 0.0f) continue;                       // LookupReflect.slang:73-77
                if (isnan(bx) || isinf(bx)) continue;
                finalValue += bx / pdf;
            } else {
                const float F = dielectric_fresnel(fabsf(dot(V, H)), m.Eta);
                float val = 0.0f;
                if (rng.next() < F) {
                    const float3 L = normalize(reflect3(-V, H));
                    if (L.z > 0.0f) { bake_reflection(m, c, V, L, bx, pdf); if (pdf > 0.0f && !isnan(bx) && !isinf(bx)) val += bx / pdf; }
                } else {
                    const float3 L = normalize(refract3(-V, H, m.Eta));
                    if (L.z < 0.0f) { bake_refraction(m, c, V, L, bx, pdf); if (pdf > 0.0f && !isnan(bx) && !isinf(bx)) val += bx / pdf; }
                }
                if (!isnan(val) && !isinf(val)) finalValue += val;
            }
        }
        table += finalValue / 20.0f;
    }
    partial[(size_t)blockIdx.y * n + idx] = table;
}

__global__ void k_bake_finish(const float *__restrict__ partial, float *__restrict__ table, uint32_t n, uint32_t slices, uint32_t loops) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    float s = 0.0f;
    for (uint32_t k = 0; k < slices; k++) s += partial[(size_t)k * n + idx];
    table[idx] = loops ? s / (float)loops : 0.0f;
}

void launch_bake_lut(int kind, float *partial, float *table, uint32_t SX, uint32_t SY, uint32_t SZ, uint32_t sample_count, uint32_t seed,
                     uint32_t slices, cudaStream_t st) {
    const uint32_t n = SX * SY * SZ, loops = sample_count / 20u;
    if (slices < 1) slices = 1;
    if (slices > loops && loops) slices = loops;
    const uint32_t per = loops ? (loops + slices - 1) / slices : 0;
    // pcg_hash(seed) on the host: the same integer hash (SH/Sampler.slang:4-9 == PT/LookupTableCalculator.cpp:97-101)
    uint32_t state = seed * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    const uint32_t hs = (word >> 22u) ^ word;
    dim3 grid((n + 127) / 128, slices);
    if (kind == 0) k_bake_lut<0><<<grid, 128, 0, st>>>(partial, SX, SY, SZ, sample_count, hs, loops, per);
    else if (kind == 1) k_bake_lut<1><<<grid, 128, 0, st>>>(partial, SX, SY, SZ, sample_count, hs, loops, per);
    else k_bake_lut<2><<<grid, 128, 0, st>>>(partial, SX, SY, SZ, sample_count, hs, loops, per);
    k_bake_finish<<<(n + 255) / 256, 256, 0, st>>>(partial, table, n, slices, loops);
}

} // namespace b200pt
