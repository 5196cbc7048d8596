// post_kernels.cu -- bloom mip chain + ACES tonemap (sm_100a), HBM-streaming kernels.
// Restates PathTracer/Shaders/PostProcess/{BloomDownSample,BloomUpSample,Tonemap}.slang as dispatched by
// PathTracer/PostProcessor.cpp:193-246.  Images are float4 (RGBA32F) row-major; every thread handles one pixel with
// 16-B coalesced accesses; the 16 taps of the down/up filters hit L1/L2 (each source texel is read by 4 neighbours).
#include "shading.cuh"
#include "kernels.h"

namespace b200pt {

__device__ __forceinline__ float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// BloomDownSample.slang:32-45 (FirstDispatch)
__global__ void __launch_bounds__(256) k_bloom_threshold(const float4 *__restrict__ hdr, float4 *__restrict__ mip0, uint32_t npix, PostParams p) {
    const float start = p.BloomThreshold - p.FalloffRange, end = p.BloomThreshold + p.FalloffRange;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        const float4 s = hdr[i];
        const float br = s.x * 0.2126f + s.y * 0.7152f + s.z * 0.0722f;
        const float f = smoothstepf(start, end, br);
        mip0[i] = make_float4(s.x * f, s.y * f, s.z * f, 1.0f);
    }
}

// BloomDownSample.slang:46-64 : taps 2*xy + (a,b), a,b in [-2,1], clamp, /25, *strength (Q14)
__global__ void __launch_bounds__(256) k_bloom_down(const float4 *__restrict__ src, int sw, int sh, float4 *__restrict__ dst, int dw, int dh, PostParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    float3 acc = f3(0.0f);
    #pragma unroll
    for (int a = -2; a < 2; a++) {
        #pragma unroll
        for (int b = -2; b < 2; b++) {
            const int sx = clampi(x * 2 + a, 0, sw - 1), sy = clampi(y * 2 + b, 0, sh - 1);
            const float4 s = __ldg(src + (size_t)sy * sw + sx);
            acc.x += s.x; acc.y += s.y; acc.z += s.z;
        }
    }
    dst[(size_t)y * dw + x] = make_float4((acc.x / 25.0f) * p.BloomStrength, (acc.y / 25.0f) * p.BloomStrength, (acc.z / 25.0f) * p.BloomStrength, 1.0f);
}

// BloomUpSample.slang:21-48 : taps xy/2 + (a,b) + 1, clamp, /25, *strength, added to the finer mip in place
__global__ void __launch_bounds__(256) k_bloom_up(const float4 *__restrict__ src, int sw, int sh, float4 *__restrict__ dst, int dw, int dh, PostParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    float3 acc = f3(0.0f);
    #pragma unroll
    for (int a = -2; a < 2; a++) {
        #pragma unroll
        for (int b = -2; b < 2; b++) {
            const int sx = clampi(x / 2 + a + 1, 0, sw - 1), sy = clampi(y / 2 + b + 1, 0, sh - 1);
            const float4 s = __ldg(src + (size_t)sy * sw + sx);
            acc.x += s.x; acc.y += s.y; acc.z += s.z;
        }
    }
    const float4 cur = dst[(size_t)y * dw + x];
    dst[(size_t)y * dw + x] = make_float4((acc.x / 25.0f) * p.BloomStrength + cur.x, (acc.y / 25.0f) * p.BloomStrength + cur.y, (acc.z / 25.0f) * p.BloomStrength + cur.z, 1.0f);
}

// Tonemap.slang:20-55
__device__ __forceinline__ float3 aces_fitted(float3 c) {
    float3 v = f3(0.59719f * c.x + 0.35458f * c.y + 0.04823f * c.z,
                  0.07600f * c.x + 0.90834f * c.y + 0.01566f * c.z,
                  0.02840f * c.x + 0.13383f * c.y + 0.83777f * c.z);
    float3 a = f3(v.x * (v.x + 0.0245786f) - 0.000090537f, v.y * (v.y + 0.0245786f) - 0.000090537f, v.z * (v.z + 0.0245786f) - 0.000090537f);
    float3 b = f3(v.x * (0.983729f * v.x + 0.4329510f) + 0.238081f, v.y * (0.983729f * v.y + 0.4329510f) + 0.238081f, v.z * (0.983729f * v.z + 0.4329510f) + 0.238081f);
    float3 r = a / b;
    return f3(clampf(1.60475f * r.x + -0.53108f * r.y + -0.07367f * r.z, 0.0f, 1.0f),
              clampf(-0.10208f * r.x + 1.10813f * r.y + -0.00605f * r.z, 0.0f, 1.0f),
              clampf(-0.00327f * r.x + -0.07276f * r.y + 1.07602f * r.z, 0.0f, 1.0f));
}
__device__ __forceinline__ unsigned char unorm8(float q) {
    if (!(q > 0.0f)) q = 0.0f;
    if (q > 1.0f) q = 1.0f;
    return (unsigned char)__float2int_rn(q * 255.0f);
}
// Tonemap.slang:159-175 : HDR + bilinear(bloom, uv = xy/size, CLAMP_TO_EDGE) -> *exposure -> pow(1/gamma) -> ACES -> RGBA8 (Q15)
__global__ void __launch_bounds__(256) k_tonemap(const float4 *__restrict__ hdr, const float4 *__restrict__ bloom0, uchar4 *__restrict__ ldr, int W, int H, PostParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float u = (float)x / (float)W, v = (float)y / (float)H;
    const float fxp = u * (float)W - 0.5f, fyp = v * (float)H - 0.5f;
    const float fx = floorf(fxp), fy = floorf(fyp);
    const float ax = fxp - fx, ay = fyp - fy;
    const int x0 = clampi((int)fx, 0, W - 1), x1 = clampi((int)fx + 1, 0, W - 1);
    const int y0 = clampi((int)fy, 0, H - 1), y1 = clampi((int)fy + 1, 0, H - 1);
    const float4 t00 = __ldg(bloom0 + (size_t)y0 * W + x0), t10 = __ldg(bloom0 + (size_t)y0 * W + x1);
    const float4 t01 = __ldg(bloom0 + (size_t)y1 * W + x0), t11 = __ldg(bloom0 + (size_t)y1 * W + x1);
    const float4 h4 = hdr[(size_t)y * W + x];
    float3 bl = f3(tex_mix(tex_mix(t00.x, t10.x, ax), tex_mix(t01.x, t11.x, ax), ay),
                   tex_mix(tex_mix(t00.y, t10.y, ax), tex_mix(t01.y, t11.y, ax), ay),
                   tex_mix(tex_mix(t00.z, t10.z, ax), tex_mix(t01.z, t11.z, ax), ay));
    float3 c = (f3(h4) + bl) * p.Exposure;
    const float ig = 1.0f / p.Gamma;
    c = f3(powf(c.x, ig), powf(c.y, ig), powf(c.z, ig));
    const float3 o = aces_fitted(c);
    ldr[(size_t)y * W + x] = make_uchar4(unorm8(o.x), unorm8(o.y), unorm8(o.z), 255);
}

void launch_bloom_threshold(const float4 *hdr, float4 *mip0, uint32_t npix, PostParams p, int grid, cudaStream_t st) {
    k_bloom_threshold<<<grid, 256, 0, st>>>(hdr, mip0, npix, p);
}
void launch_bloom_down(const float4 *src, uint32_t sw, uint32_t sh, float4 *dst, uint32_t dw, uint32_t dh, PostParams p, cudaStream_t st) {
    dim3 g((dw + 31) / 32, (dh + 7) / 8);
    k_bloom_down<<<g, 256, 0, st>>>(src, (int)sw, (int)sh, dst, (int)dw, (int)dh, p);
}
void launch_bloom_up(const float4 *src, uint32_t sw, uint32_t sh, float4 *dst, uint32_t dw, uint32_t dh, PostParams p, cudaStream_t st) {
    dim3 g((dw + 31) / 32, (dh + 7) / 8);
    k_bloom_up<<<g, 256, 0, st>>>(src, (int)sw, (int)sh, dst, (int)dw, (int)dh, p);
}
void launch_tonemap(const float4 *hdr, const float4 *bloom0, uchar4 *ldr, uint32_t W, uint32_t H, PostParams p, cudaStream_t st) {
    dim3 g((W + 31) / 32, (H + 7) / 8);
    k_tonemap<<<g, 256, 0, st>>>(hdr, bloom0, ldr, (int)W, (int)H, p);
}

} // namespace b200pt
