// wavefront_kernels.cu -- the per-bounce kernels of the B200 wavefront integrator (sm_100a).
//
// The reference runs ONE ray-gen megakernel per frame that loops over bounces and calls TraceRay /
// closest-hit / miss inline (SH/RayGen.slang:9-160).  Here the same estimator is a wavefront:
//
//   k_raygen   camera sample + payload init                    SH/RayGen.slang:12-63
//   per bounce:
//     k_extend   closest-hit traversal of the live paths       SH/RayGen.slang:68-72,90 (TraceRay)
//                + sort into the miss queue and one hit queue per MATERIAL CLASS (lobe set)
//     k_shade_miss  miss shading of the miss queue             SH/Miss.slang:8-76
//     k_shade_hit<CLASS>  closest-hit shading of one class's   SH/ClosestHit.slang:20-378
//                   hit queue, emits <=2 NEE shadow requests
//     k_connect  shadow rays, payload.Emitted assembly,        SH/ClosestHit.slang:139,171-176,326-372
//                luminance clamp, throughput, Russian          SH/RayGen.slang:92-113
//                roulette, ballot/prefix-sum compaction
//   scenes whose BVH is staged in shared memory by TMA run the bounce as ONE kernel per class (k_shade_hit<CLASS, ., 2>):
//   shading + both NEE shadow queries + roulette + the next segment's closest-hit query + queue append.
//   k_resolve  NaN/Inf rejection + running mean                SH/RayGen.slang:116-159
//
// Every path owns one sample slot, so radiance accumulation needs no atomics and is deterministic.
// All per-path state is SoA float4 (coalesced 16-B lanes); live paths are kept dense by compaction.
#include "bvh_traverse.cuh"
#include "bvh_dynfetch.cuh"
#include "atmosphere.cuh"
#include "kernels.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b200pt {

// ------------------------------------------------------------------------------------------------
// partition helpers (rank r owns rows y with ((y / band) % world) == r)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline uint32_t part_global_row(uint32_t local_row, uint32_t rank, uint32_t world, uint32_t band) {
    uint32_t blk = local_row / band, in = local_row % band;
    return (blk * world + rank) * band + in;
}

// control block (device memory, u32): device_types.h (CTRL_Q)

// ------------------------------------------------------------------------------------------------
// k_raygen
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_raygen(DevConfig cfg, const DevDispatch *__restrict__ disp, uint32_t n_disp, uint32_t P,
                                                 uint32_t first_sample, const uint32_t *__restrict__ rng_carry,
                                                 PathState ps, float4 *__restrict__ sample_buf, uint32_t *__restrict__ ctrl,
                                                 WaveCounters *ctr) {
    const uint32_t n = n_disp * P;
    const uint32_t S = cfg.ScreenSplitCount;
    const uint32_t LW = (cfg.W + S - 1) / S;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t b = j / P, p = j - b * P;
        const DevDispatch dd = disp[b];
        uint32_t x, y;
        if (S == 1) {
            const uint32_t lr = p / cfg.W; x = p - lr * cfg.W;
            y = part_global_row(lr, cfg.rank, cfg.world, cfg.band_rows);
        } else {                                                           // SH/RayGen.slang:17-22
            const uint32_t ly = p / LW, lx = p - ly * LW;
            x = lx * S + dd.ChunkIndex % S; y = ly * S + dd.ChunkIndex / S;
        }
        const bool inside = (x < cfg.W && y < cfg.H);                      // :24-25
        Rng rng;
        if (first_sample) { rng.s = y + cfg.W * x + dd.Seed; sample_buf[j] = make_float4(0, 0, 0, 0); }   // :28 (Q13)
        else rng.s = rng_carry[j];
        // :35-50
        const float jx = rng.next() * (0.5f - -0.5f) + -0.5f;
        const float jy = rng.next() * (0.5f - -0.5f) + -0.5f;
        const float pcx = (float)x + 0.5f + jx, pcy = (float)y + 0.5f + jy;
        const float dx = (pcx / (float)cfg.W) * 2.0f - 1.0f, dy = (pcy / (float)cfg.H) * 2.0f - 1.0f;
        const float *VI = cfg.VI, *PI = cfg.PI;
        float3 origin = f3(VI[0] * 0.0f + VI[4] * 0.0f + VI[8] * 0.0f + VI[12] * 1.0f,
                           VI[1] * 0.0f + VI[5] * 0.0f + VI[9] * 0.0f + VI[13] * 1.0f,
                           VI[2] * 0.0f + VI[6] * 0.0f + VI[10] * 0.0f + VI[14] * 1.0f);
        float3 target = f3(PI[0] * dx + PI[4] * dy + PI[8] * 1.0f + PI[12] * 1.0f,
                           PI[1] * dx + PI[5] * dy + PI[9] * 1.0f + PI[13] * 1.0f,
                           PI[2] * dx + PI[6] * dy + PI[10] * 1.0f + PI[14] * 1.0f);
        float3 direction = mat4_dir(VI, normalize(target));
        const float3 focus = origin + direction * fmaxf(cfg.FocusDistance, 0.001f);
        const float u1 = rng.next(), u2 = rng.next();                       // RandomCircleVec, SH/Sampler.slang:103-112
        const float theta = 2.0f * PT_PI * u1, rad = sqrtf(u2);
        const float ox = (rad * cosf(theta)) * 0.5f * cfg.DepthOfFieldStrength, oy = (rad * sinf(theta)) * 0.5f * cfg.DepthOfFieldStrength;
        origin = origin + (f3(VI[0], VI[1], VI[2]) * ox + f3(VI[4], VI[5], VI[6]) * oy);
        direction = normalize(focus - origin);
        // payload init :52-63.  Pixels outside the image (chunked dispatch overhang) are born dead (Depth = MAX).
        ps.org_pdf[j] = make_float4(origin.x, origin.y, origin.z, 1.0f);
        ps.dir_rng[j] = make_float4(direction.x, direction.y, direction.z, __uint_as_float(rng.s));
        ps.thr_depth[j] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(inside ? 0u : PT_MAX_DEPTH));
        ps.rad_slot[j] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(j));
        if (ps.vol_depth) ps.vol_depth[j] = 0u;                             // payload.VolumeDepth = 0 (SH/RayGen.slang:61)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { for (uint32_t w = 0; w < CTRL_WORDS; w++) ctrl[w] = 0; ctrl[0] = n; atomicAdd(&ctr->paths, (unsigned long long)n); }
}

// ------------------------------------------------------------------------------------------------
// queue helpers (device_types.h: Queues, CTRL_Q).  Queue codes: 0 = miss, 1 + c = hit of material class c, Q_NONE = inactive lane.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ctrl_reset_parity(uint32_t *ctrl, uint32_t q) {   // counters of a bounce with parity q (called by ONE thread, before any writer of that bounce)
    ctrl[q] = 0; ctrl[8u + q] = 0; ctrl[10u + q] = 0;
    #pragma unroll
    for (uint32_t c = 0; c < 1u + MC_COUNT; c++) ctrl[CTRL_Q + 8u * q + c] = 0;
}
// All hit queues of one bounce seen as one list of n entries (k_connect, k_shadow_dyn, k_shade_volume walk every hit whatever its class)
struct HitSpan { uint32_t p1, p2, p3, n; };
__device__ __forceinline__ HitSpan hit_span(const uint32_t *ctrl, uint32_t parity) {
    const uint32_t *c = ctrl + CTRL_Q + 8u * parity + 1u;
    HitSpan h; h.p1 = c[0]; h.p2 = h.p1 + c[1]; h.p3 = h.p2 + c[2]; h.n = h.p3 + c[3];
    return h;
}
__device__ __forceinline__ uint32_t hit_entry(const Queues &q, const HitSpan &h, uint32_t j) {
    const uint32_t c = (uint32_t)(j >= h.p1) + (uint32_t)(j >= h.p2) + (uint32_t)(j >= h.p3);
    const uint32_t base = c == 0u ? 0u : (c == 1u ? h.p1 : (c == 2u ? h.p2 : h.p3));
    return q.hit[(size_t)c * q.cap + (j - base)];
}
__device__ __forceinline__ uint32_t hit_code(const DevScene &sc, uint32_t gid) {   // queue code of a hit on triangle gid
    return 1u + (gid == VOLUME_EVENT ? (uint32_t)MC_GENERAL : (uint32_t)__ldg(sc.tri_class + gid));
}
// Warp-level append of one entry per lane (code Q_NONE = nothing): the lowest lane of every code group reserves room for its group with one
// atomic -- a single atomic instruction with up to five distinct addresses per warp.  Must be called by all 32 lanes.
__device__ __forceinline__ void queue_append(uint32_t code, uint32_t entry, uint32_t *qc, const Queues &q, uint32_t lane) {
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, code);
    const int leader = __ffs((int)peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader && code != Q_NONE) base = atomicAdd(qc + code, (uint32_t)__popc(peers));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    const uint32_t k = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (code == 0u) q.miss[k] = entry;
    else if (code != Q_NONE) q.hit[(size_t)(code - 1u) * q.cap + k] = entry;
}

// ------------------------------------------------------------------------------------------------
// k_extend : closest hit for every live path; sorts the live list into the miss queue and one hit queue per material class
//            (material-sorted shading: every k_shade_hit<CLASS> warp runs one lobe set, converged)
// ------------------------------------------------------------------------------------------------
template <bool SMEM, bool PRIMARY>   // PRIMARY: camera rays (origin anywhere) -> exact slab arithmetic; later bounces start on scene surfaces -> FMA slabs
__global__ void __launch_bounds__(512) k_extend(DevScene sc, PathState ps, float4 *__restrict__ hit_out, uint32_t *__restrict__ ctrl, uint32_t parity,
                                                 Queues q, int max_stack, WaveCounters *ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    const int stride = blockDim.x;
    BvhView bv;
    if (SMEM) bv = stage_bvh_smem(sc, smem + (size_t)max_stack * blockDim.x * sizeof(int), &bar);
    else bv = global_bvh(sc);
    const uint32_t n = ctrl[parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctrl_reset_parity(ctrl, parity ^ 1u);   // counters of the NEXT bounce (last used two bounces ago)
    uint32_t *qc = ctrl + CTRL_Q + 8u * parity;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    // Queue append without block barriers and with few atomics (ncu round 1: returning per-warp atomics on hot addresses were
    // 51 % of this kernel's stalls; the block-level scan that replaced them left 2.9 warps/issue parked on __syncthreads).
    // A block owns SEGMENTS of K*256 consecutive paths (grid-stride over segments); in iteration `it` its 8 warps cover 256
    // consecutive paths (neighbouring pixels -> shared BVH nodes in L1), each warp traces its K rays without ever waiting for the
    // others.  Every lane keeps the queue codes of its K rays (3 bits each), lane `it` keeps the per-queue counts of iteration `it`
    // (6 bits x 5 queues, one redux), then ONE atomic instruction per warp (lanes 0..4, one address each) reserves room in all five
    // queues for all K iterations.  K adapts to the live count so small waves still fill the GPU.
    const uint32_t K = min(8u, max(1u, n / (gridDim.x * blockDim.x * 2u)));
    const uint32_t seg_paths = K * blockDim.x, n_seg = (n + seg_paths - 1u) / seg_paths;
    // Scenes whose materials all fall into ONE class (and have no volumes) need no per-hit class lookup and only two queues: lane `it` keeps the
    // hit / miss ballots of iteration `it`, one 64-bit atomic reserves room in both (the miss count and the class's hit count are not adjacent
    // words, so two lanes issue one 32-bit atomic each in the same instruction).
    const bool uniform = sc.uniform_class != 0xFFu && sc.pre_pass == 0u;
    if (uniform) {
        uint32_t *const q_hit_u = q.hit + (size_t)sc.uniform_class * q.cap;
        for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
            const uint32_t base_i = seg * seg_paths + (threadIdx.x & ~31u);
            uint32_t my_bh = 0, my_bm = 0, ch = 0, cm = 0;
            uint32_t i = base_i + lane;
            float4 o4 = make_float4(0, 0, 0, 0), d4 = o4;
            if (i < n) { o4 = ps.org_pdf[i]; d4 = ps.dir_rng[i]; }
            for (uint32_t it = 0; it < K; it++, i += blockDim.x) {
                if (base_i + it * blockDim.x >= n) break;                           // warp-uniform
                const bool active = i < n;
                bool hit = false;
                float4 o4n = make_float4(0, 0, 0, 0), d4n = o4n;
                if (it + 1u < K && i + blockDim.x < n) { o4n = ps.org_pdf[i + blockDim.x]; d4n = ps.dir_rng[i + blockDim.x]; }
                if (active) {
                    const float3 rd = normalize_ray(f3(d4));                    // SH/RayGen.slang:70
                    HitRec h;
                    hit = bvh_trace<SMEM, false, false, false, !PRIMARY>(bv, f3(o4), rd, 0.01f, 100000.0f, h, stack, stride, max_stack);   // :71-72
                    hit_out[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.gid));
                }
                o4 = o4n; d4 = d4n;
                const uint32_t bh = __ballot_sync(0xFFFFFFFFu, hit), bm = __ballot_sync(0xFFFFFFFFu, active && !hit);
                if (lane == it) { my_bh = bh; my_bm = bm; }
                ch += (uint32_t)__popc(bh); cm += (uint32_t)__popc(bm);
            }
            uint32_t base = 0u;
            if (lane == 0u && cm) base = atomicAdd(qc, cm);
            else if (lane == 1u && ch) base = atomicAdd(qc + 1u + sc.uniform_class, ch);
            uint32_t om = __shfl_sync(0xFFFFFFFFu, base, 0), oh = __shfl_sync(0xFFFFFFFFu, base, 1);
            i = base_i + lane;
            for (uint32_t it = 0; it < K; it++, i += blockDim.x) {
                const uint32_t bh = __shfl_sync(0xFFFFFFFFu, my_bh, (int)it), bm = __shfl_sync(0xFFFFFFFFu, my_bm, (int)it);
                if ((bh >> lane) & 1u) q_hit_u[oh + __popc(bh & lt)] = i;
                else if ((bm >> lane) & 1u) q.miss[om + __popc(bm & lt)] = i;
                oh += (uint32_t)__popc(bh); om += (uint32_t)__popc(bm);
            }
        }
    } else
    for (uint32_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        const uint32_t base_i = seg * seg_paths + (threadIdx.x & ~31u);     // first path of this warp in iteration 0
        uint32_t codes = 0u, my_cnts = 0u, totA = 0u, totB = 0u;            // totA: queues 0..2, totB: queues 3..4 (10-bit fields, <= 256 each)
        uint32_t i = base_i + lane;
        float4 o4 = make_float4(0, 0, 0, 0), d4 = o4;
        if (i < n) { o4 = ps.org_pdf[i]; d4 = ps.dir_rng[i]; }
        for (uint32_t it = 0; it < K; it++, i += blockDim.x) {
            if (base_i + it * blockDim.x >= n) { codes |= Q_NONE << (3u * it); continue; }   // warp-uniform
            const bool active = i < n;
            uint32_t code = Q_NONE;
            float4 o4n = make_float4(0, 0, 0, 0), d4n = o4n;
            if (it + 1u < K && i + blockDim.x < n) { o4n = ps.org_pdf[i + blockDim.x]; d4n = ps.dir_rng[i + blockDim.x]; }   // software pipelining of the state loads
            const uint32_t pre = (active && sc.pre_pass) ? __float_as_uint(hit_out[i].w) : 0u;   // k_volume_decide's verdict on this segment
            if (pre == VOLUME_EVENT) {
                code = 1u + MC_GENERAL;                                     // scattered inside a volume / the atmosphere: no TraceRay, SH/RayGen.slang:86-90
            } else if (pre == DEAD_EVENT) {
                code = Q_NONE;                                              // finished in k_volume_decide (below the planet surface): nothing to queue
            } else if (active) {
                const float3 rd = normalize_ray(f3(d4));                    // SH/RayGen.slang:70
                HitRec h;
                const bool hit = bvh_trace<SMEM, false, false, false, !PRIMARY>(bv, f3(o4), rd, 0.01f, 100000.0f, h, stack, stride, max_stack);   // :71-72
                hit_out[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.gid));
                code = hit ? hit_code(sc, h.gid) : 0u;
            }
            o4 = o4n; d4 = d4n;
            const uint32_t cn = __reduce_add_sync(0xFFFFFFFFu, code != Q_NONE ? 1u << (6u * code) : 0u);
            codes |= code << (3u * it);
            if (lane == it) my_cnts = cn;
            totA += (cn & 63u) | (((cn >> 6) & 63u) << 10) | (((cn >> 12) & 63u) << 20);
            totB += ((cn >> 18) & 63u) | (((cn >> 24) & 63u) << 10);
        }
        uint32_t base = 0u;
        {
            const uint32_t mytot = lane < 3u ? (totA >> (10u * lane)) & 1023u : (lane < 5u ? (totB >> (10u * (lane - 3u))) & 1023u : 0u);
            if (mytot) base = atomicAdd(qc + lane, mytot);
        }
        uint32_t runA = 0u, runB = 0u;
        i = base_i + lane;
        for (uint32_t it = 0; it < K; it++, i += blockDim.x) {
            const uint32_t code = (codes >> (3u * it)) & 7u;
            const uint32_t cn = __shfl_sync(0xFFFFFFFFu, my_cnts, (int)it);
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, code);
            const uint32_t b = __shfl_sync(0xFFFFFFFFu, base, (int)(code < 5u ? code : 0u));
            if (code != Q_NONE) {
                const uint32_t run = code < 3u ? (runA >> (10u * code)) & 1023u : (runB >> (10u * (code - 3u))) & 1023u;
                const uint32_t k = b + run + (uint32_t)__popc(peers & lt);
                if (code == 0u) q.miss[k] = i; else q.hit[(size_t)(code - 1u) * q.cap + k] = i;
            }
            runA += (cn & 63u) | (((cn >> 6) & 63u) << 10) | (((cn >> 12) & 63u) << 20);
            runB += ((cn >> 18) & 63u) | (((cn >> 24) & 63u) << 10);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&ctr->extend_rays, (unsigned long long)n);
}

// SH/RayGen.slang:116-128: an unsplit path adds all of pathLight, a path split by an atmosphere event only its colour channel
__device__ __forceinline__ void add_path_light(float4 &acc, float3 rad, int channel) {
    if (channel < 0) { acc.x += rad.x; acc.y += rad.y; acc.z += rad.z; }
    else if (channel == 0) acc.x += rad.x; else if (channel == 1) acc.y += rad.y; else acc.z += rad.z;
}
// SH/Miss.slang:8-76 + the ray-gen epilogue for a path whose segment missed (shared by k_shade_miss and the fused bounce kernel's tail)
__device__ __forceinline__ void finish_missed_path(const DevScene &sc, const DevConfig &cfg, float3 dir, uint32_t rng_state, float4 t4, float4 r4, float payPDF,
                                                   float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry) {
    const uint32_t depth = __float_as_uint(t4.w) & DEPTH_MASK;
    const int channel = dflags_channel(__float_as_uint(t4.w));
    float4 c;
    if (cfg.EnableAtmosphere) c = make_float4(0, 0, 0, 1);                  // SH/Miss.slang:11-14: with the atmosphere a miss emits nothing
    else if (cfg.ShowEnvMapDirectly || depth > 0) {
        float3 r = rotate3_cs(dir, f3(1, 0, 0), cfg.cosAl, -cfg.sinAl);    // Rotate(dir, X, -altitude): cos is even, sin odd
        r = rotate3_cs(r, f3(0, 1, 0), cfg.cosAz, -cfg.sinAz);
        float u, v; direction_to_uv(r, u, v);
        c = env_sample(sc, u, v);
    } else c = make_float4(0, 0, 0, 1);
    float3 em = f3(c.x * cfg.EnvironmentIntensity, c.y * cfg.EnvironmentIntensity, c.z * cfg.EnvironmentIntensity);
    if (cfg.FurnaceTestMode && !cfg.EnableAtmosphere) em = f3(1.0f);
    if (cfg.EnableSkyMIS && depth > 0 && !cfg.EnableAtmosphere) em = em * power_heuristic(payPDF, c.w);
    // SH/RayGen.slang:92-102 with payload.Depth == MAX_DEPTH (!= 1): the contribution is always luminance-clamped (Q3)
    float3 contribution = em * f3(t4);
    const float lum = dot(contribution, f3(0.212671f, 0.715160f, 0.072169f));
    contribution = contribution * (cfg.MaxLuminance / fmaxf(lum, cfg.MaxLuminance));
    const float3 rad = f3(r4) + contribution;
    Rng rng; rng.s = rng_state;
    (void)rng.next();                                                       // the Russian-roulette draw of this segment (:111) still advances the stream
    const uint32_t slot = __float_as_uint(r4.w);
    const bool ok = !isinf(rad.x) && !isinf(rad.y) && !isinf(rad.z) && !isnan(rad.x) && !isnan(rad.y) && !isnan(rad.z);   // :116
    float4 acc = sample_buf[slot];
    if (ok) add_path_light(acc, rad, channel);
    sample_buf[slot] = acc;
    rng_carry[slot] = rng.s;
}

// ------------------------------------------------------------------------------------------------
// k_shade_miss : SH/Miss.slang:8-76 for the miss queue.  A miss always ends the path (Depth = MAX_DEPTH), so the
//                ray-gen epilogue (SH/RayGen.slang:92-128) is applied here and the path never reaches k_connect.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shade_miss(DevScene sc, DevConfig cfg, PathState ps, uint32_t *__restrict__ ctrl, uint32_t parity, uint32_t reset_next,
                                                     const uint32_t *__restrict__ q_miss, float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry,
                                                     WaveCounters *ctr) {
    const uint32_t n = ctrl[CTRL_Q + 8u * parity];
    // fused pipeline (k_bounce): this is the first kernel of the bounce, so it clears the counters the bounce kernels are about to fill
    if (reset_next && blockIdx.x == 0 && threadIdx.x == 0) ctrl_reset_parity(ctrl, parity ^ 1u);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = q_miss[j];
        const float4 d4 = ps.dir_rng[i], t4 = ps.thr_depth[i], r4 = ps.rad_slot[i];
        finish_missed_path(sc, cfg, f3(d4), __float_as_uint(d4.w), t4, r4, ps.org_pdf[i].w, sample_buf, rng_carry);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&ctr->misses, (unsigned long long)n); atomicAdd(&ctr->shade_invocations, (unsigned long long)n); }
}

// SH/RayGen.slang:92-113 for a path that was shaded at a hit: contribution (luminance clamp unless Depth == 1, Q3), throughput update, Russian
// roulette (Q4), loop condition.  Returns whether the path continues; a finished path is folded into its sample slot (:116-128).
__device__ __forceinline__ bool path_epilogue(const DevConfig &cfg, float3 emitted, float4 b4, uint32_t newDepth, int channel, float4 thr4, float4 r4, Rng &rng,
                                              float3 &thr, float3 &rad, float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry) {
    thr = f3(thr4); rad = f3(r4);
    float3 contribution = emitted * thr;
    if (newDepth != 1u) {                                                   // Q3
        const float lum = dot(contribution, f3(0.212671f, 0.715160f, 0.072169f));
        const float scale = cfg.MaxLuminance / fmaxf(lum, cfg.MaxLuminance);
        contribution = contribution * scale;
    }
    rad = rad + contribution;
    thr = thr * (f3(b4) / b4.w);
    float p = fmaxf(thr.x, fmaxf(thr.y, thr.z));
    p = fminf(p, 1.0f);
    const float u = rng.next();                                             // Q4
    bool alive = !(p < u);
    if (alive) thr = thr / p;
    alive = alive && (newDepth < cfg.MaxDepth);                             // loop condition :66
    if (!alive) {                                                           // path finished: :116-128 (+ carry RNG for SampleCount > 1)
        const uint32_t slot = __float_as_uint(r4.w);
        const bool ok = !isinf(rad.x) && !isinf(rad.y) && !isinf(rad.z) && !isnan(rad.x) && !isnan(rad.y) && !isnan(rad.z);
        float4 acc = sample_buf[slot];
        if (ok) add_path_light(acc, rad, channel);
        sample_buf[slot] = acc;
        rng_carry[slot] = rng.s;
    }
    return alive;
}

// ------------------------------------------------------------------------------------------------
// k_shade_hit<CLASS, VOL, FUSE> : SH/ClosestHit.slang:20-378 for the hit queue of one material class.
//   FUSE == 0  writes the new ray + payload in place and <= 2 NEE shadow requests; k_connect (or k_shadow_dyn + k_connect) finishes the bounce.
//   FUSE >= 1  "k_bounce" for scenes whose BVH is staged in shared memory (no volumes): each NEE request is traced the moment it is built
//              (one copy of the any-hit traversal inside the rolled NEE loop), then SH/RayGen.slang:92-113 runs in the same thread and the
//              survivor is written compacted to the other PathState buffer -- no request / payload round trip through HBM, no k_connect.
//   FUSE == 2  additionally traces the survivor's NEXT segment (the next bounce's k_extend) and appends it to the next bounce's miss / class
//              queues: one kernel per (bounce, class), 80 B read + 84 B written per continuing path.
// ------------------------------------------------------------------------------------------------
#ifndef SHADE_MIN_BLOCKS
#define SHADE_MIN_BLOCKS 5
#endif
#ifndef BOUNCE_MIN_BLOCKS
#define BOUNCE_MIN_BLOCKS 5
#endif
// BIG: 512-thread CTAs, one per SM -- for BVHs of 64..~170 KiB, which fit shared memory only once per SM (config 4's glass scene: 108 KiB).
template <uint32_t CLASS, bool VOL, int FUSE, bool BIG = false>   // VOL: the scene has AABB volumes (volumes.cuh) -- volume events in the hit queue are skipped, NEE terms get the transmittance
__global__ void __launch_bounds__(BIG ? 512 : 128, BIG ? 1 : (FUSE ? BOUNCE_MIN_BLOCKS : SHADE_MIN_BLOCKS)) k_shade_hit(DevScene sc, DevConfig cfg, PathState ps, PathState dst, ShadeOut so,
                                                    const float4 *__restrict__ hit_in, float4 *__restrict__ hit_out,
                                                    uint32_t *__restrict__ ctrl, uint32_t parity, Queues q, Queues q_next,
                                                    float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry, int max_stack, WaveCounters *ctr) {
    constexpr uint32_t LOBES = class_lobes(CLASS);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    const int stride = blockDim.x;
    BvhView bv;
    if (FUSE) bv = stage_bvh_smem(sc, smem + (size_t)max_stack * blockDim.x * sizeof(int), &bar);
    const uint32_t n = ctrl[CTRL_Q + 8u * parity + 1u + CLASS];
    const uint32_t *__restrict__ q_hit = q.hit + (size_t)CLASS * q.cap;
    uint32_t *qc_next = ctrl + CTRL_Q + 8u * (parity ^ 1u);
    uint32_t *n_next_ptr = ctrl + (parity ^ 1u);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t n_med = 0, n_shadow = 0, n_ext = 0;
    const float4 zero4 = make_float4(0, 0, 0, 0);
    // The queue entry and the hit record of the NEXT path of this thread are fetched one iteration ahead (two dependent DRAM
    // round trips off the critical path for 5 registers); its path-state lines are pulled towards L2 meanwhile.
    const uint32_t jstep = gridDim.x * blockDim.x;
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i_nx = 0; float4 h4_nx = zero4;
    if (j < n) { i_nx = q_hit[j]; h4_nx = hit_in[i_nx]; }
    const uint32_t n_loop = FUSE ? ((n + 31u) & ~31u) : n;                  // fused: warps stay converged for the compaction collectives
    for (; j < n_loop; j += jstep) {
        const bool valid = j < n;
        // ---- results of the shading part, consumed by the fused tail
        bool shaded = false;                                                // this lane holds a path whose bounce must be finished (fused)
        float3 emitted = f3(0.0f), newO = f3(0.0f), newD = f3(0.0f);
        float4 b4 = zero4; uint32_t newDflags = 0u; Rng rng; rng.s = 0u;
        bool entered = false;
        const DevMaterial *cmp = nullptr;
        const uint32_t i = i_nx;
        if (valid) do {
            const float4 h4 = h4_nx;
            const float4 o4 = ps.org_pdf[i], d4 = ps.dir_rng[i];
            const uint32_t dflags = __float_as_uint(ps.thr_depth[i].w);
            if (j + jstep < n) {
                i_nx = q_hit[j + jstep]; h4_nx = hit_in[i_nx];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ps.org_pdf + i_nx));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ps.dir_rng + i_nx));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ps.thr_depth + i_nx));
            }
            const uint32_t depth = dflags & DEPTH_MASK;
            const uint32_t chanBits = dflags & CHANNEL_MASK;           // colour channel of a path split by an atmosphere event (0 = unsplit)
            const bool inMedium = (dflags >> 31) != 0u;
            const float3 payOrigin = f3(o4), payDir = f3(d4);
            const float payPDF = o4.w;
            rng.s = __float_as_uint(d4.w);
            const uint32_t gid = __float_as_uint(h4.w);                        // global triangle id of the hit
            if (VOL && gid == VOLUME_EVENT) break;                          // a volume scattering event: k_shade_volume

            // The sky-NEE draws are the first draws of a hit outside a medium: take them now and put the alias-table load in flight.
            EnvPick ep;
            const bool envEarly = cfg.EnableSkyMIS && !inMedium && !cfg.EnableAtmosphere;   // (with the atmosphere the sky sample is the sun disk: no table load to hide)
            if (envEarly) sample_env_begin(sc, rng, ep);

            const float3 rd = normalize(payDir);                                // WorldRayDirection()
            float4 g[7];                                                        // one 112-B gather: vertices + ids of the hit triangle
            {
                const float4 *gp = reinterpret_cast<const float4 *>(sc.shade_tris + gid);
                #pragma unroll
                for (int qq = 0; qq < 7; qq++) g[qq] = __ldg(gp + qq);
            }
            const uint32_t inst = __float_as_uint(g[0].w);
            const DevInstance &in = sc.instances[inst];
            const DevMaterial &cm = sc.materials[__float_as_uint(g[2].w)];
            cmp = &cm;
            Surface sf;
            surface_init(sf, sc, cfg, in, g, h4.y, h4.z, rd, cm);
            Mat m;
            material_init(m, sc, cfg, cm, sf);
            const bool isLight = m.EmissiveColor.x > 0.0f || m.EmissiveColor.y > 0.0f || m.EmissiveColor.z > 0.0f;   // :65
            surface_rotate_tangents(sf, m.AnisotropyRotation);                  // :67

            bool newInMedium = inMedium;
            if (inMedium) {                                                     // :80-116 (Q6)
                const float4 med = ps.medium[i]; const float med_g = ps.medium_g[i];
                const float dist = length(payOrigin - sf.WorldPos);
                if (med_g == 1.0f) {
                    // Beer-law BxDF is overwritten below (:323) -- nothing observable happens here.
                } else {
                    const float sd = -logf(rng.next()) / med.w;
                    if (sd < dist) {
                        const float3 no = payOrigin + payDir * sd;
                        const float3 nd = sample_henyey_greenstein(rng, payDir, med_g);
                        n_med++;
                        if (FUSE) {                                             // Depth and InMedium unchanged, no NEE, stale PDF (Q6)
                            newO = no; newD = nd; b4 = make_float4(med.x, med.y, med.z, payPDF); newDflags = dflags; shaded = true;
                        } else {
                            ps.org_pdf[i] = make_float4(no.x, no.y, no.z, payPDF);
                            ps.dir_rng[i] = make_float4(nd.x, nd.y, nd.z, __uint_as_float(rng.s));
                            so.bxdf_pdf[i] = make_float4(med.x, med.y, med.z, payPDF);
                            so.e0[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(dflags));
                        }
                        break;
                    }
                }
            }

            // emission :265-317 (no RNG draws: evaluated before the NEE/BSDF loop so the corner positions and the ray origin die here)
            float3 e0 = f3(0.0f);
            if (cfg.EnableMeshMIS) {
                if (depth == 0 && isLight) e0 = e0 + m.EmissiveColor;
                else if (isLight) {
                    const float3 w1 = xf_point(in.o2w, sf.P1), w2 = xf_point(in.o2w, sf.P2), w3 = xf_point(in.o2w, sf.P3);
                    const float area = length(cross(w2 - w1, w3 - w1)) * 0.5f;
                    const float3 dl = sf.WorldPos - payOrigin;
                    const float d2 = dot(dl, dl);
                    const float cosTheta = fabsf(dot(sf.Normal, normalize(payOrigin - sf.WorldPos)));
                    float lp = (1.0f / (float)sc.n_emissive) * (1.0f / (float)in.emissive_tri_count) * (1.0f / area) * (d2 / cosTheta);
                    lp = fmaxf(lp, cfg.EmissiveMeshSamplingPDFBias);
                    e0 = e0 + m.EmissiveColor * power_heuristic(payPDF, lp);
                }
            } else e0 = e0 + m.EmissiveColor;
            emitted = e0;

            float3 V = normalize(-rd);
            V = sf.world_to_tangent(V);
            BsdfCtx bc;
            bsdf_ctx_init<LOBES>(bc, m, sc, cfg, V);
            // Three uses of EvaluateBSDF(V, .) per hit -- sky NEE (:125-147, :326-358), light NEE (:154-184, :360-372) and the sampled
            // direction (SampleBSDF :94-165) -- run as ONE rolled loop in the reference's RNG order (sky draws, light draws, BSDF draws):
            // a single copy of the BSDF code stays in the kernel, and each NEE request is built and stored (or, fused, traced) inside its
            // own iteration so its direction / radiance / pdf registers die there.
            Eval evS; evS.BxDF = f3(0.0f); evS.PDF = 0.0f;
            float3 Ls = f3(0.0f); bool validDir = false;
            uint32_t reqMask = 0u;                                              // bit 0: sky request stored, bit 1: light request stored
            #pragma unroll 1
            for (int k = 0; k < 3; k++) {
                float3 toW = f3(0.0f), dk = f3(0.0f); float4 lv = zero4; uint32_t lgid = 0xFFFFFFFFu;
                bool need = false;
                if (k == 0) {
                    if (cfg.EnableSkyMIS) {
                        if (cfg.EnableAtmosphere) sample_sun_disk(cfg, rng, toW, lv);   // ImportanceSampleSky under ENABLE_ATMOSPHERE (SH/Sampler.slang:465-476)
                        else {
                            if (!envEarly) sample_env_begin(sc, rng, ep);
                            sample_env_finish(sc, cfg, ep, toW, lv);
                        }
                        lv.x *= cfg.EnvironmentIntensity; lv.y *= cfg.EnvironmentIntensity; lv.z *= cfg.EnvironmentIntensity;   // Q7
                        dk = sf.world_to_tangent(toW);
                        need = lv.w > 0.0f;
                    }
                } else if (k == 1) {
                    if (cfg.EnableMeshMIS && !isLight) {
                        sample_emissive(sc, rng, sf.WorldPos, toW, lv, lgid);
                        if (lv.w > 0.0f) { dk = sf.world_to_tangent(toW); need = true; }
                    }
                } else {
                    const float3 H = ggx_sample_vndf(rng, V, m.Ax, m.Ay);           // :191-204
                    validDir = sample_bsdf_direction<LOBES>(m, bc, rng, V, H, Ls);
                    dk = Ls; need = validDir;
                }
                Eval e; e.BxDF = f3(0.0f); e.PDF = 0.0f;
                if (need) e = eval_bsdf<LOBES>(m, bc, cfg, V, dk);
                if (k == 2) { evS = e; break; }
                // NEE term (added iff the shadow query allows) :326-372.  With the atmosphere on, the sky query is issued even when its term is zero:
                // k_connect draws the transmittance walk whenever the sun is VISIBLE (SH/ClosestHit.slang:335-349 sits outside the pdf tests)
                const bool contributes = need && e.PDF > 0.0f;
                if (contributes || (k == 0 && (cfg.EnableAtmosphere || (VOL && sc.n_grids)) && cfg.EnableSkyMIS)) {   // grid volumes: same reason (:332-333)
                    const float3 c = contributes ? (((e.BxDF * 1.0f) * f3(lv)) / lv.w) * power_heuristic(lv.w, e.PDF) : f3(0.0f);
                    const float3 ro = (k == 0) ? sf.WorldPos + sf.Normal * 1e-5f : sf.WorldPos + toW * 1e-2f;   // :139, :171
                    if (FUSE) {
                        //   sky   (SH/ClosestHit.slang:139 + :326-358): any hit in (1e-4, 1e6) occludes;
                        //   light (:171-176 + :360-372): the closest hit must be the sampled triangle.  Equivalent occlusion form: the ray hits
                        //         that triangle at tL and nothing lies in front of it (ties at tL resolve to the lower triangle id, exactly like
                        //         the closest-hit query) -- bounded by tL and free to stop at the first occluder.
                        // Both run through the TARGET any-hit query: for the sky ray target = 0xFFFFFFFF and tmax_test = tmax = 1e6.
                        float tmax = 1000000.0f, tmax_test = 1000000.0f; bool go = true;
                        if (k == 1) {
                            const float4 *tp = bv.tris + (size_t)__ldg(sc.tri_slot + lgid) * 3;
                            const float4 ta = tp[0], tb = tp[1], tc = tp[2];
                            float tL, uL, vL;
                            go = tri_test(f3(ta), f3(tb), f3(tc), ro, toW, 0.0001f, 1000000.0f, tL, uL, vL);
                            tmax = tL; tmax_test = __uint_as_float(__float_as_uint(tL) + 1u);
                        }
                        n_shadow++;
                        if (go) {
                            HitRec hs;
                            const bool occluded = bvh_trace<true, true, false, true, true>(bv, ro, toW, 0.0001f, tmax, hs, stack, stride, max_stack, nullptr, nullptr, lgid, tmax_test);
                            if (!occluded) emitted = emitted + c;
                        }
                    } else {
                        float4 *const po = (k == 0) ? so.sky_o : so.lit_o, *const pd = (k == 0) ? so.sky_d : so.lit_d, *const pc = (k == 0) ? so.sky_c : so.lit_c;
                        po[i] = make_float4(ro.x, ro.y, ro.z, 1.0f);
                        pd[i] = make_float4(toW.x, toW.y, toW.z, __uint_as_float(lgid));        // .w of the light request: id of the sampled triangle
                        pc[i] = make_float4(c.x, c.y, c.z, 0.0f);
                        reqMask |= 1u << k;
                    }
                }
            }
            BSample ss;
            ss.L = validDir ? Ls : f3(0.0f); ss.BxDF = evS.BxDF; ss.PDF = evS.PDF;
            const bool wasRefracted = ss.L.z < 0.0f;
            const float3 scatterW = sf.tangent_to_world(ss.L);
            if (!wasRefracted && dot(scatterW, sf.GeometryNormal) < 0.0f) { ss.PDF = 0.0f; ss.BxDF = f3(0.0f); }   // :220-225
            if (wasRefracted && sf.HitFromInside) newInMedium = false;          // :227-238
            else if (wasRefracted && !sf.HitFromInside) {                       // entering: the medium parameters are re-read from the material
                newInMedium = true;                                             // record here instead of being carried through the loop above
                entered = true;
                if (!FUSE) {
                    const b200pt_material &mm = cm.m;
                    const float3 mc = cfg.FurnaceTestMode ? f3(1.0f) : f3(mm.MediumColor[0], mm.MediumColor[1], mm.MediumColor[2]);   // SH/Material.slang:78-86
                    ps.medium[i] = make_float4(mc.x, mc.y, mc.z, mm.MediumDensity); ps.medium_g[i] = mm.MediumAnisotropy;
                }
            }

            // payload write :319-324, :375-376
            const float off = -1e-3f * (wasRefracted ? 1.0f : 0.0f) + 1e-3f * (wasRefracted ? 0.0f : 1.0f);
            const float3 no = sf.WorldPos + sf.Normal * off;
            const bool invalid = ss.PDF <= 0.0f;
            const uint32_t newDepth = invalid ? PT_MAX_DEPTH + depth : depth + 1u;   // MAX_DEPTH*(invalid) + (Depth + 1*(!invalid))
            if (FUSE) {
                newO = no; newD = scatterW; b4 = make_float4(ss.BxDF.x, ss.BxDF.y, ss.BxDF.z, ss.PDF);
                newDflags = newDepth | chanBits | (newInMedium ? 0x80000000u : 0u); shaded = true;
            } else {
                if (VOL && reqMask && !sc.n_grids) {                            // volumes cast shadows on the NEE terms: transmittance from the NEW origin (:332-333, :364); with grid volumes k_connect walks it
                    if (reqMask & 1u) { const float T = volumes_transmittance(sc, no, f3(so.sky_d[i])); const float4 c = so.sky_c[i]; so.sky_c[i] = make_float4(c.x * T, c.y * T, c.z * T, 0.0f); }
                    if (reqMask & 2u) { const float T = volumes_transmittance(sc, no, f3(so.lit_d[i])); const float4 c = so.lit_c[i]; so.lit_c[i] = make_float4(c.x * T, c.y * T, c.z * T, 0.0f); }
                }
                ps.org_pdf[i] = make_float4(no.x, no.y, no.z, ss.PDF);
                ps.dir_rng[i] = make_float4(scatterW.x, scatterW.y, scatterW.z, __uint_as_float(rng.s));
                so.bxdf_pdf[i] = make_float4(ss.BxDF.x, ss.BxDF.y, ss.BxDF.z, ss.PDF);
                so.e0[i] = make_float4(e0.x, e0.y, e0.z, __uint_as_float(newDepth | chanBits | (reqMask << 29) | (newInMedium ? 0x80000000u : 0u)));   // bits 29/30: stored shadow requests
            }
        } while (false);

        if (FUSE) {
            // ---- SH/RayGen.slang:92-113, then (FUSE == 2) the next segment's TraceRay, then compaction into the other PathState buffer
            bool alive = false;
            float4 r4 = zero4; float3 thr = f3(0.0f), rad = f3(0.0f);
            if (shaded) {
                const float4 thr4 = ps.thr_depth[i]; r4 = ps.rad_slot[i];
                alive = path_epilogue(cfg, emitted, b4, newDflags & DEPTH_MASK, dflags_channel(newDflags), thr4, r4, rng, thr, rad, sample_buf, rng_carry);
            }
            uint32_t code = Q_NONE; HitRec h; h.t = 0.0f; h.u = 0.0f; h.v = 0.0f; h.gid = 0xFFFFFFFFu;
            if (FUSE == 2 && alive) {
                const float3 rd2 = normalize_ray(newD);                     // SH/RayGen.slang:70-72 of the next loop iteration
                const bool hit = bvh_trace<true, false, false, false, true>(bv, newO, rd2, 0.01f, 100000.0f, h, stack, stride, max_stack);
                code = hit ? hit_code(sc, h.gid) : 0u;
                n_ext++;
            }
            const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, alive);
            if (ballot) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(n_next_ptr, (uint32_t)__popc(ballot));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const uint32_t k = base + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                if (alive) {
                    dst.org_pdf[k] = make_float4(newO.x, newO.y, newO.z, b4.w);
                    dst.dir_rng[k] = make_float4(newD.x, newD.y, newD.z, __uint_as_float(rng.s));
                    dst.thr_depth[k] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(newDflags));
                    dst.rad_slot[k] = make_float4(rad.x, rad.y, rad.z, r4.w);
                    if (newDflags >> 31) {
                        float4 mv; float mg;
                        if (entered) {
                            const b200pt_material &mm = cmp->m;
                            const float3 mc = cfg.FurnaceTestMode ? f3(1.0f) : f3(mm.MediumColor[0], mm.MediumColor[1], mm.MediumColor[2]);   // SH/Material.slang:78-86
                            mv = make_float4(mc.x, mc.y, mc.z, mm.MediumDensity); mg = mm.MediumAnisotropy;
                        } else { mv = ps.medium[i]; mg = ps.medium_g[i]; }
                        dst.medium[k] = mv; dst.medium_g[k] = mg;
                    }
                    if (FUSE == 2) hit_out[k] = make_float4(h.t, h.u, h.v, __uint_as_float(h.gid));
                }
                if (FUSE == 2) queue_append(code, k, qc_next, q_next, lane);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { n_med += __shfl_down_sync(0xFFFFFFFFu, n_med, o); n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, o); n_ext += __shfl_down_sync(0xFFFFFFFFu, n_ext, o); }
    if (lane == 0 && n_med) atomicAdd(&ctr->medium_events, (unsigned long long)n_med);
    if (FUSE && lane == 0 && n_shadow) atomicAdd(&ctr->shadow_rays, (unsigned long long)n_shadow);
    if (FUSE == 2 && lane == 0 && n_ext) atomicAdd(&ctr->extend_rays, (unsigned long long)n_ext);
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&ctr->shade_invocations, (unsigned long long)n); atomicAdd(&ctr->surface_hits, (unsigned long long)n); }
}

// ------------------------------------------------------------------------------------------------
// k_connect : shadow queries + SH/RayGen.slang:92-113 + stream compaction of the survivors (unfused pipeline; walks every hit queue)
// ------------------------------------------------------------------------------------------------
// WALKS: the atmosphere is on or the scene has grid volumes -- NEE terms then carry random transmittance walks (atmosphere.cuh, volumes.cuh); compiled out of
// the plain instantiations, whose register budget (64) they would double
template <bool SMEM, bool TRACE, bool WALKS = false>   // TRACE = false: k_shadow_dyn already cleared the request bits of occluded rays; only the join / roulette / compaction runs here
__global__ void __launch_bounds__(512) k_connect(DevScene sc, DevConfig cfg, PathState src, PathState dst, ShadeOut so,
                                                  uint32_t *__restrict__ ctrl, uint32_t parity, Queues q,
                                                  float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry,
                                                  int max_stack, WaveCounters *ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    const int stride = blockDim.x;
    BvhView bv;
    if (SMEM && TRACE) bv = stage_bvh_smem(sc, smem + (size_t)max_stack * blockDim.x * sizeof(int), &bar);
    else bv = global_bvh(sc);
    const HitSpan span = hit_span(ctrl, parity);
    const uint32_t n = span.n;                                              // paths that hit a surface this bounce (misses ended in k_shade_miss)
    uint32_t *n_next_ptr = ctrl + (parity ^ 1u);
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t n_shadow = 0;
    const uint32_t n_round = (n + 31u) & ~31u;                              // keep warps converged for the ballots
#ifdef B200PT_CONNECT_PREFETCH
    // The queue entry of this thread's NEXT path is fetched one iteration ahead and its request header pulled towards L2; the words of the CURRENT path that
    // are only read after the shadow queries (state, contributions) are pulled towards L2 before the traversals start.
    const uint32_t jstep = gridDim.x * blockDim.x;
    uint32_t i_nx = 0;
    { const uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x; if (j0 < n) i_nx = hit_entry(q, span, j0); }
#define B200PT_PF(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#endif
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += gridDim.x * blockDim.x) {
        const bool active = j < n;
        bool alive = false;
        float4 o4, d4, r4; uint32_t newDflags = 0; Rng rng; rng.s = 0;
        float3 thr = f3(0.0f), rad = f3(0.0f);
        uint32_t i = 0;
        if (active) {
#ifdef B200PT_CONNECT_PREFETCH
            i = i_nx;
            const float4 e4 = so.e0[i];
            if (j + jstep < n) { i_nx = hit_entry(q, span, j + jstep); B200PT_PF(so.e0 + i_nx); }
            B200PT_PF(src.thr_depth + i); B200PT_PF(src.rad_slot + i); B200PT_PF(src.org_pdf + i); B200PT_PF(src.dir_rng + i); B200PT_PF(so.bxdf_pdf + i);
            if (__float_as_uint(e4.w) & (1u << 29)) { B200PT_PF(so.sky_c + i); }
            if (__float_as_uint(e4.w) & (1u << 30)) { B200PT_PF(so.lit_o + i); B200PT_PF(so.lit_d + i); B200PT_PF(so.lit_c + i); }
#else
            i = hit_entry(q, span, j);
            const float4 e4 = so.e0[i];
#endif
            uint32_t pending = (__float_as_uint(e4.w) >> 29) & 3u;         // bit 0: sky request, bit 1: light request (k_shade_hit)
            newDflags = __float_as_uint(e4.w) & 0x9FFFFFFFu;
            const uint32_t newDepth = newDflags & DEPTH_MASK;
            const int channel = dflags_channel(newDflags);
            float3 emitted = f3(e4);
            // Atmosphere: a visible sun costs the path a transmittance walk on its own RNG stream, BEFORE the roulette draw (SH/ClosestHit.slang:335-349,
            // SH/RayGen.slang:328-342,415-422) -- so the new origin and the RNG state are fetched up front instead of after the queries.
            // Grid volumes: the same for their ratio-tracking walks (SH/Volume.slang:448-517), sky then light; rayDepth is 0 after a surface hit (:333,:364),
            // VolumeDepth / VolumeDepth + 1 after a volume event (SH/RayGen.slang:326,361; k_shade_volume has already counted the event), VolumeDepth after
            // an atmosphere event, where the atmosphere walk comes FIRST (:415-421).
            const bool atm = WALKS && cfg.EnableAtmosphere != 0u, het = WALKS && sc.n_grids != 0u, walks = atm || het;
            float skyDepth = 0.0f, litDepth = 0.0f; bool atmEvent = false;
            if (walks) { o4 = src.org_pdf[i]; d4 = src.dir_rng[i]; rng.s = __float_as_uint(d4.w); }
            if (het && pending) {
                const float4 h4 = so.hit[i];
                if (__float_as_uint(h4.w) == VOLUME_EVENT) {
                    const float vd = (float)src.vol_depth[i];
                    if (__float_as_int(h4.y) >= 0) { skyDepth = vd - 1.0f; litDepth = vd; } else { skyDepth = vd; atmEvent = true; }
                }
            }
            auto sky_term = [&](float3 c, float3 dir) {
                if (atmEvent) { c = c * atm_transmittance_nee(cfg, rng, f3(o4), dir, channel); return c * volumes_transmittance_walk(sc, rng, f3(o4), dir, skyDepth); }
                if (het) c = c * volumes_transmittance_walk(sc, rng, f3(o4), dir, skyDepth);
                if (atm) c = c * atm_transmittance_nee(cfg, rng, f3(o4), dir, channel);
                return c;
            };
            // Request words are read when needed: the light request after the sky query, a contribution only if its ray came out
            // unoccluded, the path state after both queries -- little is live across the traversal loops.
            // (One resumable loop serving both rays of a lane, if-if style, was measured 1.6x SLOWER than two tight while-while loops.)
            //   sky   (SH/ClosestHit.slang:139 + :326-358): any hit in (1e-4, 1e6) occludes;
            //   light (:171-176 + :360-372): the closest hit must be the sampled triangle.  Equivalent occlusion form: the ray hits
            //         that triangle at tL and nothing lies in front of it (ties at tL resolve to the lower triangle id, exactly like
            //         the closest-hit query) -- bounded by tL and free to stop at the first occluder.
            if (TRACE) n_shadow += (pending & 1u) + (pending >> 1);
            if (!TRACE) {                                                   // surviving bits = unoccluded requests, joined in the reference's order
                if (pending & 1u) emitted = emitted + (walks ? sky_term(f3(so.sky_c[i]), f3(so.sky_d[i])) : f3(so.sky_c[i]));
                if (pending & 2u) {
                    float3 c = f3(so.lit_c[i]);
                    if (het) c = c * volumes_transmittance_walk(sc, rng, f3(o4), f3(so.lit_d[i]), litDepth);
                    emitted = emitted + c;
                }
                pending = 0u;
            }
            float4 so4 = make_float4(0, 0, 0, 0), sd4 = so4;
            if (pending & 1u) { so4 = so.sky_o[i]; sd4 = so.sky_d[i]; }
            if (pending & 1u) {
                HitRec h;
                const bool occluded = bvh_trace<SMEM, true, false, false, true>(bv, f3(so4), f3(sd4), 0.0001f, 1000000.0f, h, stack, stride, max_stack);
                if (!occluded) emitted = emitted + (walks ? sky_term(f3(so.sky_c[i]), f3(sd4)) : f3(so.sky_c[i]));
            }
            if (pending & 2u) {
                const float4 lo4 = so.lit_o[i], ld4_ = so.lit_d[i];
                const uint32_t lgid = __float_as_uint(ld4_.w);
                const float4 *tp = bv.tris + (size_t)__ldg(sc.tri_slot + lgid) * 3;
                const float4 ta = ld4<SMEM>(tp), tb = ld4<SMEM>(tp + 1), tc = ld4<SMEM>(tp + 2);
                float tL, uL, vL;
                if (tri_test(f3(ta), f3(tb), f3(tc), f3(lo4), f3(ld4_), 0.0001f, 1000000.0f, tL, uL, vL)) {
                    HitRec h;
                    const bool occluded = bvh_trace<SMEM, true, false, true, true>(bv, f3(lo4), f3(ld4_), 0.0001f, tL, h, stack, stride, max_stack, nullptr, nullptr, lgid);
                    if (!occluded) {
                        float3 c = f3(so.lit_c[i]);
                        if (het) c = c * volumes_transmittance_walk(sc, rng, f3(o4), f3(ld4_), litDepth);
                        emitted = emitted + c;
                    }
                }
            }
            // the path state is fetched only now: nothing but the emission and the request bits is live across the traversal loop
            const float4 thr4 = src.thr_depth[i]; r4 = src.rad_slot[i];
            if (!walks) { o4 = src.org_pdf[i]; d4 = src.dir_rng[i]; rng.s = __float_as_uint(d4.w); }
            const float4 b4 = so.bxdf_pdf[i];
            alive = path_epilogue(cfg, emitted, b4, newDepth, channel, thr4, r4, rng, thr, rad, sample_buf, rng_carry);
        }
        // ---- warp ballot + prefix-sum compaction of the live paths
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, alive);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(n_next_ptr, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (alive) {
                const uint32_t k = base + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                dst.org_pdf[k] = o4;
                dst.dir_rng[k] = make_float4(d4.x, d4.y, d4.z, __uint_as_float(rng.s));
                dst.thr_depth[k] = make_float4(thr.x, thr.y, thr.z, __uint_as_float(newDflags));
                dst.rad_slot[k] = make_float4(rad.x, rad.y, rad.z, r4.w);
                if (newDflags >> 31) { dst.medium[k] = src.medium[i]; dst.medium_g[k] = src.medium_g[i]; }
                if (src.vol_depth) dst.vol_depth[k] = src.vol_depth[i];
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, o);
    if (lane == 0 && n_shadow) atomicAdd(&ctr->shadow_rays, (unsigned long long)n_shadow);
}

// ------------------------------------------------------------------------------------------------
// k_extend_dyn / k_shadow_dyn : the traversal kernels for scenes whose BVH does not fit in shared memory (bvh_dynfetch.cuh).
//   control block additions: ctrl[8+p] = fetch counter of k_extend_dyn, ctrl[10+p] = fetch counter of k_shadow_dyn (p = bounce parity)
// ------------------------------------------------------------------------------------------------
template <bool SMEM, bool PRIMARY, bool WIDE>   // WIDE: BVH4 nodes (sc.nodes4) instead of the BVH2
__global__ void __launch_bounds__(256) k_extend_dyn(DevScene sc, PathState ps, float4 *__restrict__ hit_out, uint32_t *__restrict__ ctrl, uint32_t parity,
                                                     Queues q, int max_stack, int thresh, int n_top, const uint32_t *__restrict__ order, WaveCounters *ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    stack[0] = DYN_DONE;
    BvhView bv;
    const float4 *top = reinterpret_cast<const float4 *>(smem + (size_t)(max_stack + 1) * blockDim.x * sizeof(int));   // BVH4 treelet (WIDE, n_top > 0)
    if (SMEM) bv = stage_bvh_smem(sc, smem + (size_t)(max_stack + 1) * blockDim.x * sizeof(int), &bar);
    else bv = global_bvh(sc);
    if (WIDE && n_top > 0) stage_bytes_smem(smem + (size_t)(max_stack + 1) * blockDim.x * sizeof(int), sc.nodes4, (uint32_t)n_top * (uint32_t)sizeof(Bvh4Node), &bar);
    const uint32_t n = ctrl[parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) {                              // counters of the NEXT bounce (last used two bounces ago)
        ctrl_reset_parity(ctrl, parity ^ 1u);
        atomicAdd(&ctr->extend_rays, (unsigned long long)n);
    }
    uint32_t *qc = ctrl + CTRL_Q + 8u * parity;
    uint32_t *fetch = ctrl + 8u + parity;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t s_base = smem_u32(stack), s_step = blockDim.x * 4u, s_limit = s_base + (uint32_t)(max_stack + 1) * s_step;
    DynPool pool; pool.init(n, gridDim.x * (blockDim.x >> 5));
    DynRay r; r.cur = DYN_DONE; r.gid = DYN_NONE; r.t = 0.0f; r.u = 0.0f; r.v = 0.0f; r.s_top = s_base; r.n_spill = 0;
    int spill[WIDE ? DYN_SPILL : 1];
    const float4 *nodes4 = reinterpret_cast<const float4 *>(sc.nodes4);
    if (WIDE) bv.root = 0;                                                  // the BVH4 root is node 0 (bvh4_collapse_host)
    uint32_t i = DYN_NONE;
    while (true) {
        // ---- commit finished rays: hit record + miss / class queue entry (one multi-address atomic per refill event)
        const bool fin = (r.cur == DYN_DONE);
        const bool have = fin && i != DYN_NONE;
        const bool hit = have && r.gid != DYN_NONE;
        if (have && r.gid != VOLUME_EVENT) hit_out[i] = make_float4(hit ? r.t : -1.0f, r.u, r.v, __uint_as_float(r.gid));
        const uint32_t code = hit ? hit_code(sc, r.gid) : (have ? 0u : Q_NONE);
        if (__ballot_sync(0xFFFFFFFFu, have)) queue_append(code, i, qc, q, lane);
        // ---- refill
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, fin);
        if (need) {
            const uint32_t idx = pool.take(need, lane, fetch);
            if (fin) {
                i = (order != nullptr && idx != DYN_NONE) ? order[idx] : idx;   // sorted bounce: the pool hands out positions of the sorted list (k_ray_sort_*)
                const uint32_t pre = (i != DYN_NONE && sc.pre_pass) ? __float_as_uint(hit_out[i].w) : 0u;
                if (pre == VOLUME_EVENT) {
                    r.gid = VOLUME_EVENT;                                   // scattered inside a volume / the atmosphere (k_volume_decide): queued as a hit at the next commit, record kept
                } else if (pre == DEAD_EVENT) {
                    i = DYN_NONE;                                           // finished in k_volume_decide: the lane takes no ray this round
                } else if (i != DYN_NONE) {
                    const float4 o4 = ps.org_pdf[i], d4 = ps.dir_rng[i];
                    const float3 rd = normalize_ray(f3(d4));                // SH/RayGen.slang:70
                    dyn_init(r, f3(o4), rd, 0.01f, 100000.0f, 100000.0f, DYN_NONE, bv.root, s_base, s_step);   // :71-72
                }
            }
        }
        uint32_t act = __ballot_sync(0xFFFFFFFFu, r.cur != DYN_DONE);
        if (act == 0u) {                                                    // nobody traverses: drained, unless lanes hold volume events to commit
            if (__ballot_sync(0xFFFFFFFFu, i != DYN_NONE) == 0u) break;
            continue;
        }
        const int thr_now = pool.empty() ? 1 : thresh;
        do {
            if (WIDE) { while (r.cur >= 0) dyn_node4_step<!PRIMARY>(nodes4, top, n_top, r, s_step, s_limit, spill); }
            else { while (r.cur >= 0) dyn_node_step<SMEM, !PRIMARY>(bv, r, s_step, s_limit); }
            __syncwarp();
            while (r.cur < 0 && r.cur != DYN_DONE) dyn_leaf_step<SMEM, false, WIDE>(bv, r, s_step, spill);
            act = __ballot_sync(0xFFFFFFFFu, r.cur != DYN_DONE);
        } while (__popc(act) >= thr_now);
    }
}

// ------------------------------------------------------------------------------------------------
// Ray sort (scenes traversed out of L2): after the first bounce the live list is in compaction order -- neighbouring lanes hold unrelated rays,
// every warp of k_extend_dyn touches 32 different parts of a 40 MB BVH.  A counting sort of the path INDICES by an 18-bit key (15-bit Morton
// code of the origin's cell in a 32^3 grid over the scene box | direction octant) costs three light passes (~0.3 ms per 16 M paths) and hands
// k_extend_dyn spatially coherent rays; the hit queues it fills inherit the order, so k_shadow_dyn's origins are coherent too.  The path state
// itself is not moved.  Order inside a bin is whatever the atomics return: paths are independent, the image does not depend on it.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t morton_spread5(uint32_t v) {           // 5 bits -> every third bit
    v = (v | (v << 8)) & 0x0000100Fu; v = (v | (v << 4)) & 0x000010C3u; v = (v | (v << 2)) & 0x00001249u; return v;
}
__global__ void __launch_bounds__(256) k_ray_sort_keys(DevScene sc, PathState ps, const uint32_t *__restrict__ ctrl, uint32_t parity,
                                                        uint2 *__restrict__ key_rank, uint32_t *__restrict__ hist) {
    const uint32_t n = ctrl[parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ps.org_pdf[i], d4 = ps.dir_rng[i];
        const int cx = min(31, max(0, (int)((o4.x - sc.sort_lo[0]) * sc.sort_scale[0])));
        const int cy = min(31, max(0, (int)((o4.y - sc.sort_lo[1]) * sc.sort_scale[1])));
        const int cz = min(31, max(0, (int)((o4.z - sc.sort_lo[2]) * sc.sort_scale[2])));
        const uint32_t cell = morton_spread5((uint32_t)cx) | (morton_spread5((uint32_t)cy) << 1) | (morton_spread5((uint32_t)cz) << 2);
        const uint32_t oct = (d4.x < 0.0f ? 1u : 0u) | (d4.y < 0.0f ? 2u : 0u) | (d4.z < 0.0f ? 4u : 0u);
        const uint32_t key = (cell << 3) | oct;
        key_rank[i] = make_uint2(key, atomicAdd(hist + key, 1u));
    }
}
__global__ void __launch_bounds__(1024) k_ray_sort_scan(uint32_t *__restrict__ hist, uint32_t *__restrict__ offs) {   // one block: exclusive scan of SORT_BINS counts, hist cleared
    __shared__ uint32_t part[1024];
    constexpr uint32_t PER = SORT_BINS / 1024u;
    const uint32_t t = threadIdx.x, b0 = t * PER;
    uint32_t sum = 0;
    for (uint32_t k = 0; k < PER; k++) sum += hist[b0 + k];
    part[t] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) { const uint32_t v = t >= o ? part[t - o] : 0u; __syncthreads(); part[t] += v; __syncthreads(); }
    uint32_t run = part[t] - sum;
    for (uint32_t k = 0; k < PER; k++) { const uint32_t c = hist[b0 + k]; offs[b0 + k] = run; run += c; hist[b0 + k] = 0u; }
}
__global__ void __launch_bounds__(256) k_ray_sort_scatter(const uint32_t *__restrict__ ctrl, uint32_t parity, const uint2 *__restrict__ key_rank,
                                                           const uint32_t *__restrict__ offs, uint32_t *__restrict__ order) {
    const uint32_t n = ctrl[parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 kr = key_rank[i];
        order[__ldg(offs + kr.x) + kr.y] = i;
    }
}

// Shadow requests of the hit queues, compacted per warp: ring entry = (path index << 1) | kind (0 sky, 1 light).
// An occluded ray clears its request bit in e0.w (bit 29 sky, bit 30 light); k_connect<.., false> joins the surviving requests.
constexpr size_t DYN_RING_BYTES = 8 * 128 * sizeof(uint32_t);              // per-CTA request rings of k_shadow_dyn (dynamic shared memory, after the stacks)
template <bool SMEM, bool WIDE>
__global__ void __launch_bounds__(256, 4) k_shadow_dyn(DevScene sc, ShadeOut so, uint32_t *__restrict__ ctrl, uint32_t parity,
                                                     Queues q, int max_stack, int thresh, int n_top, WaveCounters *ctr) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    stack[0] = DYN_DONE;
    const size_t stack_bytes = (size_t)(max_stack + 1) * blockDim.x * sizeof(int);
    uint32_t *ring_all = reinterpret_cast<uint32_t *>(smem + stack_bytes);  // 128 entries per warp
    BvhView bv;
    const float4 *top = reinterpret_cast<const float4 *>(smem + stack_bytes + DYN_RING_BYTES);   // BVH4 treelet (WIDE, n_top > 0)
    if (SMEM) bv = stage_bvh_smem(sc, smem + stack_bytes + DYN_RING_BYTES, &bar);
    else bv = global_bvh(sc);
    if (WIDE && n_top > 0) stage_bytes_smem(smem + stack_bytes + DYN_RING_BYTES, sc.nodes4, (uint32_t)n_top * (uint32_t)sizeof(Bvh4Node), &bar);
    const HitSpan span = hit_span(ctrl, parity);
    const uint32_t n = span.n;                                              // total hit-queue length of this bounce
    uint32_t *fetch = ctrl + 10u + parity;
    const uint32_t lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    uint32_t *ring = ring_all + (threadIdx.x >> 5) * 128u;
    const uint32_t s_base = smem_u32(stack), s_step = blockDim.x * 4u, s_limit = s_base + (uint32_t)(max_stack + 1) * s_step;
    DynPool pool; pool.init(n, gridDim.x * (blockDim.x >> 5));
    DynRay r; r.cur = DYN_DONE; r.s_top = s_base; r.gid = DYN_NONE; r.t = 0.0f; r.n_spill = 0;
    int spill[WIDE ? DYN_SPILL : 1];
    const float4 *nodes4 = reinterpret_cast<const float4 *>(sc.nodes4);
    if (WIDE) bv.root = 0;
    uint32_t head = 0, tail = 0;                                            // warp-uniform ring cursors
    uint32_t n_shadow = 0;
    while (true) {
        const bool fin = (r.cur == DYN_DONE);
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, fin);
        const uint32_t cnt = (uint32_t)__popc(need);
        // ---- top the ring up with the requests of the next 32 hit-queue entries until it covers the idle lanes
        while (tail - head < cnt && !pool.empty()) {
            const uint32_t j = pool.take(0xFFFFFFFFu, lane, fetch);
            uint32_t pend = 0, i = 0;
            if (j != DYN_NONE) { i = hit_entry(q, span, j); pend = (__float_as_uint(so.e0[i].w) >> 29) & 3u; }
            const uint32_t c = (pend & 1u) + (pend >> 1);
            uint32_t incl = c;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((int)lane >= o) incl += y; }
            uint32_t w = tail + incl - c;
            if (pend & 1u) { ring[w & 127u] = i << 1; w++; }
            if (pend & 2u) ring[w & 127u] = (i << 1) | 1u;
            tail += __shfl_sync(0xFFFFFFFFu, incl, 31);
            n_shadow += c;
            __syncwarp();
        }
        // ---- hand requests to the idle lanes
        if (fin) {
            const uint32_t rank = (uint32_t)__popc(need & lt);
            if (rank < tail - head) {
                const uint32_t e = ring[(head + rank) & 127u];
                const uint32_t i = e >> 1;
                if (e & 1u) {                                               // light: SH/ClosestHit.slang:171-176 in its bounded any-hit form (k_connect)
                    const float4 lo4 = so.lit_o[i], ld4_ = so.lit_d[i];
                    const uint32_t lgid = __float_as_uint(ld4_.w);
                    const float4 *tp = bv.tris + (size_t)__ldg(sc.tri_slot + lgid) * 3;
                    const float4 ta = ld4<SMEM>(tp), tb = ld4<SMEM>(tp + 1), tc = ld4<SMEM>(tp + 2);
                    float tL, uL, vL;
                    if (tri_test(f3(ta), f3(tb), f3(tc), f3(lo4), f3(ld4_), 0.0001f, 1000000.0f, tL, uL, vL)) {
                        dyn_init(r, f3(lo4), f3(ld4_), 0.0001f, tL, __uint_as_float(__float_as_uint(tL) + 1u), lgid, bv.root, s_base, s_step);
                        r.gid = e;
                    } else {
                        atomicAnd(reinterpret_cast<unsigned int *>(&so.e0[i].w), ~(1u << 30));
                    }
                } else {                                                    // sky: SH/ClosestHit.slang:139 -- any hit in (1e-4, 1e6) occludes
                    const float4 so4 = so.sky_o[i], sd4 = so.sky_d[i];
                    dyn_init(r, f3(so4), f3(sd4), 0.0001f, 1000000.0f, 1000000.0f, DYN_NONE, bv.root, s_base, s_step);
                    r.gid = e;
                }
            }
        }
        { const uint32_t avail = tail - head; head += cnt < avail ? cnt : avail; }
        __syncwarp();
        uint32_t act = __ballot_sync(0xFFFFFFFFu, r.cur != DYN_DONE);
        if (act == 0u) { if (tail == head && pool.empty()) break; else continue; }
        const int thr_now = (tail == head && pool.empty()) ? 1 : thresh;
        do {
            if (WIDE) { while (r.cur >= 0) dyn_node4_step<true>(nodes4, top, n_top, r, s_step, s_limit, spill); }
            else { while (r.cur >= 0) dyn_node_step<SMEM, true>(bv, r, s_step, s_limit); }
            __syncwarp();
            while (r.cur < 0 && r.cur != DYN_DONE) {
                if (dyn_leaf_step<SMEM, true, WIDE>(bv, r, s_step, spill))  // occluded: drop the request
                    atomicAnd(reinterpret_cast<unsigned int *>(&so.e0[r.gid >> 1].w), ~(1u << (29u + (r.gid & 1u))));
            }
            act = __ballot_sync(0xFFFFFFFFu, r.cur != DYN_DONE);
        } while (__popc(act) >= thr_now);
    }
    for (int o = 16; o > 0; o >>= 1) n_shadow += __shfl_down_sync(0xFFFFFFFFu, n_shadow, o);
    if (lane == 0 && n_shadow) atomicAdd(&ctr->shadow_rays, (unsigned long long)n_shadow);
}

// ------------------------------------------------------------------------------------------------
// k_volume_decide / k_shade_volume : homogeneous AABB volumes (volumes.cuh; SH/RayGen.slang:162-380)
// ------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(256) k_volume_decide(DevScene sc, DevConfig cfg, PathState ps, ShadeOut so, const uint32_t *__restrict__ ctrl, uint32_t parity, int max_stack,
                                                        float4 *__restrict__ sample_buf, uint32_t *__restrict__ rng_carry) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    BvhView bv;
    if (SMEM) bv = stage_bvh_smem(sc, smem + (size_t)max_stack * blockDim.x * sizeof(int), &bar);
    else bv = global_bvh(sc);
    const uint32_t n = ctrl[parity];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 o4 = ps.org_pdf[i], d4 = ps.dir_rng[i];
        const float3 o = f3(o4), d = f3(d4);
        Rng rng; rng.s = __float_as_uint(d4.w);
        if (cfg.EnableAtmosphere && atm_height(cfg, o) < 0.0f) {            // SH/RayGen.slang:76-84: below the surface of the planet -> the path loop breaks (no roulette draw)
            const float4 t4 = ps.thr_depth[i], r4 = ps.rad_slot[i];
            const float3 rad = f3(r4);
            const uint32_t slot = __float_as_uint(r4.w);
            const bool ok = !isinf(rad.x) && !isinf(rad.y) && !isinf(rad.z) && !isnan(rad.x) && !isnan(rad.y) && !isnan(rad.z);
            float4 acc = sample_buf[slot];
            if (ok) add_path_light(acc, rad, dflags_channel(__float_as_uint(t4.w)));
            sample_buf[slot] = acc; rng_carry[slot] = rng.s;
            so.hit[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(DEAD_EVENT));
            continue;
        }
        HitRec h;                                                           // GetDistanceToGeometry (SH/RTCommon.slang:86-100): payload.Direction as is, tmin 1e-5, tmax 1e6
        const bool found = bvh_trace<SMEM, false>(bv, o, d, 0.00001f, 1000000.0f, h, stack, (int)blockDim.x, max_stack);
        const float distanceToGeometry = found ? h.t : -1.0f;
        int vi = -1;
        const uint32_t dflags0 = __float_as_uint(ps.thr_depth[i].w);
        float sd = volumes_free_flight(sc, o, d, rng, (float)(dflags0 & DEPTH_MASK), vi);   // :164-209 (heterogeneous volumes: payload.Depth scales the density, :199)
        if (cfg.EnableAtmosphere) {                                         // :212-236: channel pick (while unsplit), delta tracking on that channel
            const uint32_t dflags = dflags0;
            int channel = dflags_channel(dflags);
            if (channel < 0) { const float pick = rng.next(); channel = pick < 0.33333f ? 0 : (pick < 0.66666f ? 1 : 2); }
            int component = -1;
            const float ta = atm_sample_scatter_distance(cfg, rng, o, d, channel, component);
            if (ta >= 0.0f && (ta < sd || sd < 0.0f)) {
                sd = ta; vi = -2 - component;                               // the atmosphere scattered first: event code -2 (Rayleigh) / -3 (Mie) / -4 (ozone)
                if (distanceToGeometry < 0.0f || sd < distanceToGeometry) {  // the event will happen: the ray is split to this channel from now on (:244)
                    float4 t4 = ps.thr_depth[i];
                    t4.w = __uint_as_float((dflags & ~CHANNEL_MASK) | ((uint32_t)(channel + 1) << CHANNEL_SHIFT));
                    ps.thr_depth[i] = t4;
                }
            }
        }
        ps.dir_rng[i] = make_float4(d4.x, d4.y, d4.z, __uint_as_float(rng.s));
        const bool ev = sd >= 0.0f && (distanceToGeometry < 0.0f || sd < distanceToGeometry);   // :236
        so.hit[i] = ev ? make_float4(sd, __int_as_float(vi), 0.0f, __uint_as_float(VOLUME_EVENT)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// EvaluateVolumeScatteringEvent (SH/RayGen.slang:265-380) / EvaluateAtmosphereScatteringEvent (:382-471) for the flagged entries of the hit queue
__global__ void __launch_bounds__(128) k_shade_volume(DevScene sc, DevConfig cfg, PathState ps, ShadeOut so, const uint32_t *__restrict__ ctrl, uint32_t parity,
                                                      Queues q, WaveCounters *ctr) {
    const uint32_t n = ctrl[CTRL_Q + 8u * parity + 1u + MC_GENERAL];          // volume events are queued with the general class (hit_code)
    const uint32_t *__restrict__ q_hit = q.hit + (size_t)MC_GENERAL * q.cap;
    uint32_t n_ev = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = q_hit[j];
        const float4 h4 = so.hit[i];
        if (__float_as_uint(h4.w) != VOLUME_EVENT) continue;
        n_ev++;
        const float4 o4 = ps.org_pdf[i], d4 = ps.dir_rng[i];
        const uint32_t dflags = __float_as_uint(ps.thr_depth[i].w);
        Rng rng; rng.s = __float_as_uint(d4.w);
        const float3 dir = f3(d4);
        const float3 origin = f3(o4) + dir * h4.x;                          // :267 / :384
        const int code = __float_as_int(h4.y);
        if (code <= -2) {                                                   // ---- atmosphere event
            const int component = -2 - code;
            float3 newDir;
            if (component == 0) newDir = sample_rayleigh(rng, dir);
            else if (component == 1) newDir = sample_henyey_greenstein(rng, dir, 0.85f);
            else newDir = dir;                                              // ozone only absorbs
            const float C_MIE_ABS = 4.40f * 1e-6f, C_MIE = 3.996f * 1e-6f + 4.40f * 1e-6f;
            float3 bx; float pdf; uint32_t reqMask = 0u;
            if (cfg.EnableSkyMIS) {                                         // :403-446
                float3 toSun; float4 cp;
                sample_sun_disk(cfg, rng, toSun, cp);
                cp.x *= cfg.EnvironmentIntensity; cp.y *= cfg.EnvironmentIntensity; cp.z *= cfg.EnvironmentIntensity;
                const float3 col = f3(cp) / cp.w;
                float ph = 0.0f;
                if (component == 0) { ph = phase_rayleigh(dir, toSun); const float pn = phase_rayleigh(dir, newDir); bx = f3(pn); pdf = pn; }
                else if (component == 1) { ph = phase_hg(dir, toSun, 0.85f); const float att = C_MIE_ABS / C_MIE; const float pn = phase_hg(dir, newDir, 0.85f); bx = f3(pn * (1.0f - att)); pdf = pn; }
                else { bx = f3(0.0f); pdf = 1.0f; }
                // the sun query is always issued: its transmittance walk is drawn whenever the sun is visible (:415-422); k_connect multiplies by it
                const float Tv = volumes_transmittance_shade(sc, origin, toSun);
                const float3 c = (f3(Tv) * ph) * col;
                so.sky_o[i] = make_float4(origin.x, origin.y, origin.z, 1.0f);
                so.sky_d[i] = make_float4(toSun.x, toSun.y, toSun.z, __uint_as_float(0xFFFFFFFFu));
                so.sky_c[i] = make_float4(c.x, c.y, c.z, 0.0f);
                reqMask = 1u;
            } else {                                                        // :447-461
                if (component == 0) { const float pn = phase_rayleigh(dir, newDir); bx = f3(pn); pdf = pn; }
                else { const float att = C_MIE_ABS / C_MIE; bx = f3(phase_mie_approx(dir, newDir, 0.85f) * att); pdf = phase_hg(dir, newDir, 0.85f); }
            }
            const uint32_t newDepth = (dflags & DEPTH_MASK) + 1u;           // :470 (VolumeDepth untouched)
            ps.org_pdf[i] = make_float4(origin.x, origin.y, origin.z, pdf);
            ps.dir_rng[i] = make_float4(newDir.x, newDir.y, newDir.z, __uint_as_float(rng.s));
            so.bxdf_pdf[i] = make_float4(bx.x, bx.y, bx.z, pdf);
            so.e0[i] = make_float4(0.0f, 0.0f, 0.0f, __uint_as_float(newDepth | (dflags & (0x80000000u | CHANNEL_MASK)) | (reqMask << 29)));
            continue;
        }
        const DevVolume v = sc.volumes[code];
        const float3 color = f3(v.color_alpha);
        const int vdepth = (int)ps.vol_depth[i];
        const float3 emitted = f3(v.emis_droplet) + vol_temperature_emission(sc, v, rng, origin);   // :268
        float3 toSky = f3(0.0f), toLight = f3(0.0f); float4 sky = make_float4(0, 0, 0, 0), light = sky; uint32_t lgid = 0xFFFFFFFFu;
        if (cfg.EnableSkyMIS) {                                             // :273-287
            if (cfg.EnableAtmosphere) sample_sun_disk(cfg, rng, toSky, sky);
            else { EnvPick ep; sample_env_begin(sc, rng, ep); sample_env_finish(sc, cfg, ep, toSky, sky); }
            sky.x *= cfg.EnvironmentIntensity; sky.y *= cfg.EnvironmentIntensity; sky.z *= cfg.EnvironmentIntensity;   // Q7 again (:277)
        }
        if (cfg.EnableMeshMIS) sample_emissive(sc, rng, origin, toLight, light, lgid);   // :294-308
        const float3 newDir = vol_scatter_direction(sc.phase_function, v, rng, dir, vdepth);   // :311
        const float phaseS = vol_phase(sc.phase_function, v, dir, newDir, vdepth);             // :314
        uint32_t reqMask = 0u;
        // the two visibility queries of :280-306 become shadow requests; the phase-weighted, transmittance-attenuated MIS terms of
        // :319-369 are added by k_connect iff the query comes out clear (sky: any hit occludes; light: the sampled triangle must be the closest hit)
        if (cfg.EnableSkyMIS && sky.w > 0.0f) {
            const float ph = vol_phase(sc.phase_function, v, dir, toSky, vdepth);
            if (ph > 0.0f || cfg.EnableAtmosphere || sc.n_grids) {          // atmosphere / grid volumes: the query decides whether the transmittance walks are drawn (:326-342), even for a zero term
                const float T = volumes_transmittance_shade(sc, origin, toSky);
                const float3 c = ph > 0.0f ? ((f3(T) * (color * ph)) * (f3(sky) / sky.w)) * power_heuristic(sky.w, ph) : f3(0.0f);
                so.sky_o[i] = make_float4(origin.x, origin.y, origin.z, 1.0f);
                so.sky_d[i] = make_float4(toSky.x, toSky.y, toSky.z, __uint_as_float(0xFFFFFFFFu));
                so.sky_c[i] = make_float4(c.x, c.y, c.z, 0.0f);
                reqMask |= 1u;
            }
        }
        if (cfg.EnableMeshMIS && light.w > 0.0f) {
            const float ph = vol_phase(sc.phase_function, v, dir, toLight, vdepth);
            if (ph > 0.0f || sc.n_grids) {                                  // :361 sits outside the phase test
                const float T = volumes_transmittance_shade(sc, origin, toLight);
                const float3 c = ph > 0.0f ? ((f3(T) * (color * ph)) * (f3(light) / light.w)) * power_heuristic(light.w, ph) : f3(0.0f);
                so.lit_o[i] = make_float4(origin.x, origin.y, origin.z, 1.0f);
                so.lit_d[i] = make_float4(toLight.x, toLight.y, toLight.z, __uint_as_float(lgid));
                so.lit_c[i] = make_float4(c.x, c.y, c.z, 0.0f);
                reqMask |= 2u;
            }
        }
        const uint32_t newDepth = (dflags & DEPTH_MASK) + 1u;               // :377
        ps.org_pdf[i] = make_float4(origin.x, origin.y, origin.z, phaseS);  // payload.PDF = phase (:374)
        ps.dir_rng[i] = make_float4(newDir.x, newDir.y, newDir.z, __uint_as_float(rng.s));
        so.bxdf_pdf[i] = make_float4(color.x * phaseS, color.y * phaseS, color.z * phaseS, phaseS);
        so.e0[i] = make_float4(emitted.x, emitted.y, emitted.z, __uint_as_float(newDepth | (dflags & (0x80000000u | CHANNEL_MASK)) | (reqMask << 29)));
        ps.vol_depth[i] = (uint32_t)(vdepth + 1);                           // :378
    }
    for (int o = 16; o > 0; o >>= 1) n_ev += __shfl_down_sync(0xFFFFFFFFu, n_ev, o);
    if ((threadIdx.x & 31) == 0 && n_ev) atomicAdd(&ctr->medium_events, (unsigned long long)n_ev);
}

// parity hook (b200pt_volume_walks): the two volume walks on caller-supplied rays and seeds
__global__ void __launch_bounds__(128) k_volume_walks(DevScene sc, uint32_t n, const float *__restrict__ org, const float *__restrict__ dir, const uint32_t *__restrict__ seeds,
                                                      float ray_depth, float *__restrict__ T_out, float *__restrict__ scatter_out, int32_t *__restrict__ vol_out, uint32_t *__restrict__ rng_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float3 o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
    Rng r; r.s = seeds[i];
    T_out[i] = sc.n_grids ? volumes_transmittance_walk(sc, r, o, d, ray_depth) : volumes_transmittance(sc, o, d);
    rng_out[2 * i] = r.s;
    r.s = seeds[i];
    int vi = -1;
    scatter_out[i] = volumes_free_flight(sc, o, d, r, ray_depth, vi);
    vol_out[i] = vi; rng_out[2 * i + 1] = r.s;
}

// ------------------------------------------------------------------------------------------------
// k_prepare_materials : per-material constants of materials whose textures are all 1x1 (DevMaterial::pre0..pre3)
// ------------------------------------------------------------------------------------------------
__global__ void k_prepare_materials(DevMaterial *mats, uint32_t first, uint32_t count) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    DevMaterial &dm = mats[first + j];
    if ((dm.const_mask & 29u) != 29u) { dm.const_mask &= ~32u; return; }     // base, roughness, metallic, emissive textures must all be constant
    Mat m;
    material_load_constants(m, dm.m);
    material_apply_texels(m, dm.cbase, dm.crough, dm.cmetal, dm.cemis);
    dm.pre0 = make_float4(m.BaseColor.x, m.BaseColor.y, m.BaseColor.z, m.Roughness);
    dm.pre1 = make_float4(m.EmissiveColor.x, m.EmissiveColor.y, m.EmissiveColor.z, m.Metallic);
    dm.pre2 = make_float4(m.Ax, m.Ay, m.IOR, 0.0f);
    dm.pre3 = make_float4(m.pm, m.pd, m.pg, 0.0f);
    dm.const_mask |= 32u;
}
void launch_prepare_materials(DevMaterial *mats, uint32_t first, uint32_t count, cudaStream_t st) {
    if (count) k_prepare_materials<<<(count + 127) / 128, 128, 0, st>>>(mats, first, count);
}

// ------------------------------------------------------------------------------------------------
// k_resolve : SH/RayGen.slang:130-159 for every dispatch of the wave, in dispatch order
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resolve(DevConfig cfg, const DevDispatch *__restrict__ disp, uint32_t n_disp, uint32_t P,
                                                  const float4 *__restrict__ sample_buf, float4 *__restrict__ image) {
    const uint32_t S = cfg.ScreenSplitCount;
    const uint32_t npix = cfg.W * cfg.local_rows;
    const uint32_t LW = (cfg.W + S - 1) / S;
    const float inv_spp = (float)cfg.SampleCount;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < npix; q += gridDim.x * blockDim.x) {
        float4 px = image[q];
        float3 color = f3(px);
        if (S == 1) {
            for (uint32_t b = 0; b < n_disp; b++) {
                const float3 acc = f3(sample_buf[(size_t)b * P + q]) / inv_spp;                         // :130
                const uint32_t fc = disp[b].FrameCount;
                if (fc > 0) { const float a = 1.0f / (float)(fc + 1u); color = mix3(color, acc, a); }   // :133-137
                else color = acc;
            }
        } else {
            const uint32_t y = q / cfg.W, x = q - y * cfg.W;
            const uint32_t chunk = (x % S) + (y % S) * S;
            const uint32_t sidx = (y / S) * LW + (x / S);
            for (uint32_t b = 0; b < n_disp; b++) {
                const DevDispatch dd = disp[b];
                const float3 acc = f3(sample_buf[(size_t)b * P + sidx]) / inv_spp;
                if (dd.ChunkIndex == chunk) {
                    if (dd.FrameCount > 0) { const float a = 1.0f / (float)(dd.FrameCount + 1u); color = mix3(color, acc, a); }
                    else color = acc;
                } else if (dd.FrameCount == 0 && dd.ChunkIndex == 0) color = acc;                       // splat :144-157
            }
        }
        image[q] = make_float4(color.x, color.y, color.z, 1.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// test hook: closest-hit for arbitrary rays (b200pt_trace_closest)
// ------------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(256) k_trace_rays(DevScene sc, uint32_t n, const float *__restrict__ org, const float *__restrict__ dir,
                                                     float tmin, float tmax, float *t_out, uint32_t *prim_out, uint32_t *inst_out,
                                                     float *uv_out, int max_stack, uint32_t *stats) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    int *stack = reinterpret_cast<int *>(smem) + threadIdx.x;
    BvhView bv;
    if (SMEM) bv = stage_bvh_smem(sc, smem + (size_t)max_stack * blockDim.x * sizeof(int), &bar);
    else bv = global_bvh(sc);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        HitRec h;
        uint32_t nn = 0, nt = 0;
        const bool f = stats ? bvh_trace<SMEM, false, true>(bv, f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]),
                                                            tmin, tmax, h, stack, (int)blockDim.x, max_stack, &nn, &nt)
                             : bvh_trace<SMEM, false>(bv, f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]), f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]),
                                                      tmin, tmax, h, stack, (int)blockDim.x, max_stack);
        if (stats) { stats[2 * i] = nn; stats[2 * i + 1] = nt; }
        t_out[i] = f ? h.t : -1.0f;
        uint32_t pi = 0xFFFFFFFFu, ii = 0xFFFFFFFFu;
        if (f) { const float4 *tp = bv.tris + (size_t)h.slot * 3; ii = __float_as_uint(ld4<SMEM>(tp + 1).w); pi = __float_as_uint(ld4<SMEM>(tp + 2).w); }
        prim_out[i] = pi; inst_out[i] = ii;
        uv_out[2 * i] = f ? h.u : 0.0f; uv_out[2 * i + 1] = f ? h.v : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// host-callable launchers
// ------------------------------------------------------------------------------------------------
static constexpr size_t DYN_RING_BYTES_HOST = 8 * 128 * sizeof(uint32_t);
static size_t trace_smem_bytes(const DevScene &sc, int max_stack, int threads, bool smem) {
    return (size_t)max_stack * threads * sizeof(int) + (smem ? sc.bvh_bytes : 0);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, and the C-ABI lets one process drive several GPUs
// (b200pt_create(device, ..)): the opt-in is tracked per device ordinal.
static unsigned long long g_attr_done_mask = 0ull;
static int set_attrs_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev >= 0 && dev < 64 && ((g_attr_done_mask >> dev) & 1ull)) return 0;
    // opt-in dynamic shared memory: the 227 KB per-CTA limit covers static + dynamic, so each kernel gets 227 KB minus its static part
    int rc = 0;
    auto optin = [&rc](const void *f) {
        cudaFuncAttributes a{};
        cudaError_t e = cudaFuncGetAttributes(&a, f);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)a.sharedSizeBytes);
        if (e != cudaSuccess && rc == 0) rc = (int)e;
    };
    optin((const void *)k_extend<true, true>); optin((const void *)k_extend<true, false>);
    optin((const void *)k_extend<false, true>); optin((const void *)k_extend<false, false>);
    optin((const void *)k_connect<true, true>); optin((const void *)k_connect<false, true>); optin((const void *)k_connect<false, false>);
    optin((const void *)k_connect<true, true, true>); optin((const void *)k_connect<false, true, true>); optin((const void *)k_connect<false, false, true>);
    optin((const void *)k_extend_dyn<true, true, false>); optin((const void *)k_extend_dyn<true, false, false>);
    optin((const void *)k_extend_dyn<false, true, false>); optin((const void *)k_extend_dyn<false, false, false>);
    optin((const void *)k_extend_dyn<false, true, true>); optin((const void *)k_extend_dyn<false, false, true>);
    optin((const void *)k_shadow_dyn<true, false>); optin((const void *)k_shadow_dyn<false, false>); optin((const void *)k_shadow_dyn<false, true>);
    optin((const void *)k_trace_rays<true>); optin((const void *)k_trace_rays<false>);
    optin((const void *)k_volume_decide<true>); optin((const void *)k_volume_decide<false>);
#define B200PT_OPTIN_BOUNCE(C) optin((const void *)k_shade_hit<C, false, 1>); optin((const void *)k_shade_hit<C, false, 2>); optin((const void *)k_shade_hit<C, false, 2, true>);
    B200PT_OPTIN_BOUNCE(MC_DIFFUSE) B200PT_OPTIN_BOUNCE(MC_METAL) B200PT_OPTIN_BOUNCE(MC_GLASS) B200PT_OPTIN_BOUNCE(MC_GENERAL)
#undef B200PT_OPTIN_BOUNCE
    cudaGetLastError();
    if (rc == 0 && dev >= 0 && dev < 64) g_attr_done_mask |= 1ull << dev;
    return rc;
}

// grid sizing: persistent grids = SM count x resident CTAs per SM for the chosen shared-memory footprint
int query_launch_cfg(const DevScene &sc, int bvh_max_depth, int bvh4_depth, LaunchCfg *lc) {
    { const int a = set_attrs_for_current_device(); if (a != 0) return a; }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop; cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    const int sms = prop.multiProcessorCount;
    lc->max_stack = bvh_max_depth + 2; if (lc->max_stack < 4) lc->max_stack = 4; if (lc->max_stack > 64) lc->max_stack = 64;
    // Whole BVH in shared memory: <= 64 KiB -> several CTAs per SM (Cornell: 1.3 KiB); up to what fits beside the stacks of ONE 512-thread CTA per SM
    // -> "big" mode (config 4's glass scene: 108 KiB): k_extend, k_connect and the fused bounce kernels run 512-thread CTAs.
    const bool aligned = sc.bvh_bytes > 0 && (sc.bvh_bytes % 16u) == 0;
    const bool fits_small = aligned && sc.bvh_bytes <= 64u * 1024u;
    // Measured on config 4's glass scene (profiles/r02_variants.txt): 856 Mpaths/s with the one-ray-per-thread kernels over the shared-memory BVH (8-16 warps
    // per SM, divergent LDS) against 940 with the dynamic-fetch kernels over L2 -> big mode is OPT-IN (B200PT_SMEM_BIG=1).
    bool fits_big = false;
    if (const char *e = getenv("B200PT_SMEM_BIG")) { if (atoi(e) == 1) fits_big = aligned && !fits_small && (size_t)sc.bvh_bytes + (size_t)lc->max_stack * 512u * sizeof(int) <= 226u * 1024u; }
    lc->bvh_in_smem = fits_small || fits_big;
    lc->big = fits_big;
    lc->extend_threads = fits_big ? 512 : 256;
    const size_t sh = trace_smem_bytes(sc, lc->max_stack, lc->extend_threads, lc->bvh_in_smem);
    const size_t sh256 = trace_smem_bytes(sc, lc->max_stack, 256, lc->bvh_in_smem);
    // traversal shape: dynamic-fetch while-while kernels (bvh_dynfetch.cuh) over the BVH4 when the BVH lives in L2/HBM, one ray per thread
    // over the BVH2 when the whole BVH sits in shared memory (tiny scenes: no divergence or latency to recover).
    // Overrides: B200PT_TRAV=classic|dyn, B200PT_WIDE=0|1, B200PT_DYN_THRESH=1..32, B200PT_WIDE_STACK=<shared-memory stack entries>.
    lc->trav_dyn = !lc->bvh_in_smem;
    if (const char *e = getenv("B200PT_TRAV")) { if (!strcmp(e, "dyn")) lc->trav_dyn = true; else if (!strcmp(e, "classic")) lc->trav_dyn = false; }
    lc->dyn_thresh = 20;
    if (const char *e = getenv("B200PT_DYN_THRESH")) { int v = atoi(e); if (v >= 1 && v <= 32) lc->dyn_thresh = v; }
    lc->wide = lc->trav_dyn && !lc->bvh_in_smem && sc.nodes4 != nullptr && sc.bvh_bytes >= (8u << 20);   // small BVHs stay L1/L2-hot: BVH2 is cheaper there (profiles/r01_variants.txt)
    if (const char *e = getenv("B200PT_WIDE")) { if (atoi(e) == 1 && lc->trav_dyn && !lc->bvh_in_smem && sc.nodes4 != nullptr) lc->wide = true; }
    if (const char *e = getenv("B200PT_WIDE")) { if (atoi(e) == 0) lc->wide = false; }
    lc->dyn_stack = lc->max_stack;                                          // BVH2: depth + 2 entries always suffice
    if (lc->wide) {
        const int need = 3 * bvh4_depth + 2;                                // worst case: three pushes per level
        int cap = 24;
        if (const char *e = getenv("B200PT_WIDE_STACK")) { int v = atoi(e); if (v >= 4 && v <= 96) cap = v; }
        lc->dyn_stack = need < cap ? need : cap;
        if (need - lc->dyn_stack > DYN_SPILL) lc->wide = false, lc->dyn_stack = lc->max_stack;   // deeper than shared column + overflow array: BVH2
    }
    // fused bounce kernel (k_shade_hit<., ., 2>): scenes whose BVH is staged in shared memory and walked one ray per thread.  B200PT_FUSE=0|1|2 overrides
    // (0: k_shade_hit + k_connect, 1: NEE queries and the path epilogue fused, k_extend separate, 2: next segment's TraceRay fused as well).
    // Measured (profiles/r02_variants.txt): on the Cornell box the fused kernels are SLOWER than k_extend + k_shade_hit + k_connect (87.4 / 82.5 vs 77.2 ms
    // per step) -- the traffic they save was never the bound (DRAM 13 % busy), while the shadow / extend traversals now run inside half-empty warps
    // (22 of 32 lanes per instruction) of a 65 KB kernel.  They stay available as an opt-in.
    lc->fuse = 0;
    if (const char *e = getenv("B200PT_FUSE")) { const int v = atoi(e); if (v >= 0 && v <= 2 && lc->bvh_in_smem && !lc->trav_dyn) lc->fuse = (lc->big && v == 1) ? 2 : v; }
    // BVH4 treelet in shared memory (bvh_dynfetch.cuh): the first n_top4 nodes (breadth-first order = the top levels).  B200PT_TOP_KB sizes it (0 = off).
    lc->n_top4 = 0;
    if (lc->wide) {
        // Measured on BreakfastRoom (profiles/r02_variants.txt): 1376 Mpaths/s without, 1292 / 1161 / 881 with 40 / 80 / 120 KiB -- the generic loads that serve
        // treelet and L2 lanes in one instruction cost more than the L1 hits they replace, and the footprint costs occupancy -> OPT-IN.
        int kb = 0;
        if (const char *e = getenv("B200PT_TOP_KB")) { const int v = atoi(e); if (v >= 0 && v <= 160) kb = v; }
        const uint32_t want = (uint32_t)kb * 1024u / (uint32_t)sizeof(Bvh4Node);
        lc->n_top4 = (int)(want < sc.n_nodes4 ? want : sc.n_nodes4);
    }
    const size_t sh_dyn = trace_smem_bytes(sc, lc->dyn_stack + 1, 256, lc->bvh_in_smem) + (size_t)lc->n_top4 * sizeof(Bvh4Node);
    int occ_e = 0, occ_c = 0, occ_s = 0, occ_sh = 0, occ_b = 0;
    if (lc->trav_dyn) {
        if (lc->wide) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend_dyn<false, false, true>, 256, sh_dyn);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sh, k_shadow_dyn<false, true>, 256, sh_dyn + DYN_RING_BYTES_HOST);
        } else if (lc->bvh_in_smem) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend_dyn<true, false, false>, 256, sh_dyn);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sh, k_shadow_dyn<true, false>, 256, sh_dyn + DYN_RING_BYTES_HOST);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend_dyn<false, false, false>, 256, sh_dyn);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sh, k_shadow_dyn<false, false>, 256, sh_dyn + DYN_RING_BYTES_HOST);
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_connect<false, false>, 256, 0);
    } else if (lc->bvh_in_smem) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend<true, false>, lc->extend_threads, sh);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_connect<true, true>, lc->extend_threads, sh);
    } else {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_e, k_extend<false, false>, 256, sh);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, k_connect<false, true>, 256, sh);
    }
    if (occ_sh < 1) occ_sh = 1;
    lc->grid_shadow = sms * occ_sh;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_shade_hit<MC_GENERAL, false, 0>, 128, 0);
    if (lc->big) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_shade_hit<MC_GENERAL, false, 2, true>, 512, trace_smem_bytes(sc, lc->max_stack, 512, true));
    else if (lc->bvh_in_smem) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_shade_hit<MC_GENERAL, false, 2>, 128, trace_smem_bytes(sc, lc->max_stack, 128, true));
    if (occ_e < 1) occ_e = 1; if (occ_c < 1) occ_c = 1; if (occ_s < 1) occ_s = 1; if (occ_b < 1) occ_b = 1;
    lc->grid_extend = sms * occ_e; lc->grid_connect = sms * occ_c;
    { int occ_t = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, lc->bvh_in_smem ? k_trace_rays<true> : k_trace_rays<false>, 256, sh256); lc->grid_trace = sms * (occ_t < 1 ? 1 : occ_t); }
    lc->grid_shade = sms * occ_s;
    lc->grid_bounce = sms * occ_b;
    lc->grid_light = sms * 8;
    if (getenv("B200PT_DEBUG"))
        fprintf(stderr, "[b200pt] launch cfg: bvh depth %d (bvh4 %d) max_stack %d dyn_stack %d bvh_bytes %u smem %d dyn %d wide %d (treelet %d nodes) thresh %d fuse %d big %d | CTAs/SM extend %d connect %d shadow %d shade %d bounce %d\n",
                bvh_max_depth, bvh4_depth, lc->max_stack, lc->dyn_stack, sc.bvh_bytes, (int)lc->bvh_in_smem, (int)lc->trav_dyn, (int)lc->wide, lc->n_top4, lc->dyn_thresh, lc->fuse, (int)lc->big, occ_e, occ_c, occ_sh, occ_s, occ_b);
    return 0;
}

void launch_raygen(const LaunchCfg &lc, const DevConfig &cfg, const DevDispatch *disp, uint32_t n_disp, uint32_t P, uint32_t first_sample,
                   const uint32_t *rng_carry, PathState ps, float4 *sample_buf, uint32_t *ctrl, WaveCounters *ctr, cudaStream_t st) {
    k_raygen<<<lc.grid_light, 256, 0, st>>>(cfg, disp, n_disp, P, first_sample, rng_carry, ps, sample_buf, ctrl, ctr);
}
void launch_ray_sort(const LaunchCfg &lc, const DevScene &sc, PathState ps, const uint32_t *ctrl, uint32_t parity, uint2 *key_rank, uint32_t *hist, uint32_t *offs,
                     uint32_t *order, cudaStream_t st) {
    k_ray_sort_keys<<<lc.grid_light, 256, 0, st>>>(sc, ps, ctrl, parity, key_rank, hist);
    k_ray_sort_scan<<<1, 1024, 0, st>>>(hist, offs);
    k_ray_sort_scatter<<<lc.grid_light, 256, 0, st>>>(ctrl, parity, key_rank, offs, order);
}
void launch_extend(const LaunchCfg &lc, const DevScene &sc, PathState ps, float4 *hit_out, uint32_t *ctrl, uint32_t parity, Queues q,
                   WaveCounters *ctr, bool primary, const uint32_t *order, cudaStream_t st) {
    set_attrs_for_current_device();
    const bool smem = lc.bvh_in_smem;
    if (lc.trav_dyn) {
        const size_t sh = trace_smem_bytes(sc, lc.dyn_stack + 1, 256, smem) + (size_t)lc.n_top4 * sizeof(Bvh4Node);
#define B200PT_EXT_DYN(S, P, W) k_extend_dyn<S, P, W><<<lc.grid_extend, 256, sh, st>>>(sc, ps, hit_out, ctrl, parity, q, lc.dyn_stack, lc.dyn_thresh, lc.n_top4, order, ctr)
        if (lc.wide) { if (primary) B200PT_EXT_DYN(false, true, true); else B200PT_EXT_DYN(false, false, true); }
        else if (smem) { if (primary) B200PT_EXT_DYN(true, true, false); else B200PT_EXT_DYN(true, false, false); }
        else { if (primary) B200PT_EXT_DYN(false, true, false); else B200PT_EXT_DYN(false, false, false); }
#undef B200PT_EXT_DYN
        return;
    }
    const size_t sh = trace_smem_bytes(sc, lc.max_stack, lc.extend_threads, smem);
    if (smem) {
        if (primary) k_extend<true, true><<<lc.grid_extend, lc.extend_threads, sh, st>>>(sc, ps, hit_out, ctrl, parity, q, lc.max_stack, ctr);
        else k_extend<true, false><<<lc.grid_extend, lc.extend_threads, sh, st>>>(sc, ps, hit_out, ctrl, parity, q, lc.max_stack, ctr);
    } else {
        if (primary) k_extend<false, true><<<lc.grid_extend, 256, sh, st>>>(sc, ps, hit_out, ctrl, parity, q, lc.max_stack, ctr);
        else k_extend<false, false><<<lc.grid_extend, 256, sh, st>>>(sc, ps, hit_out, ctrl, parity, q, lc.max_stack, ctr);
    }
}
// k_shade_miss + one k_shade_hit<CLASS> per material class present in the scene (+ k_shade_volume).  fuse (0 / 1 / 2): see k_shade_hit.
// Returns the number of kernels launched.
int launch_shade(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState ps, PathState dst, ShadeOut so, const float4 *hit_in, float4 *hit_out,
                 uint32_t *ctrl, uint32_t parity, Queues q, Queues q_next, float4 *sample_buf, uint32_t *rng_carry, WaveCounters *ctr,
                 int fuse, uint32_t class_mask, cudaStream_t st) {
    set_attrs_for_current_device();
    int launched = 1;
    k_shade_miss<<<lc.grid_light, 256, 0, st>>>(sc, cfg, ps, ctrl, parity, fuse == 2 ? 1u : 0u, q.miss, sample_buf, rng_carry, ctr);
    const size_t shb = trace_smem_bytes(sc, lc.max_stack, lc.big ? 512 : 128, true);
#define B200PT_SHADE(C) do { if (class_mask & (1u << C)) { launched++;                                                                                              \
        if (fuse == 2 && lc.big) k_shade_hit<C, false, 2, true><<<lc.grid_bounce, 512, shb, st>>>(sc, cfg, ps, dst, so, hit_in, hit_out, ctrl, parity, q, q_next, sample_buf, rng_carry, lc.max_stack, ctr); \
        else if (fuse == 2) k_shade_hit<C, false, 2><<<lc.grid_bounce, 128, shb, st>>>(sc, cfg, ps, dst, so, hit_in, hit_out, ctrl, parity, q, q_next, sample_buf, rng_carry, lc.max_stack, ctr);      \
        else if (fuse == 1) k_shade_hit<C, false, 1><<<lc.grid_bounce, 128, shb, st>>>(sc, cfg, ps, dst, so, hit_in, hit_out, ctrl, parity, q, q_next, sample_buf, rng_carry, lc.max_stack, ctr); \
        else if (sc.pre_pass) k_shade_hit<C, true, 0><<<lc.grid_shade, 128, 0, st>>>(sc, cfg, ps, dst, so, hit_in, hit_out, ctrl, parity, q, q_next, sample_buf, rng_carry, lc.max_stack, ctr);   \
        else k_shade_hit<C, false, 0><<<lc.grid_shade, 128, 0, st>>>(sc, cfg, ps, dst, so, hit_in, hit_out, ctrl, parity, q, q_next, sample_buf, rng_carry, lc.max_stack, ctr); } } while (0)
    B200PT_SHADE(MC_DIFFUSE); B200PT_SHADE(MC_METAL); B200PT_SHADE(MC_GLASS); B200PT_SHADE(MC_GENERAL);
#undef B200PT_SHADE
    if (sc.pre_pass) { launched++; k_shade_volume<<<lc.grid_light, 128, 0, st>>>(sc, cfg, ps, so, ctrl, parity, q, ctr); }
    return launched;
}
void launch_volume_decide(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState ps, ShadeOut so, const uint32_t *ctrl, uint32_t parity,
                          float4 *sample_buf, uint32_t *rng_carry, cudaStream_t st) {
    set_attrs_for_current_device();
    const bool smem = lc.bvh_in_smem;
    const size_t sh = trace_smem_bytes(sc, lc.max_stack, 256, smem);
    if (smem) k_volume_decide<true><<<lc.grid_trace, 256, sh, st>>>(sc, cfg, ps, so, ctrl, parity, lc.max_stack, sample_buf, rng_carry);
    else k_volume_decide<false><<<lc.grid_trace, 256, sh, st>>>(sc, cfg, ps, so, ctrl, parity, lc.max_stack, sample_buf, rng_carry);
}
void launch_connect(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState src, PathState dst, ShadeOut so,
                    uint32_t *ctrl, uint32_t parity, Queues q, float4 *sample_buf, uint32_t *rng_carry, WaveCounters *ctr, cudaStream_t st) {
    set_attrs_for_current_device();
    const bool smem = lc.bvh_in_smem;
    const bool walks = cfg.EnableAtmosphere != 0u || sc.n_grids != 0u;      // k_connect<.., WALKS>
    if (lc.trav_dyn) {                                                      // shadow rays in their own dynamic-fetch kernel, then the join without tracing
        const size_t shd = trace_smem_bytes(sc, lc.dyn_stack + 1, 256, smem) + DYN_RING_BYTES_HOST + (size_t)lc.n_top4 * sizeof(Bvh4Node);
#define B200PT_SH_DYN(S, W) k_shadow_dyn<S, W><<<lc.grid_shadow, 256, shd, st>>>(sc, so, ctrl, parity, q, lc.dyn_stack, lc.dyn_thresh, lc.n_top4, ctr)
        if (lc.wide) B200PT_SH_DYN(false, true); else if (smem) B200PT_SH_DYN(true, false); else B200PT_SH_DYN(false, false);
#undef B200PT_SH_DYN
        if (walks) k_connect<false, false, true><<<lc.grid_connect, 256, 0, st>>>(sc, cfg, src, dst, so, ctrl, parity, q, sample_buf, rng_carry, lc.max_stack, ctr);
        else k_connect<false, false><<<lc.grid_connect, 256, 0, st>>>(sc, cfg, src, dst, so, ctrl, parity, q, sample_buf, rng_carry, lc.max_stack, ctr);
        return;
    }
    const size_t sh = trace_smem_bytes(sc, lc.max_stack, smem ? lc.extend_threads : 256, smem);
#define B200PT_CONNECT(S, W, T) k_connect<S, true, W><<<lc.grid_connect, T, sh, st>>>(sc, cfg, src, dst, so, ctrl, parity, q, sample_buf, rng_carry, lc.max_stack, ctr)
    if (smem) { if (walks) B200PT_CONNECT(true, true, lc.extend_threads); else B200PT_CONNECT(true, false, lc.extend_threads); }
    else { if (walks) B200PT_CONNECT(false, true, 256); else B200PT_CONNECT(false, false, 256); }
#undef B200PT_CONNECT
}
void launch_resolve(const LaunchCfg &lc, const DevConfig &cfg, const DevDispatch *disp, uint32_t n_disp, uint32_t P,
                    const float4 *sample_buf, float4 *image, cudaStream_t st) {
    k_resolve<<<lc.grid_light, 256, 0, st>>>(cfg, disp, n_disp, P, sample_buf, image);
}
void launch_volume_walks(const DevScene &sc, uint32_t n, const float *org, const float *dir, const uint32_t *seeds, float ray_depth,
                         float *T_out, float *scatter_out, int32_t *vol_out, uint32_t *rng_out, cudaStream_t st) {
    k_volume_walks<<<(n + 127u) / 128u, 128, 0, st>>>(sc, n, org, dir, seeds, ray_depth, T_out, scatter_out, vol_out, rng_out);
}
void launch_trace_rays(const LaunchCfg &lc, const DevScene &sc, uint32_t n, const float *org, const float *dir, float tmin, float tmax,
                       float *t_out, uint32_t *prim_out, uint32_t *inst_out, float *uv_out, uint32_t *stats, cudaStream_t st) {
    set_attrs_for_current_device();
    const bool smem = lc.bvh_in_smem;
    const size_t sh = trace_smem_bytes(sc, lc.max_stack, 256, smem);
    if (smem) k_trace_rays<true><<<lc.grid_trace, 256, sh, st>>>(sc, n, org, dir, tmin, tmax, t_out, prim_out, inst_out, uv_out, lc.max_stack, stats);
    else k_trace_rays<false><<<lc.grid_trace, 256, sh, st>>>(sc, n, org, dir, tmin, tmax, t_out, prim_out, inst_out, uv_out, lc.max_stack, stats);
}

} // namespace b200pt
