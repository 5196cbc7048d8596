// kernels.h -- host-callable launchers of the sm_100a kernels (wavefront_kernels.cu, lbvh.cu, post_kernels.cu)
#pragma once
#include "device_types.h"
#include <cuda_runtime.h>

namespace b200pt {

struct LaunchCfg {
    int grid_light;     // grid for streaming kernels (raygen / resolve / post)
    int grid_trace;     // persistent grid for the ray-query test hook
    int grid_extend, grid_connect;   // persistent grids of the traversal kernels (SM count x resident CTAs of each kernel)
    int grid_shade;     // persistent grid for the shading kernel
    int max_stack;      // traversal stack entries (BVH depth + 2)
    bool bvh_in_smem;   // whole BVH staged to shared memory by TMA
    bool trav_dyn;      // dynamic-fetch while-while traversal kernels (k_extend_dyn / k_shadow_dyn) instead of one ray per thread
    int dyn_thresh;     // a warp refills its idle lanes when fewer than this many lanes are still traversing
    int grid_shadow;    // persistent grid of k_shadow_dyn
    bool wide;          // the dynamic-fetch kernels walk the BVH4 (DevScene::nodes4)
    int dyn_stack;      // shared-memory stack entries per thread of the dynamic-fetch kernels (BVH4: overflow goes to local memory)
    bool big;           // bvh_in_smem with a BVH > 64 KiB: one 512-thread CTA per SM for k_extend and the fused bounce kernels
    int extend_threads; // CTA size of k_extend (256, big: 512)
    int n_top4;         // BVH4 nodes (top levels, breadth-first prefix) staged into shared memory by the dynamic-fetch kernels
    int fuse;           // 0: k_shade_hit + k_connect; 1: NEE queries + path epilogue inside k_shade_hit; 2: + the next segment's TraceRay (one kernel per bounce and class)
    int grid_bounce;    // persistent grid of the fused k_shade_hit instantiations (BVH + stacks in shared memory)
};

int query_launch_cfg(const DevScene &sc, int bvh_max_depth, int bvh4_depth, LaunchCfg *lc);
void launch_raygen(const LaunchCfg &lc, const DevConfig &cfg, const DevDispatch *disp, uint32_t n_disp, uint32_t P, uint32_t first_sample,
                   const uint32_t *rng_carry, PathState ps, float4 *sample_buf, uint32_t *ctrl, WaveCounters *ctr, cudaStream_t st);
void launch_extend(const LaunchCfg &lc, const DevScene &sc, PathState ps, float4 *hit_out, uint32_t *ctrl, uint32_t parity, Queues q,
                   WaveCounters *ctr, bool primary, const uint32_t *order, cudaStream_t st);
// Ray sort for the dynamic-fetch traversal of incoherent bounces: order[0..n) = the live paths of parity `parity` sorted by (origin cell, direction octant).
// key_rank: n x uint2 scratch, hist / offs: SORT_BINS words each (hist must be zero on entry and is zero again on exit).
constexpr uint32_t SORT_BINS = 1u << 18;
void launch_ray_sort(const LaunchCfg &lc, const DevScene &sc, PathState ps, const uint32_t *ctrl, uint32_t parity, uint2 *key_rank, uint32_t *hist, uint32_t *offs,
                     uint32_t *order, cudaStream_t st);
// k_shade_miss + k_shade_hit<CLASS> for every class in class_mask (+ k_shade_volume); returns the number of kernels launched.
// fuse != 0: the bounce is finished inside k_shade_hit (dst, q_next, hit_out are written), launch_connect is not called.
int launch_shade(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState ps, PathState dst, ShadeOut so, const float4 *hit_in, float4 *hit_out,
                 uint32_t *ctrl, uint32_t parity, Queues q, Queues q_next, float4 *sample_buf, uint32_t *rng_carry, WaveCounters *ctr,
                 int fuse, uint32_t class_mask, cudaStream_t st);
void launch_volume_decide(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState ps, ShadeOut so, const uint32_t *ctrl, uint32_t parity,
                          float4 *sample_buf, uint32_t *rng_carry, cudaStream_t st);   // before launch_extend when the scene has volumes or the atmosphere is on (DevScene::pre_pass)
void launch_connect(const LaunchCfg &lc, const DevScene &sc, const DevConfig &cfg, PathState src, PathState dst, ShadeOut so,
                    uint32_t *ctrl, uint32_t parity, Queues q, float4 *sample_buf, uint32_t *rng_carry, WaveCounters *ctr, cudaStream_t st);
void launch_resolve(const LaunchCfg &lc, const DevConfig &cfg, const DevDispatch *disp, uint32_t n_disp, uint32_t P,
                    const float4 *sample_buf, float4 *image, cudaStream_t st);
void launch_volume_walks(const DevScene &sc, uint32_t n, const float *org, const float *dir, const uint32_t *seeds, float ray_depth,
                         float *T_out, float *scatter_out, int32_t *vol_out, uint32_t *rng_out, cudaStream_t st);
void launch_trace_rays(const LaunchCfg &lc, const DevScene &sc, uint32_t n, const float *org, const float *dir, float tmin, float tmax,
                       float *t_out, uint32_t *prim_out, uint32_t *inst_out, float *uv_out, uint32_t *stats, cudaStream_t st);

// ---- LBVH build (lbvh.cu): world-space flattening of the two-level TLAS/BLAS ----
struct LbvhResult { ShadeTri *shade; uint32_t *tri_slot; BvhNode *nodes; BvhTri *tris; uint32_t n_nodes, n_tris; int32_t root; int max_depth; size_t bytes;
                    Bvh4Node *nodes4; uint32_t n_nodes4; int depth4;
                    float *h_ref_box; uint32_t n_prims;
                    float scene_bounds[6]; };                              // lo.xyz, hi.xyz of all triangles (world space)                 // host copy of the per-slot reference boxes (only when asked for), triangle count
// Builds into ONE contiguous allocation [nodes | tris] (so small scenes can be staged to smem with one bulk copy).
// Returns cudaError_t as int.
int lbvh_build(const b200pt_vertex *d_verts, const uint32_t *d_indices, const DevMesh *d_meshes, const DevInstance *d_instances,
               const DevInstance *h_instances, const DevMesh *h_meshes, uint32_t n_instances, uint32_t n_tris, LbvhResult *out, cudaStream_t st, bool keep_ref_boxes = false);
void lbvh_free(LbvhResult *r);
// BVH2 -> BVH4 on the host (pure CPU code, unit-tested without a GPU through b200pt_bvh4_collapse).  `out` must hold n_nodes2 entries.
// Returns the number of BVH4 nodes written (0 if the root is a leaf), *depth4 = depth of the BVH4 (root = 1).
uint32_t bvh4_collapse_host(const BvhNode *nodes2, uint32_t n_nodes2, int32_t root2, Bvh4Node *out, int *depth4);
// Opt-in binned-SAH rebuild of the inner nodes above the LBVH's leaves (pure CPU code, unit-tested through b200pt_bvh2_sah_rebuild).
// `out` must hold n_nodes2 entries; returns the node count (0 = nothing rebuilt), *depth, sah[0/1] = SAH cost before / after.
uint32_t bvh2_sah_rebuild_host(const BvhNode *nodes2, uint32_t n_nodes2, int32_t root2, BvhNode *out, int *depth, double sah[2]);
// Insertion-based refinement of any BVH2 (leaves kept); same output conventions as bvh2_sah_rebuild_host.
uint32_t bvh2_reinsert_host(const BvhNode *nodes2, uint32_t n_nodes2, int32_t root2, BvhNode *out, int passes, float fraction, int *depth, double sah[2]);
// Full binned-SAH build from per-slot reference boxes (n x 6 floats); leaves are re-formed, perm[new slot] = old slot; see lbvh.cu.
uint32_t bvh2_sah_build_host(const float *ref_boxes, uint32_t n, float trav_cost, BvhNode *out, uint32_t *perm, int *depth, double *sah_cost);
// mode 1: inner nodes only (download, rebuild, upload in place); mode 2: full build from r->h_ref_box, reference slots and tri_slot permuted
// (falls back to mode 1 when the boxes were not kept); mode 3: mode 2 + insertion-based refinement; 4: mode 1 + refinement; 5: refinement only.  Returns cudaError_t as int.
int lbvh_refine_sah(LbvhResult *r, cudaStream_t st, double sah[2], int mode = 1);
// Downloads r->nodes, collapses, uploads r->nodes4 (own allocation).  Returns cudaError_t as int; leaves nodes4 = nullptr if the root is a leaf.
int lbvh_build_wide(LbvhResult *r, cudaStream_t st);

// ---- post chain (post_kernels.cu) ----
struct PostParams { float Exposure, Gamma, BloomThreshold, BloomStrength, FalloffRange; };
void launch_accumulate(const float4 *frame, float4 *image, uint32_t first, uint32_t count, uint32_t frame_index, int grid, cudaStream_t st);   // running mean of pixels [first, first + count)
void launch_bloom_threshold(const float4 *hdr, float4 *mip0, uint32_t npix, PostParams p, int grid, cudaStream_t st);
// rows (optional): {first, end} row range of the destination the launch computes (multi-GPU post pass); nullptr = all rows
void launch_bloom_down(const float4 *src, uint32_t sw, uint32_t sh, float4 *dst, uint32_t dw, uint32_t dh, PostParams p, cudaStream_t st, const int *rows = nullptr);
// fused chain: threshold folded into the first down pass (mip 0 is never written) and [last up pass + threshold + tonemap] in one kernel
void launch_bloom_down_first(const float4 *hdr, uint32_t W, uint32_t H, float4 *mip1, uint32_t dw, uint32_t dh, PostParams p, cudaStream_t st, const int *rows = nullptr);
void launch_bloom_final(const float4 *hdr, const float4 *mip1, uint32_t mw, uint32_t mh, uchar4 *ldr, float4 *mip0_out, uint32_t W, uint32_t H, PostParams p, cudaStream_t st, const int *rows = nullptr);
// passes first..last of the chain (down first..last, then up last..first) in one cluster launch; mip[i] / w[i] / h[i] for i in [first-1, last]
struct SmallMips { float4 *mip[16]; int w[16], h[16]; int first, last; };
void launch_bloom_small(const SmallMips &m, PostParams p, cudaStream_t st);
void launch_bloom_up(const float4 *src, uint32_t sw, uint32_t sh, float4 *dst, uint32_t dw, uint32_t dh, PostParams p, cudaStream_t st, const int *rows = nullptr);
void launch_tonemap(const float4 *hdr, const float4 *bloom0, uchar4 *ldr, uint32_t W, uint32_t H, PostParams p, cudaStream_t st);
void launch_prepare_materials(DevMaterial *mats, uint32_t first, uint32_t count, cudaStream_t st);   // fills DevMaterial::pre0..pre3
// lut_baker.cu : kind 0 = reflect, 1 = refract hit-from-outside, 2 = refract hit-from-inside; partial holds slices * SX*SY*SZ floats
void launch_bake_lut(int kind, float *partial, float *table, uint32_t SX, uint32_t SY, uint32_t SZ, uint32_t sample_count, uint32_t seed,
                     uint32_t slices, cudaStream_t st);

} // namespace b200pt
