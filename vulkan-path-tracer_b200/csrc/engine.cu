// engine.cu -- host runtime of the wavefront path tracer (see engine.h).
// Orchestration per b200pt_path_trace call (all on one CUDA stream, no host sync inside a wave unless the
// scene can scatter inside a medium or MaxDepth is large enough that early exit pays):
//   upload dispatch table -> for each wave of F dispatches: raygen -> [extend, shade, connect] x bounces -> resolve
#include "engine.h"
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <cmath>

namespace b200pt {

void Engine::check(cudaError_t e, const char *what) const {
    if (e != cudaSuccess) throw CudaError{ (int)e, std::string(what) + ": " + cudaGetErrorString(e) };
}
#define CK(x) check((x), #x)

static const uint32_t kMaxDispatchTable = 1u << 16;
template <class T> static void dfree(T *&p) { if (p) { cudaFree(p); p = nullptr; } }

Engine::Engine(int device) : device_(device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) throw CudaError{ B200PT_ERR_NO_DEVICE, "no CUDA device available (the product has no CPU fallback)" };
    if (device < 0 || device >= n) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "device ordinal out of range" };
    CK(cudaSetDevice(device_));
    CK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking)); stream_ = own_stream_;
    CK(cudaEventCreate(&ev_[0])); CK(cudaEventCreate(&ev_[1]));
    b200pt_default_config(&cfg_);
    memset(view_inv_, 0, sizeof view_inv_); memset(proj_inv_, 0, sizeof proj_inv_);
    for (int i = 0; i < 4; i++) view_inv_[i * 5] = proj_inv_[i * 5] = 1.0f;
    CK(cudaMalloc(&d_ctr_, sizeof(WaveCounters)));
    CK(cudaMemset(d_ctr_, 0, sizeof(WaveCounters)));
    for (int c = 0; c < 2; c++) {
        CK(cudaMalloc(&wb_[c].counts, CTRL_WORDS * sizeof(uint32_t))); CK(cudaMallocHost(&wb_[c].h_count, 4 * sizeof(uint32_t)));
        CK(cudaStreamCreateWithFlags(&aux_stream_[c], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&resolved_ev_[c], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&wb_[c].disp_ev, cudaEventDisableTiming));
        CK(cudaMalloc(&wb_[c].d_disp, kMaxDispatchTable * sizeof(DevDispatch))); CK(cudaMallocHost(&wb_[c].h_disp, kMaxDispatchTable * sizeof(DevDispatch)));
    }
    CK(cudaEventCreateWithFlags(&image_free_ev_, cudaEventDisableTiming));
    b200pt_default_atmosphere(&atmosphere_);
    memset(&last_, 0, sizeof last_);
}

Engine::~Engine() {
    cudaSetDevice(device_);
    for (auto &st : aux_stream_) if (st) cudaStreamSynchronize(st);
    if (stream_) cudaStreamSynchronize(stream_);
    free_scene(); free_wave(); free_post();
    dfree(d_env_); dfree(d_alias_); dfree(d_env_row_cos_); for (auto &l : d_luts_) dfree(l);
    dfree(d_image_); dfree(d_ctr_); dfree(d_volumes_); dfree(d_grids_); grids_.clear(); dfree(d_tri_class_);
    for (int c = 0; c < 2; c++) {
        dfree(wb_[c].counts); if (wb_[c].h_count) cudaFreeHost(wb_[c].h_count);
        if (aux_stream_[c]) { cudaStreamSynchronize(aux_stream_[c]); cudaStreamDestroy(aux_stream_[c]); }
        if (resolved_ev_[c]) cudaEventDestroy(resolved_ev_[c]);
        if (wb_[c].disp_ev) cudaEventDestroy(wb_[c].disp_ev);
        dfree(wb_[c].d_disp); if (wb_[c].h_disp) cudaFreeHost(wb_[c].h_disp);
    }
    if (image_free_ev_) cudaEventDestroy(image_free_ev_);
    if (ev_[0]) cudaEventDestroy(ev_[0]); if (ev_[1]) cudaEventDestroy(ev_[1]);
    for (auto &e : prof_ev_) cudaEventDestroy(e);
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

void Engine::sync_all() {
    for (auto &st : aux_stream_) if (st) CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(stream_));
}
void Engine::join_waves() {
    if (wave_seq_ == 0) return;
    for (auto &e : resolved_ev_) CK(cudaStreamWaitEvent(stream_, e, 0));    // an event that was never recorded counts as complete
}

void Engine::free_scene() {
    dfree(d_verts_); dfree(d_indices_); dfree(d_meshes_); dfree(d_instances_); dfree(d_materials_); dfree(d_textures_); dfree(d_emissive_); dfree(d_em_tris_); dfree(d_em_tri_base_); dfree(d_tri_class_);
    for (auto &p : d_texdata_) if (p) cudaFree(p);
    d_texdata_.clear();
    lbvh_free(&bvh_);
    has_scene_ = false;
}

// PathTracer::SetScene (PathTracer.cpp:158-676)
void Engine::set_scene(HostScene &&scene) {
    CK(cudaSetDevice(device_));
    sync_all();
    free_scene();
    scene_ = std::move(scene);
    // capacity limits of the reference (PathTracer.h:192-195, asserts PathTracer.cpp:182-184)
    if (scene_.meshes.size() >= 10000 || scene_.materials.size() >= 10000 || scene_.instances.size() >= 100000)
        throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "scene exceeds the reference's MAX_ENTITIES / MAX_INSTANCES limits" };
    upload_scene();
    // camera: PathTracer.cpp:187-188,578 then the Editor/FlyCamera round trip
    camera_from_view(scene_.camera_view, scene_.camera_aspect, view_inv_, proj_inv_);
    // output image: W = (uint)(1080 * aspect), H = 1080 (PathTracer.cpp:509-512)
    W_ = (uint32_t)((float)1080 * scene_.camera_aspect); H_ = 1080;
    local_rows_ = partition_local_rows(H_, rank_, world_, band_);
    ensure_image();
    has_scene_ = true;
    reset();
}

void Engine::upload_scene() {
    size_t nv = 0, ni = 0;
    h_meshes_.clear();
    for (auto &m : scene_.meshes) { h_meshes_.push_back({ (uint32_t)nv, (uint32_t)ni, (uint32_t)(m.indices.size() / 3), 0 }); nv += m.vertices.size(); ni += m.indices.size(); }
    std::vector<b200pt_vertex> verts(nv); std::vector<uint32_t> idx(ni);
    for (size_t i = 0; i < scene_.meshes.size(); i++) {
        memcpy(&verts[h_meshes_[i].vbase], scene_.meshes[i].vertices.data(), scene_.meshes[i].vertices.size() * sizeof(b200pt_vertex));
        memcpy(&idx[h_meshes_[i].ibase], scene_.meshes[i].indices.data(), scene_.meshes[i].indices.size() * 4);
    }
    CK(cudaMalloc(&d_verts_, nv * sizeof(b200pt_vertex))); CK(cudaMemcpy(d_verts_, verts.data(), nv * sizeof(b200pt_vertex), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_indices_, ni * 4)); CK(cudaMemcpy(d_indices_, idx.data(), ni * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_meshes_, h_meshes_.size() * sizeof(DevMesh))); CK(cudaMemcpy(d_meshes_, h_meshes_.data(), h_meshes_.size() * sizeof(DevMesh), cudaMemcpyHostToDevice));
    // instances: one per MeshInstance in loader order == InstanceIndex (TLASImpl.cpp:29-35)
    h_instances_.assign(scene_.instances.size(), DevInstance{});
    uint32_t tri_base = 0;
    for (size_t i = 0; i < scene_.instances.size(); i++) {
        const b200pt_instance &in = scene_.instances[i]; DevInstance &d = h_instances_[i];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) d.o2w[r * 4 + c] = in.Transform[c * 4 + r];
        mat3_inverse_from_o2w(d.o2w, d.w2o);
        d.mesh = in.MeshIndex; d.material = in.MaterialIndex; d.tri_base = tri_base;
        tri_base += h_meshes_[in.MeshIndex].tri_count;
    }
    n_tris_ = tri_base;
    CK(cudaMalloc(&d_instances_, h_instances_.size() * sizeof(DevInstance)));
    {
        std::vector<DevMaterial> dm; dm.reserve(scene_.materials.size());
        for (const auto &m : scene_.materials) dm.push_back(make_dev_material(m));
        CK(cudaMalloc(&d_materials_, dm.size() * sizeof(DevMaterial)));
        CK(cudaMemcpy(d_materials_, dm.data(), dm.size() * sizeof(DevMaterial), cudaMemcpyHostToDevice));
        launch_prepare_materials(d_materials_, 0, (uint32_t)dm.size(), stream_);
        CK(cudaStreamSynchronize(stream_));
    }
    CK(cudaMalloc(&d_emissive_, std::max<size_t>(1, scene_.instances.size()) * sizeof(DevEmissive)));
    rebuild_emissive();
    // textures
    std::vector<DevTexture> dt(scene_.textures.size());
    d_texdata_.assign(scene_.textures.size(), nullptr);
    for (size_t i = 0; i < scene_.textures.size(); i++) {
        const HostTexture &t = scene_.textures[i];
        CK(cudaMalloc(&d_texdata_[i], std::max<size_t>(t.data.size(), 4)));
        CK(cudaMemcpy(d_texdata_[i], t.data.data(), t.data.size(), cudaMemcpyHostToDevice));
        dt[i] = { d_texdata_[i], t.width, t.height, t.channels, 0 };
    }
    CK(cudaMalloc(&d_textures_, dt.size() * sizeof(DevTexture))); CK(cudaMemcpy(d_textures_, dt.data(), dt.size() * sizeof(DevTexture), cudaMemcpyHostToDevice));
    // acceleration structure: GPU LBVH over the flattened instances
    // tree-quality pass over the GPU LBVH: same hits, fewer node visits / triangle tests.  Level 3 (full binned-SAH build + insertion-based
    // refinement) measured best on every workload (profiles/r02_sah_levels.txt: BreakfastRoom +12 %, viking_room +28 %, glass +11 %, Cornell +5 %);
    // B200PT_BVH_SAH=0..5 overrides (0 = plain LBVH).
    int sah_mode = 3;
    if (const char *e = getenv("B200PT_BVH_SAH")) sah_mode = (e[0] >= '0' && e[0] <= '5' && e[1] == 0) ? e[0] - '0' : 3;
    int r = lbvh_build(d_verts_, d_indices_, d_meshes_, d_instances_, h_instances_.data(), h_meshes_.data(), (uint32_t)h_instances_.size(), n_tris_, &bvh_, stream_, sah_mode == 2 || sah_mode == 3);
    if (r != 0) throw CudaError{ B200PT_ERR_CUDA, std::string("lbvh_build failed: ") + cudaGetErrorString((cudaError_t)r) };
    if (sah_mode) {
        double sah[2];
        r = lbvh_refine_sah(&bvh_, stream_, sah, sah_mode);
        if (r != 0) throw CudaError{ B200PT_ERR_CUDA, std::string("lbvh_refine_sah failed: ") + cudaGetErrorString((cudaError_t)r) };
        if (getenv("B200PT_DEBUG")) fprintf(stderr, "[b200pt] SAH pass (mode %d): cost %.3f -> %.3f, depth %d\n", sah_mode, sah[0], sah[1], bvh_.max_depth);
    }
    ds_.verts = d_verts_; ds_.indices = d_indices_; ds_.meshes = d_meshes_; ds_.instances = d_instances_; ds_.materials = d_materials_;
    CK(cudaMalloc(&d_tri_class_, std::max<size_t>(1, n_tris_)));
    rebuild_tri_class();
    ds_.textures = d_textures_; ds_.emissive = d_emissive_; ds_.nodes = bvh_.nodes; ds_.tris = bvh_.tris; ds_.shade_tris = bvh_.shade; ds_.tri_slot = bvh_.tri_slot;
    ds_.n_tris = bvh_.n_tris; ds_.n_nodes = bvh_.n_nodes; ds_.root = bvh_.root; ds_.bvh_bytes = bvh_.bytes <= 0xFFFFFFFFull ? (uint32_t)bvh_.bytes : 0;
    ds_.nodes4 = nullptr; ds_.n_nodes4 = 0;
    if (bvh_.bytes > 64u * 1024u) {                                         // scenes that traverse out of L2/HBM get the BVH4 (host collapse of the GPU-built BVH2)
        r = lbvh_build_wide(&bvh_, stream_);
        if (r != 0) throw CudaError{ B200PT_ERR_CUDA, std::string("lbvh_build_wide failed: ") + cudaGetErrorString((cudaError_t)r) };
        ds_.nodes4 = bvh_.nodes4; ds_.n_nodes4 = bvh_.n_nodes4;
    }
    // A handful of triangle slots (the 12-triangle Cornell box): shared-memory traversals test them in order, no hierarchy (bvh_traverse.cuh).
    {
        uint32_t flat_max = 16;
        if (const char *e = getenv("B200PT_FLAT_MAX")) { const int v = atoi(e); if (v >= 0 && v <= 64) flat_max = (uint32_t)v; }
        ds_.n_flat = (bvh_.n_tris <= flat_max) ? bvh_.n_tris : 0u;
    }
    for (int k = 0; k < 3; k++) {                                           // ray-sort grid over the scene box (k_ray_sort_keys)
        const float ext = bvh_.scene_bounds[3 + k] - bvh_.scene_bounds[k];
        ds_.sort_lo[k] = bvh_.scene_bounds[k]; ds_.sort_scale[k] = ext > 0.0f ? 32.0f / ext : 0.0f;
    }
    int q = query_launch_cfg(ds_, bvh_.max_depth, bvh_.depth4, &lc_);
    // Measured (profiles/r02_variants.txt): feeding k_extend_dyn rays sorted by (origin cell, direction octant) does not shorten it (BreakfastRoom bounce 1..3:
    // 1283 / 644 / 310 us unsorted, 1335 / 586 / 319 us sorted) -- incoherent traversal is bound by node visits per ray, not by cache misses -- while the sort
    // and the scattered state reads it causes cost 30-45 %.  OPT-IN: B200PT_SORT=1.
    sort_rays_ = false;
    if (const char *e = getenv("B200PT_SORT")) sort_rays_ = atoi(e) != 0 && lc_.trav_dyn;
    if (q != 0) throw CudaError{ B200PT_ERR_CUDA, "query_launch_cfg failed" };
}

// 1x1 textures (the defaults of PathTracer.cpp:1557-1621, or any constant texture) are resolved to their texel value here
DevMaterial Engine::make_dev_material(const b200pt_material &m) const {
    DevMaterial d; memset(&d, 0, sizeof d);
    d.m = m;
    auto texel = [&](uint32_t idx, float4 &out) -> bool {
        if (idx >= scene_.textures.size()) return false;
        const HostTexture &t = scene_.textures[idx];
        if (t.width != 1 || t.height != 1) return false;
        if (t.channels == 4) out = make_float4((float)t.data[0] / 255.0f, (float)t.data[1] / 255.0f, (float)t.data[2] / 255.0f, (float)t.data[3] / 255.0f);
        else out = make_float4((float)t.data[0] / 255.0f, 0.0f, 0.0f, 1.0f);
        return true;
    };
    float4 v;
    if (texel(m.BaseColorTextureIndex, v)) { d.const_mask |= 1u; d.cbase = v; }
    if (texel(m.NormalTextureIndex, v)) { d.const_mask |= 2u; d.cnormal = v; }
    if (texel(m.RoughnessTextureIndex, v)) { d.const_mask |= 4u; d.crough = v.x; }
    if (texel(m.MetallicTextureIndex, v)) { d.const_mask |= 8u; d.cmetal = v.x; }
    if (texel(m.EmissiveTextureIndex, v)) { d.const_mask |= 16u; d.cemis = v; }
    return d;
}

// Shading class of every triangle (device_types.h: MaterialClass): decides the hit queue k_extend appends to and the k_shade_hit<CLASS>
// instantiation that shades it.  Uses the same products the device evaluates (Metallic * texel, material_apply_texels in shading.cuh).
void Engine::rebuild_tri_class() {
    std::vector<uint8_t> cls(scene_.materials.size(), (uint8_t)MC_GENERAL);
    for (size_t i = 0; i < scene_.materials.size(); i++) {
        const DevMaterial dm = make_dev_material(scene_.materials[i]);
        if (!(dm.const_mask & 8u)) continue;                                // metallic texture is not 1x1: the lobe set depends on the texel
        const float metal = dm.m.Metallic * dm.cmetal, T = dm.m.Transmission;
        if (metal == 1.0f) cls[i] = (uint8_t)MC_METAL;
        else if (metal == 0.0f && T == 0.0f) cls[i] = (uint8_t)MC_DIFFUSE;
        else if (metal == 0.0f && T == 1.0f) cls[i] = (uint8_t)MC_GLASS;
    }
    if (const char *e = getenv("B200PT_CLASSES")) { if (atoi(e) == 0) std::fill(cls.begin(), cls.end(), (uint8_t)MC_GENERAL); }   // A/B: one general queue
    std::vector<uint8_t> tc(std::max<size_t>(1, n_tris_), (uint8_t)MC_GENERAL);
    class_mask_ = 0;
    for (size_t i = 0; i < scene_.instances.size(); i++) {
        const uint8_t c = cls[scene_.instances[i].MaterialIndex];
        const uint32_t first = h_instances_[i].tri_base, cnt = h_meshes_[scene_.instances[i].MeshIndex].tri_count;
        if (cnt) { class_mask_ |= 1u << c; std::fill(tc.begin() + first, tc.begin() + first + cnt, c); }
    }
    CK(cudaMemcpy(d_tri_class_, tc.data(), tc.size(), cudaMemcpyHostToDevice));
    ds_.tri_class = d_tri_class_;
    ds_.uniform_class = 0xFFu;
    for (uint32_t c = 0; c < MC_COUNT; c++) if (class_mask_ == (1u << c)) ds_.uniform_class = c;
}

// emissive-mesh list: PathTracer.cpp:458-469 (+ SetMaterial maintenance :712-810); keyed on constant EmissiveColor != 0 (Q16)
void Engine::rebuild_emissive() {
    std::vector<DevEmissive> em;
    for (size_t i = 0; i < scene_.instances.size(); i++) {
        const b200pt_instance &in = scene_.instances[i];
        const b200pt_material &m = scene_.materials[in.MaterialIndex];
        h_instances_[i].emissive_tri_count = 0;
        if (m.EmissiveColor[0] != 0.0f || m.EmissiveColor[1] != 0.0f || m.EmissiveColor[2] != 0.0f) {
            DevEmissive e; e.mesh = in.MeshIndex; e.material = in.MaterialIndex; e.tri_count = h_meshes_[in.MeshIndex].tri_count; e.instance = (uint32_t)i;
            memcpy(e.xf, in.Transform, sizeof e.xf);
            em.push_back(e);
            h_instances_[i].emissive_tri_count = e.tri_count;
        }
    }
    if (em.size() >= 10000) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "too many emissive meshes (MAX_EMISSIVE_MESHES)" };
    n_emissive_ = (uint32_t)em.size();
    if (!em.empty()) CK(cudaMemcpy(d_emissive_, em.data(), em.size() * sizeof(DevEmissive), cudaMemcpyHostToDevice));
    // world-space emissive triangles (SH/Sampler.slang:376-391: vertices fetched and transformed per sample in the reference)
    std::vector<EmTri> et; std::vector<uint32_t> base;
    for (const DevEmissive &e : em) {
        base.push_back((uint32_t)et.size());
        const HostMesh &hm = scene_.meshes[e.mesh];
        auto xf = [&](const b200pt_vertex &v) {
            const float *t = e.xf; const float x = v.Position[0], y = v.Position[1], z = v.Position[2];
            return make_float3(t[0] * x + t[4] * y + t[8] * z + t[12] * 1.0f, t[1] * x + t[5] * y + t[9] * z + t[13] * 1.0f, t[2] * x + t[6] * y + t[10] * z + t[14] * 1.0f);
        };
        for (uint32_t t = 0; t < e.tri_count; t++) {
            const b200pt_vertex &A = hm.vertices[hm.indices[t * 3]], &B = hm.vertices[hm.indices[t * 3 + 1]], &C = hm.vertices[hm.indices[t * 3 + 2]];
            const float3 p0 = xf(A), p1 = xf(B), p2 = xf(C);
            EmTri r;
            r.r[0] = make_float4(p0.x, p0.y, p0.z, A.TexCoord[0]); r.r[1] = make_float4(p1.x, p1.y, p1.z, A.TexCoord[1]);
            r.r[2] = make_float4(p2.x, p2.y, p2.z, B.TexCoord[0]);
            const uint32_t gid = h_instances_[e.instance].tri_base + t; float gidf; memcpy(&gidf, &gid, 4);
            r.r[3] = make_float4(B.TexCoord[1], C.TexCoord[0], C.TexCoord[1], gidf);
            et.push_back(r);
        }
    }
    dfree(d_em_tris_); dfree(d_em_tri_base_);
    CK(cudaMalloc(&d_em_tris_, std::max<size_t>(1, et.size()) * sizeof(EmTri))); CK(cudaMalloc(&d_em_tri_base_, std::max<size_t>(1, base.size()) * sizeof(uint32_t)));
    if (!et.empty()) CK(cudaMemcpy(d_em_tris_, et.data(), et.size() * sizeof(EmTri), cudaMemcpyHostToDevice));
    if (!base.empty()) CK(cudaMemcpy(d_em_tri_base_, base.data(), base.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ds_.em_tris = d_em_tris_; ds_.em_tri_base = d_em_tri_base_;
    CK(cudaMemcpy(d_instances_, h_instances_.data(), h_instances_.size() * sizeof(DevInstance), cudaMemcpyHostToDevice));
    ds_.n_emissive = n_emissive_;
}

void Engine::set_material(uint32_t idx, const b200pt_material &m) {
    CK(cudaSetDevice(device_));
    if (!has_scene_) throw CudaError{ B200PT_ERR_NO_SCENE, "no scene" };
    if (idx >= scene_.materials.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "material index out of range" };
    const uint32_t tix[5] = { m.BaseColorTextureIndex, m.NormalTextureIndex, m.RoughnessTextureIndex, m.MetallicTextureIndex, m.EmissiveTextureIndex };
    for (uint32_t t : tix) if (t >= scene_.textures.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "texture index out of range" };
    sync_all();
    scene_.materials[idx] = m;
    { const DevMaterial dm = make_dev_material(m); CK(cudaMemcpy(d_materials_ + idx, &dm, sizeof dm, cudaMemcpyHostToDevice)); launch_prepare_materials(d_materials_, idx, 1, stream_); CK(cudaStreamSynchronize(stream_)); }
    rebuild_emissive();
    rebuild_tri_class();
    reset();
}

// PathTracer::AddVolume / SetVolume / RemoveVolume / AddDensityDataToVolume / RemoveDensityDataFromVolume (PathTracer.cpp:1334-1555)
static void check_volume(const b200pt_volume &v) {
    for (int k = 0; k < 3; k++) if (!(v.CornerMin[k] <= v.CornerMax[k]) || !(v.Scale[k] >= 0.0f)) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume AABB: CornerMin > CornerMax or negative Scale" };
    if (!(v.Density >= 0.0f)) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume density must be >= 0" };
}
Engine::GridData::~GridData() { cudaFree(d_values); cudaFree(d_max_densities); }
void Engine::upload_volumes() {
    CK(cudaSetDevice(device_));
    sync_all();
    if (!d_volumes_) CK(cudaMalloc(&d_volumes_, B200PT_MAX_VOLUMES * sizeof(DevVolume)));
    if (!d_grids_) CK(cudaMalloc(&d_grids_, B200PT_MAX_VOLUMES * sizeof(DevGrid)));
    std::vector<DevVolume> dv(volumes_.size());
    std::vector<DevGrid> dg;
    for (size_t i = 0; i < volumes_.size(); i++) {
        const b200pt_volume &v = volumes_[i];
        float lo[3], hi[3];                                                     // VolumeGPU's constructor, PathTracer.h:396-397
        for (int k = 0; k < 3; k++) { lo[k] = v.Position[k] + (v.CornerMin[k] * v.Scale[k]); hi[k] = v.Position[k] + (v.CornerMax[k] * v.Scale[k]); }
        dv[i].mn_density = make_float4(lo[0], lo[1], lo[2], v.Density);
        dv[i].mx_g = make_float4(hi[0], hi[1], hi[2], v.Anisotropy);
        dv[i].color_alpha = make_float4(v.Color[0], v.Color[1], v.Color[2], v.Alpha);
        dv[i].emis_droplet = make_float4(v.EmissiveColor[0], v.EmissiveColor[1], v.EmissiveColor[2], v.DropletSize);
        dv[i].tcol_gamma = make_float4(v.TemperatureColor[0], v.TemperatureColor[1], v.TemperatureColor[2], v.TemperatureGamma);
        dv[i].tparams = make_float4(v.TemperatureScale, v.EmissiveColorGamma, v.ApproximatedScatteringFalloff, v.GridSharpness);
        dv[i].kelvin = make_float4((float)v.KelvinMin, (float)(v.KelvinMax - v.KelvinMin), v.MaxDensityInTheGrid, 0.0f);
        uint32_t slot = 0xFFFFFFFFu;
        if (grids_[i]) {
            const GridData &G = *grids_[i];
            DevGrid g{};
            g.values = G.d_values; g.max_densities = G.d_max_densities;
            for (int k = 0; k < 3; k++) { g.imin[k] = G.host.imin[k]; g.dim[k] = G.host.dim[k]; g.wmin[k] = G.host.wmin[k]; g.wext[k] = G.host.wext[k]; g.inv_vs[k] = G.host.inv_vs[k]; g.trans[k] = G.host.trans[k]; }
            slot = (uint32_t)dg.size(); dg.push_back(g);
        }
        dv[i].flags = make_uint4(v.ApproximatedScatteringForClouds, (uint32_t)(v.HasTemperatureData != 0), (uint32_t)(v.UseBlackbody != 0), slot);
    }
    if (!dv.empty()) CK(cudaMemcpy(d_volumes_, dv.data(), dv.size() * sizeof(DevVolume), cudaMemcpyHostToDevice));
    if (!dg.empty()) CK(cudaMemcpy(d_grids_, dg.data(), dg.size() * sizeof(DevGrid), cudaMemcpyHostToDevice));
    ds_.volumes = d_volumes_; ds_.n_volumes = (uint32_t)dv.size(); ds_.phase_function = phase_function_;
    ds_.grids = d_grids_; ds_.n_grids = (uint32_t)dg.size();
    reset();
}
// the three READ-ONLY members of b200pt_volume follow the density data attached to the index, whatever the caller's struct says
static b200pt_volume with_grid_fields(b200pt_volume v, const b200pt_volume *keep) {
    v.DensityDataIndex = keep ? keep->DensityDataIndex : -1;
    v.MaxDensityInTheGrid = keep ? keep->MaxDensityInTheGrid : 0.0f;
    v.HasTemperatureData = keep ? keep->HasTemperatureData : 0;
    return v;
}
uint32_t Engine::add_volume(const b200pt_volume &v) {
    check_volume(v);
    if (volumes_.size() >= B200PT_MAX_VOLUMES) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "too many volumes (B200PT_MAX_VOLUMES)" };
    volumes_.push_back(with_grid_fields(v, nullptr)); grids_.push_back(nullptr); upload_volumes();
    return (uint32_t)volumes_.size() - 1u;
}
void Engine::set_volume(uint32_t idx, const b200pt_volume &v) {
    check_volume(v);
    if (idx >= volumes_.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume index out of range" };
    const b200pt_volume old = volumes_[idx];
    volumes_[idx] = with_grid_fields(v, grids_[idx] ? &old : nullptr); upload_volumes();
}
void Engine::remove_volume(uint32_t idx) {
    if (idx >= volumes_.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume index out of range" };
    sync_all();                                                                 // the grid's device buffers may still be read by waves in flight
    volumes_.erase(volumes_.begin() + idx); grids_.erase(grids_.begin() + idx); upload_volumes();
}
void Engine::add_density_grid(uint32_t idx, const b200pt_density_grid &g) {
    if (idx >= volumes_.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume index out of range" };   // PathTracer.cpp:1348-1352
    auto G = std::make_shared<GridData>();
    prepare_density_grid(g, G->host);
    CK(cudaSetDevice(device_));
    sync_all();
    CK(cudaMalloc(&G->d_values, G->host.values.size() * sizeof(float)));
    CK(cudaMalloc(&G->d_max_densities, G->host.max_densities.size() * sizeof(float)));
    CK(cudaMemcpy(G->d_values, G->host.values.data(), G->host.values.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(G->d_max_densities, G->host.max_densities.data(), G->host.max_densities.size() * sizeof(float), cudaMemcpyHostToDevice));
    b200pt_volume &v = volumes_[idx];
    for (int k = 0; k < 3; k++) { v.CornerMin[k] = G->host.corner_min[k]; v.CornerMax[k] = G->host.corner_max[k]; }
    v.MaxDensityInTheGrid = G->host.max_density;
    v.HasTemperatureData = G->host.has_temperature ? 1 : 0;
    v.DensityDataIndex = density_data_counter_;                                 // :1512-1513
    density_data_counter_ = (density_data_counter_ + 1) % (int)B200PT_MAX_VOLUMES;
    grids_[idx] = std::move(G);
    upload_volumes();
}
void Engine::remove_density_data(uint32_t idx) {                                // PathTracer.cpp:1518-1528
    if (idx >= volumes_.size()) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "volume index out of range" };
    sync_all();
    grids_[idx] = nullptr;
    b200pt_volume &v = volumes_[idx];
    v.DensityDataIndex = -1; v.MaxDensityInTheGrid = 0.0f; v.HasTemperatureData = 0;
    for (int k = 0; k < 3; k++) { v.CornerMin[k] = -1.0f; v.CornerMax[k] = 1.0f; }
    upload_volumes();
}
void Engine::set_phase_function(uint32_t pf) {
    if (pf > 2u) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "phase function: 0 HG, 1 Draine, 2 HG + Draine" };
    phase_function_ = pf; ds_.phase_function = pf; reset();
}

// PathTracer::LoadEnvironmentMap (PathTracer.cpp:1137-1332)
void Engine::set_env_map(uint32_t w, uint32_t h, const float *rgba) {
    CK(cudaSetDevice(device_));
    if (!w || !h || !rgba || (uint64_t)w * h > 0x7FFFFFFFull) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "bad environment map" };
    sync_all();
    std::vector<float> px(rgba, rgba + (size_t)w * h * 4);
    std::vector<uint2> alias((size_t)w * h);
    build_env_alias(px.data(), w, h, alias.data());
    dfree(d_env_); dfree(d_alias_); dfree(d_env_row_cos_);
    {   // SH/Sampler.slang:329-331: stepTheta = M_PI / height; theta0 = py * stepTheta; cos(theta0), cos(theta0 + stepTheta)
        std::vector<float2> rc(h); const float stepTheta = 3.1415926535897F / (float)h;
        for (uint32_t y = 0; y < h; y++) { const float t0 = (float)y * stepTheta; rc[y] = make_float2(cosf(t0), cosf(t0 + stepTheta)); }
        CK(cudaMalloc(&d_env_row_cos_, rc.size() * sizeof(float2))); CK(cudaMemcpy(d_env_row_cos_, rc.data(), rc.size() * sizeof(float2), cudaMemcpyHostToDevice));
        ds_.env_row_cos = d_env_row_cos_;
    }
    CK(cudaMalloc(&d_env_, px.size() * 4)); CK(cudaMemcpy(d_env_, px.data(), px.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_alias_, alias.size() * sizeof(uint2))); CK(cudaMemcpy(d_alias_, alias.data(), alias.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    ds_.env = d_env_; ds_.alias = d_alias_; ds_.envW = w; ds_.envH = h;
    has_env_ = true;
    reset();
}

void Engine::set_luts(const float *refl, const float *rout, const float *rin) {
    CK(cudaSetDevice(device_));
    if (!refl || !rout || !rin) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "null LUT" };
    sync_all();
    const size_t n[3] = { 64 * 64 * 32, 128 * 128 * 32, 128 * 128 * 32 }; const float *src[3] = { refl, rout, rin };
    for (int i = 0; i < 3; i++) { dfree(d_luts_[i]); CK(cudaMalloc(&d_luts_[i], n[i] * 4)); CK(cudaMemcpy(d_luts_[i], src[i], n[i] * 4, cudaMemcpyHostToDevice)); }
    ds_.lut_reflect = d_luts_[0]; ds_.lut_refract_out = d_luts_[1]; ds_.lut_refract_in = d_luts_[2];
    has_luts_ = true;
    reset();
}

void Engine::set_config(const b200pt_config &c) {
    if (c.SamplesPerFrame == 0 || c.ScreenChunkCount == 0 || c.ScreenChunkCount > 64) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "SamplesPerFrame and ScreenChunkCount must be >= 1" };
    if (c.ScreenChunkCount > 1 && world_ > 1) throw CudaError{ B200PT_ERR_NOT_IMPLEMENTED, "ScreenChunkCount > 1 cannot be combined with a multi-GPU partition" };
    cfg_ = c;
    reset();
}
void Engine::set_camera(const float vi[16], const float pi[16]) { memcpy(view_inv_, vi, sizeof view_inv_); memcpy(proj_inv_, pi, sizeof proj_inv_); reset(); }
void Engine::get_camera(float vi[16], float pi[16]) const { memcpy(vi, view_inv_, sizeof view_inv_); memcpy(pi, proj_inv_, sizeof proj_inv_); }
void Engine::reset() { dispatch_count_ = 0; frame_count_ = 0; samples_accumulated_ = 0; }   // PathTracer.h:183

void Engine::resize(uint32_t w, uint32_t h) {                                                  // PathTracer::ResizeImage
    if (!w || !h || w > 65535 || h > 65535) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "bad image size" };
    CK(cudaSetDevice(device_));
    sync_all();
    W_ = w; H_ = h; local_rows_ = partition_local_rows(H_, rank_, world_, band_);
    ensure_image(); reset();
}
void Engine::set_partition(uint32_t rank, uint32_t world, uint32_t band) {
    if (!world || rank >= world || !band) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "bad partition" };
    if (world > 1 && cfg_.ScreenChunkCount > 1) throw CudaError{ B200PT_ERR_NOT_IMPLEMENTED, "ScreenChunkCount > 1 cannot be combined with a multi-GPU partition" };
    CK(cudaSetDevice(device_));
    sync_all();
    rank_ = rank; world_ = world; band_ = band;
    if (W_ && H_) { local_rows_ = partition_local_rows(H_, rank_, world_, band_); ensure_image(); }
    reset();
}
void Engine::ensure_image() {
    const size_t need = (size_t)W_ * std::max<uint32_t>(local_rows_, 1);
    if (need != image_pixels_) { dfree(d_image_); CK(cudaMalloc(&d_image_, need * sizeof(float4))); image_pixels_ = need; }
    CK(cudaMemsetAsync(d_image_, 0, image_pixels_ * sizeof(float4), stream_));
}

void Engine::free_wave() {
    for (auto &B : wb_) {
        for (auto &p : B.ps) { dfree(p.org_pdf); dfree(p.dir_rng); dfree(p.thr_depth); dfree(p.rad_slot); dfree(p.medium); dfree(p.medium_g); dfree(p.vol_depth); }
        dfree(B.so.hit); dfree(B.so.bxdf_pdf); dfree(B.so.e0); dfree(B.so.sky_o); dfree(B.so.sky_d); dfree(B.so.sky_c); dfree(B.so.lit_o); dfree(B.so.lit_d); dfree(B.so.lit_c);
        dfree(B.sample_buf); dfree(B.rng_carry); dfree(B.q_hit[0]); dfree(B.q_hit[1]); dfree(B.q_miss[0]); dfree(B.q_miss[1]);
        B.cap = 0;
    }
    dfree(d_sort_key_rank_); dfree(d_sort_hist_); dfree(d_sort_offs_); dfree(d_order_);
    wave_cap_ = 0;
}
void Engine::ensure_wave(size_t cap, int contexts) {
    bool ok = cap <= wave_cap_ && (!sort_rays_ || d_order_ != nullptr);
    for (int c = 0; c < contexts; c++) ok = ok && wb_[c].cap >= cap;
    if (ok) return;
    sync_all();
    cap = std::max(cap, wave_cap_);
    free_wave();
    auto a4 = [&](float4 *&p) { CK(cudaMalloc(&p, cap * sizeof(float4))); };
    for (int c = 0; c < contexts; c++) {
        WaveBuf &B = wb_[c];
        for (auto &p : B.ps) { a4(p.org_pdf); a4(p.dir_rng); a4(p.thr_depth); a4(p.rad_slot); a4(p.medium); CK(cudaMalloc(&p.medium_g, cap * sizeof(float))); CK(cudaMalloc(&p.vol_depth, cap * sizeof(uint32_t))); }
        a4(B.so.hit); a4(B.so.bxdf_pdf); a4(B.so.e0); a4(B.so.sky_o); a4(B.so.sky_d); a4(B.so.sky_c); a4(B.so.lit_o); a4(B.so.lit_d); a4(B.so.lit_c);
        a4(B.sample_buf);
        CK(cudaMalloc(&B.rng_carry, cap * sizeof(uint32_t)));
        for (int i = 0; i < 2; i++) { CK(cudaMalloc(&B.q_hit[i], (size_t)MC_COUNT * cap * sizeof(uint32_t))); CK(cudaMalloc(&B.q_miss[i], cap * sizeof(uint32_t))); }   // one hit queue per material class, ping-pong for the fused bounce kernel
        B.cap = cap;
    }
    if (sort_rays_) {                                                        // opt-in ray sort (B200PT_SORT=1, single context): 12 B per path + two bin tables
        CK(cudaMalloc(&d_sort_key_rank_, cap * sizeof(uint2))); CK(cudaMalloc(&d_order_, cap * sizeof(uint32_t)));
        CK(cudaMalloc(&d_sort_hist_, SORT_BINS * sizeof(uint32_t))); CK(cudaMalloc(&d_sort_offs_, SORT_BINS * sizeof(uint32_t)));
        CK(cudaMemsetAsync(d_sort_hist_, 0, SORT_BINS * sizeof(uint32_t), stream_));
    }
    wave_cap_ = cap;
}

DevConfig Engine::make_dev_config() const {
    DevConfig d; memset(&d, 0, sizeof d);
    memcpy(d.VI, view_inv_, sizeof d.VI); memcpy(d.PI, proj_inv_, sizeof d.PI);
    d.SampleCount = cfg_.SamplesPerFrame; d.MaxDepth = cfg_.MaxDepth; d.MaxLuminance = cfg_.MaxLuminance; d.FocusDistance = cfg_.FocusDistance;
    d.DepthOfFieldStrength = cfg_.DepthOfFieldStrength; d.SkyRotationAzimuth = cfg_.SkyRotationAzimuth; d.SkyRotationAltitude = cfg_.SkyRotationAltitude;
    d.EnvironmentIntensity = cfg_.SkyIntensity; d.EmissiveMeshSamplingPDFBias = cfg_.EmissiveMeshSamplingPDFBias; d.ScreenSplitCount = cfg_.ScreenChunkCount;
    d.EnableSkyMIS = cfg_.EnableSkyMIS; d.EnableMeshMIS = cfg_.EnableMeshMIS; d.ShowEnvMapDirectly = cfg_.ShowEnvMapDirectly;
    d.UseOnlyGeometryNormals = cfg_.UseOnlyGeometryNormals; d.UseEnergyCompensation = cfg_.UseEnergyCompensation; d.FurnaceTestMode = cfg_.FurnaceTestMode;
    {   // atmosphere block (PathTracer.h:221-232 -> SH/Bindings.slang:26-37)
        const b200pt_atmosphere &a = atmosphere_;
        d.EnableAtmosphere = a.Enable ? 1u : 0u; d.PlanetRadius = a.PlanetRadius; d.AtmosphereHeight = a.AtmosphereHeight;
        for (int k = 0; k < 3; k++) { d.PlanetPosition[k] = a.PlanetPosition[k]; d.RayleighMult[k] = a.RayleighScatteringCoefficientMultiplier[k]; d.MieMult[k] = a.MieScatteringCoefficientMultiplier[k];
                                      d.OzoneMult[k] = a.OzoneAbsorptionCoefficientMultiplier[k]; d.SunColor[k] = a.SunColor[k]; }
        d.RayleighDensityFalloff = a.RayleighDensityFalloff; d.MieDensityFalloff = a.MieDensityFalloff; d.OzoneDensityFalloff = a.OzoneDensityFalloff; d.OzonePeak = a.OzonePeak;
        d.cosSunTheta = cosf(0.004675f);
    }
    { const float az = cfg_.SkyRotationAzimuth / 180.0f * 3.1415926535897F, al = cfg_.SkyRotationAltitude / 180.0f * 3.1415926535897F;   // SH/Sampler.slang:338-339
      d.cosAz = cosf(az); d.sinAz = sinf(az); d.cosAl = cosf(al); d.sinAl = sinf(al); }
    d.W = W_; d.H = H_; d.rank = rank_; d.world = world_; d.band_rows = band_; d.local_rows = local_rows_;
    return d;
}

// PathTracer::PathTrace x dispatches (PathTracer.cpp:122-156)
bool Engine::path_trace(uint32_t dispatches, uint32_t base_seed) {
    CK(cudaSetDevice(device_));
    if (!has_scene_) throw CudaError{ B200PT_ERR_NO_SCENE, "PathTrace before SetScene" };
    if (!has_env_) throw CudaError{ B200PT_ERR_NO_SCENE, "no environment map (PathTracer always loads one, PathTracer.cpp:202)" };
    if (!has_luts_) throw CudaError{ B200PT_ERR_NO_SCENE, "no energy-compensation lookup tables (PathTracer.cpp:199-201)" };
    const uint32_t S = cfg_.ScreenChunkCount, S2 = S * S;
    // how many dispatches until MaxSamplesAccumulated (PathTracer.cpp:124-125,151-153)
    uint32_t todo = 0;
    { uint64_t d = dispatch_count_; uint32_t acc = samples_accumulated_;
      while (todo < dispatches && acc < cfg_.MaxSamplesAccumulated) { d++; todo++; acc = (uint32_t)(d / S2) * cfg_.SamplesPerFrame; } }
    if (todo == 0) return samples_accumulated_ >= cfg_.MaxSamplesAccumulated;
    if (todo > kMaxDispatchTable) todo = kMaxDispatchTable;

    const uint32_t P = (S == 1) ? W_ * local_rows_ : ((W_ + S - 1) / S) * ((H_ + S - 1) / S);
    if (P == 0) { dispatch_count_ += todo; frame_count_ = (uint32_t)(dispatch_count_ / S2); samples_accumulated_ = frame_count_ * cfg_.SamplesPerFrame; return samples_accumulated_ >= cfg_.MaxSamplesAccumulated; }
    // Two waves in flight (two contexts, two internal streams): the late bounces of a wave are bound by the latency chain of their longest rays (a ~150-200 us floor
    // per traversal launch on BreakfastRoom, profiles/r02_variants.txt), during which most SMs idle -- the next wave's full-machine bounces fill them.  Waves are
    // numbered across calls (wave_seq_): wave w runs in context w & 1 on internal stream w & 1 and its k_resolve waits for wave w - 1's, so the running mean folds
    // the frames in dispatch order whatever overlaps.  The join with the caller's stream is LAZY: path_trace returns with the waves in flight, and whatever touches
    // the image afterwards (get_hdr, post_process, set_hdr, checkpoints, ...) first makes the caller's stream wait for the resolves (join_waves); the next call's
    // resolves in turn wait for what the caller enqueued on the image in between (image_free_ev_).  So consecutive path_trace calls -- the steps of a multi-GPU
    // run, each followed by an NCCL gather of the bands -- overlap as well: the gather of step k runs beside the first bounces of step k + 1.
    // B200PT_OVERLAP=0 runs one wave at a time on the caller's stream (also used while profiling per-kernel times and with the opt-in ray sort).
    bool overlap = !profiling_ && !sort_rays_;
    if (const char *e = getenv("B200PT_OVERLAP")) { if (atoi(e) == 0) overlap = false; }
    uint32_t F = cfg_.FramesInFlight;
    if (F == 0) {
        // Auto wave size: up to 64 M paths per context (~24 GB of wavefront state at ~380 B per path; all contexts capped to a quarter of the free HBM): large enough
        // to amortise the floors above.  Measured 16 M -> 32 M -> 64 M paths: BreakfastRoom 1385 -> 1593 -> 1727 Mpaths/s, glass 929 -> 991 -> 1002, Cornell 3415 -> 3535 -> 3590.
        uint64_t target = 64ull << 20;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            uint64_t have = 0; for (const auto &B : wb_) have += (uint64_t)B.cap * 400ull;       // what this engine already holds counts as available
            target = std::min<uint64_t>(target, std::max<uint64_t>(1ull << 20, ((uint64_t)free_b + have) / 4ull / 400ull / (overlap ? 2ull : 1ull)));
        }
        F = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(256, target / P));
    }
    F = std::min(F, todo);
    { const uint32_t n_waves = (todo + F - 1) / F; F = (todo + n_waves - 1) / n_waves; }          // equal waves: 128 frames at F = 123 become 64 + 64, not 123 + 5
    const int n_ctx = overlap ? 2 : 1;
    ensure_wave((size_t)F * P, n_ctx);
    if (!overlap) join_waves();                               // single-stream mode runs on the caller's stream, behind whatever the internal streams still hold
    else CK(cudaEventRecord(image_free_ev_, stream_));        // what the caller has enqueued on the image so far (gather copy, post chain, set_hdr) precedes this call's resolves
    const DevConfig dc = make_dev_config();
    auto pcg = [](uint32_t in) { uint32_t st = in * 747796405u + 2891336453u; uint32_t w = ((st >> ((st >> 28u) + 4u)) ^ st) * 277803737u; return (w >> 22u) ^ w; };

    const uint32_t n_vol = ds_.n_volumes;
    const bool pre = n_vol != 0u || atmosphere_.Enable != 0u;   // k_volume_decide + k_shade_volume run: volumes and / or the atmosphere (csrc/atmosphere.cuh)
    ds_.pre_pass = pre ? 1u : 0u;
    bool medium = false;                                      // can a path random-walk inside a mesh without gaining Depth?
    // Conservative: the device-side metallic value is Metallic * texel, so the constant factor alone cannot rule refraction out, and a negative
    // density (not validated by the reference either) scatters on every segment.  A false positive only costs a 4-byte read-back per 16 bounces.
    for (const auto &m : scene_.materials) if (m.Transmission > 0.0f && m.MediumDensity != 0.0f && m.MediumAnisotropy != 1.0f) medium = true;

    // Fused bounce kernels (LaunchCfg::fuse) need the BVH in shared memory and no volumes; they ping-pong the queues and the hit records
    // (the kernel that consumes bounce k's queue fills bounce k+1's), the unfused pipeline uses buffer 0 only.
    const int fuse = pre ? 0 : lc_.fuse;
    const uint32_t cmask = pre ? (class_mask_ | (1u << MC_GENERAL)) : class_mask_;   // volume / atmosphere events ride in the general queue
    const uint32_t qcap = (uint32_t)wave_cap_;
    const int qsel[2] = { 0, fuse == 2 ? 1 : 0 };

    uint64_t launches = 0; uint32_t waves = 0, bounces_total = 0;
    std::vector<int> prof_kind; size_t prof_n = 0;            // kind: 0 raygen 1 extend 2 shade 3 connect 4 resolve
    auto mark = [&](int kind) {
        if (!profiling_) return;
        if (prof_n >= prof_ev_.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); prof_ev_.push_back(e); }
        CK(cudaEventRecord(prof_ev_[prof_n++], stream_)); prof_kind.push_back(kind);
    };
    mark(-1);
    cudaStream_t last_st = stream_;
    for (uint32_t w0 = 0; w0 < todo; w0 += F) {
        const uint32_t nd = std::min(F, todo - w0);
        const int ctx = overlap ? (int)(wave_seq_ & 1ull) : 0;
        WaveBuf &B = wb_[ctx];
        cudaStream_t st = overlap ? aux_stream_[ctx] : stream_;
        if (w0 == 0) CK(cudaEventRecord(ev_[0], st));
        // this wave's slice of the dispatch table (PathTracer.cpp:127-150: FrameCount, Seed = PCG(seed + dispatch), ChunkIndex) travels on the wave's own stream
        CK(cudaEventSynchronize(B.disp_ev));                  // the pinned copy is reusable once ITS last upload has executed
        for (uint32_t i = 0; i < nd; i++) {
            const uint64_t d = dispatch_count_ + w0 + i;
            B.h_disp[i].FrameCount = (uint32_t)(d / S2); B.h_disp[i].Seed = pcg(base_seed + (uint32_t)d); B.h_disp[i].ChunkIndex = (uint32_t)(d % S2); B.h_disp[i]._pad = 0;
        }
        CK(cudaMemcpyAsync(B.d_disp, B.h_disp, nd * sizeof(DevDispatch), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(B.disp_ev, st));
        PathState pst[2] = { B.ps[0], B.ps[1] };              // payload.VolumeDepth travels only while the scene has volumes
        if (!n_vol) { pst[0].vol_depth = nullptr; pst[1].vol_depth = nullptr; }
        float4 *const hitb[2] = { B.so.hit, fuse == 2 ? B.so.bxdf_pdf : B.so.hit };
        for (uint32_t s = 0; s < cfg_.SamplesPerFrame; s++) {
            launch_raygen(lc_, dc, B.d_disp, nd, P, s == 0 ? 1u : 0u, B.rng_carry, pst[0], B.sample_buf, B.counts, d_ctr_, st);
            launches++; mark(0);
            int cur = 0; uint32_t k = 0;
            for (;;) {
                uint32_t chunk = 16;
                if (!medium && cfg_.MaxDepth - std::min(cfg_.MaxDepth, k) < chunk) chunk = cfg_.MaxDepth - std::min(cfg_.MaxDepth, k);
                if (chunk == 0) break;
                for (uint32_t b = 0; b < chunk; b++, k++) {
                    const uint32_t par = k & 1u;
                    const Queues qc{ B.q_miss[qsel[par]], B.q_hit[qsel[par]], qcap }, qn{ B.q_miss[qsel[par ^ 1u]], B.q_hit[qsel[par ^ 1u]], qcap };
                    if (pre) { launch_volume_decide(lc_, ds_, dc, pst[cur], B.so, B.counts, par, B.sample_buf, B.rng_carry, st); launches++; }
                    const bool sorted = sort_rays_ && d_order_ != nullptr && k >= 1;   // camera rays are pixel-coherent already
                    if (sorted) { launch_ray_sort(lc_, ds_, pst[cur], B.counts, par, d_sort_key_rank_, d_sort_hist_, d_sort_offs_, d_order_, st); launches += 3; }
                    if (fuse != 2 || k == 0) { launch_extend(lc_, ds_, pst[cur], hitb[par], B.counts, par, qc, d_ctr_, k == 0, sorted ? d_order_ : nullptr, st); launches++; } mark(1);
                    launches += launch_shade(lc_, ds_, dc, pst[cur], pst[cur ^ 1], B.so, hitb[par], hitb[par ^ 1u], B.counts, par, qc, qn, B.sample_buf, B.rng_carry, d_ctr_,
                                             fuse, cmask, st); mark(2);
                    if (!fuse) { launch_connect(lc_, ds_, dc, pst[cur], pst[cur ^ 1], B.so, B.counts, par, qc, B.sample_buf, B.rng_carry, d_ctr_, st); launches += lc_.trav_dyn ? 2 : 1; } mark(3);
                    cur ^= 1;
                }
                if (!medium && k >= cfg_.MaxDepth) break;    // every surviving path has Depth >= MaxDepth: provably empty
                CK(cudaMemcpyAsync(B.h_count, B.counts + (k & 1u), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (B.h_count[0] == 0) break;
                if (k > (1u << 20)) break;                    // runaway guard
            }
            bounces_total += k;
        }
        // the running mean folds the waves in dispatch order: this wave's resolve waits for the previous wave's (the other stream, possibly an earlier call)
        // and for whatever the caller enqueued on the image before this call
        if (overlap) {
            if (wave_seq_ > 0) CK(cudaStreamWaitEvent(st, resolved_ev_[ctx ^ 1], 0));
            CK(cudaStreamWaitEvent(st, image_free_ev_, 0));
        }
        launch_resolve(lc_, dc, B.d_disp, nd, P, B.sample_buf, d_image_, st);
        CK(cudaEventRecord(resolved_ev_[ctx], st));
        wave_seq_++;
        launches++; waves++; mark(4);
        last_st = st;
    }
    CK(cudaEventRecord(ev_[1], last_st));                     // ms_total: first kernel of the first wave .. resolve of the last (counters() synchronises before reading it)
    CK(cudaGetLastError());
    dispatch_count_ += todo;
    frame_count_ = (uint32_t)(dispatch_count_ / S2);                     // PathTracer.cpp:152
    samples_accumulated_ = frame_count_ * cfg_.SamplesPerFrame;          // :153
    last_.kernel_launches += launches; last_.waves = waves; last_.bounces = bounces_total;
    last_.ms_raygen = last_.ms_extend = last_.ms_shade = last_.ms_connect = last_.ms_resolve = 0.0f;
    if (profiling_ && prof_n > 1) {
        CK(cudaStreamSynchronize(stream_));
        float *acc[5] = { &last_.ms_raygen, &last_.ms_extend, &last_.ms_shade, &last_.ms_connect, &last_.ms_resolve };
        for (size_t i = 1; i < prof_n; i++) { float ms = 0.0f; CK(cudaEventElapsedTime(&ms, prof_ev_[i - 1], prof_ev_[i])); if (prof_kind[i] >= 0) *acc[prof_kind[i]] += ms; }
    }
    return samples_accumulated_ >= cfg_.MaxSamplesAccumulated;
}

void Engine::set_stream(cudaStream_t s) {
    CK(cudaSetDevice(device_));
    sync_all();
    stream_ = s ? s : own_stream_;
}

void Engine::synchronize() { CK(cudaSetDevice(device_)); sync_all(); CK(cudaGetLastError()); }

b200pt_counters Engine::counters() {
    CK(cudaSetDevice(device_));
    sync_all();
    WaveCounters wc; CK(cudaMemcpy(&wc, d_ctr_, sizeof wc, cudaMemcpyDeviceToHost));
    b200pt_counters c = last_;
    c.paths = wc.paths; c.extend_rays = wc.extend_rays; c.shade_invocations = wc.shade_invocations; c.surface_hits = wc.surface_hits;
    c.misses = wc.misses; c.shadow_rays = wc.shadow_rays; c.medium_events = wc.medium_events;
    float ms = 0.0f; if (cudaEventElapsedTime(&ms, ev_[0], ev_[1]) == cudaSuccess) c.ms_total = ms;
    return c;
}

void Engine::get_hdr(float *dst, bool dev) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !dst) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "no image" };
    join_waves();
    CK(cudaMemcpyAsync(dst, d_image_, (size_t)W_ * local_rows_ * sizeof(float4), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream_));
    if (!dev) CK(cudaStreamSynchronize(stream_));        // a device destination stays asynchronous on the engine's stream: the NCCL gather that follows is stream-ordered,
                                                          // and the host can already enqueue the next wave (no per-step bubble on multi-GPU runs)
}
void Engine::set_hdr(const float *src, bool dev) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !src) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "no image" };
    join_waves();
    CK(cudaMemcpyAsync(d_image_, src, (size_t)W_ * local_rows_ * sizeof(float4), dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream_));
    CK(cudaStreamSynchronize(stream_));
}

// ---------------------------------------------------------------- checkpoint / resume
// The whole progressive state of the reference is the RGBA32F image + three counters (PT/PathTracer.h:199-201); frame f always uses
// Seed_f = PCG(base_seed + dispatch index), so a render resumed from a checkpoint continues the SAME sample sequence: N frames in one go
// and k frames + checkpoint + (N - k) frames give bit-identical images (tested).  File: 64-byte header + local rows of float4.
namespace {
struct CkptHeader { char magic[8]; uint32_t version, W, H, rank, world, band, local_rows, frame_count, samples_accumulated, samples_per_frame, chunk_count, _pad; uint64_t dispatch_count; };
static_assert(sizeof(CkptHeader) == 64, "checkpoint header is 64 B");
}
void Engine::save_checkpoint(const char *path) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !path) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "no image to checkpoint" };
    std::vector<float> img((size_t)W_ * local_rows_ * 4);
    get_hdr(img.data(), false);
    CkptHeader h{}; memcpy(h.magic, "B2PTCKPT", 8); h.version = 1; h.W = W_; h.H = H_; h.rank = rank_; h.world = world_; h.band = band_; h.local_rows = local_rows_;
    h.frame_count = frame_count_; h.samples_accumulated = samples_accumulated_; h.samples_per_frame = cfg_.SamplesPerFrame; h.chunk_count = cfg_.ScreenChunkCount;
    h.dispatch_count = dispatch_count_;
    FILE *f = fopen(path, "wb");
    if (!f) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, std::string("cannot open checkpoint for writing: ") + path };
    const bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(img.data(), sizeof(float), img.size(), f) == img.size();
    if (fclose(f) != 0 || !ok) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "short write of the checkpoint" };
}
void Engine::load_checkpoint(const char *path) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !path) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "resize the image before loading a checkpoint" };
    FILE *f = fopen(path, "rb");
    if (!f) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, std::string("cannot open checkpoint: ") + path };
    CkptHeader h{};
    std::vector<float> img;
    bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "B2PTCKPT", 8) == 0 && h.version == 1;
    const bool fits = ok && h.W == W_ && h.H == H_ && h.rank == rank_ && h.world == world_ && h.band == band_ && h.local_rows == local_rows_ &&
                      h.samples_per_frame == cfg_.SamplesPerFrame && h.chunk_count == cfg_.ScreenChunkCount;
    if (fits) { img.resize((size_t)W_ * local_rows_ * 4); ok = fread(img.data(), sizeof(float), img.size(), f) == img.size(); }
    fclose(f);
    if (!ok) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "not a b200pt checkpoint (bad magic / version / truncated)" };
    if (!fits) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "checkpoint was taken with another image size, partition, SamplesPerFrame or ScreenChunkCount" };
    set_hdr(img.data(), false);
    dispatch_count_ = h.dispatch_count; frame_count_ = h.frame_count; samples_accumulated_ = h.samples_accumulated;
}

// ---------------------------------------------------------------- PostProcessor (PostProcessor.cpp:128-246)
void Engine::free_post() { for (auto &m : d_mips_) if (m) cudaFree(m); d_mips_.clear(); mip_wh_.clear(); dfree(d_ldr_); post_w_ = post_h_ = 0; }
void Engine::ensure_post() {
    if (post_w_ == W_ && post_h_ == H_ && !d_mips_.empty()) return;
    free_post();
    uint32_t wh[20]; const uint32_t levels = bloom_mip_sizes(W_, H_, wh);                     // SetInputImage :128-158
    for (uint32_t i = 0; i < levels; i++) { float4 *p = nullptr; CK(cudaMalloc(&p, (size_t)wh[2 * i] * wh[2 * i + 1] * sizeof(float4))); d_mips_.push_back(p); mip_wh_.push_back(wh[2 * i]); mip_wh_.push_back(wh[2 * i + 1]); }
    CK(cudaMalloc(&d_ldr_, (size_t)W_ * H_ * sizeof(uchar4)));
    post_w_ = W_; post_h_ = H_;
}
void Engine::accumulate_rows(const float4 *d_frame, uint32_t frame_index, uint32_t y0, uint32_t y1) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !d_frame || y0 >= y1 || y1 > H_ || world_ != 1) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "accumulate_rows: no image / bad rows / partitioned image" };
    join_waves();
    launch_accumulate(d_frame, d_image_, y0 * W_, (y1 - y0) * W_, frame_index, lc_.grid_light > 0 ? lc_.grid_light : 1184, stream_);
    CK(cudaGetLastError());
}
// HDR rows read by post_process_rows(y0, y1): the final kernel's rows y0-1 .. y1-1 and the first down pass's taps 2y-2 .. 2y+1 over D[1]
void Engine::post_input_rows(uint32_t y0, uint32_t y1, uint32_t *in0, uint32_t *in1) {
    if (y0 >= y1 || y1 > H_) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_input_rows: bad row range" };
    ensure_post();
    const uint32_t levels = (uint32_t)d_mips_.size(), mips = std::min(std::max(bloom_.MipCount, 1u), levels);
    int lo = std::max((int)y0 - 1, 0), hi = (int)y1 - 1;
    if (mips >= 2) {
        int U[24][2], D[24][2];
        auto clampr = [&](int v, uint32_t i) { return std::min(std::max(v, 0), (int)mip_wh_[2 * i + 1] - 1); };
        U[1][0] = clampr(lo / 2 - 1, 1); U[1][1] = clampr(hi / 2 + 2, 1);
        for (uint32_t i = 2; i < mips; i++) { U[i][0] = clampr(U[i - 1][0] / 2 - 1, i); U[i][1] = clampr(U[i - 1][1] / 2 + 2, i); }
        D[mips - 1][0] = U[mips - 1][0]; D[mips - 1][1] = U[mips - 1][1];
        for (uint32_t i = mips - 1; i-- > 1;) { D[i][0] = std::min(U[i][0], clampr(2 * D[i + 1][0] - 2, i)); D[i][1] = std::max(U[i][1], clampr(2 * D[i + 1][1] + 1, i)); }
        lo = std::min(lo, std::max(2 * D[1][0] - 2, 0)); hi = std::max(hi, std::min(2 * D[1][1] + 1, (int)H_ - 1));
    }
    *in0 = (uint32_t)lo; *in1 = (uint32_t)hi + 1u;
}
void Engine::post_process() { post_process_rows(0, H_); }

// Row-block post pass (BASELINE config 5 on 2 / 4 / 8 GPUs): rank r holds the full HDR image and produces output rows [y0, y1) only.  The bloom chain is
// evaluated on exactly the rows those outputs depend on -- per mip i the rows U[i] whose post-up value is read and the rows D[i] whose down
// result is read (U[i] plus the taps of the next down pass) -- so the block equals the same rows of the full-image pass bit for bit (tested)
// without any exchange: the halo is recomputed, ~4 rows per level next to a block of H / (N 2^i) rows, the whole mip once it is a few rows tall.
void Engine::post_process_rows(uint32_t y0, uint32_t y1) {
    CK(cudaSetDevice(device_));
    if (!d_image_ || !W_ || !H_) throw CudaError{ B200PT_ERR_NO_SCENE, "PostProcess without an input image" };
    if (world_ != 1) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_process needs the full image: gather the bands first (world must be 1)" };
    if (y0 >= y1 || y1 > H_) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_process_rows: bad row range" };
    ensure_post();
    join_waves();
    const bool whole = (y0 == 0 && y1 == H_);
    const uint32_t levels = (uint32_t)d_mips_.size();
    const uint32_t mips = std::min(std::max(bloom_.MipCount, 1u), levels);                      // :195
    const PostParams p{ tonemap_.Exposure, tonemap_.Gamma, bloom_.BloomThreshold, bloom_.BloomStrength, bloom_.FalloffRange };
    // Fused chain (default): mip 0 is never materialised -- the threshold is evaluated inside the first down pass and inside the final
    // [up + tonemap] kernel.  B200PT_POST_FUSED=0 (or a 1-level chain) runs the reference's pass-per-pass structure; both produce the
    // same RGBA8 image bit for bit (tested).
    bool fused = mips >= 2;
    if (const char *e = getenv("B200PT_POST_FUSED")) { if (atoi(e) == 0) fused = false; }
    if (!fused && !whole) throw CudaError{ B200PT_ERR_NOT_IMPLEMENTED, "post_process_rows needs the fused chain (MipCount >= 2, B200PT_POST_FUSED unset)" };
    if (fused) {
        // needed row ranges, inclusive: U[i] post-up rows of mip i, D[i] down-result rows of mip i (i = 1 .. mips-1)
        int U[24][2], D[24][2];
        auto hgt = [&](uint32_t i) { return (int)mip_wh_[2 * i + 1]; };
        auto clampr = [&](int v, uint32_t i) { return std::min(std::max(v, 0), hgt(i) - 1); };
        {   // mip 0 rows the final kernel builds for outputs [y0, y1): y-1 .. y (bilinear bloom fetch); their up-sample term reads mip 1 rows yy/2 - 1 .. yy/2 + 2
            const int a = std::max((int)y0 - 1, 0), b = (int)y1 - 1;
            U[1][0] = clampr(a / 2 - 1, 1); U[1][1] = clampr(b / 2 + 2, 1);
        }
        for (uint32_t i = 2; i < mips; i++) { U[i][0] = clampr(U[i - 1][0] / 2 - 1, i); U[i][1] = clampr(U[i - 1][1] / 2 + 2, i); }   // up pass i: dst rows U[i-1] read src rows y/2 - 1 .. y/2 + 2
        D[mips - 1][0] = U[mips - 1][0]; D[mips - 1][1] = U[mips - 1][1];
        for (uint32_t i = mips - 1; i-- > 1;) {                                                 // down pass i+1: dst rows D[i+1] read src rows 2y - 2 .. 2y + 1
            D[i][0] = std::min(U[i][0], clampr(2 * D[i + 1][0] - 2, i)); D[i][1] = std::max(U[i][1], clampr(2 * D[i + 1][1] + 1, i));
        }
        auto rows_of = [&](const int r[2], int out[2]) { out[0] = r[0]; out[1] = r[1] + 1; };
        // the small end of the chain inside one cluster launch (k_bloom_small): measured slower at 3840x2160 (0.323 vs 0.302 ms, profiles/r02_variants.txt) -> OPT-IN, whole image only
        uint32_t ls = mips;
        if (whole) if (const char *e = getenv("B200PT_POST_SMALL")) { if (atoi(e) == 1) for (uint32_t i = 2; i < mips; i++) if ((uint64_t)mip_wh_[2 * (i - 1)] * mip_wh_[2 * (i - 1) + 1] <= 160u * 1024u) { ls = i; break; } }
        if (mips > 16) ls = mips;
        int r[2];
        rows_of(D[1], r);
        launch_bloom_down_first(d_image_, W_, H_, d_mips_[1], mip_wh_[2], mip_wh_[3], p, stream_, whole ? nullptr : r);                               // :200-226, i == 0 and i == 1
        for (uint32_t i = 2; i < ls; i++) { rows_of(D[i], r); launch_bloom_down(d_mips_[i - 1], mip_wh_[2 * (i - 1)], mip_wh_[2 * (i - 1) + 1], d_mips_[i], mip_wh_[2 * i], mip_wh_[2 * i + 1], p, stream_, whole ? nullptr : r); }
        if (ls < mips) {
            SmallMips sm{}; sm.first = (int)ls; sm.last = (int)mips - 1;
            for (uint32_t i = ls - 1; i < mips; i++) { sm.mip[i] = d_mips_[i]; sm.w[i] = (int)mip_wh_[2 * i]; sm.h[i] = (int)mip_wh_[2 * i + 1]; }
            launch_bloom_small(sm, p, stream_);
        }
        for (uint32_t i = ls - 1; i > 1; i--) { rows_of(U[i - 1], r); launch_bloom_up(d_mips_[i], mip_wh_[2 * i], mip_wh_[2 * i + 1], d_mips_[i - 1], mip_wh_[2 * (i - 1)], mip_wh_[2 * (i - 1) + 1], p, stream_, whole ? nullptr : r); }   // :229-235
        r[0] = (int)y0; r[1] = (int)y1;
        launch_bloom_final(d_image_, d_mips_[1], mip_wh_[2], mip_wh_[3], d_ldr_, (keep_bloom_ && whole) ? d_mips_[0] : nullptr, W_, H_, p, stream_, whole ? nullptr : r);   // last up pass + :238-245
        bloom_valid_ = keep_bloom_ && whole;
    } else {
        launch_bloom_threshold(d_image_, d_mips_[0], W_ * H_, p, lc_.grid_light > 0 ? lc_.grid_light : 1184, stream_);            // :200-226, i == 0
        for (uint32_t i = 1; i < mips; i++) launch_bloom_down(d_mips_[i - 1], mip_wh_[2 * (i - 1)], mip_wh_[2 * (i - 1) + 1], d_mips_[i], mip_wh_[2 * i], mip_wh_[2 * i + 1], p, stream_);
        for (uint32_t i = mips - 1; i > 0; i--) launch_bloom_up(d_mips_[i], mip_wh_[2 * i], mip_wh_[2 * i + 1], d_mips_[i - 1], mip_wh_[2 * (i - 1)], mip_wh_[2 * (i - 1) + 1], p, stream_);   // :229-235
        launch_tonemap(d_image_, d_mips_[0], d_ldr_, W_, H_, p, stream_);                           // :238-245
        bloom_valid_ = true;
    }
    CK(cudaGetLastError());
}
void Engine::get_ldr_rows(uint32_t y0, uint32_t y1, uint8_t *dst, bool dev) {
    CK(cudaSetDevice(device_));
    if (!d_ldr_ || !dst || y0 >= y1 || y1 > H_) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_process has not run / bad row range" };
    CK(cudaMemcpyAsync(dst, reinterpret_cast<const uint8_t *>(d_ldr_) + (size_t)y0 * W_ * 4, (size_t)(y1 - y0) * W_ * 4, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream_));
    if (!dev) CK(cudaStreamSynchronize(stream_));
}
void Engine::get_ldr(uint8_t *dst, bool dev) {
    CK(cudaSetDevice(device_));
    if (!d_ldr_ || !dst) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_process has not run" };
    CK(cudaMemcpyAsync(dst, d_ldr_, (size_t)W_ * H_ * 4, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
}
void Engine::get_bloom(float *dst) {
    CK(cudaSetDevice(device_));
    if (d_mips_.empty() || !dst) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "post_process has not run" };
    if (!bloom_valid_) { keep_bloom_ = true; post_process(); }              // the fused chain skips mip 0 until somebody asks for it
    CK(cudaMemcpyAsync(dst, d_mips_[0], (size_t)W_ * H_ * sizeof(float4), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
}

void Engine::volume_walks(uint32_t n, const float *org, const float *dir, const uint32_t *seeds, float ray_depth, float *T, float *scatter, int32_t *vol, uint32_t *rng2) {
    CK(cudaSetDevice(device_));
    if (!n) return;
    sync_all();
    float *d_o = nullptr, *d_d = nullptr, *d_T = nullptr, *d_s = nullptr; uint32_t *d_seed = nullptr, *d_r = nullptr; int32_t *d_v = nullptr;
    CK(cudaMalloc(&d_o, (size_t)n * 12)); CK(cudaMalloc(&d_d, (size_t)n * 12)); CK(cudaMalloc(&d_seed, (size_t)n * 4)); CK(cudaMalloc(&d_T, (size_t)n * 4));
    CK(cudaMalloc(&d_s, (size_t)n * 4)); CK(cudaMalloc(&d_v, (size_t)n * 4)); CK(cudaMalloc(&d_r, (size_t)n * 8));
    CK(cudaMemcpyAsync(d_o, org, (size_t)n * 12, cudaMemcpyHostToDevice, stream_)); CK(cudaMemcpyAsync(d_d, dir, (size_t)n * 12, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(d_seed, seeds, (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    launch_volume_walks(ds_, n, d_o, d_d, d_seed, ray_depth, d_T, d_s, d_v, d_r, stream_);
    CK(cudaMemcpyAsync(T, d_T, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_)); CK(cudaMemcpyAsync(scatter, d_s, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(vol, d_v, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_)); CK(cudaMemcpyAsync(rng2, d_r, (size_t)n * 8, cudaMemcpyDeviceToHost, stream_));
    cudaError_t e = cudaStreamSynchronize(stream_);
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_seed); cudaFree(d_T); cudaFree(d_s); cudaFree(d_v); cudaFree(d_r);
    CK(e); CK(cudaGetLastError());
}
void Engine::trace_closest(uint32_t n, const float *org, const float *dir, float tmin, float tmax, float *t, uint32_t *prim, uint32_t *inst, float *uv, uint32_t *stats) {
    CK(cudaSetDevice(device_));
    if (!has_scene_) throw CudaError{ B200PT_ERR_NO_SCENE, "no scene" };
    if (!n) return;
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr, *d_uv = nullptr; uint32_t *d_p = nullptr, *d_i = nullptr;
    CK(cudaMalloc(&d_o, (size_t)n * 12)); CK(cudaMalloc(&d_d, (size_t)n * 12)); CK(cudaMalloc(&d_t, (size_t)n * 4)); CK(cudaMalloc(&d_uv, (size_t)n * 8));
    CK(cudaMalloc(&d_p, (size_t)n * 4)); CK(cudaMalloc(&d_i, (size_t)n * 4));
    CK(cudaMemcpyAsync(d_o, org, (size_t)n * 12, cudaMemcpyHostToDevice, stream_)); CK(cudaMemcpyAsync(d_d, dir, (size_t)n * 12, cudaMemcpyHostToDevice, stream_));
    uint32_t *d_st = nullptr; if (stats) CK(cudaMalloc(&d_st, (size_t)n * 8));
    launch_trace_rays(lc_, ds_, n, d_o, d_d, tmin, tmax, d_t, d_p, d_i, d_uv, d_st, stream_);
    if (stats) CK(cudaMemcpyAsync(stats, d_st, (size_t)n * 8, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(t, d_t, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_)); CK(cudaMemcpyAsync(prim, d_p, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(inst, d_i, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_)); CK(cudaMemcpyAsync(uv, d_uv, (size_t)n * 8, cudaMemcpyDeviceToHost, stream_));
    cudaError_t e = cudaStreamSynchronize(stream_);
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_uv); cudaFree(d_p); cudaFree(d_i); if (d_st) cudaFree(d_st);
    CK(e); CK(cudaGetLastError());
}
// LookupTableCalculator::CalculateTable (PT/LookupTableCalculator.cpp:44-157) on this engine's device and stream.
// slices == 0 picks enough dispatch slices to give every SM ~2048 resident threads.
void Engine::bake_lut(int kind, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t sample_count, uint32_t seed, uint32_t slices, float *out_host, float *elapsed_ms) {
    CK(cudaSetDevice(device_));
    if (kind < 0 || kind > 2 || !sx || !sy || !sz || !out_host || (uint64_t)sx * sy * sz > (1ull << 28)) throw CudaError{ B200PT_ERR_WRONG_ARGUMENTS, "bake_lut: bad arguments" };
    const size_t n = (size_t)sx * sy * sz;
    if (!slices) {
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device_));
        const size_t want = (size_t)prop.multiProcessorCount * 2048;
        slices = (uint32_t)std::max<size_t>(1, (want + n - 1) / n);
    }
    slices = std::min<uint32_t>(slices, 4096u);
    float *d_part = nullptr, *d_tab = nullptr;
    CK(cudaMalloc(&d_part, n * slices * sizeof(float)));
    if (cudaMalloc(&d_tab, n * sizeof(float)) != cudaSuccess) { cudaFree(d_part); throw CudaError{ B200PT_ERR_OUT_OF_MEMORY, "bake_lut: out of device memory" }; }
    cudaEvent_t e0 = nullptr, e1 = nullptr; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, stream_);
    launch_bake_lut(kind, d_part, d_tab, sx, sy, sz, sample_count, seed, slices, stream_);
    cudaEventRecord(e1, stream_);
    cudaError_t e = cudaMemcpyAsync(out_host, d_tab, n * sizeof(float), cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.0f; if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    if (elapsed_ms) *elapsed_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_part); cudaFree(d_tab);
    CK(e);
}
void Engine::scene_stats(uint32_t *tris, uint32_t *nodes, uint32_t *emissive, uint32_t *textures) const {
    if (tris) *tris = n_tris_; if (nodes) *nodes = bvh_.n_nodes; if (emissive) *emissive = n_emissive_; if (textures) *textures = (uint32_t)scene_.textures.size();
}

} // namespace b200pt
