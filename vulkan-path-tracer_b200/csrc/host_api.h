// host_api.h -- host-side (C++17) pieces of the runtime above the C-ABI: codecs, glTF loader,
// environment-map preprocessing, camera, and the PathTracer / PostProcessor mirror classes.
#pragma once
#include "device_types.h"
#include <string>
#include <vector>
#include <memory>

namespace b200pt {

struct CudaError { int code; std::string what; };   // every failure crosses the C-ABI as a code + message (capi.cpp guard)

// ---- density_grid.cpp: AddDensityDataToVolume after the file read (PT/PathTracer.cpp:1391-1452) ----
struct PreparedGrid {
    std::vector<float> values, max_densities;       // temperature-patched values [z][y][x]; 32^3 majorants
    int imin[3], dim[3];
    float corner_min[3], corner_max[3], max_density;
    float wmin[3], wext[3], inv_vs[3], trans[3];    // SampleNanoVDBBuffer's per-grid constants (DevGrid)
    bool has_temperature = false;
};
void prepare_density_grid(const b200pt_density_grid &g, PreparedGrid &out);

// ---- image_codecs.cpp ----
bool decode_image_rgba8(const std::string &path, uint32_t &W, uint32_t &H, std::vector<uint8_t> &rgba, std::string &err);
bool decode_hdr_rgba32f(const std::string &path, uint32_t &W, uint32_t &H, std::vector<float> &rgba, std::string &err);
bool write_png_rgba8(const std::string &path, uint32_t W, uint32_t H, const uint8_t *rgba);

// ---- scene_loader.cpp : what AssetImporter::ImportScene + PathTracer::SetScene keep (Asset.h:131-138) ----
struct HostMesh { std::vector<b200pt_vertex> vertices; std::vector<uint32_t> indices; std::string name; };
struct HostTexture { uint32_t width = 0, height = 0, channels = 0; std::vector<uint8_t> data; };
struct HostScene {
    std::vector<HostMesh> meshes;
    std::vector<b200pt_material> materials;
    std::vector<std::string> material_names;
    std::vector<HostTexture> textures;
    std::vector<b200pt_instance> instances;
    float camera_view[16];
    float camera_aspect = 1.0f;
};
bool load_gltf_scene(const std::string &path, HostScene &out, std::string &err);   // .gltf / .glb
bool load_obj_scene(const std::string &path, HostScene &out, std::string &err);    // .obj + .mtl (assimp semantics from upstream knowledge)
bool load_scene_file(const std::string &path, HostScene &out, std::string &err);   // by extension, like AssetImporter::ImportScene
bool scene_from_desc(const b200pt_scene_desc *d, HostScene &out, std::string &err);

// ---- env_camera.cpp ----
// PathTracer::LoadEnvironmentMap (PathTracer.cpp:1137-1332): alias table + pdf in alpha. Returns the importance sum.
float build_env_alias(float *rgba, uint32_t width, uint32_t height, uint2 *alias_out);
// Editor/FlyCamera round trip (Editor.cpp:45-48,1042-1051; FlyCamera.cpp:84-140)
void camera_from_view(const float view[16], float aspect, float viewInv[16], float projInv[16]);
void mat3_inverse_from_o2w(const float o2w[12], float w2o[9]);

// ---- bloom mip chain sizes (PostProcessor.cpp:128-158) ----
uint32_t bloom_mip_sizes(uint32_t W, uint32_t H, uint32_t wh[20]);

// partition helpers
uint32_t partition_global_row(uint32_t local_row, uint32_t rank, uint32_t world, uint32_t band);
uint32_t partition_local_rows(uint32_t H, uint32_t rank, uint32_t world, uint32_t band);

} // namespace b200pt
