// engine.h -- host runtime: the B200 counterparts of the reference's PathTracer (PathTracer/PathTracer.h:83-183)
// and PostProcessor (PathTracer/PostProcessor.h:25-33) classes.  One Engine == one GPU.
#pragma once
#include "host_api.h"
#include "kernels.h"
#include <cuda_runtime.h>

namespace b200pt {

class Engine {
public:
    explicit Engine(int device);
    ~Engine();

    // PathTracer::SetScene
    void set_scene(HostScene &&scene);
    void set_env_map(uint32_t w, uint32_t h, const float *rgba);
    void set_luts(const float *refl, const float *rout, const float *rin);

    void set_config(const b200pt_config &c);
    const b200pt_config &config() const { return cfg_; }
    void set_material(uint32_t idx, const b200pt_material &m);
    // PathTracer::AddVolume / SetVolume / RemoveVolume / SetPhaseFunction / AddDensityDataToVolume / RemoveDensityDataFromVolume (volumes.cuh)
    uint32_t add_volume(const b200pt_volume &v);
    void set_volume(uint32_t idx, const b200pt_volume &v);
    void remove_volume(uint32_t idx);
    void add_density_grid(uint32_t idx, const b200pt_density_grid &g);
    void remove_density_data(uint32_t idx);
    const std::vector<b200pt_volume> &volumes() const { return volumes_; }
    void set_phase_function(uint32_t pf);
    uint32_t phase_function() const { return phase_function_; }
    void set_atmosphere(const b200pt_atmosphere &a) { atmosphere_ = a; reset(); }     // parameters only (PathTracer.h:170-181); rendering with it is not built
    const b200pt_atmosphere &atmosphere() const { return atmosphere_; }
    const HostScene &scene() const { return scene_; }
    bool has_scene() const { return has_scene_; }
    void set_camera(const float vi[16], const float pi[16]);
    void get_camera(float vi[16], float pi[16]) const;
    void resize(uint32_t w, uint32_t h);
    void reset();                                   // ResetPathTracing
    void set_partition(uint32_t rank, uint32_t world, uint32_t band);
    uint32_t width() const { return W_; }
    uint32_t height() const { return H_; }
    uint32_t local_rows() const { return local_rows_; }
    uint32_t samples_accumulated() const { return samples_accumulated_; }

    // PathTracer::PathTrace x dispatches; returns true when all samples are accumulated
    bool path_trace(uint32_t dispatches, uint32_t base_seed);
    void synchronize();
    void flush() { check(cudaSetDevice(device_), "cudaSetDevice"); join_waves(); }   // device-side: the caller's stream waits for every launched wave
    void set_stream(cudaStream_t s);
    void set_profiling(bool on) { profiling_ = on; }
    void get_hdr(float *dst, bool dst_is_device);
    void set_hdr(const float *src, bool src_is_device);
    float4 *hdr_device() { return d_image_; }
    // checkpoint / resume of the accumulation (SURVEY 5 "checkpoint / resume", 8f row 4): image + dispatch bookkeeping
    void save_checkpoint(const char *path);
    void load_checkpoint(const char *path);
    b200pt_counters counters();

    // PostProcessor
    void set_tonemap(const b200pt_tonemap &t) { tonemap_ = t; }
    void set_bloom(const b200pt_bloom &b) { bloom_ = b; }
    // HDR accumulate as a stand-alone pass (config 5): rows [y0, y1) of a full-size device frame are folded into the accumulation image (frame_index = frames already in it)
    void accumulate_rows(const float4 *d_frame, uint32_t frame_index, uint32_t y0, uint32_t y1);
    void post_input_rows(uint32_t y0, uint32_t y1, uint32_t *in0, uint32_t *in1);   // HDR rows that post_process_rows(y0, y1) reads
    void post_process();
    void post_process_rows(uint32_t y0, uint32_t y1);             // output rows [y0, y1) only (multi-GPU post pass); equals the same rows of post_process()
    void get_ldr_rows(uint32_t y0, uint32_t y1, uint8_t *dst, bool dst_is_device);   // device destination: asynchronous on the engine's stream
    void get_ldr(uint8_t *dst, bool dst_is_device);
    void get_bloom(float *dst);

    void volume_walks(uint32_t n, const float *org, const float *dir, const uint32_t *seeds, float ray_depth, float *T, float *scatter, int32_t *vol, uint32_t *rng2);
    void trace_closest(uint32_t n, const float *org, const float *dir, float tmin, float tmax, float *t, uint32_t *prim, uint32_t *inst, float *uv, uint32_t *stats = nullptr);
    void scene_stats(uint32_t *tris, uint32_t *nodes, uint32_t *emissive, uint32_t *textures) const;
    void bake_lut(int kind, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t sample_count, uint32_t seed, uint32_t slices, float *out_host, float *elapsed_ms);

    std::string last_error;

private:
    void free_scene();
    void upload_scene();
    void rebuild_emissive();
    void rebuild_tri_class();                       // per-triangle shading class (MaterialClass) + the set of classes present in the scene
    void upload_volumes();
    DevMaterial make_dev_material(const b200pt_material &m) const;
    void ensure_image();
    void ensure_wave(size_t capacity, int contexts);
    void free_wave();
    void ensure_post();
    void free_post();
    DevConfig make_dev_config() const;
    void check(cudaError_t e, const char *what) const;

    int device_;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    bool profiling_ = false;
    std::vector<cudaEvent_t> prof_ev_;
    cudaEvent_t ev_[2] = { nullptr, nullptr };

    HostScene scene_;
    bool has_scene_ = false, has_env_ = false, has_luts_ = false;
    b200pt_config cfg_;
    float view_inv_[16], proj_inv_[16];
    uint32_t W_ = 0, H_ = 0, rank_ = 0, world_ = 1, band_ = 16, local_rows_ = 0;
    uint64_t dispatch_count_ = 0;
    uint32_t frame_count_ = 0, samples_accumulated_ = 0;

    // device scene
    DevScene ds_{};
    b200pt_vertex *d_verts_ = nullptr; uint32_t *d_indices_ = nullptr; DevMesh *d_meshes_ = nullptr; DevInstance *d_instances_ = nullptr;
    DevMaterial *d_materials_ = nullptr; DevTexture *d_textures_ = nullptr; DevEmissive *d_emissive_ = nullptr; EmTri *d_em_tris_ = nullptr; uint32_t *d_em_tri_base_ = nullptr;
    std::vector<uint8_t *> d_texdata_;
    std::vector<DevInstance> h_instances_; std::vector<DevMesh> h_meshes_;
    float4 *d_env_ = nullptr; uint2 *d_alias_ = nullptr; float2 *d_env_row_cos_ = nullptr; float *d_luts_[3] = { nullptr, nullptr, nullptr };
    std::vector<b200pt_volume> volumes_; uint32_t phase_function_ = 0; DevVolume *d_volumes_ = nullptr;
    struct GridData { PreparedGrid host; float *d_values = nullptr, *d_max_densities = nullptr; ~GridData(); };
    std::vector<std::shared_ptr<GridData>> grids_;      // parallel to volumes_: density data of the heterogeneous ones (nullptr: homogeneous)
    DevGrid *d_grids_ = nullptr; int density_data_counter_ = 0;   // PathTracer.cpp:1508-1513: slots are handed out round-robin
    b200pt_atmosphere atmosphere_{};
    LbvhResult bvh_{};
    LaunchCfg lc_{};
    uint32_t n_tris_ = 0, n_emissive_ = 0;
    uint8_t *d_tri_class_ = nullptr; uint32_t class_mask_ = 0;   // bit c: some triangle's material is of MaterialClass c

    // framebuffer + post
    float4 *d_image_ = nullptr; size_t image_pixels_ = 0;
    b200pt_tonemap tonemap_{ 1.0f, 2.2f };
    b200pt_bloom bloom_{ 2.0f, 1.0f, 10, 5.0f };
    std::vector<float4 *> d_mips_; std::vector<uint32_t> mip_wh_; uchar4 *d_ldr_ = nullptr; uint32_t post_w_ = 0, post_h_ = 0;
    bool keep_bloom_ = false, bloom_valid_ = false;   // fused post chain: mip 0 is written only once get_bloom() has been used

    // wave buffers: one context per wave in flight.  Two contexts on two internal streams let the latency-bound tail of wave w (late bounces: few, long rays)
    // overlap the full-machine head of wave w+1; k_resolve runs in dispatch order (events), so the image does not depend on it.
    struct WaveBuf {
        PathState ps[2]{}; ShadeOut so{};
        float4 *sample_buf = nullptr; uint32_t *rng_carry = nullptr; uint32_t *q_hit[2] = { nullptr, nullptr }, *q_miss[2] = { nullptr, nullptr };
        uint32_t *counts = nullptr, *h_count = nullptr;
        DevDispatch *d_disp = nullptr, *h_disp = nullptr; cudaEvent_t disp_ev = nullptr;   // this context's slice of the dispatch table (pinned host copy reusable once disp_ev has fired)
        size_t cap = 0;
    };
    WaveBuf wb_[2];
    cudaStream_t aux_stream_[2] = { nullptr, nullptr }; cudaEvent_t resolved_ev_[2] = { nullptr, nullptr }, image_free_ev_ = nullptr;
    uint64_t wave_seq_ = 0;                         // waves launched so far: wave w runs in context w & 1 and resolves after wave w - 1
    size_t wave_cap_ = 0;                           // capacity (paths) of every allocated context
    void sync_all();                                // the caller's stream AND the internal wave streams are idle
    void join_waves();                              // work enqueued on the caller's stream from here on sees every launched wave resolved into the image
    WaveCounters *d_ctr_ = nullptr;
    uint2 *d_sort_key_rank_ = nullptr; uint32_t *d_sort_hist_ = nullptr, *d_sort_offs_ = nullptr, *d_order_ = nullptr;   // ray sort of incoherent bounces (launch_ray_sort)
    bool sort_rays_ = false;
    b200pt_counters last_{};
};

} // namespace b200pt
