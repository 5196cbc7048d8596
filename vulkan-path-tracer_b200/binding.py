"""ctypes binding of the product's C-ABI (libb200pt.so, include/b200pt.h).

This is only a test/bench harness convenience: the product is the C-ABI library itself.  The class mirrors the
reference's `PathTracer` + `PostProcessor` pair (PathTracer/PathTracer.h:83-183, PathTracer/PostProcessor.h:25-33):
same verbs, same argument meaning.  There is NO fallback: if the CUDA library is missing or no GPU is present,
loading / creating a handle raises.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200PT_LIB") or os.path.join(_HERE, "libb200pt.so")   # B200PT_LIB: A/B a differently built library (profiles/)
_LIB = None

OK = 0
ERR_NOT_IMPLEMENTED = -15001
ERR_NO_SCENE = -15002
ERR_WRONG_ARGUMENTS = -15000
ERR_NO_DEVICE = -15004
ERR_INIT_FAILED = -3


class B200ptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200pt error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("BaseColor", C.c_float * 3), ("EmissiveColor", C.c_float * 3), ("SpecularColor", C.c_float * 3),
                ("MediumColor", C.c_float * 3), ("MediumEmissiveColor", C.c_float * 3),
                ("Metallic", C.c_float), ("Roughness", C.c_float), ("IOR", C.c_float), ("Transmission", C.c_float),
                ("Anisotropy", C.c_float), ("AnisotropyRotation", C.c_float), ("MediumDensity", C.c_float), ("MediumAnisotropy", C.c_float),
                ("BaseColorTextureIndex", C.c_uint32), ("NormalTextureIndex", C.c_uint32), ("RoughnessTextureIndex", C.c_uint32),
                ("MetallicTextureIndex", C.c_uint32), ("EmissiveTextureIndex", C.c_uint32)]


class Mesh(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("indices", C.c_void_p), ("vertex_count", C.c_uint32), ("index_count", C.c_uint32)]


class Instance(C.Structure):
    _fields_ = [("Transform", C.c_float * 16), ("MeshIndex", C.c_uint32), ("MaterialIndex", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32), ("_pad", C.c_uint32), ("data", C.c_void_p)]


class SceneDesc(C.Structure):
    _fields_ = [("meshes", C.c_void_p), ("mesh_count", C.c_uint32), ("_p0", C.c_uint32),
                ("materials", C.c_void_p), ("material_count", C.c_uint32), ("_p1", C.c_uint32),
                ("textures", C.c_void_p), ("texture_count", C.c_uint32), ("_p2", C.c_uint32),
                ("instances", C.c_void_p), ("instance_count", C.c_uint32), ("_p3", C.c_uint32),
                ("camera_view", C.c_float * 16), ("camera_aspect", C.c_float), ("_p4", C.c_uint32)]


class Config(C.Structure):
    _fields_ = [("SamplesPerFrame", C.c_uint32), ("MaxDepth", C.c_uint32), ("MaxLuminance", C.c_float), ("FocusDistance", C.c_float),
                ("DepthOfFieldStrength", C.c_float), ("SkyRotationAzimuth", C.c_float), ("SkyRotationAltitude", C.c_float),
                ("SkyIntensity", C.c_float), ("EmissiveMeshSamplingPDFBias", C.c_float), ("ScreenChunkCount", C.c_uint32),
                ("EnableSkyMIS", C.c_uint32), ("EnableMeshMIS", C.c_uint32), ("ShowEnvMapDirectly", C.c_uint32),
                ("UseOnlyGeometryNormals", C.c_uint32), ("UseEnergyCompensation", C.c_uint32), ("FurnaceTestMode", C.c_uint32),
                ("MaxSamplesAccumulated", C.c_uint32), ("FramesInFlight", C.c_uint32)]


class Volume(C.Structure):     # b200pt_volume (PathTracer::Volume, PT/PathTracer.h:36-70)
    _fields_ = [("CornerMin", C.c_float * 3), ("CornerMax", C.c_float * 3), ("Position", C.c_float * 3), ("Scale", C.c_float * 3),
                ("Color", C.c_float * 3), ("EmissiveColor", C.c_float * 3), ("TemperatureColor", C.c_float * 3),
                ("Density", C.c_float), ("Anisotropy", C.c_float), ("Alpha", C.c_float), ("DropletSize", C.c_float),
                ("DensityDataIndex", C.c_int32), ("MaxDensityInTheGrid", C.c_float), ("UseBlackbody", C.c_int32), ("HasTemperatureData", C.c_int32),
                ("TemperatureGamma", C.c_float), ("TemperatureScale", C.c_float), ("EmissiveColorGamma", C.c_float),
                ("KelvinMin", C.c_int32), ("KelvinMax", C.c_int32),
                ("ApproximatedScatteringForClouds", C.c_uint32), ("ApproximatedScatteringFalloff", C.c_float), ("GridSharpness", C.c_float)]


class DensityGrid(C.Structure):   # b200pt_density_grid
    _fields_ = [("IndexMin", C.c_int32 * 3), ("Dim", C.c_uint32 * 3), ("Density", C.c_void_p), ("Temperature", C.c_void_p),
                ("TemperatureMin", C.c_float), ("TemperatureMax", C.c_float), ("VoxelSize", C.c_double), ("Translation", C.c_double * 3)]


def make_density_grid(density, index_min=(0, 0, 0), temperature=None, voxel_size=1.0, translation=(0.0, 0.0, 0.0), temperature_range=None):
    """b200pt_density_grid over numpy arrays [z][y][x] (float32); returns (struct, keep-alive tuple)."""
    d = np.ascontiguousarray(density, dtype=np.float32)
    t = None if temperature is None else np.ascontiguousarray(temperature, dtype=np.float32)
    g = DensityGrid()
    for k in range(3): g.IndexMin[k] = int(index_min[k]); g.Dim[k] = d.shape[2 - k]; g.Translation[k] = float(translation[k])
    g.Density = d.ctypes.data; g.Temperature = None if t is None else t.ctypes.data
    g.TemperatureMin, g.TemperatureMax = (0.0, 0.0) if temperature_range is None else temperature_range
    g.VoxelSize = float(voxel_size)
    return g, (d, t)


def prepare_density_grid(density, **kw):
    """b200pt_prepare_density_grid (host only): (values, max_densities[32768], corner_min, corner_max, max_density)"""
    g, keep = make_density_grid(density, **kw)
    vals = np.empty(keep[0].shape, np.float32); maj = np.empty(32768, np.float32); cmin = (C.c_float * 3)(); cmax = (C.c_float * 3)(); mx = C.c_float()
    r = lib().b200pt_prepare_density_grid(C.byref(g), vals.ctypes.data_as(C.c_void_p), maj.ctypes.data_as(C.c_void_p), cmin, cmax, C.byref(mx))
    if r != OK: raise B200ptError(r, "prepare_density_grid")
    return vals, maj, tuple(cmin), tuple(cmax), mx.value


MAX_VOLUMES = 100


class Tonemap(C.Structure):
    _fields_ = [("Exposure", C.c_float), ("Gamma", C.c_float)]


class Bloom(C.Structure):
    _fields_ = [("BloomThreshold", C.c_float), ("BloomStrength", C.c_float), ("MipCount", C.c_uint32), ("FalloffRange", C.c_float)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "extend_rays", "shade_invocations", "surface_hits", "misses", "shadow_rays", "medium_events", "kernel_launches")] + \
               [(n, C.c_float) for n in ("ms_total", "ms_raygen", "ms_extend", "ms_shade", "ms_connect", "ms_resolve")] + \
               [("waves", C.c_uint32), ("bounces", C.c_uint32)]


def build(force=False):
    """Compile libb200pt.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))] + [os.path.join(_HERE, "..", "include", "b200pt.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "-s"])
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.b200pt_version.restype = C.c_char_p
        L.b200pt_last_error.restype = C.c_char_p; L.b200pt_last_error.argtypes = [C.c_void_p]
        L.b200pt_partition_global_row.restype = C.c_uint32; L.b200pt_partition_global_row.argtypes = [C.c_uint32] * 4
        L.b200pt_partition_local_row_count.restype = C.c_uint32; L.b200pt_partition_local_row_count.argtypes = [C.c_uint32] * 4
        L.b200pt_free.argtypes = [C.c_void_p]; L.b200pt_free.restype = None
        sig = {
            "b200pt_create": [C.c_int32, C.POINTER(C.c_void_p)], "b200pt_destroy": [C.c_void_p],
            "b200pt_set_scene_file": [C.c_void_p, C.c_char_p], "b200pt_set_scene_arrays": [C.c_void_p, C.POINTER(SceneDesc)],
            "b200pt_set_env_map_file": [C.c_void_p, C.c_char_p], "b200pt_set_env_map": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p],
            "b200pt_set_luts": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], "b200pt_set_luts_dir": [C.c_void_p, C.c_char_p],
            "b200pt_default_config": [C.POINTER(Config)], "b200pt_set_config": [C.c_void_p, C.POINTER(Config)], "b200pt_get_config": [C.c_void_p, C.POINTER(Config)],
            "b200pt_material_count": [C.c_void_p, C.POINTER(C.c_uint32)], "b200pt_get_material": [C.c_void_p, C.c_uint32, C.POINTER(Material)],
            "b200pt_set_material": [C.c_void_p, C.c_uint32, C.POINTER(Material)], "b200pt_get_material_name": [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32],
            "b200pt_set_camera": [C.c_void_p, C.c_void_p, C.c_void_p], "b200pt_get_camera": [C.c_void_p, C.c_void_p, C.c_void_p],
            "b200pt_camera_from_view": [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p],
            "b200pt_resize": [C.c_void_p, C.c_uint32, C.c_uint32], "b200pt_get_size": [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)],
            "b200pt_reset": [C.c_void_p], "b200pt_add_volume": [C.c_void_p, C.c_void_p],
            "b200pt_save_checkpoint": [C.c_void_p, C.c_char_p], "b200pt_load_checkpoint": [C.c_void_p, C.c_char_p],
            "b200pt_default_volume": [C.c_void_p], "b200pt_set_volume": [C.c_void_p, C.c_uint32, C.c_void_p], "b200pt_remove_volume": [C.c_void_p, C.c_uint32],
            "b200pt_volume_count": [C.c_void_p, C.POINTER(C.c_uint32)], "b200pt_get_volume": [C.c_void_p, C.c_uint32, C.c_void_p],
            "b200pt_add_density_data_to_volume": [C.c_void_p, C.c_uint32, C.c_char_p], "b200pt_remove_density_data_from_volume": [C.c_void_p, C.c_uint32],
            "b200pt_add_density_grid_to_volume": [C.c_void_p, C.c_uint32, C.c_void_p],
            "b200pt_prepare_density_grid": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "b200pt_set_phase_function": [C.c_void_p, C.c_uint32], "b200pt_get_phase_function": [C.c_void_p, C.POINTER(C.c_uint32)],
            "b200pt_set_partition": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32], "b200pt_local_rows": [C.c_void_p, C.POINTER(C.c_uint32)],
            "b200pt_path_trace": [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)], "b200pt_samples_accumulated": [C.c_void_p, C.POINTER(C.c_uint32)],
            "b200pt_synchronize": [C.c_void_p], "b200pt_set_stream": [C.c_void_p, C.c_void_p], "b200pt_set_profiling": [C.c_void_p, C.c_int32], "b200pt_get_hdr": [C.c_void_p, C.c_void_p, C.c_int32], "b200pt_hdr_device_ptr": [C.c_void_p, C.POINTER(C.c_void_p)],
            "b200pt_set_hdr": [C.c_void_p, C.c_void_p, C.c_int32], "b200pt_get_counters": [C.c_void_p, C.POINTER(Counters)],
            "b200pt_post_set_tonemap": [C.c_void_p, C.POINTER(Tonemap)], "b200pt_post_set_bloom": [C.c_void_p, C.POINTER(Bloom)],
            "b200pt_post_process": [C.c_void_p], "b200pt_get_ldr": [C.c_void_p, C.c_void_p, C.c_int32], "b200pt_get_bloom": [C.c_void_p, C.c_void_p],
            "b200pt_bloom_mip_sizes": [C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)], "b200pt_save_png": [C.c_void_p, C.c_char_p],
            "b200pt_volume_walks": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "b200pt_trace_closest": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "b200pt_scene_stats": [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4,
            "b200pt_trace_stats": [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p],
            "b200pt_bake_lut": [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p],
            "b200pt_bake_luts_to_dir": [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int32],
            "b200pt_decode_image_file": [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)],
            "b200pt_decode_hdr_file": [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_void_p)],
            "b200pt_write_png": [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p],
            "b200pt_build_env_alias": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_float)],
            "b200pt_bvh4_collapse": [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)],
            "b200pt_bvh2_sah_build": [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_double)],
            "b200pt_bvh2_reinsert": [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_int32, C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_void_p],
            "b200pt_bvh2_sah_rebuild": [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_void_p],
            "b200pt_load_gltf": [C.c_char_p, C.POINTER(C.POINTER(SceneDesc))], "b200pt_free_scene": [C.POINTER(SceneDesc)],
        }
        for name, args in sig.items():
            fn = getattr(L, name); fn.restype = C.c_int32; fn.argtypes = args
        _LIB = L
    return _LIB


class Atmosphere(C.Structure):      # b200pt_atmosphere (include/b200pt.h)
    _fields_ = [("Enable", C.c_uint32), ("PlanetPosition", C.c_float * 3), ("PlanetRadius", C.c_float), ("AtmosphereHeight", C.c_float),
                ("RayleighScatteringCoefficientMultiplier", C.c_float * 3), ("MieScatteringCoefficientMultiplier", C.c_float * 3),
                ("OzoneAbsorptionCoefficientMultiplier", C.c_float * 3), ("RayleighDensityFalloff", C.c_float), ("MieDensityFalloff", C.c_float),
                ("OzoneDensityFalloff", C.c_float), ("OzonePeak", C.c_float), ("SunColor", C.c_float * 3)]


def declared_symbols():
    """Every function name declared in include/b200pt.h."""
    import re
    txt = open(os.path.join(_HERE, "..", "include", "b200pt.h")).read()
    return sorted(set(re.findall(r"\b(b200pt_[a-z0-9_]+)\s*\(", txt)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def default_config(**kw):
    c = Config(); lib().b200pt_default_config(C.byref(c))
    for k, v in kw.items(): setattr(c, k, v)
    return c


def camera_from_view(view16, aspect):
    view16 = np.ascontiguousarray(view16, np.float32); vi = np.zeros(16, np.float32); pi = np.zeros(16, np.float32)
    r = lib().b200pt_camera_from_view(_p(view16), C.c_float(float(aspect)), _p(vi), _p(pi))
    if r != OK: raise B200ptError(r, "camera_from_view")
    return vi, pi


def partition_rows(H, rank, world, band):
    n = lib().b200pt_partition_local_row_count(H, rank, world, band)
    return np.array([lib().b200pt_partition_global_row(i, rank, world, band) for i in range(n)], dtype=np.int64)


def decode_image(path):
    w, h, p = C.c_uint32(), C.c_uint32(), C.c_void_p()
    r = lib().b200pt_decode_image_file(path.encode(), C.byref(w), C.byref(h), C.byref(p))
    if r != OK: raise B200ptError(r, lib().b200pt_last_error(None).decode(errors="replace"))
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 4)).copy()
    lib().b200pt_free(p)
    return a


def decode_hdr(path):
    w, h, p = C.c_uint32(), C.c_uint32(), C.c_void_p()
    r = lib().b200pt_decode_hdr_file(path.encode(), C.byref(w), C.byref(h), C.byref(p))
    if r != OK: raise B200ptError(r, lib().b200pt_last_error(None).decode(errors="replace"))
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(h.value, w.value, 4)).copy()
    lib().b200pt_free(p)
    return a


def write_png(path, rgba):
    rgba = np.ascontiguousarray(rgba, np.uint8)
    r = lib().b200pt_write_png(path.encode(), rgba.shape[1], rgba.shape[0], _p(rgba))
    if r != OK: raise B200ptError(r, "write_png")


BVH2_DTYPE = np.dtype([("lo0", "<f4", 3), ("hi0", "<f4", 3), ("lo1", "<f4", 3), ("hi1", "<f4", 3), ("c0", "<i4"), ("c1", "<i4"), ("pad", "<u4", 2)])
BVH4_DTYPE = np.dtype([("lox", "<f4", 4), ("loy", "<f4", 4), ("loz", "<f4", 4), ("hix", "<f4", 4), ("hiy", "<f4", 4), ("hiz", "<f4", 4), ("child", "<i4", 4), ("pad", "<u4", 4)])
BVH4_EMPTY = -(1 << 31)


def bvh4_collapse(nodes2, root2=0):
    """Host BVH2 -> BVH4 collapse (csrc/lbvh.cu: bvh4_collapse_host).  nodes2: BVH2_DTYPE array; returns (BVH4_DTYPE array, depth)."""
    nodes2 = np.ascontiguousarray(nodes2, BVH2_DTYPE)
    out = np.zeros(len(nodes2), BVH4_DTYPE)
    n4, d = C.c_uint32(), C.c_int32()
    r = lib().b200pt_bvh4_collapse(_p(nodes2), len(nodes2), int(root2), _p(out), C.byref(n4), C.byref(d))
    if r != OK: raise B200ptError(r, "bvh4_collapse")
    return out[:n4.value].copy(), d.value


def bvh2_sah_rebuild(nodes2, root2=0):
    """Opt-in binned-SAH rebuild of the inner nodes (csrc/lbvh.cu: bvh2_sah_rebuild_host).  Returns (BVH2_DTYPE array, depth, (sah_before, sah_after))."""
    nodes2 = np.ascontiguousarray(nodes2, BVH2_DTYPE)
    out = np.zeros(len(nodes2), BVH2_DTYPE)
    n, d = C.c_uint32(), C.c_int32(); sah = np.zeros(2, np.float64)
    r = lib().b200pt_bvh2_sah_rebuild(_p(nodes2), len(nodes2), int(root2), _p(out), C.byref(n), C.byref(d), _p(sah))
    if r != OK: raise B200ptError(r, "bvh2_sah_rebuild")
    return out[:n.value].copy(), d.value, (float(sah[0]), float(sah[1]))


def bvh2_sah_build(ref_boxes, trav_cost=1.0):
    """Full binned-SAH build from reference boxes [n, 6] (csrc/lbvh.cu: bvh2_sah_build_host).  Returns (nodes, perm, depth, sah_cost)."""
    rb = np.ascontiguousarray(ref_boxes, np.float32).reshape(-1, 6)
    out = np.zeros(max(len(rb) - 1, 1), BVH2_DTYPE); perm = np.zeros(len(rb), np.uint32)
    n, d, c = C.c_uint32(), C.c_int32(), C.c_double()
    r = lib().b200pt_bvh2_sah_build(_p(rb), len(rb), C.c_float(trav_cost), _p(out), _p(perm), C.byref(n), C.byref(d), C.byref(c))
    if r != OK: raise B200ptError(r, "bvh2_sah_build")
    return out[:n.value].copy(), perm, d.value, c.value


def bvh2_reinsert(nodes2, root2=0, passes=2, fraction=0.25):
    """Insertion-based refinement (csrc/lbvh.cu: bvh2_reinsert_host).  Returns (BVH2_DTYPE array, depth, (sah_before, sah_after))."""
    nodes2 = np.ascontiguousarray(nodes2, BVH2_DTYPE)
    out = np.zeros(len(nodes2), BVH2_DTYPE)
    n, d = C.c_uint32(), C.c_int32(); sah = np.zeros(2, np.float64)
    r = lib().b200pt_bvh2_reinsert(_p(nodes2), len(nodes2), int(root2), _p(out), int(passes), C.c_float(fraction), C.byref(n), C.byref(d), _p(sah))
    if r != OK: raise B200ptError(r, "bvh2_reinsert")
    return out[:n.value].copy(), d.value, (float(sah[0]), float(sah[1]))


def build_env_alias(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32).copy(); h, w = rgba.shape[:2]
    alias = np.zeros(h * w, dtype=np.dtype([("Alias", "<u4"), ("Importance", "<f4")]))
    s = C.c_float()
    r = lib().b200pt_build_env_alias(_p(rgba), w, h, _p(alias), C.byref(s))
    if r != OK: raise B200ptError(r, "build_env_alias")
    return rgba, alias, s.value


def bloom_mip_sizes(W, H):
    wh = np.zeros(20, np.uint32); n = C.c_uint32()
    lib().b200pt_bloom_mip_sizes(W, H, _p(wh), C.byref(n))
    return [(int(wh[2 * i]), int(wh[2 * i + 1])) for i in range(n.value)]


def load_gltf(path):
    """C++ loader -> python dict in the same shape as oracle.gltf_ref.load_gltf (for loader parity tests)."""
    pd = C.POINTER(SceneDesc)()
    r = lib().b200pt_load_gltf(path.encode(), C.byref(pd))
    if r != OK: raise B200ptError(r, lib().b200pt_last_error(None).decode(errors="replace"))
    d = pd.contents
    vdt = np.dtype([("pos", "<f4", 3), ("nrm", "<f4", 3), ("uv", "<f4", 2)])
    meshes = []
    ms = C.cast(d.meshes, C.POINTER(Mesh))
    for i in range(d.mesh_count):
        v = np.ctypeslib.as_array(C.cast(ms[i].vertices, C.POINTER(C.c_float)), shape=(ms[i].vertex_count, 8)).copy().view(vdt).reshape(-1)
        ix = np.ctypeslib.as_array(C.cast(ms[i].indices, C.POINTER(C.c_uint32)), shape=(ms[i].index_count,)).copy()
        meshes.append((v, ix))
    mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_uint8)), shape=(d.material_count * 112,)).copy()
    ts = C.cast(d.textures, C.POINTER(Texture)); textures = []
    for i in range(d.texture_count):
        textures.append(np.ctypeslib.as_array(C.cast(ts[i].data, C.POINTER(C.c_uint8)), shape=(ts[i].height, ts[i].width, ts[i].channels)).copy())
    ins = C.cast(d.instances, C.POINTER(Instance)); instances = []
    for i in range(d.instance_count):
        instances.append((np.array(list(ins[i].Transform), np.float32), int(ins[i].MeshIndex), int(ins[i].MaterialIndex)))
    out = dict(meshes=meshes, materials_bytes=mats, textures=textures, instances=instances,
               camera_view=np.array(list(d.camera_view), np.float32), aspect=np.float32(d.camera_aspect))
    lib().b200pt_free_scene(pd)
    return out


class PathTracer:
    """One GPU: the reference's PathTracer + PostProcessor behind the C-ABI."""

    def __init__(self, device=0):
        self.L = lib(); self.h = C.c_void_p()
        r = self.L.b200pt_create(device, C.byref(self.h))
        if r != OK: raise B200ptError(r, self.L.b200pt_last_error(None).decode(errors="replace"))
        self._keep = []

    def close(self):
        if self.h: self.L.b200pt_destroy(self.h); self.h = C.c_void_p()

    def __del__(self):
        try: self.close()
        except Exception: pass

    def _ck(self, r):
        if r != OK: raise B200ptError(r, self.L.b200pt_last_error(self.h).decode(errors="replace"))

    # ---- volumes: PathTracer::AddVolume / SetVolume / RemoveVolume / SetPhaseFunction (homogeneous AABB volumes)
    @staticmethod
    def make_volume(**kw):
        v = Volume(); r = lib().b200pt_default_volume(C.byref(v))
        if r != OK: raise B200ptError(r, "default_volume")
        for k, val in kw.items():
            if k == "ApproximatedScattering": k = "ApproximatedScatteringForClouds"
            if k == "Grid": continue                                   # density data: add_density_grid after add_volume (tests/util.py)
            if k in ("CornerMin", "CornerMax", "Position", "Scale", "Color", "EmissiveColor", "TemperatureColor"):
                for j in range(3): getattr(v, k)[j] = float(val[j])
            else: setattr(v, k, val)
        return v

    def add_density_grid(self, idx, density, **kw):
        """AddDensityDataToVolume after the file read: density / temperature as numpy [z][y][x] (see make_density_grid)"""
        g, keep = make_density_grid(density, **kw)
        self._ck(self.L.b200pt_add_density_grid_to_volume(self.h, idx, C.byref(g)))

    def remove_density_data(self, idx): self._ck(self.L.b200pt_remove_density_data_from_volume(self.h, idx))

    def add_volume(self, vol=None, **kw):
        v = vol if vol is not None else self.make_volume(**kw)
        self._ck(self.L.b200pt_add_volume(self.h, C.byref(v)))

    def set_volume(self, idx, vol): self._ck(self.L.b200pt_set_volume(self.h, idx, C.byref(vol)))

    def remove_volume(self, idx): self._ck(self.L.b200pt_remove_volume(self.h, idx))

    def volume_count(self):
        n = C.c_uint32(); self._ck(self.L.b200pt_volume_count(self.h, C.byref(n))); return n.value

    def get_volume(self, idx):
        v = Volume(); self._ck(self.L.b200pt_get_volume(self.h, idx, C.byref(v))); return v

    def set_phase_function(self, pf): self._ck(self.L.b200pt_set_phase_function(self.h, pf))

    def get_phase_function(self):
        n = C.c_uint32(); self._ck(self.L.b200pt_get_phase_function(self.h, C.byref(n))); return n.value

    # ---- PathTracer::SetScene
    def set_scene_file(self, path): self._ck(self.L.b200pt_set_scene_file(self.h, path.encode()))

    def set_scene(self, sc):
        """sc: dict as produced by oracle.gltf_ref.load_gltf / load_scene_npz."""
        keep = []
        meshes = (Mesh * len(sc["meshes"]))()
        for i, (v, idx) in enumerate(sc["meshes"]):
            v = np.ascontiguousarray(v); idx = np.ascontiguousarray(idx, np.uint32); keep += [v, idx]
            meshes[i] = Mesh(v.ctypes.data, idx.ctypes.data, len(v), len(idx))
        mats = np.ascontiguousarray(sc["materials"]); keep.append(mats)
        texs = (Texture * len(sc["textures"]))()
        for i, t in enumerate(sc["textures"]):
            t = np.ascontiguousarray(t, np.uint8); keep.append(t)
            texs[i] = Texture(t.shape[1], t.shape[0], t.shape[2], 0, t.ctypes.data)
        insts = (Instance * len(sc["instances"]))()
        for i, (xf, m, mat) in enumerate(sc["instances"]):
            insts[i].MeshIndex = m; insts[i].MaterialIndex = mat
            for k in range(16): insts[i].Transform[k] = float(xf[k])
        d = SceneDesc()
        d.meshes = C.addressof(meshes); d.mesh_count = len(sc["meshes"])
        d.materials = mats.ctypes.data; d.material_count = len(mats)
        d.textures = C.addressof(texs); d.texture_count = len(sc["textures"])
        d.instances = C.addressof(insts); d.instance_count = len(sc["instances"])
        for k in range(16): d.camera_view[k] = float(sc["camera_view"][k])
        d.camera_aspect = float(sc["aspect"])
        self._ck(self.L.b200pt_set_scene_arrays(self.h, C.byref(d)))

    def set_env_map(self, rgba):
        rgba = np.ascontiguousarray(rgba, np.float32)
        self._ck(self.L.b200pt_set_env_map(self.h, rgba.shape[1], rgba.shape[0], _p(rgba)))

    def set_env_map_file(self, path): self._ck(self.L.b200pt_set_env_map_file(self.h, path.encode()))

    def set_luts(self, refl, rout, rin):
        a, b, c = [np.ascontiguousarray(x, np.float32) for x in (refl, rout, rin)]
        self._ck(self.L.b200pt_set_luts(self.h, _p(a), _p(b), _p(c)))

    def set_luts_dir(self, d): self._ck(self.L.b200pt_set_luts_dir(self.h, d.encode()))

    # ---- parameters
    def save_checkpoint(self, path): self._ck(self.L.b200pt_save_checkpoint(self.h, str(path).encode()))

    def load_checkpoint(self, path): self._ck(self.L.b200pt_load_checkpoint(self.h, str(path).encode()))

    def set_config(self, cfg): self._ck(self.L.b200pt_set_config(self.h, C.byref(cfg)))

    def get_config(self):
        c = Config(); self._ck(self.L.b200pt_get_config(self.h, C.byref(c))); return c

    def material_count(self):
        n = C.c_uint32(); self._ck(self.L.b200pt_material_count(self.h, C.byref(n))); return n.value

    def get_material(self, i):
        m = Material(); self._ck(self.L.b200pt_get_material(self.h, i, C.byref(m))); return m

    def set_material(self, i, m): self._ck(self.L.b200pt_set_material(self.h, i, C.byref(m)))

    def get_material_name(self, i):
        b = C.create_string_buffer(256); self._ck(self.L.b200pt_get_material_name(self.h, i, b, 256)); return b.value.decode()

    def set_camera(self, vi, pi):
        vi = np.ascontiguousarray(vi, np.float32); pi = np.ascontiguousarray(pi, np.float32)
        self._ck(self.L.b200pt_set_camera(self.h, _p(vi), _p(pi)))

    def get_camera(self):
        vi = np.zeros(16, np.float32); pi = np.zeros(16, np.float32)
        self._ck(self.L.b200pt_get_camera(self.h, _p(vi), _p(pi))); return vi, pi

    def resize(self, w, h): self._ck(self.L.b200pt_resize(self.h, w, h))

    def size(self):
        w, h = C.c_uint32(), C.c_uint32(); self._ck(self.L.b200pt_get_size(self.h, C.byref(w), C.byref(h))); return w.value, h.value

    def reset(self): self._ck(self.L.b200pt_reset(self.h))

    def set_partition(self, rank, world, band): self._ck(self.L.b200pt_set_partition(self.h, rank, world, band))

    def local_rows(self):
        n = C.c_uint32(); self._ck(self.L.b200pt_local_rows(self.h, C.byref(n))); return n.value

    # ---- hot path
    def path_trace(self, dispatches, base_seed):
        done = C.c_int32(); self._ck(self.L.b200pt_path_trace(self.h, dispatches, base_seed & 0xFFFFFFFF, C.byref(done))); return bool(done.value)

    def samples_accumulated(self):
        n = C.c_uint32(); self._ck(self.L.b200pt_samples_accumulated(self.h, C.byref(n))); return n.value

    def synchronize(self): self._ck(self.L.b200pt_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr): self._ck(self.L.b200pt_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_profiling(self, on): self._ck(self.L.b200pt_set_profiling(self.h, 1 if on else 0))

    def get_hdr(self, out=None):
        w, h = self.size(); rows = self.local_rows()
        if out is None: out = np.empty((rows, w, 4), np.float32)
        self._ck(self.L.b200pt_get_hdr(self.h, _p(out), 0)); return out

    def get_hdr_into_device(self, ptr): self._ck(self.L.b200pt_get_hdr(self.h, C.c_void_p(ptr), 1))

    def set_hdr(self, img):
        img = np.ascontiguousarray(img, np.float32); self._ck(self.L.b200pt_set_hdr(self.h, _p(img), 0))

    def set_hdr_from_device(self, ptr): self._ck(self.L.b200pt_set_hdr(self.h, C.c_void_p(ptr), 1))

    def counters(self):
        c = Counters(); self._ck(self.L.b200pt_get_counters(self.h, C.byref(c)))
        return {n: getattr(c, n) for n, _ in Counters._fields_}

    # ---- PostProcessor
    def set_tonemap(self, exposure=1.0, gamma=2.2): self._ck(self.L.b200pt_post_set_tonemap(self.h, C.byref(Tonemap(exposure, gamma))))

    def set_bloom(self, threshold=2.0, strength=1.0, mips=10, falloff=5.0): self._ck(self.L.b200pt_post_set_bloom(self.h, C.byref(Bloom(threshold, strength, mips, falloff))))

    def get_atmosphere(self):
        a = Atmosphere(); self._ck(self.L.b200pt_get_atmosphere(self.h, C.byref(a))); return a

    def set_atmosphere(self, **kw):
        """fields of b200pt_atmosphere (the reference's twelve atmosphere setters); unnamed fields keep their current value"""
        a = self.get_atmosphere()
        for k, v in kw.items():
            if k in ("PlanetPosition", "RayleighScatteringCoefficientMultiplier", "MieScatteringCoefficientMultiplier", "OzoneAbsorptionCoefficientMultiplier", "SunColor"):
                for i in range(3): getattr(a, k)[i] = float(v[i])
            else: setattr(a, k, v)
        self._ck(self.L.b200pt_set_atmosphere(self.h, C.byref(a)))

    def flush(self): self._ck(self.L.b200pt_flush(self.h))

    def post_process(self): self._ck(self.L.b200pt_post_process(self.h))

    def get_ldr(self, out=None):
        w, h = self.size()
        if out is None: out = np.empty((h, w, 4), np.uint8)
        self._ck(self.L.b200pt_get_ldr(self.h, _p(out), 0)); return out

    def accumulate_rows(self, frame_device_ptr, frame_index, y0, y1):
        self._ck(self.L.b200pt_accumulate_rows(self.h, C.c_void_p(frame_device_ptr), C.c_uint32(frame_index), C.c_uint32(y0), C.c_uint32(y1)))

    def post_input_rows(self, y0, y1):
        a, b = C.c_uint32(), C.c_uint32()
        self._ck(self.L.b200pt_post_input_rows(self.h, C.c_uint32(y0), C.c_uint32(y1), C.byref(a), C.byref(b))); return a.value, b.value

    def post_process_rows(self, y0, y1): self._ck(self.L.b200pt_post_process_rows(self.h, C.c_uint32(y0), C.c_uint32(y1)))

    def get_ldr_rows(self, y0, y1, out=None):
        w, h = self.size()
        if out is None: out = np.empty((y1 - y0, w, 4), np.uint8)
        self._ck(self.L.b200pt_get_ldr_rows(self.h, C.c_uint32(y0), C.c_uint32(y1), _p(out), 0)); return out

    def get_ldr_rows_into_device(self, y0, y1, device_ptr):
        self._ck(self.L.b200pt_get_ldr_rows(self.h, C.c_uint32(y0), C.c_uint32(y1), C.c_void_p(device_ptr), 1))

    def get_bloom(self):
        w, h = self.size(); out = np.empty((h, w, 4), np.float32)
        self._ck(self.L.b200pt_get_bloom(self.h, _p(out))); return out

    def save_png(self, path): self._ck(self.L.b200pt_save_png(self.h, path.encode()))

    # ---- test hooks
    def volume_walks(self, org, dirs, seeds, ray_depth=0.0):
        """(T, scatter distance, scattering volume, sampler state after each walk [n, 2]) for every ray / seed"""
        org = np.ascontiguousarray(org, np.float32); dirs = np.ascontiguousarray(dirs, np.float32); seeds = np.ascontiguousarray(seeds, np.uint32); n = len(org)
        T = np.zeros(n, np.float32); sd = np.zeros(n, np.float32); vol = np.zeros(n, np.int32); rng = np.zeros((n, 2), np.uint32)
        self._ck(self.L.b200pt_volume_walks(self.h, n, _p(org), _p(dirs), _p(seeds), C.c_float(ray_depth), _p(T), _p(sd), _p(vol), _p(rng)))
        return T, sd, vol, rng

    def trace_closest(self, org, dirs, tmin, tmax):
        org = np.ascontiguousarray(org, np.float32); dirs = np.ascontiguousarray(dirs, np.float32); n = len(org)
        t = np.zeros(n, np.float32); prim = np.zeros(n, np.uint32); inst = np.zeros(n, np.uint32); uv = np.zeros((n, 2), np.float32)
        self._ck(self.L.b200pt_trace_closest(self.h, n, _p(org), _p(dirs), C.c_float(tmin), C.c_float(tmax), _p(t), _p(prim), _p(inst), _p(uv)))
        return t, prim, inst, uv

    LUT_SIZES = {0: (64, 64, 32), 1: (128, 128, 32), 2: (128, 128, 32)}   # Application.cpp:41,54,67

    def bake_lut(self, kind, sample_count, seed=0, size=None, slices=0):
        """LookupTableCalculator::CalculateTable on the GPU -> (table[z][y][x] float32, device milliseconds)."""
        sx, sy, sz = size or self.LUT_SIZES[kind]
        out = np.zeros((sz, sy, sx), np.float32); ms = C.c_float(0)
        self._ck(self.L.b200pt_bake_lut(self.h, kind, sx, sy, sz, sample_count, seed, slices, _p(out), C.byref(ms)))
        return out, ms.value

    def bake_luts_to_dir(self, path, sample_count=10_000_000, seed=0, overwrite=False):
        self._ck(self.L.b200pt_bake_luts_to_dir(self.h, str(path).encode(), sample_count, seed, int(overwrite)))

    def trace_stats(self, org, dirs, tmin, tmax):
        org = np.ascontiguousarray(org, np.float32); dirs = np.ascontiguousarray(dirs, np.float32); n = len(org)
        st = np.zeros((n, 2), np.uint32)
        self._ck(self.L.b200pt_trace_stats(self.h, n, _p(org), _p(dirs), C.c_float(tmin), C.c_float(tmax), _p(st)))
        return st

    def scene_stats(self):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self.L.b200pt_scene_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(triangles=a.value, bvh_nodes=b.value, emissive_meshes=c.value, textures=d.value)
