"""
ORACLE-SIDE scene loader (TEST INFRASTRUCTURE ONLY -- the product has its own C++ loader).

Restates what `AssetImporter::ImportScene` + `PathTracer::SetScene` hand to the shaders
(reference: VulkanHelper/Source/Utility/AssetImporterImpl.cpp:76-642, PathTracer/PathTracer.cpp:158-502)
for glTF 2.0 input, *without* assimp (assimp v6.0.2 is a network dependency that is not vendored in
/root/reference; its glTF2 importer's published behaviour is restated):

  * one mesh per glTF primitive, in mesh order; vertex = {pos, normalize(normal), uv} (32 B), u32 indices
  * one instance per (node, primitive) in depth-first scene order,
    Transform = diag(1,-1,1,1) * nodeWorld              (AssetImporterImpl.cpp:233-247)
  * material key mapping and defaults                    (AssetImporterImpl.cpp:353-455)
  * camera: ViewMatrix = inverse(flipY * nodeWorld * local), aspect = perspective.aspectRatio (:547-642)
  * texture table built in PathTracer::SetScene order with 1x1 defaults (PathTracer.cpp:228-408, 1557-1621)

assimp's OptimizeMeshes/OptimizeGraph/JoinIdenticalVertices only re-partition meshes; the un-merged glTF
partition is used (does not change the estimator's expectation).
"""
import json
import os
import numpy as np

MATERIAL_DTYPE = np.dtype([
    ("BaseColor", "<f4", 3), ("EmissiveColor", "<f4", 3), ("SpecularColor", "<f4", 3),
    ("MediumColor", "<f4", 3), ("MediumEmissiveColor", "<f4", 3),
    ("Metallic", "<f4"), ("Roughness", "<f4"), ("IOR", "<f4"), ("Transmission", "<f4"),
    ("Anisotropy", "<f4"), ("AnisotropyRotation", "<f4"), ("MediumDensity", "<f4"), ("MediumAnisotropy", "<f4"),
    ("BaseColorTextureIndex", "<u4"), ("NormalTextureIndex", "<u4"), ("RoughnessTextureIndex", "<u4"),
    ("MetallicTextureIndex", "<u4"), ("EmissiveTextureIndex", "<u4"),
])
assert MATERIAL_DTYPE.itemsize == 112
VERTEX_DTYPE = np.dtype([("pos", "<f4", 3), ("nrm", "<f4", 3), ("uv", "<f4", 2)])
assert VERTEX_DTYPE.itemsize == 32

_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}
FLIP_Y = np.diag([1.0, -1.0, 1.0, 1.0]).astype(np.float32)


def _accessor(g, buffers, idx):
    a = g["accessors"][idx]
    bv = g["bufferViews"][a["bufferView"]]
    dt = np.dtype(_COMP[a["componentType"]]).newbyteorder("<")
    nc = _NCOMP[a["type"]]
    off = bv.get("byteOffset", 0) + a.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * nc
    buf = buffers[bv["buffer"]]
    out = np.empty((a["count"], nc), dtype=dt)
    for i in range(nc):
        out[:, i] = np.ndarray((a["count"],), dtype=dt, buffer=buf, offset=off + i * dt.itemsize, strides=(stride,))
    if a.get("normalized", False) and dt.kind in "ui":
        out = out.astype(np.float32) / np.float32(np.iinfo(dt).max)
    return out


def _quat_to_mat(q):
    x, y, z, w = [np.float32(v) for v in q]
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = 1 - 2 * (y * y + z * z); m[0, 1] = 2 * (x * y - z * w); m[0, 2] = 2 * (x * z + y * w)
    m[1, 0] = 2 * (x * y + z * w); m[1, 1] = 1 - 2 * (x * x + z * z); m[1, 2] = 2 * (y * z - x * w)
    m[2, 0] = 2 * (x * z - y * w); m[2, 1] = 2 * (y * z + x * w); m[2, 2] = 1 - 2 * (x * x + y * y)
    return m


def _node_local(n):
    if "matrix" in n:
        return np.array(n["matrix"], dtype=np.float32).reshape(4, 4).T  # glTF column-major
    t = np.eye(4, dtype=np.float32); r = np.eye(4, dtype=np.float32); s = np.eye(4, dtype=np.float32)
    if "translation" in n: t[:3, 3] = np.array(n["translation"], dtype=np.float32)
    if "rotation" in n: r = _quat_to_mat(n["rotation"])
    if "scale" in n: s[0, 0], s[1, 1], s[2, 2] = [np.float32(v) for v in n["scale"]]
    return (t @ r @ s).astype(np.float32)


def _decode_image(path):
    from PIL import Image
    im = Image.open(path).convert("RGBA")      # stbi_load(..., STBI_rgb_alpha)
    return np.asarray(im, dtype=np.uint8).copy()


def load_gltf(path, decode_image=_decode_image):
    """Returns a dict: meshes [(verts(VERTEX_DTYPE), indices u32)], materials (MATERIAL_DTYPE),
    material_names, textures [HxWxC u8], instances [(transform col-major f32[16], mesh, material)],
    camera_view (f32[16] col-major) , aspect."""
    raw = open(path, "rb").read()
    glb_bin = None
    if raw[:4] == b"glTF":                                  # binary glTF container: JSON chunk + optional BIN chunk (= buffer 0)
        import struct
        total = min(struct.unpack_from("<I", raw, 8)[0], len(raw)); o = 12; jtxt = None
        while o + 8 <= total:
            ln, ty = struct.unpack_from("<II", raw, o); o += 8
            if ty == 0x4E4F534A and jtxt is None: jtxt = raw[o:o + ln]
            elif ty == 0x004E4942 and glb_bin is None: glb_bin = raw[o:o + ln]
            o += (ln + 3) & ~3
        g = json.loads(jtxt.decode("utf-8"))
    else:
        g = json.loads(raw.decode("utf-8"))
    base = os.path.dirname(os.path.abspath(path))

    def _buffer(i, b):
        uri = b.get("uri")
        if uri is None: return glb_bin
        if uri.startswith("data:"):
            import base64
            return base64.b64decode(uri.split(";base64,", 1)[1])
        return open(os.path.join(base, uri), "rb").read()
    buffers = [_buffer(i, b) for i, b in enumerate(g["buffers"])]

    # ---- meshes: one per primitive
    meshes, prim_of_mesh, mesh_material = [], [], []
    for m in g["meshes"]:
        ids = []
        for p in m["primitives"]:
            at = p["attributes"]
            pos = _accessor(g, buffers, at["POSITION"]).astype(np.float32)
            n = len(pos)
            v = np.zeros(n, dtype=VERTEX_DTYPE)
            v["pos"] = pos
            if "indices" in p:
                idx = _accessor(g, buffers, p["indices"]).astype(np.uint32).reshape(-1)
            else:
                idx = np.arange(n, dtype=np.uint32)
            mode = p.get("mode", 4)
            if mode == 5:                                    # TRIANGLE_STRIP -> faces, odd faces swapped to keep the orientation (assimp glTF2 importer)
                idx = np.array([(idx[f + 1], idx[f], idx[f + 2]) if (f + 1) % 2 == 0 else (idx[f], idx[f + 1], idx[f + 2]) for f in range(max(len(idx) - 2, 0))], np.uint32).reshape(-1)
            elif mode == 6:                                  # TRIANGLE_FAN
                idx = np.array([(idx[0], idx[f + 1], idx[f + 2]) for f in range(max(len(idx) - 2, 0))], np.uint32).reshape(-1)
            idx = idx[:len(idx) // 3 * 3]
            if len(idx) == 0 or n == 0: raise ValueError("primitive without triangles")
            if "NORMAL" in at:
                nr = _accessor(g, buffers, at["NORMAL"]).astype(np.float32)
            else:
                # aiProcess_GenNormals = assimp 6.0.2 GenFaceNormalsProcess (upstream knowledge): face normal written to the
                # face's three vertices in face order -- a shared vertex keeps the LAST face's normal; unreferenced vertices stay 0
                nr = np.zeros((n, 3), np.float32)
                for f in range(0, len(idx), 3):
                    a, b, c = pos[idx[f]], pos[idx[f + 1]], pos[idx[f + 2]]
                    e1 = (b - a).astype(np.float32); e2 = (c - a).astype(np.float32)
                    nx = np.float32(e1[1] * e2[2]) - np.float32(e1[2] * e2[1]); ny = np.float32(e1[2] * e2[0]) - np.float32(e1[0] * e2[2]); nz = np.float32(e1[0] * e2[1]) - np.float32(e1[1] * e2[0])
                    ln = np.sqrt(np.float32(np.float32(nx * nx) + np.float32(ny * ny)) + np.float32(nz * nz), dtype=np.float32)
                    if ln > 0: nx, ny, nz = np.float32(nx / ln), np.float32(ny / ln), np.float32(nz / ln)
                    nr[idx[f]] = nr[idx[f + 1]] = nr[idx[f + 2]] = (nx, ny, nz)
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = (np.float32(1.0) / np.sqrt((nr[:, 0] * nr[:, 0] + nr[:, 1] * nr[:, 1] + nr[:, 2] * nr[:, 2]).astype(np.float32))).astype(np.float32)
                v["nrm"] = nr * inv[:, None]                      # glm::normalize (AssetImporterImpl.cpp:165)
            if "TEXCOORD_0" in at:
                v["uv"] = _accessor(g, buffers, at["TEXCOORD_0"]).astype(np.float32)   # importer flip + FlipUVs = identity
            if p.get("mode", 4) not in (4, 5, 6):
                raise NotImplementedError("only TRIANGLES / TRIANGLE_STRIP / TRIANGLE_FAN primitives")
            ids.append(len(meshes))
            meshes.append((v, idx))
            mesh_material.append(p.get("material", None))
        prim_of_mesh.append(ids)

    # ---- materials (AssetImporterImpl.cpp:353-455 on top of assimp's glTF2 key mapping)
    images = g.get("images", [])
    textures_g = g.get("textures", [])

    def tex_path(ti):
        if ti is None: return ""
        src = textures_g[ti["index"]].get("source")
        if src is None: return ""
        return base + "/" + images[src]["uri"]

    mats_src = list(g.get("materials", []))
    needs_default = any(m is None for m in mesh_material)
    if needs_default: mats_src.append({"name": "DefaultMaterial"})
    mats = np.zeros(len(mats_src), dtype=MATERIAL_DTYPE)
    tex_paths = []
    names = []
    for i, m in enumerate(mats_src):
        pbr = m.get("pbrMetallicRoughness", {})
        ext = m.get("extensions", {})
        o = mats[i]
        o["BaseColor"] = np.array(pbr.get("baseColorFactor", [1, 1, 1, 1])[:3], dtype=np.float32)
        strength = ext.get("KHR_materials_emissive_strength", {}).get("emissiveStrength", 1.0)
        o["EmissiveColor"] = np.array(m.get("emissiveFactor", [0, 0, 0]), dtype=np.float32) * np.float32(strength)
        o["SpecularColor"] = np.array(ext["KHR_materials_specular"].get("specularColorFactor", [1, 1, 1]), dtype=np.float32) \
            if "KHR_materials_specular" in ext else np.ones(3, np.float32)
        o["MediumColor"] = 1.0; o["MediumEmissiveColor"] = 0.0          # PT/PathTracer.h:16-17 defaults
        o["Metallic"] = np.float32(pbr.get("metallicFactor", 1.0))       # glTF default 1.0 (assimp always sets the key)
        o["Roughness"] = np.float32(pbr.get("roughnessFactor", 1.0))
        o["IOR"] = np.float32(ext.get("KHR_materials_ior", {}).get("ior", 1.5))
        o["Transmission"] = np.float32(ext.get("KHR_materials_transmission", {}).get("transmissionFactor", 0.0))
        an = ext.get("KHR_materials_anisotropy", {})
        o["Anisotropy"] = np.float32(an.get("anisotropyStrength", 0.0))
        o["AnisotropyRotation"] = np.float32(an.get("anisotropyRotation", 0.0)) * (np.float32(180.0) / np.float32(np.pi))
        o["MediumDensity"] = 0.0; o["MediumAnisotropy"] = 0.0
        mr = tex_path(pbr.get("metallicRoughnessTexture"))
        tex_paths.append(dict(base=tex_path(pbr.get("baseColorTexture")), normal=tex_path(m.get("normalTexture")),
                              rough=mr, metal=mr, emissive=tex_path(m.get("emissiveTexture"))))
        names.append(m.get("name", ""))

    # ---- texture table, PathTracer.cpp:228-332 order; defaults PathTracer.cpp:1557-1621
    textures, index_of = [], {}

    def get_tex(path, key_default, default_px, single):
        key = path if path else key_default
        if key not in index_of:
            index_of[key] = len(textures)
            if path:
                img = decode_image(path)
                if single: img = img[:, :, :1].copy()             # R channel only (PathTracer.cpp:826-836, Q8)
            else:
                img = np.array(default_px, dtype=np.uint8).reshape(1, 1, -1)
            textures.append(np.ascontiguousarray(img))
        return index_of[key]

    for i, tp in enumerate(tex_paths):
        mats[i]["BaseColorTextureIndex"] = get_tex(tp["base"], "EMPTY_BASECOLOR_TEXTURE", [255, 255, 255, 255], False)
        mats[i]["NormalTextureIndex"] = get_tex(tp["normal"], "EMPTY_NORMAL_TEXTURE", [128, 128, 255, 255], False)
        mats[i]["RoughnessTextureIndex"] = get_tex(tp["rough"], "EMPTY_ROUGHNESS_TEXTURE", [255], True)
        mats[i]["MetallicTextureIndex"] = get_tex(tp["metal"], "EMPTY_METALLIC_TEXTURE", [255], True)
        mats[i]["EmissiveTextureIndex"] = get_tex(tp["emissive"], "EMPTY_EMISSIVE_TEXTURE", [255, 255, 255, 255], False)

    # ---- nodes -> instances + camera
    instances = []
    cam = {"view": None, "aspect": None}
    nodes = g["nodes"]
    default_mat = len(mats_src) - 1

    def visit(ni, parent):
        n = nodes[ni]
        world = (parent @ _node_local(n)).astype(np.float32)
        if "mesh" in n:
            for mi in prim_of_mesh[n["mesh"]]:
                mm = mesh_material[mi]
                instances.append(((FLIP_Y @ world).astype(np.float32).T.reshape(-1).copy(), mi, default_mat if mm is None else mm))
        if "camera" in n and cam["view"] is None and n["camera"] == 0:
            c = g["cameras"][n["camera"]]
            # assimp camera: lookAt (0,0,-1), up (0,1,0), pos 0 -> right (1,0,0), up' = -cross(right, lookAt) = (0,-1,0)
            local = np.eye(4, dtype=np.float32)
            local[:3, 0] = [1, 0, 0]; local[:3, 1] = [0, -1, 0]; local[:3, 2] = [0, 0, 1]
            final = (FLIP_Y @ world @ local).astype(np.float32)
            cam["view"] = np.linalg.inv(final.astype(np.float64)).astype(np.float32).T.reshape(-1).copy()
            asp = c.get("perspective", {}).get("aspectRatio", 0.0)
            cam["aspect"] = float(asp) if asp and asp > 0 else 1.0
        for ch in n.get("children", []):
            visit(ch, world)

    scene = g["scenes"][g.get("scene", 0)]
    for r in scene["nodes"]:
        visit(r, np.eye(4, dtype=np.float32))

    if cam["view"] is None:                                     # PathTracer.cpp:171-178 default camera
        eye = np.array([0, 0, 5.0]); f = np.array([0, 0, -1.0]); s = np.array([1.0, 0, 0]); u = np.array([0, 1.0, 0])
        view = np.eye(4); view[0, :3] = s; view[1, :3] = u; view[2, :3] = -f
        view[0, 3] = -s @ eye; view[1, 3] = -u @ eye; view[2, 3] = f @ eye
        cam["view"] = view.astype(np.float32).T.reshape(-1).copy(); cam["aspect"] = 16.0 / 9.0

    return dict(meshes=meshes, materials=mats, material_names=names, textures=textures, instances=instances,
                camera_view=cam["view"], aspect=np.float32(cam["aspect"]))


# ------------------------------------------------------------------------------------------------
# Wavefront OBJ / MTL: oracle-side mirror of csrc/scene_loader.cpp:load_obj_scene (assimp ObjFileParser / ObjFileMtlImporter semantics from
# upstream knowledge -- PARITY UNPINNED, see the C++ header comment for the mapping)
# ------------------------------------------------------------------------------------------------
def _obj_lines(path):
    raw = open(path, "rb").read().decode("latin-1")
    out, cur = [], ""
    for ln in raw.split("\n"):
        if ln.endswith("\r"): ln = ln[:-1]
        if ln.endswith("\\"): cur += ln[:-1] + " "; continue
        cur += ln
        h = cur.find("#")
        if h >= 0: cur = cur[:h]
        out.append(cur); cur = ""
    if cur: out.append(cur)
    return out


def _obj_default_material():
    m = np.zeros(1, dtype=MATERIAL_DTYPE)[0]
    m["BaseColor"] = 0.6; m["SpecularColor"] = 0.0; m["MediumColor"] = 1.0; m["Roughness"] = 1.0; m["IOR"] = 1.0
    return m


def load_obj(path, decode_image=_decode_image):
    base = os.path.dirname(os.path.abspath(path))
    atof = lambda t: np.float32(_c_atof(t))
    P, N, T = [], [], []
    mats = [dict(name="DefaultMaterial", m=_obj_default_material(), tp=dict(base="", normal="", rough="", metal="", emissive=""))]

    def load_mtl(fp):
        cur = None
        for ln in _obj_lines(fp):
            t = ln.split()
            if not t: continue
            if t[0] == "newmtl":
                cur = dict(name=t[1] if len(t) > 1 else "", m=_obj_default_material(), tp=dict(base="", normal="", rough="", metal="", emissive="")); mats.append(cur); continue
            if cur is None or len(t) < 2: continue
            k = t[0]; f3 = lambda: np.array([atof(t[min(1 + c, len(t) - 1)]) for c in range(3)], np.float32); fp2 = base + "/" + t[-1]
            if k == "Kd": cur["m"]["BaseColor"] = f3()
            elif k == "Ke": cur["m"]["EmissiveColor"] = f3()
            elif k == "Ks": cur["m"]["SpecularColor"] = f3()
            elif k == "Ni": cur["m"]["IOR"] = atof(t[1])
            elif k == "Pr": cur["m"]["Roughness"] = atof(t[1])
            elif k == "Pm": cur["m"]["Metallic"] = atof(t[1])
            elif k == "aniso": cur["m"]["Anisotropy"] = atof(t[1])
            elif k == "anisor": cur["m"]["AnisotropyRotation"] = atof(t[1]) * (np.float32(180.0) / np.float32(3.14159265358979323846))
            elif k == "map_Kd": cur["tp"]["base"] = fp2
            elif k in ("norm", "map_Kn"): cur["tp"]["normal"] = fp2
            elif k == "map_Pr": cur["tp"]["rough"] = fp2
            elif k == "map_Pm": cur["tp"]["metal"] = fp2
            elif k == "map_Ke": cur["tp"]["emissive"] = fp2

    builds = []; group = ""; material = 0; need_new = True
    for ln in _obj_lines(path):
        t = ln.split()
        if not t: continue
        k = t[0]
        if k == "v" and len(t) >= 4: P.append([atof(x) for x in t[1:4]])
        elif k == "vn" and len(t) >= 4: N.append([atof(x) for x in t[1:4]])
        elif k == "vt" and len(t) >= 2: T.append([atof(t[1]), atof(t[2]) if len(t) >= 3 else np.float32(0)])
        elif k in ("o", "g"):
            group = t[1] if len(t) > 1 else ""
            if builds and len(builds[-1]["idx"]): need_new = True
            elif builds: builds[-1]["name"] = group
        elif k == "mtllib":
            for f in t[1:]: load_mtl(base + "/" + f)
        elif k == "usemtl":
            found = 0
            for i in range(1, len(mats)):
                if len(t) > 1 and mats[i]["name"] == t[1]: found = i
            if found != material:
                material = found
                if builds and len(builds[-1]["idx"]): need_new = True
                elif builds: builds[-1]["material"] = material
        elif k == "f" and len(t) >= 4:
            if need_new or not builds:
                builds.append(dict(name=group, material=material, index={}, verts=[], idx=[])); need_new = False
            b = builds[-1]
            fv = []
            for tok in t[1:]:
                parts = (tok.split("/") + ["", ""])[:3]
                v, vt, vn = [int(x) if x not in ("",) and _is_int(x) else 0 for x in parts]
                if v < 0: v = len(P) + v + 1
                if vt < 0: vt = len(T) + vt + 1
                if vn < 0: vn = len(N) + vn + 1
                if v < 1 or v > len(P) or vt > len(T) or vn > len(N): raise ValueError("OBJ face index out of range")
                fv.append((v, vt, vn))
            for i in range(1, len(fv) - 1):
                tri = [fv[0], fv[i], fv[i + 1]]
                gn = (np.float32(0), np.float32(0), np.float32(0))
                if any(q[2] == 0 for q in tri):
                    a, bb, c = (np.array(P[q[0] - 1], np.float32) for q in tri)
                    e1 = (bb - a).astype(np.float32); e2 = (c - a).astype(np.float32)
                    g = np.array([np.float32(e1[1] * e2[2]) - np.float32(e1[2] * e2[1]), np.float32(e1[2] * e2[0]) - np.float32(e1[0] * e2[2]), np.float32(e1[0] * e2[1]) - np.float32(e1[1] * e2[0])], np.float32)
                    ln2 = np.sqrt(np.float32(np.float32(g[0] * g[0]) + np.float32(g[1] * g[1])) + np.float32(g[2] * g[2]), dtype=np.float32)
                    if ln2 > 0: g = (g / ln2).astype(np.float32)
                    gn = tuple(g)
                for q in tri:
                    key = (q[0], q[1], q[2]) + (tuple(float(x) for x in gn) if q[2] == 0 else (0.0, 0.0, 0.0))
                    if key not in b["index"]:
                        b["index"][key] = len(b["verts"])
                        nrm = np.array(N[q[2] - 1], np.float32) if q[2] else np.array(gn, np.float32)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            inv = np.float32(1.0) / np.sqrt(np.float32(np.float32(nrm[0] * nrm[0]) + np.float32(nrm[1] * nrm[1])) + np.float32(nrm[2] * nrm[2]), dtype=np.float32)
                            nrm = (nrm * inv).astype(np.float32)
                        uv = (T[q[1] - 1][0], np.float32(1.0) - T[q[1] - 1][1]) if q[1] else (np.float32(0), np.float32(0))
                        b["verts"].append((tuple(P[q[0] - 1]), tuple(nrm), uv))
                    b["idx"].append(b["index"][key])
    meshes, instances = [], []
    for b in builds:
        if not b["idx"]: continue
        v = np.zeros(len(b["verts"]), dtype=VERTEX_DTYPE)
        for i, (p_, n_, uv_) in enumerate(b["verts"]): v[i]["pos"] = p_; v[i]["nrm"] = n_; v[i]["uv"] = uv_
        instances.append((FLIP_Y.astype(np.float32).T.reshape(-1).copy(), len(meshes), b["material"]))
        meshes.append((v, np.array(b["idx"], np.uint32)))
    if not meshes: raise ValueError("No meshes found in scene")
    marr = np.zeros(len(mats), dtype=MATERIAL_DTYPE)
    for i, m in enumerate(mats): marr[i] = m["m"]
    textures, index_of = [], {}

    def get_tex(path_, key_default, default_px, single):
        key = path_ if path_ else key_default
        if key not in index_of:
            index_of[key] = len(textures)
            if path_:
                img = decode_image(path_)
                if single: img = img[:, :, :1].copy()
            else:
                img = np.array(default_px, dtype=np.uint8).reshape(1, 1, -1)
            textures.append(np.ascontiguousarray(img))
        return index_of[key]
    for i, m in enumerate(mats):
        tp = m["tp"]
        marr[i]["BaseColorTextureIndex"] = get_tex(tp["base"], "EMPTY_BASECOLOR_TEXTURE", [255, 255, 255, 255], False)
        marr[i]["NormalTextureIndex"] = get_tex(tp["normal"], "EMPTY_NORMAL_TEXTURE", [128, 128, 255, 255], False)
        marr[i]["RoughnessTextureIndex"] = get_tex(tp["rough"], "EMPTY_ROUGHNESS_TEXTURE", [255], True)
        marr[i]["MetallicTextureIndex"] = get_tex(tp["metal"], "EMPTY_METALLIC_TEXTURE", [255], True)
        marr[i]["EmissiveTextureIndex"] = get_tex(tp["emissive"], "EMPTY_EMISSIVE_TEXTURE", [255, 255, 255, 255], False)
    view = np.eye(4); view[2, 3] = -5.0
    return dict(meshes=meshes, materials=marr, material_names=[m["name"] for m in mats], textures=textures, instances=instances,
                camera_view=view.astype(np.float32).T.reshape(-1).copy(), aspect=np.float32(16.0 / 9.0))


def _is_int(x):
    try: int(x); return True
    except ValueError: return False


def _c_atof(t):
    """C atof(): longest valid numeric prefix, 0.0 if none."""
    import re
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", t)
    return float(m.group(0)) if m else 0.0


def load_scene(path, decode_image=_decode_image):
    """By extension, like AssetImporter::ImportScene."""
    return load_obj(path, decode_image) if path.lower().endswith(".obj") else load_gltf(path, decode_image)


# ------------------------------------------------------------------------------------------------
# golden-fixture (de)serialisation: flat npz so GPU-box tests need neither /root/reference nor PIL
# ------------------------------------------------------------------------------------------------
def save_scene_npz(path, sc):
    d = {}
    d["n_meshes"] = np.int64(len(sc["meshes"]))
    for i, (v, idx) in enumerate(sc["meshes"]):
        d[f"mesh{i}_v"] = v.view(np.float32).reshape(-1, 8); d[f"mesh{i}_i"] = idx
    d["materials"] = np.frombuffer(sc["materials"].tobytes(), dtype=np.uint8)
    d["material_names"] = np.array(sc["material_names"])
    d["n_textures"] = np.int64(len(sc["textures"]))
    for i, t in enumerate(sc["textures"]): d[f"tex{i}"] = t
    d["inst_xf"] = np.stack([x for x, _, _ in sc["instances"]]).astype(np.float32)
    d["inst_mesh"] = np.array([m for _, m, _ in sc["instances"]], dtype=np.uint32)
    d["inst_mat"] = np.array([m for _, _, m in sc["instances"]], dtype=np.uint32)
    d["camera_view"] = sc["camera_view"]; d["aspect"] = np.float32(sc["aspect"])
    np.savez_compressed(path, **d)


def load_scene_npz(path):
    z = np.load(path, allow_pickle=False)
    meshes = []
    for i in range(int(z["n_meshes"])):
        v = np.ascontiguousarray(z[f"mesh{i}_v"]).view(VERTEX_DTYPE).reshape(-1)
        meshes.append((v, np.ascontiguousarray(z[f"mesh{i}_i"])))
    mats = np.frombuffer(z["materials"].tobytes(), dtype=MATERIAL_DTYPE).copy()
    textures = [np.ascontiguousarray(z[f"tex{i}"]) for i in range(int(z["n_textures"]))]
    inst = [(np.ascontiguousarray(z["inst_xf"][i]), int(z["inst_mesh"][i]), int(z["inst_mat"][i])) for i in range(len(z["inst_mesh"]))]
    return dict(meshes=meshes, materials=mats, material_names=[str(s) for s in z["material_names"]], textures=textures,
                instances=inst, camera_view=np.ascontiguousarray(z["camera_view"]), aspect=np.float32(z["aspect"]))


def save_luts_npz(path, refl, rout, rin):
    """Lossless, byte-plane-shuffled (compresses ~35% better)."""
    def sh(a): return np.frombuffer(np.ascontiguousarray(a, dtype="<f4").tobytes(), dtype=np.uint8).reshape(-1, 4).T.copy()
    np.savez_compressed(path, refl=sh(refl), rout=sh(rout), rin=sh(rin))


def load_luts_npz(path):
    z = np.load(path)
    def un(b, shape): return np.frombuffer(np.ascontiguousarray(b.T).tobytes(), dtype="<f4").reshape(shape).copy()
    return un(z["refl"], (32, 64, 64)), un(z["rout"], (32, 128, 128)), un(z["rin"], (32, 128, 128))


def load_luts_dir(d):
    r = np.fromfile(os.path.join(d, "ReflectionLookup.bin"), dtype="<f4").reshape(32, 64, 64)
    o = np.fromfile(os.path.join(d, "RefractionLookupHitFromOutside.bin"), dtype="<f4").reshape(32, 128, 128)
    i = np.fromfile(os.path.join(d, "RefractionLookupHitFromInside.bin"), dtype="<f4").reshape(32, 128, 128)
    return r, o, i


def synthetic_env(w=512, h=256, seed=3, sun=5000.0):
    """Procedural HDR environment for GPU-box tests/bench (the 25 MB meadow_2_4k.hdr is not committed):
    sky gradient + ground + a small very bright 'sun' disc + seeded noise; linear float RGBA, alpha 1."""
    rng = np.random.default_rng(seed)
    v = (np.arange(h, dtype=np.float32) + 0.5) / h
    u = (np.arange(w, dtype=np.float32) + 0.5) / w
    V, U = np.meshgrid(v, u, indexing="ij")
    sky = np.stack([0.35 + 0.4 * (1 - V), 0.5 + 0.35 * (1 - V), 0.9 + 0.1 * V], -1)
    ground = np.stack([0.18 + 0 * V, 0.22 + 0 * V, 0.08 + 0 * V], -1)
    img = np.where((V < 0.52)[..., None], sky, ground).astype(np.float32)
    img *= (0.8 + 0.4 * rng.random((h, w, 1), dtype=np.float32))
    d2 = ((U - 0.62) * 2.0) ** 2 + (V - 0.23) ** 2
    img += (sun * np.exp(-d2 / (2 * 0.004 ** 2)))[..., None].astype(np.float32) * np.array([1.0, 0.93, 0.8], np.float32)
    out = np.ones((h, w, 4), dtype=np.float32)
    out[..., :3] = img
    return out
