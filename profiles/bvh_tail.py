import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
name = "breakfast_room"
T = util.product_tracer(name, 64, 64)
sc = util.scene_dict(name)
vi, pi = util.camera(name); VI = vi.reshape(4, 4).T; PI = pi.reshape(4, 4).T
W, H = 256, 144
xs, ys = np.meshgrid((np.arange(W) + 0.5) / W * 2 - 1, (np.arange(H) + 0.5) / H * 2 - 1)
tgt = np.stack([xs * PI[0, 0], ys * PI[1, 1], -np.ones_like(xs)], -1).reshape(-1, 3); tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
d = (tgt @ VI[:3, :3].T).astype(np.float32); o = np.tile(VI[:3, 3].astype(np.float32), (len(d), 1))
s = T.trace_stats(o, d, 0.01, 1e5); t, prim, inst, uv = T.trace_closest(o, d, 0.01, 1e5)
ntri = [len(i) // 3 for _, i in sc["meshes"]]
print("instances:", [(k, ntri[m]) for k, (_, m, _) in enumerate(sc["instances"])])
order = np.argsort(-s[:, 1].astype(int))[:8]
for r in order: print("ray", r % W, r // W, "nodes", s[r, 0], "tris", s[r, 1], "hit inst", int(inst[r]) if t[r] > 0 else -1, "t", t[r])
for k in range(len(sc["instances"])):
    m = (inst == k) & (t > 0)
    if m.sum(): print(f"inst {k:2d} tris {ntri[sc['instances'][k][1]]:6d}: rays {m.sum():6d} nodes mean {s[m,0].mean():6.1f} tris mean {s[m,1].mean():6.1f} max {s[m,1].max()}")
tri = util.oracle_scene(name).world_triangles()[0].reshape(-1, 3, 3)
ext = (tri.max(1) - tri.min(1)).max(1); area = 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
print("extent percentiles", np.percentile(ext, [1, 50, 99, 100]), "zero-area", (area == 0).sum(), "scene", tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0))
# duplicated triangles (same centroid)?
c = np.round(tri.mean(1), 5); u = np.unique(c, axis=0); print("unique centroids", len(u), "of", len(c))
