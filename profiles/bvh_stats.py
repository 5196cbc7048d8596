#!/usr/bin/env python
"""Traversal cost of the GPU LBVH per scene: nodes visited / triangles tested per ray for camera rays and random rays."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util
for name in sys.argv[1:] or ["cornell_box", "cornell_box_glass", "viking_room", "breakfast_room"]:
    T = util.product_tracer(name, 64, 64)
    S = util.oracle_scene(name)
    tri, _, _ = S.world_triangles(); lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    vi, pi = util.camera(name); VI = vi.reshape(4, 4).T; PI = pi.reshape(4, 4).T
    W, H = 256, 144
    xs, ys = np.meshgrid((np.arange(W) + 0.5) / W * 2 - 1, (np.arange(H) + 0.5) / H * 2 - 1)
    tgt = np.stack([xs * PI[0, 0], ys * PI[1, 1], -np.ones_like(xs)], -1).reshape(-1, 3); tgt /= np.linalg.norm(tgt, axis=1, keepdims=True)
    d = (tgt @ VI[:3, :3].T).astype(np.float32); o = np.tile(VI[:3, 3].astype(np.float32), (len(d), 1))
    rng = np.random.default_rng(0); n = len(d)
    o2 = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32); d2 = rng.normal(size=(n, 3)); d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
    st = T.scene_stats()
    for label, (oo, dd) in (("camera", (o, d)), ("random", (o2, d2))):
        s = T.trace_stats(oo, dd, 0.01, 1e5)
        print(f"{name:18s} {st['triangles']:7d} tris {st['bvh_nodes']:7d} nodes | {label}: nodes/ray mean {s[:,0].mean():7.1f} p99 {np.percentile(s[:,0],99):6.0f} max {s[:,0].max():5d} | tris/ray mean {s[:,1].mean():6.1f} max {s[:,1].max():4d}")
